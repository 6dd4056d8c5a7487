// tcgen05 (kind::tf32) weight gradient of 1x1 convolutions (stride 1 or 2, no padding) -- two thirds of the layers of
// keras.applications.ResNet50 (reference utils.py:237) and the shortcut projections of wide_residual_network.py:28:
//   dW[ci, co] += sum_pixels X[pixel, ci] * dY[pixel, co]        dbias[co] += sum_pixels dY[pixel, co]
// (stride 2: X is read through a tensor map of the sub-sampled view x[:, ::2, ::2, :], tiles are rows of the output grid
// in power-of-two row slots, zero-filled past the image)
//
// A GEMM whose reduction dimension is the pixel axis of two NHWC tensors: both operands are MN-major (channels
// contiguous), read as they lie in memory through SWIZZLE_128B_ATOM_32B boxes of 32 channels x PT pixels (see
// conv_wgrad_tc.cu for the layout).  One CTA owns 128 input channels (the four 32-lane quarters of an M = 128
// instruction, LBO = one channel block) x up to 128 output channels and walks its share of the flat pixel list with the
// accumulator resident in TMEM (split-K over CTAs, 16-byte reductions into dW at the end).  The bias gradient is one more
// MMA per k-step with an all-ones A tile.
// X3 (error-compensated, see conv_tc.cu): dW = X_hi*dY_hi + X_hi*dY_lo + X_lo*dY_hi.  Warp 2 writes dY_lo into a second
// dY buffer of the stage, warps 3-5 rewrite the X blocks in place as X_lo once pass 1 has read them, the MMA warp runs
// the two passes as two cursors over the stage sequence.
// grid = (pixel-tile groups, input-channel chunks of 128, output-channel chunks of <= 128)
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace se {

using namespace tc;

struct Wg1Params {
  int Cin, Cout;
  int PT;                      // pixels per pipeline stage
  int nnb, ncols;              // dY: 32-channel blocks per CTA, MMA N = 32 * nnb
  int G;                       // accumulators: dW (+1 for the bias gradient)
  int stages, stage_bytes, x_bytes, dy_bytes;
  int tiles;
  int grid4;                   // 1: 4-d tensor maps over an (N, Ho, Wo) grid, a tile = Hb rows of one image (stride 2)
  int Hb, tpi;
  float* dw;
  float* dbias;
};

__device__ __forceinline__ float wg1_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <int X3>
__global__ void __launch_bounds__(192, 1)
conv1x1_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, Wg1Params p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;
  uint8_t* ones = tiles + (size_t)p.stages * p.stage_bytes;               // 8 pixels x 32 channels of 1.0f (G == 2)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + 1024);
  uint64_t* full = bars;
  uint64_t* empty = bars + 4;
  uint64_t* done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  uint64_t* ylo_ready = bars + 10;          // X3: dY_lo of the stage written
  uint64_t* hi_done = bars + 14;            // X3: pass-1 MMAs have read the stage
  uint64_t* lo_ready = bars + 18;           // X3: the X blocks of the stage hold X_lo

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int per_cta = (p.tiles + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per_cta;
  const int t_end = min(p.tiles, t_begin + per_cta);
  const int ci0 = blockIdx.y * 128;
  const int co0 = blockIdx.z * p.ncols;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < p.G * p.ncols) tmem_cols <<= 1;

  {
    float4* o = reinterpret_cast<float4*>(ones);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) o[i] = make_float4(1.f, 1.f, 1.f, 1.f);
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_x); prefetch_tmap(&map_dy);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    if (X3) for (int s = 0; s < p.stages; ++s) { mbar_init(&ylo_ready[s], 32); mbar_init(&hi_done[s], 1); mbar_init(&lo_ready[s], 96); }
    fence_barrier_init();
  }
  fence_proxy_async();                      // generic-proxy writes of the ones tile -> visible to the tensor core
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();                               // nothing above touches global memory (see common.cuh)

  const int blk_bytes = p.PT * 128;         // one 32-channel block of a stage
  if (t_begin < t_end) {
    if (warp == 0) {
      // ===================== TMA producer (convergent warp, one elected lane issues)
      int stage = 0, phase = 0;
      const uint32_t tx = p.x_bytes + p.dy_bytes;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sb = tiles + (size_t)stage * p.stage_bytes;
        if (elect_one()) {
          mbar_expect_tx(&full[stage], tx);
          // channel blocks past Cin / Cout are zero-filled by the hardware (no memory traffic)
          if (p.grid4) {
            const int n0 = t / p.tpi, h0 = (t - n0 * p.tpi) * p.Hb;
            for (int b = 0; b < 4; ++b) tma_load_4d(sb + b * blk_bytes, &map_x, &full[stage], ci0 + b * 32, 0, h0, n0);
            for (int nb = 0; nb < p.nnb; ++nb)
              tma_load_4d(sb + p.x_bytes + nb * blk_bytes, &map_dy, &full[stage], co0 + nb * 32, 0, h0, n0);
          } else {
            for (int b = 0; b < 4; ++b) tma_load_2d(sb + b * blk_bytes, &map_x, &full[stage], ci0 + b * 32, t * p.PT);
            for (int nb = 0; nb < p.nnb; ++nb)
              tma_load_2d(sb + p.x_bytes + nb * blk_bytes, &map_dy, &full[stage], co0 + nb * 32, t * p.PT);
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (convergent warp, one elected lane issues)
      // instruction descriptor: tf32 x tf32 -> f32, A and B both MN-major (bits 15 / 16), M = 128, N = ncols
      const uint32_t idesc = umma_idesc(2, 128, p.ncols) | (1u << 15) | (1u << 16);
      // descriptor high word: SBO = 512 B (consecutive 4-pixel atoms), version 1, SWIZZLE_128B_BASE32B
      const uint32_t hi = (512u >> 4) | (1u << 14) | (1u << 29);
      const uint32_t lbo = (((uint32_t)blk_bytes >> 4) & 0x3FFFu) << 16;          // next 32-channel block (A and B)
      const uint32_t ones_lo = (smem_u32(ones) & 0x3FFFFu) >> 4;                  // LBO 0, the same 8 pixels for every k-step
      const uint32_t tiles_u32 = smem_u32(tiles);
      const int ksteps = p.PT / 8;
      const uint32_t d_w = tmem_base, d_b = tmem_base + p.ncols;
      uint32_t acc = 0;
      if (X3 == 1) {
        const int T = t_end - t_begin;
        int u1 = 0, s1 = 0, ph1 = 0, u2 = 0, s2 = 0, ph2 = 0;
        while (u2 < T) {
          int ok1 = 0;
          if (u1 < T) ok1 = mbar_try_wait(&full[s1], ph1) && mbar_try_wait(&ylo_ready[s1], ph1);
          ok1 = __shfl_sync(0xffffffffu, ok1, 0);          // one decision for the warp (elect_one needs convergence)
          if (ok1) {
            fence_after_sync();
            const uint32_t sb = tiles_u32 + (uint32_t)s1 * p.stage_bytes;
            const uint32_t dyb = sb + p.x_bytes, dyl = dyb + p.dy_bytes;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t da = ((uint64_t)hi << 32) | (uint64_t)((((sb + ks * 1024) & 0x3FFFFu) >> 4) | lbo);
              const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)((((dyb + ks * 1024) & 0x3FFFFu) >> 4) | lbo);
              const uint64_t dl = ((uint64_t)hi << 32) | (uint64_t)((((dyl + ks * 1024) & 0x3FFFFu) >> 4) | lbo);
              if (elect_one()) {
                mma_tf32(d_w, da, db, idesc, acc);
                if (p.G == 2) mma_tf32(d_b, ((uint64_t)hi << 32) | (uint64_t)ones_lo, db, idesc, acc);
                mma_tf32(d_w, da, dl, idesc, 1);
                if (p.G == 2) mma_tf32(d_b, ((uint64_t)hi << 32) | (uint64_t)ones_lo, dl, idesc, 1);
              }
              __syncwarp();
              acc = 1;
            }
            if (elect_one()) mma_commit(&hi_done[s1]);
            __syncwarp();
            ++u1;
            if (++s1 == p.stages) { s1 = 0; ph1 ^= 1; }
          }
          int ok2 = 0;
          if (u2 < u1) ok2 = mbar_try_wait(&lo_ready[s2], ph2);
          ok2 = __shfl_sync(0xffffffffu, ok2, 0);
          if (ok2) {
            fence_after_sync();
            const uint32_t sb = tiles_u32 + (uint32_t)s2 * p.stage_bytes;
            const uint32_t dyb = sb + p.x_bytes;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t da = ((uint64_t)hi << 32) | (uint64_t)((((sb + ks * 1024) & 0x3FFFFu) >> 4) | lbo);
              const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)((((dyb + ks * 1024) & 0x3FFFFu) >> 4) | lbo);
              if (elect_one()) mma_tf32(d_w, da, db, idesc, 1);
              __syncwarp();
            }
            if (elect_one()) mma_commit(&empty[s2]);
            __syncwarp();
            ++u2;
            if (++s2 == p.stages) { s2 = 0; ph2 ^= 1; }
          }
        }
      }
      int stage = 0, phase = 0;
      for (int t = t_begin; !X3 && t < t_end; ++t) {
        mbar_wait(&full[stage], phase);
        fence_after_sync();
        const uint32_t sb = tiles_u32 + (uint32_t)stage * p.stage_bytes;
        const uint32_t dyb = sb + p.x_bytes;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t da = ((uint64_t)hi << 32) | (uint64_t)((((sb + ks * 1024) & 0x3FFFFu) >> 4) | lbo);
          const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)((((dyb + ks * 1024) & 0x3FFFFu) >> 4) | lbo);
          if (elect_one()) {
            mma_tf32(d_w, da, db, idesc, acc);
            if (p.G == 2) mma_tf32(d_b, ((uint64_t)hi << 32) | (uint64_t)ones_lo, db, idesc, acc);
          }
          __syncwarp();
          acc = 1;
        }
        if (elect_one()) mma_commit(&empty[stage]);
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) mma_commit(done);
      __syncwarp();
    } else {
      // ===================== operand splitters (X3), then the epilogue: TMEM -> reductions into dW / dbias
      const int q4 = warp & 3;                              // TMEM lane quarter == 32-channel block of the CTA's 128 input channels
      if (X3 == 1) {
        int stage = 0, phase = 0;
        for (int t = t_begin; t < t_end; ++t) {
          uint8_t* sb = tiles + (size_t)stage * p.stage_bytes;
          mbar_wait(&full[stage], phase);
          if (warp == 2) {
            const float4* src = reinterpret_cast<const float4*>(sb + p.x_bytes);
            float4* dst = reinterpret_cast<float4*>(sb + p.x_bytes + p.dy_bytes);
            const int n = p.dy_bytes >> 4;
#pragma unroll 4
            for (int i = lane; i < n; i += 32) {
              float4 v = src[i];
              v.x = wg1_lo(v.x); v.y = wg1_lo(v.y); v.z = wg1_lo(v.z); v.w = wg1_lo(v.w);
              dst[i] = v;
            }
            fence_proxy_async();
            mbar_arrive(&ylo_ready[stage]);
          } else {
            mbar_wait(&hi_done[stage], phase);
            float4* q = reinterpret_cast<float4*>(sb);
            const int n = p.x_bytes >> 4;
#pragma unroll 4
            for (int i = (warp - 3) * 32 + lane; i < n; i += 96) {
              float4 v = q[i];
              v.x = wg1_lo(v.x); v.y = wg1_lo(v.y); v.z = wg1_lo(v.z); v.w = wg1_lo(v.w);
              q[i] = v;
            }
            fence_proxy_async();
            mbar_arrive(&lo_ready[stage]);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
      mbar_wait(done, 0);
      fence_after_sync();
      const int ci = ci0 + q4 * 32 + lane;
      const int ngrp = p.ncols / 16;
      for (int g = 0; g < p.G; ++g) {
        float* dst = nullptr;
        if (g == 0) { if (ci < p.Cin) dst = p.dw + (long long)ci * p.Cout + co0; }
        else if (q4 == 0 && lane == 0 && blockIdx.y == 0 && p.dbias) dst = p.dbias + co0;
        if (g == 1 && q4 != 0) break;                                // every row of the bias accumulator is the same sum
        for (int j = 0; j < ngrp; ++j) {
          const int c0 = ((j + blockIdx.x) % ngrp) * 16;             // the CTAs of a split-K group start at different columns
          const bool col_ok = co0 + c0 < p.Cout;                     // columns past Cout come from zero-filled channels
          uint32_t v[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(tmem_base + ((uint32_t)(q4 * 32) << 16) + g * p.ncols + c0)
              : "memory");
          tmem_ld_wait();
          if (dst && col_ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 val = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                       __uint_as_float(v[4 * q + 3]));
              atomicAdd(reinterpret_cast<float4*>(dst + c0 + 4 * q), val);
            }
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

int init_conv1x1_wgrad_tc() {
  if (cudaFuncSetAttribute(conv1x1_wgrad_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(conv1x1_wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
    set_error("init_conv1x1_wgrad_tc: cannot raise the shared-memory limit");
    return SE_ERR_CUDA;
  }
  return SE_OK;
}

bool conv1x1_wgrad_tc_ok(const se_conv_desc* d) {
  static const bool off = getenv("SE_CT_NO_1X1") != nullptr;
  static const bool off2 = getenv("SE_CT_NO_S2") != nullptr;
  if (off || d->kh != 1 || d->kw != 1 || d->pad_t != 0 || d->pad_l != 0) return false;
  if (d->stride == 2) {
    if (off2 || d->Ho != (d->H + 1) / 2 || d->Wo != (d->W + 1) / 2 || d->Wo > 32) return false;
  } else if (d->stride != 1 || d->Ho != d->H || d->Wo != d->W) {
    return false;
  }
  // 16-byte reductions into dW rows of Cout floats; channel counts in whole 16-byte units for the tensor maps
  return d->Cin % 4 == 0 && d->Cout % 16 == 0 && (long long)d->N * d->H * d->W >= 32 &&
         (long long)d->N * d->H * d->W <= 0x7fffffffLL;
}

// x_view_w / x_view_h (stride 2 only, 0 = the full Wo x Ho grid): extent of the sub-sampled view that starts at x -- a
// caller that passes x + (r*W + s)*Cin reads x[:, r::2, s::2, :], whose last column / row may be missing; the missing
// entries are zero-filled (one tap of a 3x3 / stride 2 weight gradient, conv3x3s2_wgrad_tc below)
int conv1x1_wgrad_tc(const se_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias, int x3, cudaStream_t st,
                     int x_view_w = 0, int x_view_h = 0) {
  if (!conv1x1_wgrad_tc_ok(d)) return SE_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(dw) & 15) != 0 || (dbias && (reinterpret_cast<uintptr_t>(dbias) & 15) != 0))
    return SE_ERR_UNSUPPORTED;
  const int Cin = d->Cin, Cout = d->Cout;
  const long long npx = (long long)d->N * d->H * d->W;
  Wg1Params p;
  p.Cin = Cin; p.Cout = Cout;
  p.PT = 32;
  p.grid4 = d->stride == 2; p.Hb = 1; p.tpi = 1;
  int Wb = 8;
  while (Wb < d->Wo) Wb <<= 1;            // (stride 2) row pitch of the boxes
  const int gz = ceil_div(Cout, 128);
  p.nnb = ceil_div(ceil_div(Cout, gz), 32);
  p.ncols = 32 * p.nnb;
  p.G = dbias ? 2 : 1;
  p.x_bytes = 4 * p.PT * 128;
  p.dy_bytes = p.nnb * p.PT * 128;
  p.stage_bytes = ceil_div(p.x_bytes + (1 + x3) * p.dy_bytes, 1024) * 1024;
  p.stages = min(4, (200 * 1024) / p.stage_bytes);
  p.tiles = (int)ceil_div<long long>(npx, p.PT);
  if (p.grid4) { p.Hb = p.PT / Wb; p.tpi = ceil_div(d->Ho, p.Hb); p.tiles = d->N * p.tpi; }
  p.dw = dw; p.dbias = dbias;
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 + 24 * 8 + 1024 + 64;

  CUtensorMap mx, mdy;
  {
    uint64_t dims[2] = {(uint64_t)Cin, (uint64_t)npx};
    uint64_t strides[1] = {(uint64_t)Cin * 4};
    uint32_t box[2] = {32u, (uint32_t)p.PT};
    if (!make_tmap(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return SE_ERR_CUDA;
    uint64_t ydims[2] = {(uint64_t)Cout, (uint64_t)npx};
    uint64_t ystrides[1] = {(uint64_t)Cout * 4};
    if (!p.grid4 && !make_tmap(&mdy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(dy), ydims, ystrides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return SE_ERR_CUDA;
    if (p.grid4) {
      // x[:, ::2, ::2, :] (pixel and row strides doubled) and dy, both as (channels, Wo, Ho, N)
      uint64_t d4[4] = {(uint64_t)Cin, (uint64_t)(x_view_w ? x_view_w : d->Wo), (uint64_t)(x_view_h ? x_view_h : d->Ho), (uint64_t)d->N};
      uint64_t s4[3] = {(uint64_t)2 * Cin * 4, (uint64_t)2 * d->W * Cin * 4, (uint64_t)d->H * d->W * Cin * 4};
      uint32_t b4[4] = {32u, (uint32_t)Wb, (uint32_t)p.Hb, 1u};
      if (!make_tmap(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), d4, s4, b4, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
        return SE_ERR_CUDA;
      uint64_t y4[4] = {(uint64_t)Cout, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->N};
      uint64_t ys4[3] = {(uint64_t)Cout * 4, (uint64_t)d->Wo * Cout * 4, (uint64_t)d->Ho * d->Wo * Cout * 4};
      if (!make_tmap(&mdy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dy), y4, ys4, b4, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
        return SE_ERR_CUDA;
    }
  }
  const int gy = ceil_div(Cin, 128);
  const int gx = max(1, min(p.tiles, sm_count() / (gy * gz)));
  if (x3) launch(conv1x1_wgrad_tc_kernel<1>, dim3(gx, gy, gz), dim3(192), smem, st, mx, mdy, p);
  else launch(conv1x1_wgrad_tc_kernel<0>, dim3(gx, gy, gz), dim3(192), smem, st, mx, mdy, p);
  return check_launch("conv1x1_wgrad_tc_kernel");
}

// 3x3 / stride 2 / no leading padding (the down-sampling layers of wide_residual_network.py:20-31 on even image sizes:
// 'same' puts the one padding row / column AFTER the image) as nine 1x1 / stride 2 weight gradients, one per filter tap:
//   dW[r, s] = x[:, r::2, s::2, :]^T dY            (views of x, no gather pass; the taps' outputs are disjoint)
// Worth it for wide layers only (160 -> 320 at 32x32, batch 64: 1.85 ms on the fp32 kernel); the 16..64-channel layers of
// the CIFAR ResNets are faster on the fp32 kernel than in nine latency-bound launches.
bool conv3x3s2_tc_ok(const se_conv_desc* d) {
  static const bool off = getenv("SE_CT_NO_S2") != nullptr || getenv("SE_CT_NO_3X3S2") != nullptr;
  return !off && d->kh == 3 && d->kw == 3 && d->stride == 2 && d->pad_t == 0 && d->pad_l == 0 && d->H % 2 == 0 && d->W % 2 == 0 &&
         d->Ho == d->H / 2 && d->Wo == d->W / 2 && d->Wo <= 32 && d->Wo >= 2 && d->Ho >= 2 && d->Cin >= 128 && d->Cout >= 128 &&
         d->Cin % 16 == 0 && d->Cout % 32 == 0;
}

int conv3x3s2_wgrad_tc(const se_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias, int x3, cudaStream_t st) {
  if (!conv3x3s2_tc_ok(d)) return SE_ERR_UNSUPPORTED;
  se_conv_desc d1 = *d;
  d1.kh = 1; d1.kw = 1;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      const int tap = r * 3 + s;
      int rc = conv1x1_wgrad_tc(&d1, x + ((long long)r * d->W + s) * d->Cin, dy, dw + (long long)tap * d->Cin * d->Cout,
                                tap == 0 ? dbias : nullptr, x3, st, d->Wo - (s == 2), d->Ho - (r == 2));
      if (rc != SE_OK) return rc;
    }
  return SE_OK;
}

}  // namespace se
