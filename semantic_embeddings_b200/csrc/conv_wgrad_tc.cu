// tcgen05 (kind::tf32) weight gradient of 3x3 / stride 1 / 'same' convolutions:
//   dW[r, s, ci, co] += sum_pixels X[pixel + (r-1, s-1), ci] * dY[pixel, co]      (autodiff of Conv2D, reference
//   models/cifar_resnet.py:96-105 etc. under learn_image_embeddings.py:238)
//
// The reduction dimension is the PIXEL axis, so both operands are "MN-major" for the tensor core (channels are
// contiguous, pixels are strided): no transposition of activations is needed, the NHWC tiles that TMA drops into
// shared memory are consumed as they are.  For 32-bit (TF32) operands the tensor core accepts MN-major data only in
// the SWIZZLE_128B_BASE32B layout (32-byte swizzle granules, atoms of 32 channels x 4 pixels), which TMA produces
// with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Channel counts that are not multiples of 32 (the 16-channel stage of
// ResNet-110) are handled by letting the TMA box be 32 channels wide: the out-of-range channels are zero-filled
// by the hardware, so the tile still has 128-byte rows; the zero rows / columns of D are simply not written back.
//
//   GEMM per pixel tile:  D[(blk, ci), co] += A[(blk, ci), pix] * B[co, pix]
//     A = the nine tap-shifted input tiles (4-D TMA boxes, zero-filled halo = 'same' padding) plus one all-ones
//         tile, stacked along M in blocks of cb = min(Cin,32) channels: 128 MMA rows = 128/cb blocks per
//         instruction.  The all-ones block makes the bias gradient (column sums of dY) fall out of the same MMAs.
//     B = the dY tile, N = output channels (<= 256)
//   Accumulators stay in TMEM across ALL pixel tiles of a CTA (split-K over CTAs); one epilogue at the end adds
//   them into dW / dbias with vector atomics.
// grid = (pixel-tile groups, ci chunks of 32, 1); persistent over its pixel tiles.
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace se {

using namespace tc;

struct WgTcParams {
  int N, H, W, Cin, Cout;
  int Wb, Hb, Nb, PT;          // pixel box per stage, PT = Wb*Hb*Nb (128 or 64)
  int cb;                      // channels per M block (16 or 32)
  int ci_chunk;                // input channels handled by one CTA (== cb)
  int nblk;                    // 9 taps + 1 ones block
  int per;                     // blocks per MMA = 128 / cb
  int G;                       // MMAs per k-step
  int first[4];                // first block of each MMA (the last one may overlap its predecessor)
  int cbn, nnb;                // dY: channels per N block (<= 32), number of N blocks
  int stages, stage_bytes, xa_bytes, dy_bytes;
  int tiles_m;
  float* dw;
  float* dbias;
};

// MN-major TF32 operand: atoms of 32 channels x 4 pixels (512 B); LBO = distance between 32-channel blocks,
// SBO = distance between consecutive 4-pixel atoms (one instruction consumes two of them: K = 8).
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint64_t layout = 1ull;             // SWIZZLE_128B_BASE32B
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

__global__ void __launch_bounds__(192, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, WgTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + 4;
  uint64_t* done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_cta = (p.tiles_m + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per_cta;
  const int t_end = min(p.tiles_m, t_begin + per_cta);
  const int ci0 = blockIdx.y * p.ci_chunk;
  const int row_bytes = p.cb * 4;
  const int tiles_per_img = (p.Nb == 1) ? (p.H / p.Hb) : 1;

  // the all-ones block (slot 9 of every stage) is written once with ordinary stores
  for (int s = 0; s < p.stages; ++s) {
    float4* ones = reinterpret_cast<float4*>(tiles + (size_t)s * p.stage_bytes + 9 * p.xa_bytes);
    for (int i = threadIdx.x; i < p.xa_bytes / 16; i += blockDim.x) ones[i] = make_float4(1.f, 1.f, 1.f, 1.f);
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_x); prefetch_tmap(&map_dy);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  fence_proxy_async();                      // generic-proxy writes of the ones tiles -> visible to the tensor core
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int ncols = p.cbn * p.nnb;          // MMA N (== Cout)

  if (t_begin < t_end) {
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer: nine shifted X tiles + the dY tile per stage
      int stage = 0, phase = 0;
      const uint32_t tx = 9 * p.xa_bytes + p.dy_bytes;
      for (int t = t_begin; t < t_end; ++t) {
        int n0, h0;
        if (p.Nb == 1) { n0 = t / tiles_per_img; h0 = (t % tiles_per_img) * p.Hb; }
        else { n0 = t * p.Nb; h0 = 0; }
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], tx);
        uint8_t* sb = tiles + (size_t)stage * p.stage_bytes;
        for (int tap = 0; tap < 9; ++tap)
          tma_load_4d(sb + tap * p.xa_bytes, &map_x, &full[stage], ci0, tap % 3 - 1, h0 + tap / 3 - 1, n0);
        for (int nb = 0; nb < p.nnb; ++nb)
          tma_load_4d(sb + 10 * p.xa_bytes + nb * (p.PT * p.cbn * 4), &map_dy, &full[stage], nb * p.cbn, 0, h0, n0);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer
      // instruction descriptor: tf32 x tf32 -> f32, A and B both MN-major (bits 15 / 16), M = 128, N = Cout
      const uint32_t idesc = umma_idesc(2, 128, ncols) | (1u << 15) | (1u << 16);
      const uint32_t n_row_bytes = p.cbn * 4;
      int stage = 0, phase = 0;
      uint32_t first = 1;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&full[stage], phase);
        fence_after_sync();
        const uint32_t sb = smem_u32(tiles + (size_t)stage * p.stage_bytes);
        const uint32_t dyb = sb + 10 * p.xa_bytes;
        // descriptors = constant high word | (address, LBO) low word; accumulate flag resolved at compile time
        const uint32_t hi = ((4 * n_row_bytes) >> 4) | (1u << 14) | (1u << 29);          // SBO, version 1, BASE32B
        const uint32_t hia = ((4 * row_bytes) >> 4) | (1u << 14) | (1u << 29);
        const uint32_t lbo_b = (((uint32_t)(p.PT * n_row_bytes) >> 4) & 0x3FFFu) << 16;
        const uint32_t lbo_a = (((uint32_t)p.xa_bytes >> 4) & 0x3FFFu) << 16;
        for (int ks = 0; ks < p.PT / 8; ++ks) {
          const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)((((dyb + ks * 8 * n_row_bytes) & 0x3FFFFu) >> 4) | lbo_b);
          for (int g = 0; g < p.G; ++g) {
            const uint32_t aaddr = sb + min(g * p.per, p.nblk - p.per) * p.xa_bytes + ks * 8 * row_bytes;
            const uint64_t da = ((uint64_t)hia << 32) | (uint64_t)(((aaddr & 0x3FFFFu) >> 4) | lbo_a);
            if (first) mma_tf32_c<false>(tmem_base + g * ncols, da, db, idesc);
            else mma_tf32_c<true>(tmem_base + g * ncols, da, db, idesc);
          }
          first = 0;
        }
        mma_commit(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      mma_commit(done);
    } else if (warp >= 2) {
      // ===================== epilogue: TMEM -> atomics into dW / dbias (4 warps, one lane quarter each)
      const int q4 = warp & 3;
      mbar_wait(done, 0);
      fence_after_sync();
      const int l = q4 * 32 + lane;
      for (int g = 0; g < p.G; ++g) {
        const int fg = min(g * p.per, p.nblk - p.per), fprev = min((g - 1) * p.per, p.nblk - p.per);
        const int blk = fg + l / p.cb;
        const int ci = ci0 + l % p.cb;
        const bool fresh = (g == 0) || (blk >= fprev + p.per);               // not already covered by the previous MMA
        const bool is_bias = (blk == 9);
        float* dst = nullptr;
        if (fresh && blk < 9 && ci < p.Cin) dst = p.dw + ((long long)blk * p.Cin + ci) * p.Cout;
        if (fresh && is_bias && (l % p.cb) == 0 && blockIdx.y == 0 && p.dbias) dst = p.dbias;
        for (int c0 = 0; c0 < ncols; c0 += 16) {
          const bool col_ok = c0 < p.Cout;             // columns past Cout come from zero-filled channels
          uint32_t v[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(tmem_base + ((uint32_t)(q4 * 32) << 16) + g * ncols + c0)
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
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int init_conv_wgrad_tc() {
  if (cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
    set_error("init_conv_wgrad_tc: cannot raise the shared-memory limit");
    return SE_ERR_CUDA;
  }
  return SE_OK;
}

int conv_wgrad_tc(const se_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias, cudaStream_t st) {
  if (d->kh != 3 || d->kw != 3 || d->stride != 1 || d->pad_t != 1 || d->pad_l != 1 || d->Ho != d->H || d->Wo != d->W)
    return SE_ERR_UNSUPPORTED;
  const int Cin = d->Cin, Cout = d->Cout, W = d->W, H = d->H;
  if (Cin % 16 != 0 || Cout % 16 != 0 || Cout > 256) return SE_ERR_UNSUPPORTED;
  // 16-channel layers would run with half-empty (zero-filled) 32-wide blocks: measured slower (44 us) than the fp32
  // FFMA kernel (32 us) on the ResNet-110 stage-1 shape because the single MMA-issuing thread is the limit there
  static const char* force16 = getenv("SE_WG_TC16");
  if ((Cin < 32 || Cout < 32) && !force16) return SE_ERR_UNSUPPORTED;
  if (W > 64 || (W & (W - 1)) != 0 || W < 4) return SE_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(dw) & 15) != 0 || (dbias && (reinterpret_cast<uintptr_t>(dbias) & 15) != 0))
    return SE_ERR_UNSUPPORTED;
  static bool inited = false;
  if (!inited) { int rc = init_conv_wgrad_tc(); if (rc) return rc; inited = true; }

  WgTcParams p;
  p.N = d->N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.cb = 32;
  p.ci_chunk = 32;
  p.cbn = 32;
  p.nnb = ceil_div(Cout, 32);
  p.PT = 64;
  p.Wb = W;
  if (W * H >= p.PT) { if (H % (p.PT / W) != 0) return SE_ERR_UNSUPPORTED; p.Hb = p.PT / W; p.Nb = 1; }
  else { if (p.PT % (W * H) != 0) return SE_ERR_UNSUPPORTED; p.Hb = H; p.Nb = p.PT / (W * H); }
  p.nblk = 10;
  p.per = 128 / p.cb;
  p.G = ceil_div(p.nblk, p.per);
  for (int g = 0; g < 4; ++g) p.first[g] = 0;
  for (int g = 0; g < p.G; ++g) p.first[g] = min(g * p.per, p.nblk - p.per);
  if (p.G * p.cbn * p.nnb > 512) return SE_ERR_UNSUPPORTED;
  p.xa_bytes = p.PT * p.cb * 4;
  p.dy_bytes = p.PT * p.cbn * p.nnb * 4;
  p.stage_bytes = ceil_div(10 * p.xa_bytes + p.dy_bytes, 1024) * 1024;
  p.stages = min(4, (200 * 1024) / p.stage_bytes);
  if (p.stages < 1) return SE_ERR_UNSUPPORTED;
  p.tiles_m = (p.Nb == 1) ? d->N * (H / p.Hb) : ceil_div(d->N, p.Nb);
  p.dw = dw; p.dbias = dbias;

  CUtensorMap mx, mdy;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    uint32_t box[4] = {(uint32_t)p.cb, (uint32_t)p.Wb, (uint32_t)p.Hb, (uint32_t)p.Nb};
    if (!make_tmap(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return SE_ERR_CUDA;
    uint64_t ydims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)d->N};
    uint64_t ystrides[3] = {(uint64_t)Cout * 4, (uint64_t)W * Cout * 4, (uint64_t)H * W * Cout * 4};
    uint32_t ybox[4] = {(uint32_t)p.cbn, (uint32_t)p.Wb, (uint32_t)p.Hb, (uint32_t)p.Nb};
    if (!make_tmap(&mdy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dy), ydims, ystrides, ybox,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return SE_ERR_CUDA;
  }
  const int gy = ceil_div(Cin, p.ci_chunk);
  const int gx = max(1, min(p.tiles_m, sm_count() / gy));
  const size_t smem = (size_t)p.stages * p.stage_bytes + 16 * 8 + 1024 + 64;
  conv_wgrad_tc_kernel<<<dim3(gx, gy, 1), 192, smem, st>>>(mx, mdy, p);
  return check_launch("conv_wgrad_tc_kernel");
}

}  // namespace se
