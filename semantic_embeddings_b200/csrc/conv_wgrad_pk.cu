// "Packed" error-compensated (SE_MODE_TF32X3) weight gradient of 3x3 / stride 1 / 'same' convolutions with
// Cin <= 16 and Cout <= 16 -- the 36 first-stage layers of ResNet-110 (models/cifar_resnet.py:96-105), the most
// expensive weight gradients of a step:
//   dW[r, s, ci, co] += sum_pixels X[pixel + (r-1, s-1), ci] * dY[pixel, co]
// Operands are MN-major TF32 tiles in the SWIZZLE_128B_BASE32B layout exactly as in conv_wgrad_tc.cu (read its header
// first).  The 32-channel TMA boxes of a 16-channel tensor carry 16 zero-filled channel slots; the splitter warps put
// the low parts lo = x - tf32_trunc(x) THERE: chunk o of a 128-byte pixel row and chunk o ^ 64 are a data / zero-fill
// pair whatever the swizzle phase of the row, so new[o] = old[o] + lo(old[o ^ 64]) leaves hi in place (the tensor core
// truncates the raw word) and fills the free slot.  ONE pass of MMAs then yields hi*hi, hi*lo, lo*hi (and lo*lo) in the
// four 16 x 16 quadrants of each accumulator block, which the epilogue adds: fp32-level dW with the MMA count of the
// single-pass mode, no second pass and no extra buffer.
//
// The kernel is bound by shared-memory bandwidth (TMA fill + split read/write + MMA operand reads), so this variant
// uses the layout with the fewest bytes per pixel tile (PT = 64 pixels):
//   A: ONE box of Hb+2 rows x W pixels x 32 channel slots of X; the vertical taps r are descriptor offsets of r*W
//      pixels (the four 32-row blocks of an M = 128 instruction are r = 0, 1, 2 and an ignored block);
//   B: THREE boxes of dY, one per horizontal tap s, shifted by 1-s pixels (the shift lives on dY, which has no halo
//      rows; TMA zero-fills the row ends), stacked along N: one MMA (N = 96) per 8-pixel k-step reads the X slab once
//        D[(r, ci'), (s, co')] += A[(r, ci'), pix] * B[(s, co'), pix]     ci', co' in [0, 32): hi slots 0-15, lo slots 16-31
//      plus one MMA with an all-ones A for dbias.
//   40 KB per stage instead of 56 KB (three shifted X boxes + one dY box): measured 24.4 us vs 32.2 us per layer.
//   (For the single-pass and the two-pass kernels of conv_wgrad_tc.cu the three-X-boxes layout wins: there every k-step
//   re-uses ONE small B slab for all its MMAs, and an MMA with a new B operand costs ~100 cycles more.)
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace se {

using namespace tc;

struct WgPkParams {
  int N, H, W, Cin, Cout;
  int Hb, Nb, PT;              // pixel tile: Hb rows of one image (Nb == 1) or Nb whole images; PT = W*Hb*Nb pixels
  int img_px;                  // pixels of one image inside the tile (Hb * W)
  int img_stride;              // bytes between images inside the x buffer ((Hb + 2) * W * 128)
  int xbuf_bytes;              // the x buffer (Nb * img_stride)
  int dy_bytes;                // ONE shifted dY buffer (PT * 128); a stage holds three, right behind the x buffer
  int G;                       // accumulator blocks: 3 horizontal taps (+1 for the bias gradient), 32 columns each
  int stages, stage_bytes;
  int tiles_m;
  float* dw;
  float* dbias;
};

__device__ __forceinline__ float pk_tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void __launch_bounds__(192, 1)
conv_wgrad_pk_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, WgPkParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;
  uint8_t* ones = tiles + (size_t)p.stages * p.stage_bytes;               // 8 pixels x 32 channels of 1.0f (G == 4)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + (p.G == 4 ? 1024 : 0));
  uint64_t* full = bars;                    // TMA has filled the stage
  uint64_t* empty = bars + 4;               // the MMAs have read the stage
  uint64_t* packed = bars + 8;              // the splitters have filled the lo slots of the stage
  uint64_t* done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int per_cta = (p.tiles_m + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per_cta;
  const int t_end = min(p.tiles_m, t_begin + per_cta);
  const int tiles_per_img = (p.Nb == 1) ? (p.H / p.Hb) : 1;
  const uint32_t tmem_cols = 128;           // 4 blocks of 32 columns

  if (p.G == 4) {
    float4* o = reinterpret_cast<float4*>(ones);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) o[i] = make_float4(1.f, 1.f, 1.f, 1.f);
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_x); prefetch_tmap(&map_dy);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&packed[s], 128); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  fence_proxy_async();                      // generic-proxy writes of the ones tile -> visible to the tensor core
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();                               // nothing above touches global memory (see common.cuh)

  if (t_begin < t_end) {
    if (warp == 0) {
      // ===================== TMA producer (convergent warp, one elected lane issues)
      int stage = 0, phase = 0;
      const uint32_t tx = p.xbuf_bytes + 3 * p.dy_bytes;
      for (int t = t_begin; t < t_end; ++t) {
        int n0, h0;
        if (p.Nb == 1) { n0 = t / tiles_per_img; h0 = (t - n0 * tiles_per_img) * p.Hb; }
        else { n0 = t * p.Nb; h0 = 0; }
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sb = tiles + (size_t)stage * p.stage_bytes;
        if (elect_one()) {
          mbar_expect_tx(&full[stage], tx);
          tma_load_4d(sb, &map_x, &full[stage], 0, 0, h0 - 1, n0);
          for (int s = 0; s < 3; ++s) tma_load_4d(sb + p.xbuf_bytes + s * p.dy_bytes, &map_dy, &full[stage], 0, 1 - s, h0, n0);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (convergent warp, one elected lane issues)
      // instruction descriptors: tf32 x tf32 -> f32, A and B both MN-major (bits 15 / 16), M = 128, N = 96 / 32
      const uint32_t idesc3 = umma_idesc(2, 128, 96) | (1u << 15) | (1u << 16);
      const uint32_t idesc1 = umma_idesc(2, 128, 32) | (1u << 15) | (1u << 16);
      // descriptor high word: SBO = 512 B (consecutive 4-pixel atoms), version 1, SWIZZLE_128B_BASE32B
      const uint32_t hi = (512u >> 4) | (1u << 14) | (1u << 29);
      const uint32_t lbo_a = (((uint32_t)(p.W * 128) >> 4) & 0x3FFFu) << 16;      // next vertical tap = next image row
      const uint32_t lbo_b = (((uint32_t)p.dy_bytes >> 4) & 0x3FFFu) << 16;       // next N block = next shifted dY buffer
      const uint32_t ones_lo = (smem_u32(ones) & 0x3FFFFu) >> 4;                  // LBO 0, the same 8 pixels for every k-step
      const uint32_t tiles_u32 = smem_u32(tiles);
      const int ksteps = p.PT / 8;
      int stage = 0, phase = 0;
      uint32_t acc = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&packed[stage], phase);
        fence_after_sync();
        const uint32_t sb = tiles_u32 + (uint32_t)stage * p.stage_bytes;
        const uint32_t dyb = sb + p.xbuf_bytes;
        uint32_t img_off = 0, rem = 0;               // byte offset of the current image / pixel inside the image
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t da = ((uint64_t)hi << 32) | (uint64_t)((((sb + img_off + rem * 128) & 0x3FFFFu) >> 4) | lbo_a);
          const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)((((dyb + ks * 1024) & 0x3FFFFu) >> 4) | lbo_b);
          const uint64_t db1 = ((uint64_t)hi << 32) | (uint64_t)((((dyb + p.dy_bytes + ks * 1024) & 0x3FFFFu) >> 4) | lbo_b);
          if (elect_one()) {
            mma_tf32(tmem_base, da, db, idesc3, acc);
            if (p.G == 4) mma_tf32(tmem_base + 96, ((uint64_t)hi << 32) | (uint64_t)ones_lo, db1, idesc1, acc);
          }
          __syncwarp();
          acc = 1;
          rem += 8;
          if ((int)rem == p.img_px) { rem = 0; img_off += p.img_stride; }
        }
        if (elect_one()) mma_commit(&empty[stage]);
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) mma_commit(done);
      __syncwarp();
    } else {
      // ===================== warps 2-5: splitters during the main loop, then the epilogue
      {
        int stage = 0, phase = 0;
        const int et = (warp - 2) * 32 + lane;
        const int npair = (p.xbuf_bytes + 3 * p.dy_bytes) >> 5;             // X and the dY buffers are contiguous in a stage
        for (int t = t_begin; t < t_end; ++t) {
          float4* q = reinterpret_cast<float4*>(tiles + (size_t)stage * p.stage_bytes);
          mbar_wait(&full[stage], phase);
#pragma unroll 2
          for (int i = et; i < npair; i += 128) {
            // pair i = (128-byte pixel row i >> 2, 16-byte chunk i & 3 of one 64-byte half and the same chunk of the other
            // half); odd rows start with the upper half so that the 8 lanes of a shared-memory phase cover all 32 banks
            const int c = ((i >> 2) << 3) | (i & 3) | (((i >> 2) & 1) << 2);
            float4 a = q[c], b = q[c ^ 4];                                   // one of the two is a zero-filled slot
            float4 na, nb;
            na.x = a.x + pk_tf32_lo(b.x); na.y = a.y + pk_tf32_lo(b.y); na.z = a.z + pk_tf32_lo(b.z); na.w = a.w + pk_tf32_lo(b.w);
            nb.x = b.x + pk_tf32_lo(a.x); nb.y = b.y + pk_tf32_lo(a.y); nb.z = b.z + pk_tf32_lo(a.z); nb.w = b.w + pk_tf32_lo(a.w);
            q[c] = na; q[c ^ 4] = nb;
          }
          fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core
          mbar_arrive(&packed[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
      // ---- epilogue: TMEM -> reductions into dW / dbias (4 warps, one lane quarter each)
      const int q4 = warp & 3;                              // TMEM lane quarter == vertical tap r (quarter 3: bias row)
      if (q4 < 3 || p.G == 4) {
        mbar_wait(done, 0);
        fence_after_sync();
        const int g_first = (q4 < 3) ? 0 : 3, g_last = (q4 < 3) ? 3 : 4;
        for (int gi = g_first; gi < g_last; ++gi) {
          const int g = (q4 < 3) ? (gi + blockIdx.x) % 3 : 3;          // stagger the CTAs over the taps
          float* dst = nullptr;
          if (g < 3) { if (lane < p.Cin) dst = p.dw + ((long long)(q4 * 3 + g) * p.Cin + lane) * p.Cout; }
          else if (lane == 0 && p.dbias) dst = p.dbias;
          uint32_t v[16], v2[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(tmem_base + ((uint32_t)(q4 * 32) << 16) + g * 32)
              : "memory");
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v2[0]), "=r"(v2[1]), "=r"(v2[2]), "=r"(v2[3]), "=r"(v2[4]), "=r"(v2[5]), "=r"(v2[6]), "=r"(v2[7]),
                "=r"(v2[8]), "=r"(v2[9]), "=r"(v2[10]), "=r"(v2[11]), "=r"(v2[12]), "=r"(v2[13]), "=r"(v2[14]), "=r"(v2[15])
              : "r"(tmem_base + ((uint32_t)(q4 * 32) << 16) + g * 32 + 16)
              : "memory");
          tmem_ld_wait();
          // quadrant sum: columns c / c+16 (dY hi / lo slots), rows ci / ci+16 (X hi / lo slots; the bias row has none)
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float f = __uint_as_float(v[c]) + __uint_as_float(v2[c]);
            if (g < 3) f += __shfl_down_sync(0xffffffffu, f, 16);
            v[c] = __float_as_uint(f);
          }
          if (dst) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (4 * q >= p.Cout) break;
              float4 val = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                       __uint_as_float(v[4 * q + 3]));
              atomicAdd(reinterpret_cast<float4*>(dst + 4 * q), val);
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

int init_conv_wgrad_pk() {
  if (cudaFuncSetAttribute(conv_wgrad_pk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
    set_error("init_conv_wgrad_pk: cannot raise the shared-memory limit");
    return SE_ERR_CUDA;
  }
  return SE_OK;
}

// Geometry + shared-memory plan; two stages (80 KB) leave the rest of the SM to the backward-data kernel of the layer.
static int plan_wgrad_pk(const se_conv_desc* d, bool with_bias, WgPkParams* pp, size_t* smem_out) {
  if (d->kh != 3 || d->kw != 3 || d->stride != 1 || d->pad_t != 1 || d->pad_l != 1 || d->Ho != d->H || d->Wo != d->W)
    return SE_ERR_UNSUPPORTED;
  const int Cin = d->Cin, Cout = d->Cout, W = d->W, H = d->H;
  if (Cin > 16 || Cout > 16 || Cin % 4 != 0 || Cout % 4 != 0) return SE_ERR_UNSUPPORTED;
  // vertical taps are address offsets of r*W pixels: whole 1024-byte swizzle periods need W % 8 == 0
  if (W > 64 || (W & (W - 1)) != 0 || W < 8) return SE_ERR_UNSUPPORTED;
  WgPkParams& p = *pp;
  p.N = d->N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.PT = max(64, 2 * W);
  if (W * H >= p.PT) { if (H % (p.PT / W) != 0) return SE_ERR_UNSUPPORTED; p.Hb = p.PT / W; p.Nb = 1; }
  else { if (p.PT % (W * H) != 0) return SE_ERR_UNSUPPORTED; p.Hb = H; p.Nb = p.PT / (W * H); }
  p.img_px = p.Hb * W;
  p.img_stride = (p.Hb + 2) * W * 128;
  p.xbuf_bytes = p.Nb * p.img_stride;
  p.dy_bytes = p.PT * 128;
  p.G = with_bias ? 4 : 3;
  p.stage_bytes = p.xbuf_bytes + 3 * p.dy_bytes;          // both multiples of 1024 (W >= 8, PT >= 64)
  const int fixed = (with_bias ? 1024 : 0) + 16 * 8 + 1024 + 64;
  p.stages = min(2, (200 * 1024) / p.stage_bytes);
  if (p.stages < 1) return SE_ERR_UNSUPPORTED;
  p.tiles_m = (p.Nb == 1) ? d->N * (H / p.Hb) : ceil_div(d->N, p.Nb);
  *smem_out = (size_t)p.stages * p.stage_bytes + fixed;
  return SE_OK;
}

// dynamic shared memory / TMEM columns of the packed kernel for this layer (0 when it cannot run)
size_t conv_wgrad_pk_smem(const se_conv_desc* d, int* tmem_cols) {
  WgPkParams p;
  size_t smem = 0;
  if (plan_wgrad_pk(d, true, &p, &smem) != SE_OK) return 0;
  if (tmem_cols) *tmem_cols = 128;
  return smem;
}

int conv_wgrad_pk(const se_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias, cudaStream_t st) {
  WgPkParams p;
  size_t smem = 0;
  int rc = plan_wgrad_pk(d, dbias != nullptr, &p, &smem);
  if (rc != SE_OK) return rc;
  const int Cin = d->Cin, Cout = d->Cout, W = d->W, H = d->H;
  p.dw = dw; p.dbias = dbias;
  CUtensorMap mx, mdy;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    uint32_t box[4] = {32u, (uint32_t)W, (uint32_t)(p.Hb + 2), (uint32_t)p.Nb};
    if (!make_tmap(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return SE_ERR_CUDA;
    uint64_t ydims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)d->N};
    uint64_t ystrides[3] = {(uint64_t)Cout * 4, (uint64_t)W * Cout * 4, (uint64_t)H * W * Cout * 4};
    uint32_t ybox[4] = {32u, (uint32_t)W, (uint32_t)p.Hb, (uint32_t)p.Nb};
    if (!make_tmap(&mdy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dy), ydims, ystrides, ybox,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return SE_ERR_CUDA;
  }
  const int gx = max(1, min(p.tiles_m, sm_count()));
  launch(conv_wgrad_pk_kernel, dim3(gx), dim3(192), smem, st, mx, mdy, p);
  return check_launch("conv_wgrad_pk_kernel");
}

}  // namespace se
