// BatchNormalization (training + inference) fused with the ReLU / residual-add / pooled+padded
// shortcut that follows it in the reference graphs (models/cifar_resnet.py:100-124,220;
// models/plainnet.py:53,68,71; models/wide_residual_network.py:14-92; learn_image_embeddings.py:42-43).
// HBM-bound elementwise + per-channel reductions: 128-bit coalesced accesses along the channel
// axis, float64 cross-CTA accumulation of the statistics.
#include "common.cuh"

namespace se {

// ---------------------------------------------------------------------------------------- statistics
// x [rows, C] -> stats[c] += sum x, stats[C+c] += sum x^2.  Thread layout: C4 = C/4 float4 lanes
// along channels, the rest of the block strides over rows.
__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ x, long long rows, int C, double* __restrict__ stats, int rows_per_cta) {
  extern __shared__ double sred[];  // [2*C]
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * C; i += blockDim.x) sred[i] = 0.0;
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  if ((C & 3) == 0) {
    const int C4 = C >> 2;
    const int lanes = min(C4, (int)blockDim.x);
    const int rstep = blockDim.x / lanes;
    const int cq0 = tid % lanes, rr = tid / lanes;
    if (rr < rstep) {
      for (int cq = cq0; cq < C4; cq += lanes) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
        double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
        int cnt = 0;
        for (long long r = r0 + rr; r < r1; r += rstep) {
          float4 v = *reinterpret_cast<const float4*>(x + r * C + 4 * cq);
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
          q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
          if (++cnt == 64) {  // flush the short fp32 partials into float64
            ds[0] += s.x; ds[1] += s.y; ds[2] += s.z; ds[3] += s.w;
            dq[0] += q.x; dq[1] += q.y; dq[2] += q.z; dq[3] += q.w;
            s = make_float4(0.f, 0.f, 0.f, 0.f); q = s; cnt = 0;
          }
        }
        ds[0] += s.x; ds[1] += s.y; ds[2] += s.z; ds[3] += s.w;
        dq[0] += q.x; dq[1] += q.y; dq[2] += q.z; dq[3] += q.w;
        // lanes of a warp that share this channel quad (lanes apart by `lanes`, when lanes divides 32) combine first
        const bool pow2 = (lanes & (lanes - 1)) == 0 && lanes < 32 && (C4 <= lanes);
        if (pow2) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            for (int o = lanes; o < 32; o <<= 1) {
              ds[j] += __shfl_xor_sync(0xffffffffu, ds[j], o);
              dq[j] += __shfl_xor_sync(0xffffffffu, dq[j], o);
            }
        }
        if (!pow2 || (tid & 31) < lanes) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            atomicAdd(&sred[4 * cq + j], ds[j]);
            atomicAdd(&sred[C + 4 * cq + j], dq[j]);
          }
        }
      }
    }
  } else {
    for (int c = tid; c < C; c += blockDim.x) {
      double s = 0, q = 0;
      for (long long r = r0; r < r1; ++r) { float v = x[r * C + c]; s += v; q += (double)v * v; }
      sred[c] += s; sred[C + c] += q;
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * C; i += blockDim.x) atomicAdd(&stats[i], sred[i]);
}

// ---------------------------------------------------------------------------------------- forward
struct Res {
  const float* ptr; int C, pad_lo, pool, H, W;
};

__device__ __forceinline__ float res_value(const Res& r, long long row, int c) {
  // residual contribution for output element (row, c) of a (N,H,W,Cout) tensor
  int cr = c - r.pad_lo;
  if (cr < 0 || cr >= r.C) return 0.f;
  if (r.pool == 1) return r.ptr[row * r.C + cr];
  int w = (int)(row % r.W);
  long long t = row / r.W;
  int h = (int)(t % r.H);
  long long n = t / r.H;
  const float* b = r.ptr + ((n * (2 * r.H) + 2 * h) * (2 * r.W) + 2 * w) * (long long)r.C + cr;
  long long rs = (long long)(2 * r.W) * r.C;
  return 0.25f * (b[0] + b[r.C] + b[rs] + b[rs + r.C]);
}

// Each CTA first derives scale/shift for all channels (from the float64 sums in training mode, from the
// moving statistics in inference mode), CTA 0 also writes the saved statistics + moving averages.
template <bool TRAIN>
__global__ void __launch_bounds__(256)
bn_fwd_kernel(const float* __restrict__ x, long long rows, int C, const double* __restrict__ stats,
              const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
              float* __restrict__ moving_mean, float* __restrict__ moving_var, float* __restrict__ save_mean,
              float* __restrict__ save_invstd, Res res, int relu, float* __restrict__ y) {
  extern __shared__ float sc[];  // scale[C], shift[C]
  float* scale = sc;
  float* shift = sc + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, invstd;
    if (TRAIN) {
      double m = stats[c] / (double)rows;
      double var = stats[C + c] / (double)rows - m * m;
      if (var < 0) var = 0;
      mean = (float)m;
      invstd = (float)(1.0 / sqrt(var + (double)eps));
      if (blockIdx.x == 0) {
        save_mean[c] = mean;
        save_invstd[c] = invstd;
        if (moving_mean) {
          double n = (double)rows;
          double uvar = var * (n / (n - (1.0 + (double)eps)));
          moving_mean[c] = moving_mean[c] * momentum + mean * (1.f - momentum);
          moving_var[c] = moving_var[c] * momentum + (float)uvar * (1.f - momentum);
        }
      }
    } else {
      mean = moving_mean[c];
      invstd = rsqrtf(moving_var[c] + eps);
    }
    float g = gamma[c];
    scale[c] = g * invstd;
    shift[c] = beta[c] - mean * g * invstd;
  }
  __syncthreads();
  const long long total = rows * C;
  if ((C & 3) == 0) {
    const long long total4 = total >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
      long long e = i << 2;
      int c = (int)(e % C);
      long long row = e / C;
      float4 v = *reinterpret_cast<const float4*>(x + e);
      float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = o[j] * scale[c + j] + shift[c + j];
        if (res.ptr) t += res_value(res, row, c + j);
        if (relu) t = fmaxf(t, 0.f);
        o[j] = t;
      }
      *reinterpret_cast<float4*>(y + e) = make_float4(o[0], o[1], o[2], o[3]);
    }
  } else {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
      int c = (int)(e % C);
      long long row = e / C;
      float t = x[e] * scale[c] + shift[c];
      if (res.ptr) t += res_value(res, row, c);
      if (relu) t = fmaxf(t, 0.f);
      y[e] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------- backward
// pass 1: scratch[c] += sum g, scratch[C+c] += sum g*xhat with g = dout * (y > 0 if relu)
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dout,
                     long long rows, int C, const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                     int relu, double* __restrict__ scratch, int rows_per_cta) {
  extern __shared__ double sred[];
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * C; i += blockDim.x) sred[i] = 0.0;
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  if ((C & 3) == 0) {
    const int C4 = C >> 2;
    const int lanes = min(C4, (int)blockDim.x);
    const int rstep = blockDim.x / lanes;
    const int cq0 = tid % lanes, rr = tid / lanes;
    if (rr < rstep) {
      for (int cq = cq0; cq < C4; cq += lanes) {
        float mu[4], is[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { mu[j] = save_mean[4 * cq + j]; is[j] = save_invstd[4 * cq + j]; }
        float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
        double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
        int cnt = 0;
        for (long long r = r0 + rr; r < r1; r += rstep) {
          long long e = r * C + 4 * cq;
          float4 xv = *reinterpret_cast<const float4*>(x + e);
          float4 gv = *reinterpret_cast<const float4*>(dout + e);
          float g[4] = {gv.x, gv.y, gv.z, gv.w};
          float xx[4] = {xv.x, xv.y, xv.z, xv.w};
          if (relu) {
            float4 yv = *reinterpret_cast<const float4*>(y + e);
            if (!(yv.x > 0.f)) g[0] = 0.f;
            if (!(yv.y > 0.f)) g[1] = 0.f;
            if (!(yv.z > 0.f)) g[2] = 0.f;
            if (!(yv.w > 0.f)) g[3] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) { s[j] += g[j]; q[j] += g[j] * (xx[j] - mu[j]) * is[j]; }
          if (++cnt == 64) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { ds[j] += s[j]; dq[j] += q[j]; s[j] = 0.f; q[j] = 0.f; }
            cnt = 0;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { ds[j] += s[j]; dq[j] += q[j]; }
        const bool pow2 = (lanes & (lanes - 1)) == 0 && lanes < 32 && (C4 <= lanes);
        if (pow2) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            for (int o = lanes; o < 32; o <<= 1) {
              ds[j] += __shfl_xor_sync(0xffffffffu, ds[j], o);
              dq[j] += __shfl_xor_sync(0xffffffffu, dq[j], o);
            }
        }
        if (!pow2 || (tid & 31) < lanes) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            atomicAdd(&sred[4 * cq + j], ds[j]);
            atomicAdd(&sred[C + 4 * cq + j], dq[j]);
          }
        }
      }
    }
  } else {
    for (int c = tid; c < C; c += blockDim.x) {
      double s = 0, q = 0;
      float mu = save_mean[c], is = save_invstd[c];
      for (long long r = r0; r < r1; ++r) {
        float g = dout[r * C + c];
        if (relu && !(y[r * C + c] > 0.f)) g = 0.f;
        s += g; q += (double)g * (x[r * C + c] - mu) * is;
      }
      sred[c] += s; sred[C + c] += q;
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * C; i += blockDim.x) atomicAdd(&scratch[i], sred[i]);
}

// pass 2: dx, dres, and (CTA 0) dgamma / dbeta
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dout,
                    long long rows, int C, const float* __restrict__ gamma, const float* __restrict__ save_mean,
                    const float* __restrict__ save_invstd, int relu, int relu_in, float* __restrict__ dx, float beta_dx,
                    float* __restrict__ dres, float beta_res, float* __restrict__ dgamma, float* __restrict__ dbeta,
                    const double* __restrict__ scratch) {
  extern __shared__ float sc[];  // a[C], b[C], m[C], k[C]:  dx = a*g + b + k*x   (k = -a*... folded)
  float* ca = sc;
  float* cb = sc + C;
  float* cm = sc + 2 * C;
  float* ck = sc + 3 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double sg = scratch[c], sgx = scratch[C + c];
    float invstd = save_invstd[c], mean = save_mean[c], g = gamma[c];
    float a = g * invstd;
    // dx = a * (gr - sg/n - xhat * sgx/n),  xhat = (x-mean)*invstd
    ca[c] = a;
    cb[c] = (float)(-(double)a * sg / (double)rows);
    ck[c] = (float)(-(double)a * sgx / (double)rows) * invstd;
    cm[c] = mean;
    if (blockIdx.x == 0) {
      if (dgamma) dgamma[c] += (float)sgx;
      if (dbeta) dbeta[c] += (float)sg;
    }
  }
  __syncthreads();
  const long long total = rows * C;
  if ((C & 3) == 0) {
    const long long total4 = total >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
      long long e = i << 2;
      int c = (int)(e % C);
      float4 xv = *reinterpret_cast<const float4*>(x + e);
      float4 gv = *reinterpret_cast<const float4*>(dout + e);
      float g[4] = {gv.x, gv.y, gv.z, gv.w};
      float xx[4] = {xv.x, xv.y, xv.z, xv.w};
      if (relu) {
        float4 yv = *reinterpret_cast<const float4*>(y + e);
        if (!(yv.x > 0.f)) g[0] = 0.f;
        if (!(yv.y > 0.f)) g[1] = 0.f;
        if (!(yv.z > 0.f)) g[2] = 0.f;
        if (!(yv.w > 0.f)) g[3] = 0.f;
      }
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float d = ca[c + j] * g[j] + cb[c + j] + ck[c + j] * (xx[j] - cm[c + j]);
        if (relu_in && !(xx[j] > 0.f)) d = 0.f;
        o[j] = d;
      }
      if (beta_dx != 0.f) {
        float4 old = *reinterpret_cast<const float4*>(dx + e);
        o[0] += beta_dx * old.x; o[1] += beta_dx * old.y; o[2] += beta_dx * old.z; o[3] += beta_dx * old.w;
      }
      *reinterpret_cast<float4*>(dx + e) = make_float4(o[0], o[1], o[2], o[3]);
      if (dres) {
        if (beta_res != 0.f) {
          float4 old = *reinterpret_cast<const float4*>(dres + e);
          g[0] += beta_res * old.x; g[1] += beta_res * old.y; g[2] += beta_res * old.z; g[3] += beta_res * old.w;
        }
        *reinterpret_cast<float4*>(dres + e) = make_float4(g[0], g[1], g[2], g[3]);
      }
    }
  } else {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
      int c = (int)(e % C);
      float g = dout[e];
      if (relu && !(y[e] > 0.f)) g = 0.f;
      float xv = x[e];
      float d = ca[c] * g + cb[c] + ck[c] * (xv - cm[c]);
      if (relu_in && !(xv > 0.f)) d = 0.f;
      if (beta_dx != 0.f) d += beta_dx * dx[e];
      dx[e] = d;
      if (dres) dres[e] = (beta_res != 0.f ? beta_res * dres[e] : 0.f) + g;
    }
  }
}

// dsrc[n, 2h+i, 2w+j, c] = beta*dsrc + scale * g[n,h,w,c+pad_lo]   (pool==2: scale .25; pool==1: scale 1)
__global__ void __launch_bounds__(256)
shortcut_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ y, int relu, int N, int H, int W, int C,
                    Res res, float* __restrict__ dsrc, float beta) {
  const int SH = H * res.pool, SW = W * res.pool;
  const long long total = (long long)N * SH * SW * res.C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % res.C);
    long long t = e / res.C;
    int sw = (int)(t % SW); t /= SW;
    int sh = (int)(t % SH);
    long long n = t / SH;
    long long o = ((n * H + sh / res.pool) * W + sw / res.pool) * C + c + res.pad_lo;
    float g = dout[o];
    if (relu && !(y[o] > 0.f)) g = 0.f;
    if (res.pool == 2) g *= 0.25f;
    dsrc[e] = (beta != 0.f ? beta * dsrc[e] : 0.f) + g;
  }
}

static Res to_res(const se_residual* r) {
  Res o;
  if (r && r->ptr) { o.ptr = r->ptr; o.C = r->C; o.pad_lo = r->pad_lo; o.pool = r->pool; o.H = r->H; o.W = r->W; }
  else { o.ptr = nullptr; o.C = 0; o.pad_lo = 0; o.pool = 1; o.H = 0; o.W = 0; }
  return o;
}

static int ew_grid(long long work_items) {
  long long g = ceil_div<long long>(work_items, 256);
  long long cap = (long long)sm_count() * 8;
  return (int)max(1LL, min(g, cap));
}

// rows per CTA for the reduction kernels: enough CTAs to fill the machine, at least 32 rows each
static int red_rows_per_cta(long long rows, int* grid) {
  long long target = (long long)sm_count() * 2;
  long long per = max(32LL, ceil_div<long long>(rows, target));
  *grid = (int)ceil_div<long long>(rows, per);
  return (int)per;
}

}  // namespace se

using namespace se;

extern "C" int se_bn_stats(const float* x, int64_t rows, int C, double* stats, void* stream) {
  SE_REQUIRE(x && stats && rows > 0 && C > 0, "bad arguments");
  int grid;
  int per = red_rows_per_cta(rows, &grid);
  bn_stats_kernel<<<grid, 256, 2 * C * sizeof(double), as_stream(stream)>>>(x, rows, C, stats, per);
  return check_launch("bn_stats_kernel");
}

extern "C" int se_bn_fwd_train(const float* x, int64_t rows, int C, const double* stats, const float* gamma,
                               const float* beta, float eps, float momentum, float* moving_mean, float* moving_var,
                               float* save_mean, float* save_invstd, const se_residual* res, int relu, float* y,
                               void* stream) {
  SE_REQUIRE(x && y && stats && gamma && beta && save_mean && save_invstd && rows > 0 && C > 0, "bad arguments");
  Res r = to_res(res);
  long long items = ((C & 3) == 0) ? rows * C / 4 : rows * C;
  bn_fwd_kernel<true><<<ew_grid(items), 256, 2 * C * sizeof(float), as_stream(stream)>>>(
      x, rows, C, stats, gamma, beta, eps, momentum, moving_mean, moving_var, save_mean, save_invstd, r, relu, y);
  return check_launch("bn_fwd_kernel<train>");
}

extern "C" int se_bn_fwd_infer(const float* x, int64_t rows, int C, const float* gamma, const float* beta,
                               const float* moving_mean, const float* moving_var, float eps, const se_residual* res,
                               int relu, float* y, void* stream) {
  SE_REQUIRE(x && y && gamma && beta && moving_mean && moving_var && rows > 0 && C > 0, "bad arguments");
  Res r = to_res(res);
  long long items = ((C & 3) == 0) ? rows * C / 4 : rows * C;
  bn_fwd_kernel<false><<<ew_grid(items), 256, 2 * C * sizeof(float), as_stream(stream)>>>(
      x, rows, C, nullptr, gamma, beta, eps, 0.f, const_cast<float*>(moving_mean), const_cast<float*>(moving_var),
      nullptr, nullptr, r, relu, y);
  return check_launch("bn_fwd_kernel<infer>");
}

extern "C" int se_bn_bwd(const float* x, const float* y, const float* dout, int64_t rows, int C, const float* gamma,
                         const float* save_mean, const float* save_invstd, int relu, int relu_in, float* dx,
                         float beta_dx, float* dres, float beta_res, float* dgamma, float* dbeta, double* scratch,
                         void* stream) {
  SE_REQUIRE(x && dout && dx && gamma && save_mean && save_invstd && scratch && rows > 0 && C > 0, "bad arguments");
  SE_REQUIRE(!relu || y, "relu backward needs y");
  int grid;
  int per = red_rows_per_cta(rows, &grid);
  bn_bwd_reduce_kernel<<<grid, 256, 2 * C * sizeof(double), as_stream(stream)>>>(x, y, dout, rows, C, save_mean,
                                                                                   save_invstd, relu, scratch, per);
  int rc = check_launch("bn_bwd_reduce_kernel");
  if (rc) return rc;
  long long items = ((C & 3) == 0) ? rows * C / 4 : rows * C;
  bn_bwd_apply_kernel<<<ew_grid(items), 256, 4 * C * sizeof(float), as_stream(stream)>>>(
      x, y, dout, rows, C, gamma, save_mean, save_invstd, relu, relu_in, dx, beta_dx, dres, beta_res, dgamma, dbeta,
      scratch);
  return check_launch("bn_bwd_apply_kernel");
}

extern "C" int se_shortcut_bwd(const float* dout, const float* y, int relu, int N, int H, int W, int C,
                               const se_residual* res, float* dsrc, float beta, void* stream) {
  SE_REQUIRE(dout && res && dsrc && (!relu || y), "bad arguments");
  SE_REQUIRE(res->pool == 1 || res->pool == 2, "pool must be 1 or 2");
  Res r = to_res(res);
  r.ptr = dsrc;  // only the geometry is used
  long long total = (long long)N * H * r.pool * W * r.pool * r.C;
  shortcut_bwd_kernel<<<ew_grid(total), 256, 0, as_stream(stream)>>>(dout, y, relu, N, H, W, C, r, dsrc, beta);
  return check_launch("shortcut_bwd_kernel");
}
