// Pooling and small elementwise kernels (HBM-bound; channel-contiguous coalesced accesses).
//   AveragePooling2D((2,2))      models/plainnet.py:59
//   MaxPooling2D((3,3),(2,2))    keras.applications.ResNet50 stem (utils.py:237)
//   GlobalAveragePooling2D       models/cifar_resnet.py:228, plainnet.py:61, wide_residual_network.py:94
//   Add / Activation('relu')     wide_residual_network.py:34,56 ; learn_image_embeddings.py:42
#include <float.h>

#include "common.cuh"

namespace se {

static int ew_grid2(long long n) {
  long long g = ceil_div<long long>(n, 256);
  return (int)max(1LL, min(g, (long long)sm_count() * 8));
}

__global__ void avgpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C) {
  pdl_grid_sync();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    long long t = e / C;
    int ow = (int)(t % Wo); t /= Wo;
    int oh = (int)(t % Ho);
    long long n = t / Ho;
    const float* b = x + ((n * H + 2 * oh) * W + 2 * ow) * C + c;
    long long rs = (long long)W * C;
    y[e] = 0.25f * (b[0] + b[C] + b[rs] + b[rs + C]);
  }
}

__global__ void avgpool2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, float beta, int N, int H, int W, int C) {
  pdl_grid_sync();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * H * W * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    long long t = e / C;
    int w = (int)(t % W); t /= W;
    int h = (int)(t % H);
    long long n = t / H;
    float g = 0.f;
    if (h / 2 < Ho && w / 2 < Wo) g = 0.25f * dy[((n * Ho + h / 2) * Wo + w / 2) * C + c];
    dx[e] = (beta != 0.f ? beta * dx[e] : 0.f) + g;
  }
}

__global__ void maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int k,
                                   int stride, int pad_t, int pad_l, int Ho, int Wo) {
  pdl_grid_sync();
  const long long total = (long long)N * Ho * Wo * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    long long t = e / C;
    int ow = (int)(t % Wo); t /= Wo;
    int oh = (int)(t % Ho);
    long long n = t / Ho;
    float m = -FLT_MAX;
    for (int r = 0; r < k; ++r)
      for (int s = 0; s < k; ++s) {
        int ih = oh * stride - pad_t + r, iw = ow * stride - pad_l + s;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) m = fmaxf(m, x[((n * H + ih) * W + iw) * C + c]);
      }
    y[e] = m;
  }
}

// dx[i] = sum over windows containing i whose max equals x[i] (first-match tie-break like TF's
// MaxPoolGrad: the gradient goes to the first maximal element in window scan order).
__global__ void maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                                   float* __restrict__ dx, int N, int H, int W, int C, int k, int stride, int pad_t,
                                   int pad_l, int Ho, int Wo) {
  pdl_grid_sync();
  const long long total = (long long)N * Ho * Wo * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    long long t = e / C;
    int ow = (int)(t % Wo); t /= Wo;
    int oh = (int)(t % Ho);
    long long n = t / Ho;
    float m = y[e];
    bool done = false;
    for (int r = 0; r < k && !done; ++r)
      for (int s = 0; s < k && !done; ++s) {
        int ih = oh * stride - pad_t + r, iw = ow * stride - pad_l + s;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
          long long o = ((n * H + ih) * W + iw) * C + c;
          if (x[o] == m) { atomicAdd(&dx[o], dy[e]); done = true; }
        }
      }
  }
}

// one warp per (n, 32-channel group): lanes along channels, loop over HW
__global__ void gap_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int HW, int C) {
  pdl_grid_sync();
  const long long total = (long long)N * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    long long n = e / C;
    const float* b = x + n * HW * C + c;
    float s = 0.f;
    for (int i = 0; i < HW; ++i) s += b[(long long)i * C];
    y[e] = s / (float)HW;
  }
}

// one CTA per image: 256 threads = (256 / C) pixel groups x C channels, independent partial sums, one shared-memory
// combine (the per-(n,c) serial loop above is a 64-deep dependent chain of strided loads: 24 us for 8192 outputs)
__global__ void __launch_bounds__(256)
gap_fwd_cta_kernel(const float* __restrict__ x, float* __restrict__ y, int HW, int C) {
  pdl_grid_sync();
  __shared__ float part[256];
  const int G = 256 / C;
  const int c = threadIdx.x % C, g = threadIdx.x / C;
  const float* b = x + (long long)blockIdx.x * HW * C + c;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int i = g;
  for (; i + 3 * G < HW; i += 4 * G) {
    s0 += b[(long long)i * C]; s1 += b[(long long)(i + G) * C]; s2 += b[(long long)(i + 2 * G) * C]; s3 += b[(long long)(i + 3 * G) * C];
  }
  for (; i < HW; i += G) s0 += b[(long long)i * C];
  part[threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (g == 0) {
    float s = 0.f;
    for (int k = 0; k < G; ++k) s += part[k * C + c];
    y[(long long)blockIdx.x * C + c] = s / (float)HW;
  }
}

__global__ void gap_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, float beta, int N, int HW, int C) {
  pdl_grid_sync();
  const long long total = (long long)N * HW * C;
  const float inv = 1.f / (float)HW;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    long long n = e / ((long long)HW * C);
    dx[e] = (beta != 0.f ? beta * dx[e] : 0.f) + dy[n * C + c] * inv;
  }
}

__global__ void add_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, long long n, int relu) {
  pdl_grid_sync();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float v = a[e] + (b ? b[e] : 0.f);
    y[e] = relu ? fmaxf(v, 0.f) : v;
  }
}

__global__ void add_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, int relu, float* __restrict__ da,
                               float beta_a, float* __restrict__ db, float beta_b, long long n) {
  pdl_grid_sync();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float g = dy[e];
    if (relu && !(y[e] > 0.f)) g = 0.f;
    if (da) da[e] = (beta_a != 0.f ? beta_a * da[e] : 0.f) + g;
    if (db) db[e] = (beta_b != 0.f ? beta_b * db[e] : 0.f) + g;
  }
}

}  // namespace se

using namespace se;

extern "C" int se_avgpool2_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream) {
  SE_REQUIRE(x && y, "null pointer");
  long long total = (long long)N * (H / 2) * (W / 2) * C;
  launch(avgpool2_fwd_kernel, dim3(ew_grid2(total)), dim3(256), 0, as_stream(stream), x, y, N, H, W, C);
  return check_launch("avgpool2_fwd_kernel");
}
extern "C" int se_avgpool2_bwd(const float* dy, float* dx, float beta, int N, int H, int W, int C, void* stream) {
  SE_REQUIRE(dy && dx, "null pointer");
  launch(avgpool2_bwd_kernel, dim3(ew_grid2((long long)N * H * W * C)), dim3(256), 0, as_stream(stream), dy, dx, beta, N, H, W, C);
  return check_launch("avgpool2_bwd_kernel");
}
extern "C" int se_maxpool_fwd(const float* x, float* y, int N, int H, int W, int C, int k, int stride, int pad_t,
                              int pad_l, int Ho, int Wo, void* stream) {
  SE_REQUIRE(x && y, "null pointer");
  launch(maxpool_fwd_kernel, dim3(ew_grid2((long long)N * Ho * Wo * C)), dim3(256), 0, as_stream(stream), x, y, N, H, W, C, k, stride,
                                                                                           pad_t, pad_l, Ho, Wo);
  return check_launch("maxpool_fwd_kernel");
}
extern "C" int se_maxpool_bwd(const float* x, const float* y, const float* dy, float* dx, int N, int H, int W, int C,
                              int k, int stride, int pad_t, int pad_l, int Ho, int Wo, void* stream) {
  SE_REQUIRE(x && y && dy && dx, "null pointer");
  cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)N * H * W * C, as_stream(stream));
  if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return SE_ERR_CUDA; }
  launch(maxpool_bwd_kernel, dim3(ew_grid2((long long)N * Ho * Wo * C)), dim3(256), 0, as_stream(stream), x, y, dy, dx, N, H, W, C, k,
                                                                                           stride, pad_t, pad_l, Ho, Wo);
  return check_launch("maxpool_bwd_kernel");
}
extern "C" int se_gap_fwd(const float* x, float* y, int N, int HW, int C, void* stream) {
  SE_REQUIRE(x && y, "null pointer");
  if (C <= 256 && 256 % C == 0 && HW >= 256 / C) {
    launch(gap_fwd_cta_kernel, dim3(N), dim3(256), 0, as_stream(stream), x, y, HW, C);
    return check_launch("gap_fwd_cta_kernel");
  }
  launch(gap_fwd_kernel, dim3(ew_grid2((long long)N * C)), dim3(256), 0, as_stream(stream), x, y, N, HW, C);
  return check_launch("gap_fwd_kernel");
}
extern "C" int se_gap_bwd(const float* dy, float* dx, float beta, int N, int HW, int C, void* stream) {
  SE_REQUIRE(dy && dx, "null pointer");
  launch(gap_bwd_kernel, dim3(ew_grid2((long long)N * HW * C)), dim3(256), 0, as_stream(stream), dy, dx, beta, N, HW, C);
  return check_launch("gap_bwd_kernel");
}
extern "C" int se_add_fwd(const float* a, const float* b, float* y, int64_t n, int relu, void* stream) {
  SE_REQUIRE(a && y, "null pointer");
  launch(add_fwd_kernel, dim3(ew_grid2(n)), dim3(256), 0, as_stream(stream), a, b, y, n, relu);
  return check_launch("add_fwd_kernel");
}
extern "C" int se_add_bwd(const float* dy, const float* y, int relu, float* da, float beta_a, float* db, float beta_b,
                          int64_t n, void* stream) {
  SE_REQUIRE(dy && (!relu || y), "null pointer");
  launch(add_bwd_kernel, dim3(ew_grid2(n)), dim3(256), 0, as_stream(stream), dy, y, relu, da, beta_a, db, beta_b, n);
  return check_launch("add_bwd_kernel");
}
extern "C" int se_relu_fwd(const float* x, float* y, int64_t n, void* stream) {
  return se_add_fwd(x, nullptr, y, n, 1, stream);
}
extern "C" int se_relu_bwd(const float* dy, const float* y, float* dx, float beta, int64_t n, void* stream) {
  return se_add_bwd(dy, y, 1, dx, beta, nullptr, 0.f, n, stream);
}
