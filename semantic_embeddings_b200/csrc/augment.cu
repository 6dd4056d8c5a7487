// Batched training-input pipeline of the reference on the device:
//   TinyDatasetGenerator.compose_batch (datasets/common.py:771-796) = per image
//     keras ImageDataGenerator.random_transform with horizontal_flip, width_shift_range = height_shift_range = 0.15
//     (datasets/common.py:640) followed by .standardize with featurewise_center / featurewise_std_normalization (:639).
// Keras 2.2 (keras_preprocessing 1.0.x) semantics restated:
//   shift : out[h, w] = in(h + tx, w + ty) sampled by scipy.ndimage.affine_transform(order = 1, mode = 'nearest'):
//           linear interpolation between the two neighbouring rows / columns, coordinates clamped to the image
//           (tx = U(-0.15, 0.15) * H along rows, ty likewise along columns -- drawn by the host);
//   flip  : then the columns are reversed with probability 1/2;
//   standardize : (x - mean_c) / (std_c + 1e-7)  (K.epsilon), mean / std per channel over the training set.
// The whole dataset stays resident in HBM (CIFAR: 150 MB as uint8); a batch is gathered by index, transformed and
// written as the float32 NHWC input tensor of the network in ONE launch: one thread per output pixel (all channels),
// reads of a 32x32 image hit L1/L2, writes are coalesced.  HBM-bound: 4*C bytes written + ~C..4C bytes read per pixel.
#include "common.cuh"

namespace se {

template <typename T>
__global__ void __launch_bounds__(256)
augment_kernel(const T* __restrict__ src, const int* __restrict__ index, const float* __restrict__ tx,
               const float* __restrict__ ty, const unsigned char* __restrict__ flip, const float* __restrict__ mean,
               const float* __restrict__ istd, float* __restrict__ out, int B, int H, int W, int C) {
  pdl_grid_sync();
  const long long total = (long long)B * H * W;
  for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < total; px += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(px % W);
    const int h = (int)((px / W) % H);
    const int b = (int)(px / ((long long)W * H));
    const T* img = src + (long long)(index ? index[b] : b) * H * W * C;
    const int wsrc = (flip && flip[b]) ? W - 1 - w : w;            // flip acts on the shifted image
    // scipy 'nearest' + order 1: the sample position is used as is, the two taps are clamped into the image
    const float fy = (float)h + (tx ? tx[b] : 0.f), fx = (float)wsrc + (ty ? ty[b] : 0.f);
    const float y0f = floorf(fy), x0f = floorf(fx);
    const float ay = fy - y0f, ax = fx - x0f;
    const int y0 = min(max((int)y0f, 0), H - 1), y1 = min(max((int)y0f + 1, 0), H - 1);
    const int x0 = min(max((int)x0f, 0), W - 1), x1 = min(max((int)x0f + 1, 0), W - 1);
    const T* p00 = img + ((long long)y0 * W + x0) * C;
    const T* p01 = img + ((long long)y0 * W + x1) * C;
    const T* p10 = img + ((long long)y1 * W + x0) * C;
    const T* p11 = img + ((long long)y1 * W + x1) * C;
    float* o = out + px * C;
    for (int c = 0; c < C; ++c) {
      // separable linear interpolation in scipy's order: along the last axis first is not observable -- the two
      // orders agree to rounding; rows then columns here
      const float top = (1.f - ax) * (float)p00[c] + ax * (float)p01[c];
      const float bot = (1.f - ax) * (float)p10[c] + ax * (float)p11[c];
      const float v = (1.f - ay) * top + ay * bot;
      o[c] = (v - mean[c]) * istd[c];
    }
  }
}

}  // namespace se

using namespace se;

extern "C" int se_augment_batch(const void* src, int src_is_u8, const int32_t* index, const float* tx, const float* ty,
                                const unsigned char* flip, const float* mean, const float* inv_std, float* out, int B,
                                int H, int W, int C, void* stream) {
  SE_REQUIRE(src && mean && inv_std && out && B > 0 && H > 0 && W > 0 && C > 0, "bad arguments");
  const long long total = (long long)B * H * W;
  const int grid = (int)max(1LL, min(ceil_div<long long>(total, 256), (long long)sm_count() * 8));
  if (src_is_u8)
    launch(augment_kernel<unsigned char>, dim3(grid), dim3(256), 0, as_stream(stream), (const unsigned char*)src, index, tx, ty,
           flip, mean, inv_std, out, B, H, W, C);
  else
    launch(augment_kernel<float>, dim3(grid), dim3(256), 0, as_stream(stream), (const float*)src, index, tx, ty, flip, mean,
           inv_std, out, B, H, W, C);
  return check_launch("augment_kernel");
}
