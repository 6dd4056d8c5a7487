/* se_b200.h -- C ABI of the B200-native hot path of cvjena/semantic-embeddings.
 *
 * The reference (/root/reference, pure Python on Keras 2.2 / TF 1.x) has no FFI layer: its
 * device work is whatever the Keras graph of learn_image_embeddings.py and the numpy calls of
 * evaluate_retrieval.py:56-67 dispatch to cuDNN/cuBLAS/BLAS.  Each entry point below replaces
 * one group of those graph ops; the comment on each names the reference lines it stands for.
 * The host side (the semantic_embeddings_b200 package) binds this file with ctypes; INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - every function returns 0 on success, <0 on error (SE_ERR_*); se_last_error() gives the
 *     message of the calling thread's last failure.
 *   - device pointers are caller-owned (PyTorch tensors are used as containers); the library
 *     never allocates or frees device memory.  Kernels that need scratch take it explicitly.
 *   - `stream` is a cudaStream_t; all calls are asynchronous w.r.t. the host and capturable
 *     into a CUDA graph.
 *   - process-wide state is limited to what se_init() / se_comm_init() create: the low-priority
 *     side stream and the fork/join events se_run_ops uses for weight gradients, and the NCCL
 *     communicator with its stream.  Kernels keep no state between calls.
 *   - activations are float32 NHWC; conv kernels are float32 HWIO (Keras layout); dense
 *     kernels are (in,out).  `mode` selects the arithmetic of contraction kernels:
 *     SE_MODE_F32 = fp32 FFMA; SE_MODE_TF32 = tcgen05 kind::tf32 (operands truncated to a
 *     10-bit mantissa, fp32 accumulate in TMEM: ~1e-3 relative, outside the reference's fp32
 *     semantics, kept for comparison); SE_MODE_TF32X3 = tcgen05 kind::tf32 with error
 *     compensation (every operand split into hi + lo, hi*hi + hi*lo + lo*hi in one fp32
 *     accumulator: fp32-level results, the mode the training path runs and is benchmarked in).
 *     Shapes the tensor path does not cover fall back to the fp32 kernels -- never to the CPU.
 */
#ifndef SE_B200_H
#define SE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SE_OK 0
#define SE_ERR_ARG (-1)
#define SE_ERR_CUDA (-2)
#define SE_ERR_UNSUPPORTED (-3)

#define SE_MODE_F32 0
#define SE_MODE_TF32 1
#define SE_MODE_TF32X3 2

/* head variants: learn_image_embeddings.py --loss (lines 62, 127-130, 164-171) */
#define SE_LOSS_INV_CORR 0     /* l2norm wrapper + 1 - <t,x>      */
#define SE_LOSS_UNNORM_CORR 1  /* no wrapper     + 1 - <t,z>      */
#define SE_LOSS_MSE 2          /* no wrapper     + sum (z-t)^2    */
#define SE_LOSS_SOFTMAX_CORR 3 /* softmax wrapper + 1 - <t,x>     (learn_image_embeddings.py:129-130) */

/* pairwise modes: evaluate_retrieval.py:57-62 */
#define SE_PDIST_SQEUCLID 0    /* A + B - 2 F F^T                 */
#define SE_PDIST_NEGDOT 1      /* -(F F^T)  (caller passes normalised F, or sets normalize=1) */

const char* se_version(void);
const char* se_last_error(void);
/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
int64_t se_launch_count(void);
int se_device_sm_count(void);
/* one-time per-process setup (device query, shared-memory attributes); call before capturing CUDA graphs */
int se_init(void);
/* bit mask of the tcgen05 kernels compiled in: 1 conv fwd, 2 conv dgrad, 4 conv wgrad, 8 pairwise */
int se_tc_capabilities(void);

/* ------------------------------------------------------------------ convolution / dense
 * Conv2D of models/cifar_resnet.py:96-105,218, models/plainnet.py:52,70,
 * models/wide_residual_network.py:9,20,28,31,46,53 and keras.applications.ResNet50 (utils.py:237).
 * Explicit zero padding (pad_t, pad_l) and output size: the host computes TF 'SAME'
 * (pad_before = total//2, so k=3,s=2 on an even input gives pad 0 before / 1 after).
 * In SE_MODE_TF32 / SE_MODE_TF32X3 the tcgen05 kernels take: 3x3 / stride 1 / pad 1 on image widths 4..56 (weight
 * gradient: ..64), 1x1 / stride 1 / pad 0 on any image size (>= 128 pixels per call), 1x1 / stride 2 / pad 0 with
 * Wo <= 32 -- channel counts in multiples of 16 (the GEMM K dimension: 16 or a multiple of 32); backward data and weight
 * gradient of 3x3 / stride 2 / pad 0 on even image sizes with >= 128 channels on both sides (nine strided 1x1 GEMMs).
 * Every other shape (3-channel stems, the forward pass and the narrow layers of 3x3 / stride 2, 7x7, dense layers) runs on
 * the fp32 kernels in every mode; se_conv2d_path() tells which. */
typedef struct {
  int32_t N, H, W, Cin;   /* input  NHWC */
  int32_t Cout, kh, kw;   /* kernel HWIO */
  int32_t stride;
  int32_t pad_t, pad_l;
  int32_t Ho, Wo;         /* output NHWC = (N, Ho, Wo, Cout) */
} se_conv_desc;

/* y = conv(x, w) [+ bias] [+ residual] [relu]; optionally accumulates per-channel
 * sum(y) and sum(y^2) (of the stored values) into stats[0:Cout], stats[Cout:2Cout]
 * (float64, caller zeroes) -- the BatchNormalization statistics of the next layer. */
int se_conv2d_fwd(const se_conv_desc* d, const float* x, const float* w, const float* bias,
                  const float* residual, float* y, int relu, double* stats, int mode, void* stream);
/* same, with an optional transposed copy w_t = [kh][kw][Cout][Cin] of the kernel: the tcgen05 forward path
 * (SE_MODE_TF32) consumes K-major operands; without w_t the call uses the fp32 kernels. */
int se_conv2d_fwd_ex(const se_conv_desc* d, const float* x, const float* w, const float* w_t, const float* bias,
                     const float* residual, float* y, int relu, double* stats, int mode, void* stream);
/* Auxiliary copies of a convolution kernel w (HWIO) that the tensor-core paths consume; any may be NULL (the call
 * then uses the fp32 kernels):  w_t = [kh][kw][Cout][Cin] (K-major B operand of the forward GEMM);  w_t_lo / w_lo =
 * the low parts w - tf32_trunc(w) in the transposed / the HWIO order (SE_MODE_TF32X3).  se_split_filters writes all
 * three for every kernel of a flat parameter buffer in one launch. */
typedef struct {
  const float* w_t;
  const float* w_t_lo;
  const float* w_lo;
} se_conv_aux;
int se_conv2d_fwd_aux(const se_conv_desc* d, const float* x, const float* w, const se_conv_aux* aux, const float* bias,
                      const float* residual, float* y, int relu, double* stats, int mode, void* stream);
int se_conv2d_dgrad_aux(const se_conv_desc* d, const float* dy, const float* w, const se_conv_aux* aux, float* dx, float beta,
                        int mode, void* stream);
/* se_transpose_filters + PL[off + i] = lo(P[off + i]), PTL[off + (tap, co, ci)] = lo(P[off + (tap, ci, co)]) */
int se_split_filters(const float* P, float* PT, float* PL, float* PTL, const int64_t* table, int n, void* stream);
/* PT[off + (tap, co, ci)] = P[off + (tap, ci, co)] for n kernels of a flat buffer, one launch.
 * table: host array of n x {element offset, taps, Cin, Cout} (int64). */
int se_transpose_filters(const float* P, float* PT, const int64_t* table, int n, void* stream);
/* dx = beta*dx + conv^T(dy, w)   (gradient wrt the input; autodiff of the above) */
int se_conv2d_dgrad(const se_conv_desc* d, const float* dy, const float* w, float* dx, float beta,
                    int mode, void* stream);
/* dw += x (*) dy ; dbias += sum_pixels dy   (dbias may be NULL). Accumulates: caller zeroes. */
int se_conv2d_wgrad(const se_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias,
                    int mode, void* stream);
/* Which kernel family takes this layer in `mode` (pure host-side planning: no device work, callable without a GPU):
 * 1 = a tcgen05 kernel, 0 = an fp32 kernel; direction 0 forward, 1 backward data, 2 weight gradient.  (Forward: given the
 * auxiliary kernel copies of se_conv_aux; with BatchNorm statistics wider than 384 channels the sums come from a separate
 * se_bn_stats pass, see se_conv2d_fwd_aux.) */
int se_conv2d_path(const se_conv_desc* d, int mode, int direction);
/* Dense (models/cifar_resnet.py:233, plainnet.py:67,76, wide_residual_network.py:96, utils.py:242,
 * learn_image_embeddings.py:44): y = x W + b as a 1x1 convolution over a (B,1,1,Cin) tensor. */
int se_dense_fwd(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout,
                 int relu, double* stats, int mode, void* stream);
int se_dense_bwd(const float* x, const float* w, const float* dy, float* dx, float beta, float* dw,
                 float* dbias, int B, int Cin, int Cout, int mode, void* stream);

/* ------------------------------------------------------------------ batch normalisation
 * keras.layers.BatchNormalization (cifar_resnet.py:100,107,220; plainnet.py:53,68,71;
 * wide_residual_network.py:14,25,44,51,91; learn_image_embeddings.py:43), fused with the
 * Activation('relu') / layers.add / AveragePooling2D+ChannelPadding shortcut that follow it
 * (cifar_resnet.py:101,117-124). rows = N*H*W. */
typedef struct {
  const float* ptr;       /* NULL = no residual */
  int32_t C;              /* channels of the residual tensor (<= C of the BN output) */
  int32_t pad_lo;         /* ChannelPadding: residual channel c lands on output channel c+pad_lo */
  int32_t pool;           /* 1 = same resolution, 2 = 2x2 average pool of a (N,2H,2W,C) tensor */
  int32_t H, W;           /* spatial size of the BN output (needed when pool == 2) */
} se_residual;

/* accumulate sum(x), sum(x^2) per channel into stats (float64 [2C], caller zeroes) */
int se_bn_stats(const float* x, int64_t rows, int C, double* stats, void* stream);
/* training mode: mean/var (biased) from `stats`; y = relu?( gamma*(x-mean)*rsqrt(var+eps)+beta + res );
 * writes save_mean/save_invstd [C] for the backward pass and updates moving statistics
 * (moving = moving*momentum + batch*(1-momentum); variance fed as var*n/(n-(1+eps))). */
int se_bn_fwd_train(const float* x, int64_t rows, int C, const double* stats, const float* gamma,
                    const float* beta, float eps, float momentum, float* moving_mean, float* moving_var,
                    float* save_mean, float* save_invstd, const se_residual* res, int relu, float* y,
                    void* stream);
/* Conv2D followed by training-mode BatchNormalization (+ same-shape residual, + ReLU): the pair the reference stacks in
 * every block (models/cifar_resnet.py:96-107, models/wide_residual_network.py:28-66, models/plainnet.py), as ONE call:
 *   y      = conv(x, w) [+ bias] [relu]                (kept: the BatchNorm backward needs it)
 *   bn_out = act( gamma*(y-mean)*rsqrt(var+eps)+beta [+ res] ), statistics / moving averages as se_bn_fwd_train
 * = se_conv2d_fwd_ex (statistics accumulated in the convolution epilogue) + se_bn_fwd_train.  (Round 1 also had a
 * single-launch form with a grid barrier inside the convolution kernel; measured 0.13 ms per step SLOWER than the two
 * launches, it was removed in round 2.)  stats: float64 [2*Cout], caller zeroes; `counter` is unused. */
int se_conv_bn_fwd(const se_conv_desc* d, const float* x, const float* w, const float* w_t, const float* bias,
                   float* y, int relu, double* stats, const float* gamma, const float* beta, float eps, float momentum,
                   float* moving_mean, float* moving_var, float* save_mean, float* save_invstd, const float* res,
                   int bn_relu, float* bn_out, void* counter, int mode, void* stream);
/* inference mode (learn_image_embeddings.py:271 predict_generator): moving statistics */
int se_bn_fwd_infer(const float* x, int64_t rows, int C, const float* gamma, const float* beta,
                    const float* moving_mean, const float* moving_var, float eps, const se_residual* res,
                    int relu, float* y, void* stream);
/* backward of se_bn_fwd_train.  dout = gradient wrt y; `y` is needed when relu != 0 (mask y > 0).
 *   g  = dout * (y > 0)                                  (relu)          [also the residual gradient]
 *   dgamma += sum g*xhat ; dbeta += sum g                 (accumulate: caller zeroes)
 *   dx = beta_dx*dx + gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)) [* (x > 0) if relu_in]
 *   dres = beta_res*dres + g      (same-resolution residual only; NULL to skip)
 * scratch: float64 [2C + 1] (sums + the arrival counter of the fused single-launch path), caller zeroes.
 * relu_in: x itself is a relu output (plainnet.py:52-53). */
int se_bn_bwd(const float* x, const float* y, const float* dout, int64_t rows, int C, const float* gamma,
              const float* save_mean, const float* save_invstd, int relu, int relu_in, float* dx,
              float beta_dx, float* dres, float beta_res, float* dgamma, float* dbeta, double* scratch,
              void* stream);
/* gradient of the pooled / channel-padded shortcut (cifar_resnet.py:117-121):
 * dsrc[n,2h+i,2w+j,c] = beta*dsrc + 0.25 * g[n,h,w,c+pad_lo], g = dout*(y>0) if relu. */
int se_shortcut_bwd(const float* dout, const float* y, int relu, int N, int H, int W, int C,
                    const se_residual* res, float* dsrc, float beta, void* stream);

/* ------------------------------------------------------------------ pooling / elementwise */
int se_avgpool2_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream);       /* plainnet.py:59 */
int se_avgpool2_bwd(const float* dy, float* dx, float beta, int N, int H, int W, int C, void* stream);
int se_maxpool_fwd(const float* x, float* y, int N, int H, int W, int C, int k, int stride, int pad_t,
                   int pad_l, int Ho, int Wo, void* stream);                                    /* ResNet-50 pool1 */
int se_maxpool_bwd(const float* x, const float* y, const float* dy, float* dx, int N, int H, int W, int C,
                   int k, int stride, int pad_t, int pad_l, int Ho, int Wo, void* stream);
int se_gap_fwd(const float* x, float* y, int N, int HW, int C, void* stream);                  /* cifar_resnet.py:228 */
int se_gap_bwd(const float* dy, float* dx, float beta, int N, int HW, int C, void* stream);
/* y = a + b [relu]; backward: da = beta*da + g, db likewise, g = dy*(y>0)  (wide_residual_network.py:34,56) */
int se_add_fwd(const float* a, const float* b, float* y, int64_t n, int relu, void* stream);
int se_add_bwd(const float* dy, const float* y, int relu, float* da, float beta_a, float* db, float beta_b,
               int64_t n, void* stream);
/* y = relu(x) ; dx = beta*dx + dy*(y>0)  (learn_image_embeddings.py:42) */
int se_relu_fwd(const float* x, float* y, int64_t n, void* stream);
int se_relu_bwd(const float* dy, const float* y, float* dx, float beta, int64_t n, void* stream);

/* ------------------------------------------------------------------ embedding head (north-star item)
 * One fused kernel for utils.l2norm (utils.py:125-127), the target gather E[y]
 * (learn_image_embeddings.py:48-50), utils.inv_correlation / squared_distance (utils.py:34-46),
 * the metric utils.nn_accuracy (utils.py:57-100, k<=1) and the backward pass of all of it.
 *   z [B,ldz] raw network output; E [C,ldE] class matrix (fp32); labels int32 [B]
 *   x_out [B,ldz]  = wrapped output (l2norm(z) for INV_CORR, z otherwise)   (may be NULL)
 *   loss [B], acc [B] per-sample loss and 0/1 accuracy                       (may be NULL)
 *   dz [B,ldz]     = d( loss_scale * sum_b loss_b )/dz + Jx^T extra_dx, where extra_dx [B,ldz]
 *                    (may be NULL) is a gradient wrt x_out coming from the classifier branch
 *                    (learn_image_embeddings.py:34-44, --cls_weight)        (dz may be NULL)
 * loss_scale is 1/global_batch for Keras' mean-over-batch. */
int se_embed_head_fwd_bwd(const float* z, int ldz, const int32_t* labels, const float* E, int ldE, int B,
                          int D, int C, int loss_kind, float loss_scale, const float* extra_dx,
                          float* x_out, float* loss, float* acc, float* dz, void* stream);
/* same + rank_out [B] (may be NULL): the top-k form of the metric for every k at once.  utils.nn_accuracy(k)
 * (utils.py:85,95) is 1 iff one of the k best class scores lies within 1e-6 of the true class' score; with G = number
 * of classes better than the true score by >= 1e-6, that is `rank_out < k` (rank_out = G, or C when no class is within
 * 1e-6).  For SE_LOSS_SOFTMAX_CORR (metric = Keras categorical / top-k categorical accuracy) rank_out = number of
 * outputs strictly above the output at argmax(t).  --top_k_acc K: accuracy@K = mean(rank_out < K). */
int se_embed_head_fwd_bwd_ex(const float* z, int ldz, const int32_t* labels, const float* E, int ldE, int B,
                             int D, int C, int loss_kind, float loss_scale, const float* extra_dx,
                             float* x_out, float* loss, float* acc, float* dz, float* rank_out, void* stream);
/* softmax + Keras categorical_crossentropy (clip 1e-7) + argmax accuracy + backward
 * (learn_image_embeddings.py:44,230-231): dlogits = scale * dCE/dlogits. */
int se_softmax_xent_fwd_bwd(const float* logits, int ld, const int32_t* labels, int B, int C, float scale,
                            float* prob, float* loss, float* acc, float* dlogits, void* stream);
/* same + rank_out [B] (may be NULL): classes with a strictly larger logit than the label's -- utils.top_k_acc(k)
 * (utils.py:49-54, in_top_k) = mean(rank_out < k) */
int se_softmax_xent_fwd_bwd_ex(const float* logits, int ld, const int32_t* labels, int B, int C, float scale,
                               float* prob, float* loss, float* acc, float* dlogits, float* rank_out, void* stream);

/* ------------------------------------------------------------------ optimizer
 * keras.optimizers.SGD(lr, momentum, decay, nesterov, clipnorm) + kernel_regularizer=l2(.)
 * (learn_image_embeddings.py:229-236; cifar_resnet.py:152; plainnet.py:8) over ONE flat fp32
 * parameter / gradient / velocity buffer.  Segments give the L2 coefficient of a range.
 *   pass 1: g += 2*lambda*p on regularised ranges; out[0] = sum g^2, out[1] = sum lambda*p^2
 *   pass 2: scale = clipnorm/norm if norm >= clipnorm; v = m*v - lr*g*scale; p += v
 *           (nesterov: p += m*v - lr*g*scale)
 * `out` is float64[2] device memory (caller zeroes before pass 1). */
typedef struct {
  int64_t begin, end;     /* element range [begin,end) of the flat buffer */
  float l2;               /* lambda */
} se_l2_segment;
int se_sgd_step(float* p, float* g, float* v, int64_t n, const se_l2_segment* segs, int nsegs, float lr,
                float momentum, int nesterov, float clipnorm, double* out, void* stream);
/* the two passes separately (data-parallel runs all-reduce g between nothing and pass 1) */
int se_sgd_prepare(const float* p, float* g, int64_t n, const se_l2_segment* segs, int nsegs, double* out,
                   void* stream);
int se_sgd_apply(float* p, const float* g, float* v, int64_t n, float lr, float momentum, int nesterov,
                 float clipnorm, const double* out, void* stream);

/* Keras SGD(decay) (learn_image_embeddings.py:224-236, --max_decay): lr_state = float[4] device memory
 * {lr written by the schedule, decay, iterations, lr_t}; one call per optimizer step sets lr_t = lr / (1 + decay *
 * iterations) and increments iterations.  se_sgd_apply_devlr then reads lr_state + 3. */
int se_sgd_schedule(float* lr_state, void* stream);
/* same as se_sgd_apply with the learning rate read from device memory (float[1]) at run time, so
 * that a CUDA-graph-captured step follows the SGDR schedule (sgdr_callback.py:75-87) without re-capture */
int se_sgd_apply_devlr(float* p, const float* g, float* v, int64_t n, const float* lr_dev, float momentum,
                       int nesterov, float clipnorm, const double* out, void* stream);

/* ------------------------------------------------------------------ data-parallel gradient exchange
 * keras.utils.multi_gpu_model (learn_image_embeddings.py:133,148) -> one process per GPU + NCCL.  The communicator lives
 * inside the library so that the plan runner can issue bucketed all-reduces on its own stream while the backward pass
 * continues and capture them in the step's CUDA graph (SE_OP_ALLREDUCE in se_run_ops).  NCCL is taken from the
 * libnccl.so.2 already loaded in the process; SE_ERR_UNSUPPORTED when there is none.
 *   se_comm_unique_id: rank 0 creates the 128-byte NCCL id, the caller distributes it (any out-of-band channel);
 *   se_comm_init:      collective over all ranks, current CUDA device = this rank's GPU;
 *   se_allreduce_sum:  in-place SUM over all ranks of buf[0:n] on `stream` (per-sample losses are pre-scaled by
 *                      1/global_batch, so the sum IS the gradient of the global mean loss). */
int se_comm_unique_id(void* id_out, int bytes);
int se_comm_init(int rank, int world, const void* unique_id, int bytes);
int se_comm_world(void);
int se_allreduce_sum(float* buf, int64_t n, void* stream);
int se_comm_destroy(void);

/* ------------------------------------------------------------------ input pipeline
 * TinyDatasetGenerator.compose_batch (datasets/common.py:771-796): Keras ImageDataGenerator.random_transform with
 * horizontal_flip + width/height_shift_range 0.15 (datasets/common.py:640; shift = scipy affine_transform order 1,
 * mode 'nearest') and .standardize (featurewise mean / std, :639) for a batch gathered by index from a dataset that is
 * resident in device memory.  src [n, H, W, C] uint8 or float32 raw pixels; index [B] rows of src (NULL = 0..B-1);
 * tx / ty [B] row / column shifts in pixels (NULL = 0), flip [B] 0/1 (NULL = none) -- the random draws are the host's;
 * mean / inv_std [C] with inv_std = 1 / (std + 1e-7); out [B, H, W, C] float32. */
int se_augment_batch(const void* src, int src_is_u8, const int32_t* index, const float* tx, const float* ty,
                     const unsigned char* flip, const float* mean, const float* inv_std, float* out, int B, int H, int W,
                     int C, void* stream);

/* ------------------------------------------------------------------ retrieval
 * evaluate_retrieval.py:56-63: rows [row0,row0+rows) of the N x N distance matrix of F [N,ldF]
 * (fp32, D columns) against all N columns; out [rows, ldout].  normalize=1 applies line 58
 * (F /= ||F||) on the fly without mutating F.  `workspace` (device, >= se_pairwise_workspace_bytes)
 * holds the row norms and, for the tensor-core path, the split operands. */
int64_t se_pairwise_workspace_bytes(int N, int D, int mode);
int se_pairwise_dist(const float* F, int ldF, int N, int D, int row0, int rows, int pdist_mode,
                     int normalize, float* out, int64_t ldout, void* workspace, int mode, void* stream);

/* Fused distance + ranking: the k nearest items (ascending distance, ties by index) of
 * query rows [row0, row0+rows) WITHOUT writing the rows x N distance matrix -- sample-based per-row thresholds, a
 * tensor-core sweep that keeps only the entries below them, a per-row sort of those candidates.  Distances are the
 * ones se_pairwise_dist would store (same arithmetic), so out_idx equals se_row_topk of that matrix.
 * status [1] device int32: 0 = exact; non-zero = some row found fewer than k or more than 4096 candidates (pathological
 * distance distributions) and the caller must use se_pairwise_dist + se_row_topk instead.  k <= 1024, D <= 128.
 * out_idx [rows, ldo] int32, out_val [rows, ldo] float32 (may be NULL). */
int64_t se_pairwise_topk_workspace_bytes(int N, int D, int rows);
int se_pairwise_topk(const float* F, int ldF, int N, int D, int row0, int rows, int pdist_mode, int normalize, int k,
                     int32_t* out_idx, float* out_val, int ldo, void* workspace, int32_t* status, void* stream);

/* Ranking step of evaluate_retrieval.py:67 (`np.argsort(pdist, axis=-1)`) restricted to what the metrics read
 * (class_hierarchy.py:242-244,273,283: the first clip_ahp+1 ranks): for each of `rows` rows of dist [rows, ld] the k
 * smallest of its n values in ascending order, ties by ascending index (a stable argsort's prefix; -0.0 == +0.0).
 * out_idx [rows, ldo] int32 column indices, out_val [rows, ldo] the distances (may be NULL).  k <= 1024 and
 * n <= ~52000 (a row is staged in shared memory), else SE_ERR_UNSUPPORTED. */
int se_row_topk(const float* dist, int64_t ld, int rows, int n, int k, float* out_val, int32_t* out_idx, int ldo,
                void* stream);

/* ClassHierarchy.hierarchical_precision(retrieved, labels, ks, compute_ahp=clip, ignore_qids=True) of the reference
 * (class_hierarchy.py:211-316) on the first K1 = max(ks, clip) + 1 ranks of Q queries (query ids q0 .. q0+Q-1 are
 * database indices; ranks [Q, ldr] int32, e.g. from se_row_topk).  labels [N] int32 class indices (< C);
 * wup_lut / lcs_height_lut [C, C] float64 = wup_similarity / lcs_height of the hierarchy; best_wup / best_lcs [C, K1]
 * float64 = per query class the cumulative sums of the descending class similarities of the WHOLE database
 * (ranking independent, class_hierarchy.py:268,280).  out [Q, 2*(nks + (clip > 0))] float64 per query:
 * P@ks[0] (WUP), P@ks[0] (LCS_HEIGHT), ..., AHP@clip (WUP), AHP@clip (LCS_HEIGHT).  nks <= 8. */
int se_hier_precision(const int32_t* ranks, int ldr, int Q, int K1, int q0, const int32_t* labels, int C,
                      const double* wup_lut, const double* lcs_height_lut, const double* best_wup, const double* best_lcs,
                      const int32_t* ks, int nks, int clip, double* out, void* stream);

/* The same metrics from rankings of any length, as evaluate_retrieval.py:195 requests them (ks = 1..plot_max, compute_ahp
 * = clip or True, compute_ap = True):  ranks [Q, ldr] int32 with n_ret entries per query (n_ret = N for full rankings
 * from se_row_argsort, or max(kcurve, clip) + 1 for top-k rankings); best_wup / best_lcs [C, n_ret] float64;
 *   curve [Q, 2, kcurve] = P@k for k = 1..kcurve (WUP, LCS_HEIGHT)                       (NULL when kcurve == 0)
 *   ahp   [Q, 2]: clip > 0 -> AHP@clip, clip < 0 -> AHP over the whole list, clip == 0 -> not computed (may be NULL)
 *   ap    [Q]   : classical average precision; needs full rankings (NULL = not computed)  (class_hierarchy.py:310-314) */
int se_hier_metrics(const int32_t* ranks, int64_t ldr, int Q, int n_ret, int q0, const int32_t* labels, int C,
                    const double* wup_lut, const double* lcs_height_lut, const double* best_wup, const double* best_lcs,
                    int kcurve, int clip, double* curve, double* ahp, double* ap, void* stream);

/* Full-length ranking of evaluate_retrieval.py:67 (`np.argsort(pdist, axis=-1)`; ascending distance, ties by ascending
 * index, -0.0 == +0.0): out_idx [rows, ldo] int32 = the n column indices of every row of dist [rows, ld] in rank order.
 * workspace: se_row_argsort_workspace_bytes(rows, n) bytes of device memory (one padded row of 64-bit words per row). */
int64_t se_row_argsort_workspace_bytes(int rows, int n);
int se_row_argsort(const float* dist, int64_t ld, int rows, int n, int32_t* out_idx, int64_t ldo, void* workspace,
                   void* stream);

/* ------------------------------------------------------------------ plan runner
 * Runs a host-built array of ops (one training step is ~900 launches) in one call so that
 * neither Python nor ctypes sits between launches.  Each op is an opcode plus the argument
 * block of the entry point above it maps to (see semantic_embeddings_b200/engine.py). */
typedef struct {
  int32_t opcode;
  int32_t i[15];
  float f[8];
  void* p[16];
} se_op;
int se_run_ops(const se_op* ops, int n, int mode, void* stream);
/* profiling variant (bench.py): eager, a CUDA-event pair around every op, per-op device milliseconds */
int se_run_ops_timed(const se_op* ops, int n, int mode, void* stream, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* SE_B200_H */
