/*
 * zs3b200.h -- C ABI of libzs3b200.so: the sm_100a kernels underneath the ZS3Net training hot path.
 *
 * The reference (valeoai/ZS3) has no FFI: its hot path is nn.Module code that bottoms out in
 * torch / cuDNN / cuBLAS calls.  Every entry point below names the reference call site(s) whose
 * library call it replaces (paths relative to the reference checkout).  The Python modules in
 * zs3_b200/ (mirroring zs3.modeling.* / zs3.utils.loss) are the only callers; they bind this file
 * through ctypes (zs3_b200/_lib.py).  INTEGRATION.md shows the binding a reference maintainer adds.
 *
 * Conventions
 *   - plain pointers + sizes only, no torch types; all pointers are DEVICE pointers unless noted;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises, nothing allocates;
 *   - return 0 (ZS3_OK) or a negative ZS3_ERR_* code; zs3_last_error() gives the message;
 *   - activations are NHWC bf16 with a channel stride that is a multiple of 64 ("cpad");
 *   - packed conv weights are bf16 [cout_pad][R*S][cin_pad] (K-major rows per output channel).
 */
#ifndef ZS3B200_H
#define ZS3B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ZS3_OK 0
#define ZS3_ERR_INVALID_ARG (-1)
#define ZS3_ERR_UNSUPPORTED (-2)
#define ZS3_ERR_LAUNCH (-3)
#define ZS3_ERR_DRIVER (-4)

#define ZS3_MAX_SEGMENTS 6

/* last error message of the calling thread ("" if none) */
const char* zs3_last_error(void);
/* ABI version, bumped on any signature change */
int zs3_abi_version(void);
/* number of kernel-launching calls this library has made in this process (bench.py's gpu_launches) */
unsigned long long zs3_launch_count(void);
/* 1 if the current device is compute capability 10.x, else 0 (kernels are sm_100a only) */
int zs3_device_supported(void);
/* sizeof() of the argument struct `which` as THIS library was compiled (a binding checks its own mirror against it
 * before the first call; a short struct handed to a kernel is an out-of-bounds read).  Unknown id: returns 0. */
#define ZS3_STRUCT_CONV_ARGS 1
#define ZS3_STRUCT_WGRAD_ARGS 2
#define ZS3_STRUCT_BN_APPLY_ARGS 3
#define ZS3_STRUCT_BN_BWD_ARGS 4
#define ZS3_STRUCT_SGEMM_ARGS 5
#define ZS3_STRUCT_GMMN_ITEM 6
#define ZS3_STRUCT_GMMN_TRAIN_ARGS 7
#define ZS3_STRUCT_COMPONENTS_ARGS 8
#define ZS3_STRUCT_CONV_SEGMENT 9
#define ZS3_STRUCT_ROW_SOURCE 10
#define ZS3_STRUCT_BN_ACT_F32_ARGS 11
#define ZS3_STRUCT_BN_BWD_F32_ARGS 12
#define ZS3_STRUCT_AUG_ITEM 13
#define ZS3_STRUCT_AUGMENT_ARGS 14
unsigned long long zs3_sizeof(int which);

/* ------------------------------------------------------------------------------------------------
 * Convolution as implicit GEMM on tcgen05 (TMA im2col A-tiles, TMA weight tiles, TMEM accumulator).
 * Replaces nn.Conv2d forward at zs3/modeling/backbone/resnet.py:16-28,79,125,151,
 * zs3/modeling/aspp.py:11-19,86,97 and zs3/modeling/decoder.py:12,16,20,26 (cuDNN in the reference).
 * The same entry point computes the data gradient (conv with spatially flipped, transposed weights)
 * that autograd's convolution_backward computes for those layers.
 *
 * The reduction (K) dimension is a concatenation of up to ZS3_MAX_SEGMENTS segments; segment i
 * contributes sum over taps and over its cin_pad channels of x_i * w_i.  One segment = ordinary conv;
 * several segments = conv over a channel-concatenated input without materialising the concat
 * (torch.cat at aspp.py:110, decoder.py:37), or split-precision accumulation (hi/lo bf16 pieces).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;  /* bf16 NHWC [N][H][W][x_cstride] */
  int x_cstride;  /* channel stride of x in elements, multiple of 8 */
  const void* w;  /* bf16 [cout_pad][R*S][cin_pad] */
  int cin_pad;    /* channels reduced from this segment, multiple of 64 (zero padded in w) */
} zs3_conv_segment;

typedef struct {
  int N, H, W;    /* input batch / height / width */
  int Ho, Wo;     /* output height / width */
  int R, S;       /* filter height / width */
  int stride, pad, dil;
  int cout_pad;   /* output channels incl. zero padding, multiple of 64 */
  int num_segments;
  zs3_conv_segment seg[ZS3_MAX_SEGMENTS];
  void* y;        /* output [N*Ho*Wo][y_cstride], bf16 or fp32 */
  int y_cstride;  /* channel stride of y (elements) */
  int y_sp_stride;/* 0/1: y is [N][Ho][Wo]; s>1: output pixel (n,p,q) is stored at (n, p*s, q*s) of a
                     [N][y_H][y_W] tensor (scatter used by the data gradient of strided 1x1 convs) */
  int y_H, y_W;   /* only read when y_sp_stride > 1 */
  int y_is_f32;   /* 0: bf16 output, 1: fp32 output */
  int accumulate; /* 1: y += result (read-modify-write), 0: y = result */
  const float* bias; /* optional [cout_pad] fp32 added before store (decoder.pred_conv bias), or NULL */
  double* stat_sum;  /* optional [cout_pad]: += per-channel sum of the (fp32) conv output, or NULL */
  double* stat_sqsum;/* optional [cout_pad]: += per-channel sum of squares */
  /* inference epilogue: eval-mode / frozen BatchNorm (+ residual add, + ReLU) folded into the conv, i.e.
   * y = relu?(conv * ep_scale[n] + ep_shift[n] (+ ep_residual)); all NULL/0 = plain conv output */
  const float* ep_scale;    /* [cout_pad] gamma / sqrt(running_var + eps) (zs3_bn_eval_coeffs) */
  const float* ep_shift;    /* [cout_pad] */
  const void* ep_residual;  /* optional bf16 [N*Ho*Wo][ep_res_cstride] */
  int ep_res_cstride;
  int ep_relu;
  int w_forward_layout; /* 1 (data-gradient mode): seg[i].w is the FORWARD-packed weight of the layer being
                           differentiated, [seg.cin_pad = forward cout][R*S][cout_pad = forward cin]; the kernel reads it
                           as MN-major B tiles and flips the taps itself, so no transposed copy has to be packed */
  /* RESERVED (must be NULL / 0; zs3_conv_fprop rejects anything else): operand prologue.  Planned meaning: segment i is the
   * RAW (pre-normalisation) output y of a conv -> BatchNorm -> ReLU layer and relu?(y * pre_scale[i][c] + pre_shift[i][c])
   * is applied to every A tile in shared memory between the TMA load and the MMA, so that layer's normalised output never
   * makes a round trip through HBM (resnet.py:36-42: bn1/relu -> conv2, bn2/relu -> conv3).  The transform stage is not
   * built (only 27 % of the BatchNorm elements have a single consumer); the fields keep the struct
   * layout stable for it. */
  const float* pre_scale[ZS3_MAX_SEGMENTS];
  const float* pre_shift[ZS3_MAX_SEGMENTS];
  int pre_relu;
} zs3_conv_args;

int zs3_conv_fprop(const zs3_conv_args* a, void* stream);

/* Weight gradient: dw[cout_pad][R*S][cin_pad] (fp32) += sum over output pixels of dy * x(tap).
 * Replaces the wgrad half of convolution_backward for the same layers.  Split over pixels across
 * CTAs with fp32 atomics, so dw must be zeroed (or hold the value to accumulate onto) beforehand. */
typedef struct {
  int N, H, W, Ho, Wo, R, S, stride, pad, dil;
  const void* x;   /* bf16 NHWC input activation [N][H][W][x_cstride] */
  int x_cstride;
  int cin_pad;     /* multiple of 64 */
  const void* dy;  /* bf16 [N*Ho*Wo][dy_cstride] */
  int dy_cstride;
  int cout_pad;    /* multiple of 64 */
  float* dw;       /* fp32 [cout][R*S][dw_ld]: row (co, tap) starts at dw + (co*R*S + tap)*dw_ld */
  int k_splits;    /* number of pixel-range splits (>=1); 0 = choose automatically */
  long long dw_ld; /* 0 = cin_pad (dense packed buffer); else the total input-channel count of a KRSC
                      (torch channels_last) gradient tensor that is accumulated into directly */
  int dw_ci_offset;/* first input channel of this K-segment inside a dw row */
  int cout_valid;  /* 0 = cout_pad; rows >= cout_valid are not written */
  int cin_valid;   /* 0 = cin_pad; columns >= cin_valid are not written */
} zs3_wgrad_args;

int zs3_conv_wgrad(const zs3_wgrad_args* a, void* stream);

/* Weight (re)packing between the reference's OIHW fp32 parameters and the kernel layouts.
 *   mode 0 (fprop):  dst[co][r*S+s][ci - ci_begin]            = src[co][ci][r][s]
 *   mode 1 (dgrad):  dst[ci - ci_begin][(R-1-r)*S+(S-1-s)][co] = src[co][ci][r][s]
 * ci in [ci_begin, ci_begin+ci_count); rows/cols beyond the real sizes are zero filled.
 * dst dims: mode 0 [cout_pad][R*S][cin_pad], mode 1 [cin_pad][R*S][cout_pad].
 * modes 2 / 3 = modes 0 / 1 reading a KRSC source (torch channels_last memory format: src[co][r][s][ci]),
 * the layout this package keeps its conv parameters in. */
int zs3_pack_weight(const float* w_oihw, int Cout, int Cin, int R, int S, int ci_begin, int ci_count, void* dst_bf16,
                    int cout_pad, int cin_pad, int mode, void* stream);
/* grad_oihw[co][ci_begin+ci][r][s] (+)= dw[co][r*S+s][ci]  (fp32) */
int zs3_unpack_wgrad(const float* dw, int cout_pad, int cin_pad, float* grad_oihw, int Cout, int Cin, int R, int S,
                     int ci_begin, int ci_count, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm2d (+ReLU, +residual add, +Dropout) as fused passes around the conv kernels.
 * Replaces F.batch_norm at zs3/modeling/sync_batchnorm/batchnorm.py:48-58 (the single-device path every
 * BatchNorm / SynchronizedBatchNorm2d on the DeepLab path takes), nn.ReLU (resnet.py:35,39,51;
 * aspp.py:29,100; decoder.py:14,18,22), the residual add (resnet.py:50) and nn.Dropout (aspp.py:101,
 * decoder.py:19,23).  Training-mode statistics are accumulated by zs3_conv_fprop's epilogue.
 * ---------------------------------------------------------------------------------------------- */

/* Turn accumulated (sum, sumsq) over `count` values per channel into the per-channel affine
 *   scale = gamma * invstd, shift = beta - mean * scale     (biased variance, eps inside the sqrt)
 * and update running_mean / running_var (unbiased variance, momentum) exactly like F.batch_norm.
 * Channels in [C, Cpad) get scale = shift = 0.  If reset_stats != 0 the accumulators are zeroed. */
int zs3_bn_finalize(double* stat_sum, double* stat_sqsum, long long count, const float* gamma, const float* beta,
                    float eps, float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                    float* mean, float* invstd, int C, int Cpad, int reset_stats, void* stream);
/* standalone batch statistics of a bf16 tensor y[M][y_cstride]: stat_sum[c] += sum y, stat_sqsum[c] += sum y^2.
 * Used for short-K (1x1) convolutions, where fusing the reduction into the conv epilogue would make the epilogue
 * the bottleneck; long-K convolutions keep the fused statistics of zs3_conv_fprop. */
int zs3_bn_stats(const void* y, int y_cstride, long long M, int C, double* stat_sum, double* stat_sqsum, void* stream);
/* eval-mode / frozen BN: coefficients from the running statistics (deeplab.py:68-73 freeze_bn) */
int zs3_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float eps, float* scale, float* shift, float* mean, float* invstd, int C, int Cpad,
                       void* stream);

typedef struct {
  const void* y;        /* bf16 [M][y_cstride] pre-BN conv output */
  int y_cstride;
  const void* residual; /* optional bf16 [M][res_cstride], added before the activation */
  int res_cstride;
  void* out;            /* bf16 [M][out_cstride] */
  int out_cstride;
  const float* scale;   /* [C] */
  const float* shift;   /* [C] */
  long long M;          /* pixels */
  int C;                /* channels processed, multiple of 8 */
  int relu;
  int drop_mode;        /* 0: none; 1: keep-mask from the counter-based RNG (seed, offset); 2: explicit mask */
  float drop_p;
  unsigned long long seed, offset;
  const unsigned char* keep_mask; /* drop_mode 2: [M][C] bytes, 1 = keep */
  const unsigned long long* offset_dev; /* optional device counter added to `offset` at run time (lets a captured
                                           CUDA graph draw fresh masks on every replay) */
  /* Fused finalize (optional, training mode): if stat_sum != NULL, scale/shift above are ignored and every thread
   * derives the affine of its channels from the raw batch statistics exactly like zs3_bn_finalize; CTA 0 publishes
   * mean/invstd/scale/shift (fp32 [C], read by the backward), updates the running statistics and zeroes
   * reset_sum/reset_sqsum (the statistics buffer the NEXT layer will accumulate into: two buffers alternate). */
  const double* stat_sum;
  const double* stat_sqsum;
  long long count;
  const float* gamma;
  const float* beta;
  float eps, momentum;
  float* running_mean;
  float* running_var;
  int C_real;           /* channels with real statistics; [C_real, C) get scale = shift = 0 */
  float* mean_out; float* invstd_out; float* scale_out; float* shift_out;
  double* reset_sum; double* reset_sqsum;
  int reset_count;      /* number of leading entries of reset_sum/reset_sqsum to zero */
  unsigned char* relu_mask_out; /* optional [M][C/8]: bit j of byte (m, c/8) = [out[m][c+j] > 0]; lets the backward of
                                 * residual layers read 1 bit instead of 16 per element for the ReLU mask */
  int sync_clamp;       /* fused finalize: invstd = max(var, eps)^-1/2 instead of (var + eps)^-1/2 -- the multi-replica
                         * formula of the reference's SynchronizedBatchNorm (sync_batchnorm/batchnorm.py:124-142); the
                         * caller has all-reduced stat_sum / stat_sqsum over the ranks and passes the GLOBAL count */
} zs3_bn_apply_args;

/* out = dropout(relu?(scale*y + shift (+ residual))) */
int zs3_bn_apply(const zs3_bn_apply_args* a, void* stream);

typedef struct {
  const void* dout;     /* bf16 [M][dout_cstride] gradient wrt the forward output */
  int dout_cstride;
  const void* out;      /* bf16 forward output (ReLU/Dropout mask is out > 0); unused if !relu */
  int out_cstride;
  const void* y;        /* bf16 pre-BN conv output */
  int y_cstride;
  const float* mean;    /* [C] batch (training) or running (frozen) mean */
  const float* invstd;  /* [C] */
  const float* scale;   /* [C] gamma * invstd */
  const float* shift;   /* [C] beta - mean*scale; only read when relu == 2 */
  long long M;
  int C;                /* multiple of 8 */
  int relu;             /* 0: none; 1: mask = out > 0 (residual / dropout layers); 2: mask = scale*y + shift > 0 */
  float grad_scale;     /* 1/(1-p) when the forward applied dropout after the ReLU, else 1 */
  int training;         /* 1: batch-statistics backward; 0: frozen statistics (dy = scale * dz) */
  double* sum_dz;       /* [C] accumulators: reduce phase adds, apply phase reads */
  double* sum_dzx;      /* [C] sum of dz * xhat */
  void* dy;             /* bf16 gradient wrt y */
  int dy_cstride;
  int dy_sp_stride;     /* >1: scatter pixel (n,p,q) of [N][sp_Ho][sp_Wo] to (n, p*s, q*s) of [N][dy_H][dy_W]
                           (zero-inserted layout consumed by the data gradient of a strided 3x3 conv) */
  int sp_Ho, sp_Wo, dy_H, dy_W;
  void* dres;           /* optional bf16: gradient wrt the residual input (= dz) */
  int dres_cstride;
  int dres_accumulate;
  float* dgamma;        /* optional fp32 [C_real] */
  float* dbeta;
  int C_real;
  int param_accumulate; /* 1: dgamma/dbeta += */
  double* reset_sum_dz; /* optional: the apply phase zeroes the first reset_count entries of these two arrays (the */
  double* reset_sum_dzx;/* sums buffer the NEXT layer's reduce phase will accumulate into; two buffers alternate) */
  int reset_count;
  const unsigned char* relu_mask; /* relu == 3: the bit mask written by zs3_bn_apply (relu_mask_out); `out` unused */
  long long stat_count; /* 0: M.  Synchronised BatchNorm: the GLOBAL element count the (all-reduced) sums range over */
} zs3_bn_bwd_args;

/* phase 1: sum_dz += sum(dz), sum_dzx += sum(dz * xhat) with dz = dout * [out > 0] * grad_scale */
int zs3_bn_bwd_reduce(const zs3_bn_bwd_args* a, void* stream);
/* phase 2: dy = scale * (dz - sum_dz/M - xhat * sum_dzx/M); dres (+)= dz; dgamma/dbeta from the sums */
int zs3_bn_bwd_apply(const zs3_bn_bwd_args* a, void* stream);

/* Layout changes at the module boundary (the reference API is NCHW fp32):
 * dst_nhwc[n][hw][c] (bf16, channel stride cs, zero padded) <- src_nchw[n][c][hw] (fp32), and back. */
int zs3_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int N, int C, long long HW, int cs, void* stream);
int zs3_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int N, int C, long long HW, int cs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Remaining DeepLab glue ops (all HBM-bound, NHWC bf16 unless stated).
 * ---------------------------------------------------------------------------------------------- */

/* Stem 7x7/s2/p3 conv (zs3/modeling/backbone/resnet.py:79): im2col of the NCHW fp32 image into
 * cols[N*Ho*Wo][kpad] bf16 with k = c*R*R + r*R + s (the OIHW flattening), so that the stem runs as a
 * 1x1 zs3_conv_fprop / zs3_conv_wgrad over `cols` with the weight viewed as [Cout][C*R*R][1][1]. */
int zs3_stem_im2col(const float* x_nchw, void* cols, int N, int C, int H, int W, int R, int stride, int pad, int Ho,
                    int Wo, int kpad, int krsc /* 1: k = (r*R + s)*C + c, the channels_last flattening */,
                    void* stream);

/* nn.MaxPool2d(3, 2, 1) (resnet.py:82); argmax keeps the winning window slot (first max wins) per element */
int zs3_maxpool_fwd(const void* x, void* y, unsigned char* argmax, int N, int H, int W, int C, int Ho, int Wo, int k,
                    int stride, int pad, void* stream);
int zs3_maxpool_bwd(const void* dy, const unsigned char* argmax, void* dx, int N, int H, int W, int C, int Ho, int Wo,
                    int k, int stride, int pad, void* stream);

/* F.interpolate(mode="bilinear", align_corners=True) on NHWC bf16 (decoder.py:33-35,46-48) and its adjoint */
int zs3_bilinear_fwd(const void* x, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C, int x_cs, int y_cs,
                     void* stream);
int zs3_bilinear_bwd(const void* dy, void* dx, int N, int Hi, int Wi, int Ho, int Wo, int C, int dy_cs, int dx_cs,
                     int accumulate, void* stream);

/* Final upsample of the class scores to the input size (deeplab.py:44,55): NHWC bf16 [N][Hi][Wi][cs] ->
 * NCHW fp32 [N][C][Ho][Wo] (the tensor the reference API returns), and its adjoint. */
int zs3_upsample_logits_fwd(const void* x, float* y, int N, int C, int Hi, int Wi, int cs, int Ho, int Wo,
                            void* stream);
int zs3_upsample_logits_bwd(const float* dy, void* dx, int N, int C, int Hi, int Wi, int cs, int Ho, int Wo,
                            void* stream);

/* nn.AdaptiveAvgPool2d((1,1)) (aspp.py:84): y[n][c] = scale * sum_hw x[n][hw][c]; and the broadcast
 * y[n][hw][c] (+)= scale * x[n][c] (the 1x1 -> HxW "interpolate" of aspp.py:109 and the pool's adjoint). */
int zs3_spatial_sum(const void* x, void* y, int N, int HW, int C, int x_cs, int y_cs, float scale, void* stream);
int zs3_spatial_broadcast(const void* x, void* y, int N, int HW, int C, int x_cs, int y_cs, float scale,
                          int accumulate, void* stream);

/* SegmentationLosses.CrossEntropyLoss (zs3/utils/loss.py:31-46): logits NCHW fp32, target float [N][HW],
 * optional class weights, ignore_index; loss = sum_i w_t nll_i / sum_i w_t / div (div = batch size when
 * batch_average).  accum2 = two fp64 scratch values kept for the backward (sum w*nll, sum w). */
int zs3_ce_fwd(const float* logit, const float* target, const float* weight, int N, int C, long long HW,
               int ignore_index, float div, double* accum2, float* loss, void* stream);
int zs3_ce_bwd(const float* logit, const float* target, const float* weight, int N, int C, long long HW,
               int ignore_index, float div, const double* accum2, const float* grad_out, float* dlogit,
               void* stream);

/* Training-loss fusion of deeplab.py:44 (F.interpolate(x, size=input, bilinear, align_corners=True)) with
 * SegmentationLosses.CrossEntropyLoss (loss.py:31-46): loss = CE(upsample(x), target) straight from the low-resolution
 * class scores x (NHWC bf16 [N][Hi][Wi][cs], C <= 64 real classes: VOC 21, Pascal-Context 60), without materialising the [N][C][Ho][Wo] fp32
 * logits or their gradient.  Same arguments/semantics as zs3_ce_fwd/bwd; the backward returns d loss / d x (NHWC bf16,
 * padding channels zeroed; Wo <= 640).  Models that must RETURN the logits (evaluation) use zs3_upsample_logits_* +
 * zs3_ce_*. */
int zs3_upsample_ce_fwd(const void* x, const float* target, const float* weight, int N, int C, int Hi, int Wi, int cs,
                        int Ho, int Wo, int ignore_index, float div, double* accum2, float* loss, void* stream);
int zs3_upsample_ce_bwd(const void* x, const float* target, const float* weight, int N, int C, int Hi, int Wi, int cs,
                        int Ho, int Wo, int ignore_index, float div, const double* accum2, const float* grad_out,
                        void* dx, void* stream);
/* the same backward for the exact x4 geometry of DeepLab (Ho == 4*(Hi-1)+1, Wo == 4*(Wi-1)+1, Wo <= 544): one softmax
 * evaluation per output pixel, shared-memory accumulation without atomics, two fp32 partials per input row combined in
 * a fixed order (deterministic).  Returns ZS3_ERR_UNSUPPORTED for any other geometry.  workspace (16-byte aligned) >=
 * zs3_upsample4_ce_bwd_workspace_size(N, C, Hi, Wi) bytes. */
unsigned long long zs3_upsample4_ce_bwd_workspace_size(int N, int C, int Hi, int Wi);
int zs3_upsample4_ce_bwd(const void* x, const float* target, const float* weight, int N, int C, int Hi, int Wi, int cs,
                         int Ho, int Wo, int ignore_index, float div, const double* accum2, const float* grad_out,
                         void* dx, void* workspace, unsigned long long workspace_bytes, void* stream);

/* torch.optim.SGD (zs3/train_pascal.py:55-60) and torch.optim.Adam (zs3/train_pascal_GMMN.py:65-67) over one
 * flat fp32 buffer; grad_scale multiplies the gradient first (1/world_size after the NCCL all-reduce). */
/* bf16 shadow of a (flat) fp32 parameter buffer: conv weights are stored KRSC, so for every layer without channel
 * padding its slice of the shadow IS the packed forward weight [cout][R*S][cin] */
int zs3_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);
int zs3_sgd_step(float* p, const float* g, float* momentum_buf, long long n, float lr, float momentum,
                 float weight_decay, int nesterov, int first_step, float grad_scale, void* stream);
/* same update with the learning rate read from DEVICE memory when the kernel runs (one float): a CUDA graph that
 * captured the step keeps following the poly LR schedule the reference applies every iteration
 * (zs3/utils/lr_scheduler.py:46-67, zs3/base_trainer.py:15) */
int zs3_sgd_step_lrdev(float* p, const float* g, float* momentum_buf, long long n, const float* lr_dev, float momentum,
                       float weight_decay, int nesterov, int first_step, float grad_scale, void* stream);
int zs3_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, int step, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GMMN generator (zs3/modeling/gmmn.py) and GMMNLoss.moment_loss (zs3/utils/loss.py:84-115), fp32.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* A; long long lda; int transA; /* op(A)(i,k) = transA ? A[k*lda+i] : A[i*lda+k] */
  const int* idxA;                           /* optional gather of A's stored rows */
  const float* B; long long ldb; int transB; /* op(B)(k,j) = transB ? B[j*ldb+k] : B[k*ldb+j] */
  const int* idxB;
  float* C; long long ldc;                   /* C[M][N] (+)= op(A) op(B) + bias[j] */
  int M, N, K;
  const float* bias;
  int accumulate;
  const int* dyn_count;                      /* device int: dynamic extent ... */
  int dyn_dim;                               /* ... of M (1) or K (2); 0 = static */
} zs3_sgemm_args;

/* nn.Linear forward/backward (cuBLAS sgemm in the reference: gmmn.py:18,31,34) and pygcn's dense adj@(x@W) */
int zs3_sgemm(const zs3_sgemm_args* a, void* stream);
/* rows[0..*count) = indices of the rows of g[n][f] holding any non-zero (only the batch_size_generator=128
 * sampled rows of train_pascal_GMMN.py:229-237 carry gradient) */
int zs3_find_active_rows(const float* g, int n, int f, int* rows, int* count, void* stream);
/* out[j] (+)= sum_r A[idx[r]][j] (bias gradients) */
int zs3_col_sum(const float* A, long long lda, const int* idx, const int* count, int n, int ncols, float* out,
                int accumulate, void* stream);
/* nn.LeakyReLU(0.2) + nn.Dropout(p) (gmmn.py:19-20); drop_mode as in zs3_bn_apply_args */
int zs3_leaky_dropout_fwd(const float* x, float* y, long long n, float slope, int drop_mode, float p,
                          unsigned long long seed, unsigned long long offset, const unsigned char* keep_mask,
                          void* stream);
/* dx[r][c] = dy[r][c] * f'(.) reconstructed from the forward output h (rows gathered through idx) */
int zs3_leaky_dropout_bwd(const float* dy, const float* h, float* dx, int rows, int cols, const int* idx,
                          const int* count, float slope, float p, void* stream);
/* moment_loss forward: loss = sqrt(sum_ij s_i s_j sum_sigma exp(e_ij/sigma)); P [(M+N)^2] and loss2 are kept
 * for the backward.  `sigma` is a HOST array of nsigma <= 8 bandwidths. */
int zs3_mmd_fwd(const float* gen, const float* real, int M, int N, int D, const float* sigma, int nsigma, float* P,
                double* loss2, float* loss, void* stream);
/* analytic gradient wrt gen [M][D] and/or real [N][D] (either may be NULL) */
int zs3_mmd_bwd(const float* gen, const float* real, int M, int N, int D, const float* P, const float* loss,
                const float* grad_out, float* dgen, float* dreal, void* stream);
/* torch.cat((embd, noise), 1) (gmmn.py:44) */
int zs3_concat2(const float* a, int c1, const float* b, int c2, float* y, long long rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused generator update (csrc/gmmn_fused.cu): one persistent cooperative launch executes a work list of
 * (image, class) iterations of the step-2 inner loop, zs3/train_pascal_GMMN.py:211-240 -- gather of the
 * batch_size_generator sampled rows (`:229-237`), GMMNnetwork.forward (gmmn.py:43-49), GMMNLoss.moment_loss
 * (utils/loss.py:92-115), its backward, and torch.optim.Adam.step (`:239-240`, Adam(lr 2e-4) `:65-67`) --
 * sequentially (every update sees the weights left by the previous one, as in the reference) and without
 * returning to the host in between.  Replaces ~25 forward + ~40 backward library launches per update.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* base;   /* element (r, k) = base[row(r) * row_stride + k * col_stride] */
  const int* rows;     /* optional gather: row(r) = rows[r]; NULL: row(r) = r */
  long long row_stride, col_stride; /* in elements; e.g. an NCHW feature map: row_stride 1, col_stride H*W */
} zs3_row_source;

typedef struct {
  zs3_row_source emb;    /* [rows][embed_dim] class embedding of each sampled pixel (row_stride 0 = one shared row) */
  zs3_row_source noise;  /* [rows][noise_dim] z ~ U[0,1) (train_pascal_GMMN.py:216) */
  zs3_row_source real;   /* [rows][feat]      real decoder features of the sampled pixels */
  const unsigned char* keep_mask; /* optional Dropout keep mask [*][hidden] (bytes); NULL = counter-based RNG */
  const int* keep_rows;  /* optional row gather for keep_mask (and the RNG's row key) */
  const float* adj;      /* optional [rows][rows] row-major adjacency: the graph generator of config 5 (GMMNnetwork_GCN,
                            zs3/modeling/gmmn.py:52-67; pygcn GraphConvolution = adj @ (x @ W) + b): both layers multiply by
                            it, the rows are the image's cluster nodes (train_context_GMMN_GCNcontext.py:400-419); NULL = MLP */
  float* out;            /* optional [rows][feat] dense: the generated features of this item (the reference keeps them as
                            classifier inputs for images holding an unseen class, `:421-428`) */
  int rows;              /* M = N = number of sampled rows, 1..128 (batch_size_generator / number of nodes) */
  int flags;             /* ZS3_GMMN_FORWARD_ONLY: generate `out` and stop (no loss, no backward, no Adam step) */
} zs3_gmmn_item;
#define ZS3_GMMN_FORWARD_ONLY 1

typedef struct {
  const zs3_gmmn_item* items; /* DEVICE array [n_items] */
  int n_items;
  int max_rows;               /* host-side promise: every item has rows <= max_rows <= 128 */
  int embed_dim, noise_dim, hidden, feat;
  float* w1; float* b1;       /* model.0.weight [hidden][embed_dim+noise_dim], model.0.bias [hidden] */
  float* w2; float* b2;       /* model.3.weight [feat][hidden], model.3.bias [feat] */
  int apply_adam;             /* 1: update w/b and the Adam moments in place; 0: write gradients (n_items <= 1) */
  float* adam_m[4];           /* exp_avg of w1, b1, w2, b2 */
  float* adam_v[4];           /* exp_avg_sq */
  float* grad[4];             /* gradient outputs (apply_adam == 0) */
  float lr, beta1, beta2, eps;
  long long step0;            /* item w performs Adam step number step0 + w + 1 */
  float sigma[8]; int nsigma; /* MMD bandwidths (loss.py:86: [2, 5, 10, 20, 40, 80]) */
  float slope, drop_p;        /* LeakyReLU slope 0.2, Dropout p 0.5 (0 = eval mode) */
  unsigned long long seed, offset; /* counter RNG of the Dropout mask when keep_mask == NULL */
  float* losses;              /* [n_items] moment_loss of every update */
  void* workspace; unsigned long long workspace_bytes; /* >= zs3_gmmn_train_workspace_size(...), 16-byte aligned */
  unsigned long long* phase_stamps; /* optional [n_items][8]: %globaltimer (ns) at the start of each update and after
                                       each of its six phases, written by CTA 0 (profiling aid); NULL = off */
  int weights_in_out;         /* 0: nn.Linear layout, w1 [hidden][in], w2 [feat][hidden] (GMMNnetwork);
                                 1: pygcn layout, w1 [in][hidden], w2 [hidden][feat] (GMMNnetwork_GCN's gcn1/gcn2.weight) */
} zs3_gmmn_train_args;

unsigned long long zs3_gmmn_train_workspace_size(int embed_dim, int noise_dim, int hidden, int feat);
int zs3_gmmn_train_fused(const zs3_gmmn_train_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Semantic-cluster graph of a label map (csrc/graph.cu).  Replaces construct_adj_mat
 * (zs3/train_context_GMMN_GCNcontext.py:33-102: pure-Python DFS on the host, called per image at `:307-321`).
 * One thread block per image; 8-connected components of equal labels (255 included), node ids in raster
 * order of each component's first pixel (`:55-62`), node label / seed pixel (`:58-61,72-74`), binary
 * symmetric adjacency between touching components (`:78-88`).  Integer work: bit-exact.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* labels;    /* [B][*] class ids as floats (the reference keeps its label maps in float tensors) */
  const int* src_index;   /* optional [h*w]: pixel q of the graph grid reads labels[b][src_index[q]] (the nearest
                             down-sampling of train_context_GMMN_GCNcontext.py:293-298); NULL: labels is [B][h*w] */
  long long image_stride; /* elements between consecutive images in `labels` */
  int B, h, w;
  int max_nodes;          /* capacity of the per-image outputs below */
  int* n_nodes;           /* [B] number of components; > max_nodes means the outputs were truncated */
  int* node_label;        /* [B][max_nodes] class id of every node */
  int* node_seed;         /* [B][max_nodes] flat pixel index (graph grid) of every node's first pixel */
  int* node_map;          /* optional [B][h*w] node id of every pixel (the reference's `flag` array) */
  float* adj;             /* [B][max_nodes][max_nodes] 0/1, symmetric, zero diagonal (written in full) */
} zs3_components_args;

int zs3_label_components(const zs3_components_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Validation metrics (csrc/metrics.cu): np.argmax over the class axis + Evaluator._generate_matrix
 * (zs3/train_pascal_GMMN.py:371-375, zs3/utils/metrics.py:73-82) without copying the logits to the host.
 * logits [B][C][HW] fp32 (NCHW), target [B][HW] float class ids; pred [B][HW] bytes (optional);
 * conf [C][C] 64-bit counters, conf[gt*C + pred] += 1 for pixels with 0 <= gt < C (optional, accumulated).
 * ---------------------------------------------------------------------------------------------- */
int zs3_argmax_confusion(const float* logits, const float* target, int B, int C, long long HW, unsigned char* pred,
                         unsigned long long* conf, void* stream);
/* Evaluator.add_batch(gt_image, pre_image) (metrics.py:79-81) for predictions computed elsewhere: pred [n] int32 */
int zs3_confusion_from_pred(const int* pred, const float* target, long long n, int C, unsigned long long* conf,
                            void* stream);

/* ------------------------------------------------------------------------------------------------
 * Input transforms of a whole batch on the device (csrc/aug.cu).  Replaces the per-sample PIL pipeline of
 * zs3/dataloaders/custom_transforms.py: RandomHorizontalFlip (:47-56), RandomScaleCrop (:69-104),
 * RandomGaussianBlur (:58-66), FixScale (:107-124), Normalize (:8-27), ToTensor (:30-44), composed in
 * zs3/dataloaders/datasets/pascal.py:120-144 (transform_tr / transform_val).  The caller makes the random
 * draws (same order as the reference) and passes them per picture; results are bit-exact against Pillow.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const unsigned char* image; /* device: decoded picture, [h][w][3] RGB bytes (np.array(PIL image)) */
  const unsigned char* label; /* device: [h][w] class ids (np.array(PIL mask)); may be NULL when out_label is NULL */
  int w, h;                   /* decoded size */
  int flip;                   /* RandomHorizontalFlip (:51): mirror the columns of picture and mask first */
  int rw, rh;                 /* size after Image.resize: (ow, oh) of :83-90 / :115-122; (w, h) = no resize */
  int x1, y1;                 /* crop origin (:98-101) in the resized picture, padded on the right / bottom to the
                                 output size (:93-96); 0, 0 for transform_val */
  float blur_radius;          /* ImageFilter.GaussianBlur(radius) of :62-63; <= 0: no blur */
} zs3_aug_item;

typedef struct {
  const zs3_aug_item* items;      /* device copy of the n items (read by the kernels) */
  const zs3_aug_item* items_host; /* host copy of the same n items (validated by the call; not retained) */
  int n;
  int max_src_h;                  /* >= every item's h: row capacity of the horizontal-pass scratch */
  int out_w, out_h;               /* crop_size x crop_size (transform_tr) or the resized size (transform_val) */
  int fill_label;                 /* RandomScaleCrop fill (255): label value of the padding; the picture pads with 0 */
  const float* lut;               /* device [3][256]: Normalize applied to every byte value, per channel */
  float* out_image;               /* device [n][3][out_h][out_w] float32 (ToTensor's CHW) */
  float* out_label;               /* device [n][out_h][out_w] float32 class ids; NULL = pictures only */
  void* workspace; unsigned long long workspace_bytes; /* >= zs3_augment_workspace_size(...), 16-byte aligned */
} zs3_augment_args;

unsigned long long zs3_augment_workspace_size(int n, int max_src_h, int out_w, int out_h);
/* five launches on `stream` (plan, horizontal pass, vertical pass + labels, and the two blur groups when any item is
 * blurred); no allocation, no synchronisation; down-scaling by more than 8x per axis is rejected */
int zs3_augment_batch(const zs3_augment_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32-grade parity mode (forward only; csrc/parity.cu).  An fp32 convolution is emulated on the bf16 tensor
 * cores by splitting both operands into three bf16 pieces and reducing the six significant cross products as
 * six K-segments of one fp32 TMEM accumulator (zs3_conv_fprop).  Activations stay fp32 NHWC between layers;
 * these are the fp32-I/O versions of the glue kernels plus the splitters.  Same reference call sites as their
 * bf16 counterparts above.
 * ---------------------------------------------------------------------------------------------- */
int zs3_split3_f32(const float* x, void* hi, void* mid, void* lo, long long n, void* stream);
/* component 0/1/2 (hi/mid/lo) of an fp32 weight in the packed fprop layout [cout_pad][R*S][cin_pad] */
int zs3_pack_weight_component(const float* w, int Cout, int Cin, int R, int S, int ci_begin, int ci_count, void* dst,
                              int cout_pad, int cin_pad, int src_krsc, int component, void* stream);
int zs3_bn_apply_f32(const float* y, int y_cs, const float* residual, int res_cs, float* out, int out_cs,
                     const float* scale, const float* shift, long long M, int C, int relu, void* stream);
int zs3_maxpool_f32(const float* x, float* y, int N, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad,
                    void* stream);
/* bilinear align_corners; to_nchw = 1 writes NCHW [N][C_out][Ho][Wo] (the final logits), else NHWC stride Cs */
int zs3_bilinear_f32(const float* x, float* y, int N, int Hi, int Wi, int Ho, int Wo, int Cs, int C_out, int to_nchw,
                     void* stream);
int zs3_spatial_sum_f32(const float* x, float* y, int N, int HW, int C, float scale, void* stream);
int zs3_spatial_broadcast_f32(const float* x, float* y, int N, int HW, int C, void* stream);
int zs3_nchw_to_nhwc_f32(const float* src, float* dst, int N, int C, long long HW, int cs, void* stream);
int zs3_stem_im2col_f32(const float* x, float* cols, int N, int C, int H, int W, int R, int stride, int pad, int Ho,
                        int Wo, int kpad, int krsc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Split-precision TRAINING mode (csrc/parity_train.cu; host side zs3_b200/parity_train.py): fp32 activations and
 * gradients, every conv operand fed to tcgen05 as P bf16 pieces (P = 1, 2, 3: 8, 16, 24 mantissa bits; cuDNN's
 * default TF32 path that the reference runs on a GPU carries 10).  fprop/dgrad: the P(P+1)/2 significant cross
 * products are K-segments of one zs3_conv_fprop launch (fp32 output); wgrad: one zs3_conv_wgrad launch per product,
 * reduce-added into the same gradient.  Below: the HBM-bound glue of that mode, forward and backward.
 * Same reference call sites as the bf16 kernels above (zs3/base_trainer.py:16-20 drives them).
 * ---------------------------------------------------------------------------------------------- */
/* x[n] fp32 (n % 8 == 0, 16-byte aligned) -> n_pieces bf16 tensors with x = sum of pieces (+ dropped tail) */
int zs3_split_f32(const float* x, void* const* pieces, int n_pieces, long long n, void* stream);

typedef struct {
  const float* y; long long y_cstride;              /* conv output [M][y_cstride], pre-normalisation */
  const float* residual; long long res_cstride;     /* optional fp32 residual added before the ReLU */
  const float* scale; const float* shift;           /* per-channel affine (zs3_bn_finalize / zs3_bn_eval_coeffs) */
  float* out; long long out_cstride;                /* fp32 result (optional if pieces are written) */
  void* pieces[3]; int n_pieces; long long piece_cstride; /* bf16 pieces of the result: the next conv's operand */
  long long M; int C;                               /* rows, channels (multiple of 8; padding channels included) */
  int relu;
  int drop_mode; float drop_p;                      /* 0 none, 1 counter-based RNG (bn.cu's stream), 2 keep_mask */
  unsigned long long seed, offset; const void* offset_dev; const void* keep_mask;
} zs3_bn_act_f32_args;
/* out = dropout(relu(y * scale + shift + residual)), written as fp32 and/or as bf16 pieces, one pass */
int zs3_bn_act_f32(const zs3_bn_act_f32_args* a, void* stream);

typedef struct {
  const float* dout; long long dout_cstride;        /* gradient w.r.t. the layer output */
  const float* act; long long act_cstride;          /* forward output (fp32) for the ReLU/Dropout mask, or ... */
  const void* act_hi; long long act_hi_cstride;     /* ... its first bf16 piece (same sign) */
  const float* y; long long y_cstride;              /* saved conv output */
  const float* mean; const float* invstd; const float* scale;
  long long M; int C;                               /* C: power of two >= 64 */
  int relu; float grad_scale; int training;         /* grad_scale = 1/(1-p) of a Dropout layer; training: batch stats */
  double* sum_dz; double* sum_dzx;                  /* [C] scratch, zeroed by the call */
  float* dy; long long dy_cstride;                  /* fp32 gradient w.r.t. y (optional if pieces are written) */
  void* dy_pieces[3]; int n_pieces; long long piece_cstride;  /* bf16 pieces of dy: operands of dgrad / wgrad */
  float* dres; long long dres_cstride;              /* optional: gradient w.r.t. the residual input */
  float* dgamma; float* dbeta; int C_real; int param_accumulate;
} zs3_bn_bwd_f32_args;
/* BatchNorm(+ReLU/Dropout/residual) backward: per-channel sums, then dy (memset + two kernels on `stream`) */
int zs3_bn_bwd_f32(const zs3_bn_bwd_f32_args* a, void* stream);
/* out[c] = sum over rows of x[m][c] in fp64 (bias gradients); C a power of two >= 64 */
int zs3_channel_sums_f32(const float* x, long long x_cstride, long long M, int C, double* out, void* stream);
int zs3_maxpool_arg_f32(const float* x, float* y, unsigned char* argmax, int N, int H, int W, int C, int Ho, int Wo,
                        int k, int stride, int pad, void* stream);
int zs3_maxpool_bwd_f32(const float* dy, const unsigned char* argmax, float* dx, int N, int H, int W, int C, int Ho,
                        int Wo, int k, int stride, int pad, void* stream);
/* backward of zs3_bilinear_f32: dy NHWC [N][Ho][Wo][dy_cs] (from_nchw = 0) or NCHW [N][C][Ho][Wo] (from_nchw = 1,
 * the logits gradient) -> dx NHWC [N][Hi][Wi][dx_cs], channels < C written */
int zs3_bilinear_bwd_f32(const float* dy, float* dx, int N, int Hi, int Wi, int Ho, int Wo, int C, int dy_cs, int dx_cs,
                         int from_nchw, void* stream);
/* y[n][p][c] (+)= scale * x[n][c]: backward of the global average pool (aspp.py:84) */
int zs3_spatial_broadcast_acc_f32(const float* x, float* y, int N, int HW, int C, float scale, int accumulate,
                                  void* stream);

/* debug / tests: the output tile (MB x 128 output channels, CN input channels) zs3_conv_wgrad picks for a layer of M
 * output pixels (host-side cost model, no device work) */
int zs3_debug_wgrad_tile(long long M, int cout_pad, int cin_pad, int taps, int* mb, int* cn);

/* debug: one im2col TMA load dumped raw (tests/test_tma_probe.py) */
int zs3_debug_im2col_probe(const void* x, int N, int H, int W, int C, int pad, int upper, int stride, int cpp, int ppc,
                           int c, int w, int h, int n, int off_w, int off_h, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ZS3B200_H */
