/*
 * hg_sm100a.h -- C ABI of libhg_sm100a.so: the B200 (sm_100a) implementation of the stacked-hourglass
 * heatmap-regression hot path of Xinjie-Qiu/progressive_process_for_human_pose_estimation.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; each entry point below replaces the
 * library kernels PyTorch dispatches for one reference call site (cited as file:line into the reference
 * tree).  A maintainer binds these with ctypes; the in-tree Python host
 * (progressive_process_for_human_pose_estimation_b200/_lib.py) does exactly that.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all buffers;
 *   - activations are NHWC ("channels last"), element type HgDType (bf16 on the tensor-core path, fp32 on
 *     the CUDA-core path); parameters, statistics and gradients of parameters are fp32;
 *   - `stream` is a cudaStream_t passed as void*; kernels are only enqueued, never synchronised;
 *   - return value: HG_OK (0) or a negative HgStatus; hg_last_error_string() gives the text;
 *   - there is no CPU fallback: an unsupported shape is HG_ERR_UNSUPPORTED.
 */
#ifndef HG_SM100A_H_
#define HG_SM100A_H_

#include <stdint.h>

#if defined(__GNUC__)
#define HG_API __attribute__((visibility("default")))
#else
#define HG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum HgStatus {
  HG_OK = 0,
  HG_ERR_BAD_ARG = -1,
  HG_ERR_UNSUPPORTED = -2,
  HG_ERR_CUDA = -3
} HgStatus;

/* HG_F16 is accepted only by the decode / PCKh entry points (the reference evaluates .half() models). */
typedef enum HgDType { HG_BF16 = 0, HG_F32 = 1, HG_F16 = 2 } HgDType;

/* Convolution geometry.  Input [N,H,W,Cin] -> output [N,Ho,Wo,Cout], Ho = (H + 2*pad - dil*(R-1) - 1)/stride + 1.
 * Mirrors nn.Conv2d(Cin, Cout, (R,S), stride, pad, dil) as used by ResidualBlock / lin / creatModel
 * (try_with_torch.py:186-193,248,262,271-273). */
typedef struct HgConvDesc {
  int32_t N, H, W;
  int32_t Cin, Cout;
  int32_t R, S;
  int32_t stride, pad, dil;
  int32_t dtype; /* HgDType of activations */
} HgConvDesc;

/* ---- library ------------------------------------------------------------------------------------ */
HG_API const char* hg_last_error_string(void);
/* Number of kernel launches issued through this library since it was loaded. */
HG_API unsigned long long hg_launch_count(void);
/* 1 when the device behind the current context is compute capability 10.x. */
HG_API int hg_device_ok(void);

/* ---- convolution (nn.Conv2d: try_with_torch.py:186-193,199-207,248,253,262,271-273,291-297) -------- */
/* Repack an fp32 OIHW weight [Cout,Cin,R,S] into the two GEMM operand layouts the kernels read:
 *   w_fprop [R*S][Cout][Cin]  (B operand of y = x (*) w)        element type desc->dtype
 *   w_dgrad [R*S][Cin][Cout]  (B operand of dx = dy (*) w^T)    element type desc->dtype
 * Either destination may be NULL. */
HG_API int hg_pack_conv_weight(const HgConvDesc* d, const float* w_oihw, void* w_fprop, void* w_dgrad, void* stream);

/* y = conv(x, w) + bias [+ residual]; optional per-channel statistics of y for the BatchNorm that follows.
 * Statistics slots are [3*Cp] floats (Cp = channels padded to 64): {S1, S2, pivot} with
 *   S1[c] += sum(y_c - pivot_c),  S2[c] += sum((y_c - pivot_c)^2),  pivot READ, never written by a producer;
 * the consumers use mean = pivot + S1/n, var = S2/n - (S1/n)^2, so the variance cancels against (mean - pivot)^2
 * instead of mean^2.  The caller initialises a slot (S1 = S2 = 0, pivot = anything close to the expected mean: 0 gives
 * plain sums; hg_bn_prepare_stats does it for a whole forward pass with the consuming BatchNorm's running mean).
 * bias, residual, stats may be NULL.  residual has the shape of y. */
HG_API int hg_conv_fprop(const HgConvDesc* d, const void* x, const void* w_fprop, const float* bias,
                  const void* residual, void* y, float* stats, void* stream);

/* dx = conv_transpose(dy, w) [+ addend].  addend has the shape of dx, may be NULL.  Stride 2 (the strided blocks of
 * try_with_aspp_remove_max_pool.py:176,210,231 and train.py:411-447) runs as one tensor-core launch per input-pixel
 * parity class, stored through a strided view of dx. */
HG_API int hg_conv_dgrad(const HgConvDesc* d, const void* dy, const void* w_dgrad, const void* addend, void* dx,
                  void* stream);

/* Same as hg_conv_fprop plus an optional fp32 NCHW copy [N,Cout,H,W] of the result: the tensors
 * creatModel.forward returns to the training loop (try_with_torch.py:291-292,298). */
HG_API int hg_conv_fprop_ex(const HgConvDesc* d, const void* x, const void* w_fprop, const float* bias,
                     const void* residual, void* y, float* stats, float* out_nchw, void* stream);

/* dw_packed[R*S][Cout_p][Cin_p] (fp32, channels padded to 64) += sum over pixels of dy (x) x ;
 * dbias[Cout] += sum(dy).  Both ACCUMULATE: the reference shares one weight between many call sites
 * (try_with_torch.py:217,224-237), so one buffer collects all of them.  Either may be NULL. */
HG_API int hg_conv_wgrad(const HgConvDesc* d, const void* x, const void* dy, float* dw_packed, float* dbias,
                  void* stream);

/* dw_oihw[Cout,Cin,R,S] (=|+=) dw_packed : the layout torch.optim.Adam sees (try_with_torch.py:317,344). */
HG_API int hg_unpack_conv_wgrad(const HgConvDesc* d, const float* dw_packed, float* dw_oihw, int accumulate,
                         void* stream);

/* ---- BatchNorm folded into the convolutions around it (tensor-core path only) -----------------------------
 * The reference runs BN -> ReLU -> conv as three library kernels (try_with_torch.py:196-205).  Here the
 * normalised activation a = [relu](gamma*(x-mean)*invstd + beta) is never written to HBM: the convolution reads the
 * RAW tensor x and applies the transform to its operand tiles in shared memory (fprop, wgrad), and the data-gradient
 * kernel applies the ReLU mask and accumulates the two BatchNorm-backward sums in its epilogue.
 * HgBnFold describes that BatchNorm call site; `stats` = the [3*Cp] statistics slot of x as accumulated by the kernel
 * that produced x (training mode), or running statistics (eval mode, use_running = 1). */
typedef struct HgBnFold {
  const float* stats;
  const float* gamma;
  const float* beta;
  const float* running_mean;
  const float* running_var;
  float eps;
  int32_t relu;
  int32_t use_running;
  int32_t pad_;
} HgBnFold;

/* 1 when the tensor-core kernels take this geometry: bf16, stride 1 with "same" padding (any dilation) or stride 2
 * halving the map (3x3 pad 1 / 1x1 pad 0), power-of-two OUTPUT map at most 128 wide, <= 256 padded channels.  Every
 * other bf16 geometry is HG_ERR_UNSUPPORTED unless hg_set_option("allow_ref_conv", 1) opted into the CUDA-core kernel. */
HG_API int hg_conv_tc_eligible(const HgConvDesc* d);
/* 1 when, additionally, the *_bn entry points below take it (stride 1). */
HG_API int hg_conv_fold_eligible(const HgConvDesc* d);
/* y = conv([relu](bn(x_raw))) + bias [+ residual]; stats / out_nchw as in hg_conv_fprop_ex. */
HG_API int hg_conv_fprop_bn(const HgConvDesc* d, const HgBnFold* bn, const void* x_raw, const void* w_fprop,
                            const float* bias, const void* residual, void* y, float* stats, float* out_nchw,
                            void* stream);
/* Inference: convolution followed by an eval-mode BatchNorm(+ReLU) of its OUTPUT channels in one kernel,
 * y = [relu](gamma * (conv(x) + bias - running_mean) / sqrt(running_var + eps) + beta): the nn.Conv2d -> nn.BatchNorm2d
 * -> nn.ReLU runs inside every ResidualBlock (try_with_torch.py:196-205) and `lin` (:243-256) under model.eval().
 * bn_out->use_running must be 1; tensor-core geometries only (hg_conv_tc_eligible). */
HG_API int hg_conv_fprop_bnout(const HgConvDesc* d, const HgBnFold* bn_out, const void* x, const void* w_fprop,
                               const float* bias, void* y, float* out_nchw, void* stream);
/* dw_packed += dy (x) [relu](bn(x_raw)); dbias += sum(dy). */
HG_API int hg_conv_wgrad_bn(const HgConvDesc* d, const HgBnFold* bn, const void* x_raw, const void* dy,
                            float* dw_packed, float* dbias, void* stream);
/* g = conv_transpose(dy, w) * [bn(x_raw) > 0]  (the mask only when bn->relu), stored in the layout of x_raw;
 * red[0..Cp) += sum g, red[Cp..2Cp) += sum g * xhat  (caller zeroes red).  hg_bn_bwd_apply(da = g, ...) finishes
 * the BatchNorm backward. */
HG_API int hg_conv_dgrad_bn(const HgConvDesc* d, const HgBnFold* bn, const void* dy, const void* w_dgrad,
                            const void* x_raw, void* g, float* red, void* stream);

/* Slices of a wider weight: the convolution over torch.cat([a, b, c], 1) is evaluated as three chained
 * convolutions (residual = previous partial sum) over the input-channel slices [cin_offset, cin_offset + d->Cin)
 * of the [Cout, cin_total, R, S] weight, so the concatenated tensor never exists
 * (try_different_stack.py:316-328, try_with_aspp_remove_max_pool.py:239-240,291-303). */
HG_API int hg_pack_conv_weight_slice(const HgConvDesc* d, const float* w_oihw, int cin_total, int cin_offset,
                                     void* w_fprop, void* w_dgrad, void* stream);
HG_API int hg_unpack_conv_wgrad_slice(const HgConvDesc* d, const float* dw_packed, float* dw_oihw, int cin_total,
                                      int cin_offset, int accumulate, void* stream);

/* out[R,cols] (=|+=) T[R,R] (transpose ? ^T : ) * in[R,cols], all fp32: linear recombination of head channels.
 * The in-place limb mix of try_skeleton_and_keypoints.py:279-298 (t[:,19+l] = t[:,19+l] - t[:,0] + t[:,a_l] +
 * t[:,b_l]) is linear in the head's output, so it is folded into the head's weights (W_eff = T W, b_eff = T b)
 * and un-folded from their gradients (dW = T^T dW_eff). */
HG_API int hg_mix_rows(const float* T, const float* in, float* out, int R, int cols, int transpose, int accumulate,
                       void* stream);
/* Rectangular form: out[rows_out, cols] (=|+=) T[rows_out, rows_in] * in[rows_in, cols]; with transpose = 1,
 * out[rows_in, cols] (=|+=) T^T * in[rows_out, cols].  The gather-add limb maps of
 * try_skeleton_from_keypoints_merge.py:296-298 (tmpOut = cat(k, k[:, a_l] + k[:, b_l]): 17 keypoint maps -> 36
 * channels) are T = [I; G] applied to the 17-channel head: folded into the head's weights like the in-place mix. */
HG_API int hg_mix_rows_rect(const float* T, const float* in, float* out, int rows_out, int rows_in, int cols,
                            int transpose, int accumulate, void* stream);

/* Switches: "allow_ref_conv" = 1 lets bf16 convolutions outside the tensor-core geometry run on the CUDA-core kernels
 * (default 0: they fail with HG_ERR_UNSUPPORTED -- no silent slow path); "force_ref_conv" = 1 routes EVERY bf16
 * convolution there (validation).  Kernel selection (defaults = measured best): "persist_1x1" (1) /
 * "persist_3x3" (0): large-map stride-1 convolutions through the persistent kernel, from "persist_min_units" (512)
 * 128-pixel units on; "wgrad_halo" (1): 3x3 weight gradients of the 64x64 maps with one x box per filter column;
 * "upsample_sep" (1): separable up-sampling kernels; "pdl" (1): programmatic dependent launch. */
HG_API int hg_set_option(const char* name, int value);

/* ---- BatchNorm2d + ReLU (try_with_torch.py:184-192,196-204,249-250,254-255) ------------------------- */
/* x is [M, C] NHWC-flattened (M = N*H*W, channels padded to 64 in memory).  Training mode
 * (use_running = 0) normalises with the batch statistics slot `stats` = {S1[Cp], S2[Cp], pivot[Cp]} that the
 * producing kernel accumulated (hg_conv_fprop's `stats`, or hg_bn_stats); eval mode uses running_mean/var. */
typedef struct HgBnDesc {
  int64_t M;
  int32_t C;
  int32_t dtype;
  float eps;
  int32_t relu;        /* fuse nn.ReLU after the affine transform */
  int32_t use_running; /* 0 = batch statistics (train), 1 = running statistics (eval) */
} HgBnDesc;

HG_API int hg_bn_stats(const HgBnDesc* d, const void* x, float* stats, void* stream); /* slot [3*Cp], see hg_conv_fprop */
/* Initialise the statistics slots of a forward pass in one launch: S1 = S2 = 0, pivot = pivot_src[c] (the running mean
 * of the BatchNorm that consumes the tensor; NULL = 0).  slots_dev is a DEVICE array. */
typedef struct HgBnStatsSlot {
  float* stats;           /* [3*Cp] */
  const float* pivot_src; /* [C] or NULL */
  int32_t C, Cp;
} HgBnStatsSlot;
HG_API int hg_bn_prepare_stats(const HgBnStatsSlot* slots_dev, int num_slots, void* stream);
HG_API int hg_bn_apply(const HgBnDesc* d, const void* x, const float* stats, const float* gamma, const float* beta,
                const float* running_mean, const float* running_var, void* y, void* stream);
/* red[0..Cp) += sum g, red[Cp..2Cp) += sum g*xhat, g = da * [bn(x) > 0]  (caller zeroes red).  In eval mode xhat is
 * taken from the running statistics; the sums are then only the parameter gradients (dbeta, dgamma). */
HG_API int hg_bn_bwd_reduce(const HgBnDesc* d, const void* da, const void* x, const float* stats, const float* gamma,
                     const float* beta, const float* running_mean, const float* running_var, float* red,
                     void* stream);
/* dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)) [+ addend]; dgamma += , dbeta += ; colsum[C] (optional)
 * += sum over rows of dx without the addend = bias gradient of the convolution that produced x. */
HG_API int hg_bn_bwd_apply(const HgBnDesc* d, const void* da, const void* x, const float* stats, const float* gamma,
                    const float* beta, const float* running_mean, const float* running_var, const float* red,
                    const void* addend, void* dx, float* dgamma, float* dbeta, float* colsum, void* stream);

/* Running-statistics update for a whole forward pass: module `i` applies the EMA of its call sites
 * sites[first_site .. first_site+num_sites) in order and adds num_sites to num_batches_tracked (a shared
 * BatchNorm is called 6-8 x nStack times per forward: try_with_torch.py:224-237).  Both tables live in
 * device memory. */
typedef struct HgBnRunningSite {
  const float* stats;
  float count;
  int32_t pad_;
} HgBnRunningSite;
typedef struct HgBnRunningModule {
  float* running_mean;
  float* running_var;
  int64_t* num_batches_tracked;
  int32_t C, Cp;
  int32_t first_site, num_sites;
  float momentum;
  int32_t pad_;
} HgBnRunningModule;
HG_API int hg_bn_update_running(const void* modules_dev, const void* sites_dev, int num_modules, void* stream);

/* ---- spatial ops -------------------------------------------------------------------------------- */
/* nn.MaxPool2d(2) (try_with_torch.py:220,226,265); backward routes to the first row-major maximum.
 * stats (optional, here and in hg_upsample2x_add_fwd): the [3*Cp] statistics slot of the tensor just written,
 * accumulated for the BatchNorm that reads it, like hg_conv_fprop's `stats`. */
HG_API int hg_maxpool2_fwd(int dtype, const void* x, int N, int H, int W, int C, void* y, float* stats, void* stream);
HG_API int hg_maxpool2_bwd(int dtype, const void* x, const void* dy, const void* addend, int N, int H, int W, int C,
                    void* dx, void* stream);
/* out[N,2h,2w,C] = upsample_x2(low[N,h,w,C]) [+ skip]; mode 0 = bilinear align_corners=True
 * (try_with_torch.py:238-239), mode 1 = nearest (hourglass_compare.py:532-542). */
HG_API int hg_upsample2x_add_fwd(int dtype, int mode, const void* low, const void* skip, int N, int h, int w, int C,
                          void* out, float* stats, void* stream);
HG_API int hg_upsample2x_bwd(int dtype, int mode, const void* dout, const void* addend, int N, int h, int w, int C,
                      void* dlow, void* stream);
/* Image-level branch of ASPP (train.py:476-479,488-489): y[n,c] = scale * sum_hw x[n,h,w,c] [+ addend[n,c]]
 * (nn.AdaptiveAvgPool2d((1,1)) with scale = 1/(H*W)) and its transpose x[n,h,w,c] = scale * y[n,c] [+ addend[n,h,w,c]]
 * (F.interpolate of a 1x1 map, bilinear align_corners=True, = broadcast); each is the other's backward. */
HG_API int hg_spatial_mean(int dtype, const void* x, int N, int H, int W, int C, float scale, const void* addend, void* y,
                           void* stream);
HG_API int hg_spatial_broadcast(int dtype, const void* y, int N, int H, int W, int C, float scale, const void* addend,
                                void* x, void* stream);
/* Channel-window copy between NHWC tensors of `rows` pixels (channels padded to 64 in memory):
 * dst[r, dst_c0 + c] = src[r, src_c0 + c] [+ addend[r, dst_c0 + c]], c < channels (rounded up to 8).  One call per
 * input of a materialised torch.cat(xs, 1) (train.py:528-538,570-583); with the roles swapped, its backward. */
HG_API int hg_channel_copy(int dtype, const void* src, int src_channels, int src_c0, const void* addend, void* dst,
                           int dst_channels, int dst_c0, int channels, long long rows, void* stream);
HG_API int hg_add(int dtype, const void* a, const void* b, void* out, long long n_elems, void* stream);
/* layout changes at the module boundary (NCHW fp32 tensors of the training loop <-> NHWC activations) */
HG_API int hg_nchw_f32_to_nhwc(int dtype, const float* src_nchw, const void* addend, int N, int C, int H, int W, void* dst,
                        void* stream);
HG_API int hg_nhwc_to_nchw_f32(int dtype, const void* src, int N, int C, int H, int W, float* dst_nchw, void* stream);

/* ---- stem: Conv2d(3,64,7,2,3) + ReLU on the fp32 NCHW image batch (try_with_torch.py:262,276-277) ----
 * dtype HG_BF16: tcgen05 GEMMs on an im2col tile built in shared memory (image as a hi + lo bf16 pair in the forward, as
 * bf16 in the weight gradient; weights as bf16); dtype HG_F32, or hg_set_option("stem_tc", 0): CUDA-core kernels on the
 * fp32 operands. */
/* relu = 1: ReLU fused (try_with_torch.py:276-277); relu = 0: raw output for the BatchNorm that follows in
 * hourglass_compare.py:549-552.  bias may be NULL. */
HG_API int hg_stem_fwd(int dtype, const float* x_nchw, const float* w_oihw, const float* bias, int N, int H, int W,
                       int relu, void* y, void* stream);
/* dw_oihw += , dbias += ; with relu = 1 the ReLU mask is taken from y */
HG_API int hg_stem_bwd(int dtype, const float* x_nchw, const void* y, const void* dy, int N, int H, int W, int relu,
                       float* dw_oihw, float* dbias, void* stream);

/* ---- intermediate-supervision MSE loss (try_with_torch.py:305-308,333-341) --------------------------- */
/* loss[s] += mean((pred_s - target)^2) for s < num_stacks (caller zeroes loss[num_stacks]); optionally
 * dpred_s = 2 * grad_scale * (pred_s - target) / numel, the gradient `sum_s loss_s`.backward() sends into every
 * returned heatmap.  preds_host / dpreds_host are HOST arrays of num_stacks DEVICE pointers (fp32 tensors of
 * `numel` elements, 16-byte aligned); dpreds_host or any of its entries may be NULL. */
#define HG_MSE_MAX_STACKS 8
typedef struct HgMseDesc {
  int64_t numel;
  int32_t num_stacks;
  float grad_scale;
} HgMseDesc;
HG_API int hg_mse_multi(const HgMseDesc* d, const float* const* preds_host, const float* target,
                        float* const* dpreds_host, float* loss, void* stream);

/* cudaMemsetAsync(p, 0, bytes) on `stream` (a memset node when captured into a CUDA graph): the gradient / reduction
 * arenas of a step are cleared without an elementwise fill kernel. */
HG_API int hg_zero_async(void* p, int64_t bytes, void* stream);

/* In place t_s[i] *= scales[s] for s < num_tensors (<= HG_MSE_MAX_STACKS) fp32 tensors of `numel` elements (16-byte
 * aligned; tensors_host is a HOST array of DEVICE pointers, NULL entries are skipped; scales is a DEVICE array): the
 * upstream gradient of `losses[s]` applied to the per-stack gradients hg_mse_multi already produced -- one launch where
 * autograd's `g * gloss[s]` is one elementwise kernel per stack (try_with_torch.py:341 `loss.backward()`). */
HG_API int hg_scale_multi(int64_t numel, int32_t num_tensors, float* const* tensors_host, const float* scales,
                          void* stream);

/* ---- per-pixel class cross-entropy heads --------------------------------------------------------- */
/* nn.CrossEntropyLoss (mean over labels != ignore_index) on fp32 NCHW logits with int64 [B,H,W] labels
 * (only_one_hourgless.py:348,370; try_different_stack.py:360-361,388-389; the [:, :18] / [:, 18:] slices of every
 * stack's output in try_skeleton_and_keypoints.py:390-397,423-435).  Every term of a training step goes through one
 * call: term t reads `channels` consecutive channel planes starting at `logits` (point it at the first channel of
 * the slice; images are `logits_bstride` floats apart), adds its loss to loss[t] (caller zeroes loss[T] and count[T])
 * and, when dlogits != NULL, writes grad_scale * (softmax - onehot) / count into the same planes of the gradient
 * tensor (0 for ignored pixels).  *bad_label (may be NULL) is set to 1 if a label is outside [0, channels) and is
 * not ignore_index; such pixels are treated as ignored. */
#define HG_CE_MAX_TERMS 16
typedef struct HgCeTerm {
  const float* logits;
  float* dlogits;
  const int64_t* target;
  int64_t logits_bstride;
  int64_t dlogits_bstride;
  int32_t channels;
  float norm;                /* > 0: loss and gradient are divided by this instead of the valid-label count */
  const float* pixel_weight; /* optional [B,H,W]: per-pixel weight of loss and gradient (selection mask / mask input) */
  float* nll_out;            /* optional [B,H,W]: the per-pixel negative log-likelihood (0 for ignored labels) */
} HgCeTerm;
typedef struct HgCeDesc {
  int32_t num_terms;
  int32_t B;
  int32_t HW;
  int32_t ignore_index;
  float grad_scale;
} HgCeDesc;
HG_API int hg_ce_multi(const HgCeDesc* d, const HgCeTerm* terms_host, float* loss, int32_t* count, int32_t* bad_label,
                       void* stream);
/* Bootstrapping of train.py:343-362,394-408 (torch.topk(loss.view(B, -1), k) then mean): mask[r, i] = 1 for the k
 * largest of the n values of row r (ties at the k-th value: lowest indices), else 0; kth_value[r] (optional) receives
 * the k-th largest value.  Bootstrapped cross-entropy = hg_ce_multi (nll_out) -> hg_topk_mask -> hg_ce_multi with
 * pixel_weight = mask and norm = B * k. */
HG_API int hg_topk_mask(const float* values, int rows, int n, int k, float* mask, float* kth_value, void* stream);

/* ---- input pipeline tail (next row N1) ---------------------------------------------------------- */
/* transforms.ToTensor() + transforms.Normalize(mean, std) (try_with_torch.py:310-313) on the GPU:
 * dst[n,c,h,w] = (src[n,h,w,c] / 255 - mean[c]) / std[c] in torchvision's fp32 operation order (bit-exact).
 * src uint8 NHWC (device), mean_host / std_host HOST arrays of C floats, C <= 4. */
HG_API int hg_image_u8_to_nchw_f32(const uint8_t* src_nhwc, int N, int H, int W, int C, const float* mean_host,
                                   const float* std_host, float* dst_nchw, void* stream);

/* PIL's `image.resize([256, 256])` (try_with_torch.py:99; default filter BICUBIC) for a batch of variable-size RGB
 * images, bit-exact with Pillow's Resample.c on 8-bit data (two passes, horizontal first, 22-bit fixed-point
 * coefficients, uint8 intermediate).  images_dev: DEVICE array of descriptors (src = uint8 HWC RGB pixels of one
 * image; *_off = offsets, in ints, of its coefficient [out x ksize] and bounds [out x 2] tables inside coef_dev /
 * bounds_dev; tmp_off = byte offset of its [h, out_w, 3] intermediate image inside tmp).  The tables are Pillow's
 * precompute_coeffs + normalize_coeffs_8bpc, computed by the host.  Outputs (either may be NULL): out_nhwc uint8
 * [B, out_h, out_w, 3] and out_nchw_norm fp32 [B, 3, out_h, out_w] = ToTensor + Normalize(mean, std) of it. */
typedef struct HgResizeImage {
  const uint8_t* src;
  int32_t w, h;
  int32_t kx_off, ky_off;
  int32_t bx_off, by_off;
  int32_t ksize_x, ksize_y;
  int64_t tmp_off;
} HgResizeImage;
HG_API int hg_resize_bicubic_u8(const HgResizeImage* images_dev, int num_images, int max_h, int out_w, int out_h,
                                const int32_t* coef_dev, const int32_t* bounds_dev, uint8_t* tmp, uint8_t* out_nhwc,
                                float* out_nchw_norm, const float* mean_host, const float* std_host, void* stream);

/* ---- optimizer step (next row N4) ------------------------------------------------------------------ */
/* torch.optim.Adam (amsgrad = False) on every parameter tensor in one launch (try_with_torch.py:317,342-344).
 * chunks_dev is a DEVICE array: thread block b updates chunks_dev[b] (a slice of at most a few thousand elements of
 * one tensor; fp32 param / grad / exp_avg / exp_avg_sq).  `step` is the 1-based step count AFTER the increment;
 * lr_d / beta1_d / beta2_d carry the double-precision hyper-parameters python passes (bias corrections are evaluated
 * in double like the reference), lr / beta1 / beta2 their float copies (validation only). */
typedef struct HgAdamChunk {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t n;
} HgAdamChunk;
typedef struct HgAdamDesc {
  double lr_d, beta1_d, beta2_d;
  float lr, beta1, beta2, eps, weight_decay;
  int32_t step;
  int32_t num_chunks;
  int32_t pad_;
} HgAdamDesc;
HG_API int hg_adam_multi(const HgAdamDesc* d, const HgAdamChunk* chunks_dev, void* stream);

/* Weighted / bootstrapped MSE (train.py:379-408): sq_out[i] = (pred - target)^2 (optional; input of hg_topk_mask);
 * loss += sum w * (pred - target)^2 / norm and dpred = 2 * grad_scale * w * (pred - target) / norm (each optional),
 * w = weight[i] (per element) or weight[b, hw] broadcast over the C channels (weight_per_pixel = 1), 1 when NULL. */
HG_API int hg_mse_weighted(const float* pred, const float* target, const float* weight, int weight_per_pixel, int B, int C,
                           int HW, float norm, float grad_scale, float* sq_out, float* dpred, float* loss, void* stream);

/* ---- target rendering ---------------------------------------------------------------------------- */
/* Gaussian keypoint heatmaps, evaluated in float64 like the numpy code, stored as float32
 * (try_with_torch.py:107-132; variants try_with_torch_100.py:64-85, only_one_hourgless.py:112-132,
 * hourglass_compare.py:286-313,713-734).  keypoints [B,P,J,3] = (x, y, v) in image pixels, img_wh [B,2]. */
typedef struct HgGaussDesc {
  int32_t B, P, J, H, W;
  int32_t center_mode; /* 0: c = kp / size * W   1: c = kp * 256 / size / 4 (MPII) */
  int32_t truncate;    /* 1: centre truncated to an integer (astype(np.int)) */
  int32_t accumulate;  /* 0: only the LAST person is kept (quirk Q7)  1: sum over persons */
  double pre_scale;    /* 1, or 100 (try_with_torch_100.py:81) */
  double sigma;
  double amplitude;    /* 1, or 1/(2 pi sigma^2) */
} HgGaussDesc;
HG_API int hg_render_gauss(const HgGaussDesc* d, const double* keypoints, const int32_t* num_persons, const double* img_wh,
                    float* out, void* stream);

/* PIL ImageDraw.point / ImageDraw.line label maps, int64 [B,H,W]
 * (try_different_stack.py:114-155, try_skeleton_and_keypoints.py:93-114). */
typedef struct HgLabelDesc {
  int32_t B, P, J, L, H, W;
  int32_t center_mode;
  int32_t draw_points; /* 1: draw.point value k+1 for visible joints; 2: MPII draw.ellipse on the float centre +-0.5
                          (train.py:681-686) */
  int32_t draw_lines;  /* draw.line for limbs whose both ends are visible */
  int32_t line_value;  /* 0: limb index + 1; > 0: this constant (background map = 1); < 0: the limb index itself
                          (try_skeleton_from_keypoints_merge.py:130-133) */
} HgLabelDesc;
HG_API int hg_render_labels(const HgLabelDesc* d, const double* keypoints, const int32_t* num_persons, const double* img_wh,
                     const int32_t* limbs, int64_t* out, void* stream);

/* Annotation -> keypoint tensor on the device (next row N1: the `anno.loadAnns(...)` / `annopoints.point` loops of
 * try_with_torch.py:103-113 and hourglass_compare.py:691-703).  The dataset's annotations are uploaded ONCE:
 *   mode 0 (COCO): table[row] = one person's J*3 doubles (x, y, v) as in the JSON, offset[i .. i+1) = rows of image i;
 *                  an image with more than P persons keeps its LAST P (the last person wins in the reference, quirk Q7);
 *   mode 1 (MPII): table[row] = one annotated point (id, x, y, is_visible), offset[i .. i+1) = points of sample i,
 *                  scattered into one person: kp[id] = (x, y, is_visible != 0), later records overwrite earlier ones.
 * wh_all[i] = (width, height) of image i.  For the B samples `sample_index` the kernel writes keypoints [B,P,J,3],
 * num_persons [B] and img_wh [B,2] -- the inputs of hg_render_gauss / hg_render_labels. */
typedef struct HgAnnotDesc {
  int32_t B, P, J;
  int32_t mode;
} HgAnnotDesc;
HG_API int hg_gather_annotations(const HgAnnotDesc* d, const double* table, const int32_t* offset, const double* wh_all,
                                 const int64_t* sample_index, double* keypoints, int32_t* num_persons, double* img_wh,
                                 void* stream);

/* ---- decode + PCKh -------------------------------------------------------------------------------- */
/* first row-major (y, x) of the maximum of each of num_maps [H,W] maps
 * (hourglass_compare.py:831,1092; only_one_hourgless.py:294-295). */
HG_API int hg_decode_argmax(const void* heatmaps, int dtype, int num_maps, int H, int W, int32_t* out_yx, float* out_max,
                     void* stream);
/* PCKh threshold sweep (hourglass_compare.py:812-844; performance_compare.py:544-615). Counters are int32
 * [B,nthr] and must be zeroed by the caller; found[B,njoints] marks joints present in the label map. */
HG_API int hg_pckh_sweep(const void* x, int dtype, int B, int C, int H, int W, const int64_t* target, const float* rect,
                  int chan_offset, int njoints, const float* thresholds, int nthr, int32_t* correct, int32_t* total,
                  int32_t* predict_xy, int32_t* label_xy, int32_t* found, float* standard, void* stream);
/* PCKh "D" (calculate_parameters.py:906-937): same decode and label lookup as hg_pckh_sweep, but a joint is correct
 * when  sqrt(d2) < standard * factors[s]  (the reference uses the single factor 0.5; float32 throughout). */
HG_API int hg_pckh_abs(const void* x, int dtype, int B, int C, int H, int W, const int64_t* target, const float* rect,
                int chan_offset, int njoints, const float* factors, int nfac, int32_t* correct, int32_t* total,
                int32_t* predict_xy, int32_t* label_xy, int32_t* found, float* standard, void* stream);
/* Fused `softmax over channels -> PCKh` (next row N3: pckh.forward(softmax(result[2]), y_keypoints, rect),
 * hourglass_compare.py:1160, performance_compare.py:646-647): the probabilities are never written.
 * hg_softmax_stats: max_sum[b,h,w] = {max_c x, sum_c exp(x - max)} (float2 per pixel, fp32 NCHW logits);
 * hg_pckh_logits: hg_pckh_sweep (absolute = 0) / hg_pckh_abs (absolute = 1) whose decoded value is
 * exp(x - max) / sum, PyTorch's softmax expression, evaluated while the logits are scanned. */
HG_API int hg_softmax_stats(const float* logits, int B, int C, int H, int W, float* max_sum, void* stream);
HG_API int hg_pckh_logits(const float* logits, const float* max_sum, int absolute, int B, int C, int H, int W,
                   const int64_t* target, const float* rect, int chan_offset, int njoints, const float* thresholds,
                   int nthr, int32_t* correct, int32_t* total, int32_t* predict_xy, int32_t* label_xy, int32_t* found,
                   float* standard, void* stream);
/* PCKh "A" (only_one_hourgless.py:285-313): counts[0] += correct, counts[1] += total. */
HG_API int hg_pckh_a(const void* x, int x_dtype, const void* target, int t_dtype, int B, int Cx, int Ct, int H, int W,
              int njoints, int head_ch, int neck_ch, int32_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HG_SM100A_H_ */
