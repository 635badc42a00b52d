/*
 * segk.h — C ABI of libsegk.so: the B200 (sm_100a) segmentation hot path.
 *
 * The reference (SeunghwanByun/SemanticSegmentation_Tensorflow) has no FFI of its own:
 * its boundary is a set of Python helpers that each wrap one TensorFlow op.  Every entry
 * point below names the reference call site it replaces (paths relative to the upstream
 * repo root).  All tensors are caller-owned DEVICE pointers; every call enqueues work on
 * the caller's `stream` (a cudaStream_t passed as void*) and returns without
 * synchronising.  No torch types, no hidden allocation on the hot path.
 *
 * Layouts: activations NHWC bf16 (channel pitch = `ld*` elements where given), images
 * u8/f32 NHWC, logits f32 NHWC, master weights fp32 in the reference's HWIO order
 * (`[kh,kw,Cin,Cout]`; transposed conv `[kh,kw,Cout,Cin]`, FCN.py:125,143).
 *
 * Return value: 0 = OK, negative = error (see enum); text via segk_last_error().
 * There is NO CPU fallback: an unsupported shape returns SEGK_EINVAL.
 */
#ifndef SEGK_H
#define SEGK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEGK_ABI_VERSION 1

enum {
  SEGK_OK = 0,
  SEGK_EINVAL = -1,  /* unsupported shape / dtype / null pointer */
  SEGK_ECUDA = -2,   /* CUDA runtime or driver error */
  SEGK_ENOMEM = -3
};

typedef struct segk_ctx segk_ctx;

/* ---- context ----------------------------------------------------------------------
 * One context per device.  All work is enqueued on the caller's stream.  A context owns a few grow-only
 * device scratch buffers; calls that share one must not overlap on DIFFERENT streams:
 *   - segk_conv2d_wgrad / segk_deconv2d_wgrad (per-split partial sums)            -> one stream
 *   - split-K segk_conv2d_fwd / _dgrad, segk_conv2d_first_wgrad, any dgrad with dx_colsum,
 *     segk_maxpool2x2_bwd with dbias                                              -> one stream
 *   - segk_bias_grad                                                              -> one stream
 *   - segk_conv2d_small_wgrad / segk_deconv2d_small_wgrad                         -> one stream
 * (fcn.py: wgrad stream / main stream / side stream / main stream).  Use one context per stream otherwise. */
int segk_abi_version(void);
int segk_create(int device, segk_ctx** out);
int segk_destroy(segk_ctx* ctx);
const char* segk_last_error(segk_ctx* ctx);
/* number of kernels launched through this ctx since creation (bench.py gpu_launches) */
int64_t segk_launch_count(segk_ctx* ctx);
int segk_sm_count(segk_ctx* ctx);
/* Kernel-selection overrides (also read once at segk_create from SEGK_SLAB / SEGK_SLAB3 / SEGK_WSLAB /
 * SEGK_TMA_STORE / SEGK_TEAMK / SEGK_HYBRID / SEGK_FORCE_BN / SEGK_FORCE_KSPLIT / SEGK_FORCE_WSPLIT): key in {"slab", "slab3",
 * "wslab" (0 off, 1 auto, 2 wherever legal), "tma_store" (0|1), "teamk" (0|1: lockstep tap-split schedule instead of plain
 * split-K for few-tile / long-K layers), "hybrid" (0|1: whole waves + K-split remainder tiles), "tail_wide" (0|1: 16-byte-access forms of conv8 / conv_t1 and their gradients), "force_bn" (0|64|128|256), "force_ksplit", "force_wsplit" (0 = heuristic)}.  Every setting computes the same sums (fp32
 * accumulation order differs between kernels). */
int segk_set_tuning(segk_ctx* ctx, const char* key, int value);

/*
 * Channel-strided tensor views = zero-copy Concat (utils.py:332 `tf.concat(x, axis)`; FCDenseNet.py:141-157's decoder idiom,
 * the skip concats of the U-Net): a producer writes its [N,H,W,C] output straight into channels [c0, c0+C) of the wider
 * concat buffer [N,H,W,P], and a consumer of the concat's gradient reads its slice of it in place.  A view is the pointer to
 * its first element plus the pitch P (channels between consecutive pixels; a multiple of 8, first element 16-byte aligned).
 * segk_set_pitch applies to the NEXT call on this context only (0 = dense) and is understood by:
 *   out_pitch:  y  of segk_conv2d_fwd / segk_conv2d_fwd_pool (bf16, no residual / bits), y of segk_deconv2d_fwd (bf16, no
 *               residual), dx (+ act, residual: the same view geometry) of segk_maxpool2x2_bwd (full-resolution mask mode)
 *   in_pitch:   dy of segk_conv2d_dgrad / segk_conv2d_wgrad / segk_deconv2d_dgrad / segk_deconv2d_wgrad / segk_bias_grad
 *               (bf16), x of segk_maxpool2x2_fwd and of segk_conv2d_fwd(_pool)
 * Any other entry point fails with SEGK_EINVAL while a pitch is pending (nothing silently reads a view as dense).
 * The tensor-core kernels address views through their TMA descriptors, so a view costs nothing over a dense tensor.
 */
int segk_set_pitch(segk_ctx* ctx, int in_pitch, int out_pitch);

/* ---- epilogue flags for the conv family ------------------------------------------------ */
#define SEGK_EPI_RELU 1u      /* y = max(y,0) after bias (+residual)            */
#define SEGK_EPI_OUT_F32 2u   /* y stored as fp32 instead of bf16               */

/*
 * conv_layer forward: relu(conv2d(x, W, stride 1, SAME) + b)        (FCN.py:117-136)
 *   x [N,H,W,Cin] bf16; wk = kernel-layout weights [kh*kw][Cin/64][Cout][64] bf16 produced by
 *   segk_pack_conv_weights; bias fp32 [Cout] or NULL; residual (same shape as y, bf16) is
 *   added before the ReLU when non-NULL (the `fuse` skip-add, FCN.py:169-171).
 *   tcgen05 path: requires Cin % 64 == 0 and Cout % 64 == 0, odd kh,kw, kh*kw <= 64.
 */
int segk_conv2d_fwd(segk_ctx* ctx, const void* x, const void* wk, const float* bias,
                    const void* residual, void* y, uint32_t* relu_bits, int N, int H, int W, int Cin,
                    int Cout, int kh, int kw, unsigned flags, void* stream);
/* conv_layer to a FEW output channels on the tensor cores (a k x k conv to num_classes, e.g. SegNet.py:80's 3x3 64 -> 2
 * head): wk / bias hold the weights zero-padded to Cout (a multiple of 64) output channels; only the first out_cols
 * columns of the result are stored, as fp32 y [N,H,W,out_cols].  Backward: segk_pad_channels(dy [rows][out_cols] ->
 * bf16 [rows][Cout]) and the ordinary segk_conv2d_dgrad / segk_conv2d_wgrad on the padded tensors. */
int segk_conv2d_fwd_narrow(segk_ctx* ctx, const void* x, const void* wk, const float* bias, float* y, int out_cols,
                           int N, int H, int W, int Cin, int Cout, int kh, int kw, unsigned flags, void* stream);
/* dst bf16 [rows][C] = src (fp32 or bf16) [rows][c] zero-padded to C channels (C % 8 == 0) */
int segk_pad_channels(segk_ctx* ctx, const void* src, int src_is_f32, void* dst, int64_t rows, int c, int C,
                      void* stream);
/* the same words from a finished bf16 tensor y [rows][C] (C % 32 == 0): bits[r][w] bit i = (y[r][32 w + i] > 0) */
int segk_relu_bits(segk_ctx* ctx, const void* y, uint32_t* bits, int64_t rows, int C, void* stream);
/* relu_bits (or NULL): 1-bit ReLU mask of the bf16 output, u32 [N*H*W][Cout/32], bit i of word w = (y[.., 32 w + i] > 0),
 * written by the same epilogue (Cout % 32 == 0).  The consumer layer's segk_conv2d_dgrad takes it as relu_mask_bits:
 * the ReluGrad of this layer then reads 1/16 of the bytes of the bf16 activation with one coalesced load. */

/*
 * conv_layer followed by max_pool (FCN.py:54-56, 58-60, 62-65, 67-71, 73-76: conv -> ReLU -> max_pool 2x2/2, the
 * max_pool helper of FCN.py:161-163): the same forward conv with the pool computed in its epilogue from the staged
 * bf16 tile.  pooled [N,H/2,W/2,Cout] bf16 and idx [N,H/2,W/2,Cout] u8 (first-max window position 0..3, the layout
 * of segk_maxpool2x2_fwd) are bit-identical to segk_conv2d_fwd + segk_maxpool2x2_fwd.  pool_only != 0: y need not be
 * written (FCN-8s training never reads the pre-pool tensor again; y must still be a valid buffer -- geometries
 * whose tiles do not hold whole 2x2 windows run the two kernels one after the other).
 */
int segk_conv2d_fwd_pool(segk_ctx* ctx, const void* x, const void* wk, const float* bias, void* y,
                         void* pooled, uint8_t* idx, int pool_only, int N, int H, int W, int Cin,
                         int Cout, int kh, int kw, unsigned flags, void* stream);

/*
 * Conv2DBackpropInput of the above (part of what `optimizer.minimize` emits, FCN.py:340):
 *   dx = dy (*) rot180(W);  wd = dgrad-layout weights [kh*kw (taps reversed)][Cout/64][Cin][64] bf16.
 *   dx = ((acc + residual) masked by relu_mask > 0) * scale, where
 *   relu_mask (shape of dx, bf16, or NULL) is the forward activation whose ReluGrad is
 *   fused here, residual (or NULL) a second gradient path into the same tensor (AddN at
 *   pool3/pool4), and scale = 1/keep_prob folds the dropout backward (FCN.py:165-167).
 *   dx_colsum (fp32 [Cin] or NULL): the column sums of dx over all pixels from the same pass, i.e. the
 *   BiasAddGrad of the PRODUCER conv_layer, whose pre-activation gradient dx is (summed in the epilogue
 *   from the fp32 values, fixed order; a separate pass over dx only for split-K / weight-heavy layers).
 */
int segk_conv2d_dgrad(segk_ctx* ctx, const void* dy, const void* wd, const void* relu_mask,
                      const uint32_t* relu_mask_bits, const void* residual, void* dx, float* dx_colsum,
                      float scale, int N, int H, int W, int Cin, int Cout, int kh, int kw, void* stream);
/* relu_mask_bits (instead of relu_mask, not both): the producer's mask as written by segk_conv2d_fwd /
 * segk_conv2d_first_fwd (relu_bits), u32 [N*H*W][Cin/32]; same result bit for bit. */

/*
 * Conv2DBackpropFilter of conv_layer (FCN.py:340): dw[kh,kw,Cin,Cout] fp32 HWIO
 * (+= when accumulate != 0, else overwritten).  Split-K over pixels: every split stores its partial
 * sums into its own slice of a context-owned workspace and an ordered reduction adds them
 * (deterministic, no atomics).  Requires N*H*W tileable into 64-pixel boxes.
 */
int segk_conv2d_wgrad(segk_ctx* ctx, const void* x, const void* dy, float* dw, int N, int H,
                      int W, int Cin, int Cout, int kh, int kw, int accumulate, void* stream);

/*
 * deconv_layer forward: conv2d_transpose(x, W[k,k,Cout,Cin], stride s, SAME) + b, k = 2s
 * (FCN.py:138-159).  x [N,H,W,Cin] bf16 -> y [N,sH,sW,Cout]; wk = phase-packed weights
 * from segk_pack_deconv_weights; residual = skip tensor added in the epilogue (fuse_1 /
 * fuse_2, FCN.py:92,96).  tcgen05 path: Cin % 64 == 0 and Cout % 64 == 0.
 */
int segk_deconv2d_fwd(segk_ctx* ctx, const void* x, const void* wk, const float* bias,
                      const void* residual, void* y, int N, int H, int W, int Cin, int Cout,
                      int k, int s, unsigned flags, void* stream);

/* gradient of deconv_layer wrt its input = conv2d(dy, W, stride s);
 * dy [N,sH,sW,Cout] bf16, wd [k*k][Cout/64][Cin][64] bf16, dx [N,H,W,Cin] bf16. k=4, s=2.
 * dx_colsum: as for segk_conv2d_dgrad. */
int segk_deconv2d_dgrad(segk_ctx* ctx, const void* dy, const void* wd, const void* relu_mask,
                        void* dx, float* dx_colsum, int N, int H, int W, int Cin, int Cout, int k, int s,
                        void* stream);
/* Conv2D with a 2s x 2s kernel, stride s, SAME padding (the 4x4 stride-2 encoder convs of LidCamNet.py:28-33): the same
 * strided implicit GEMM with the conv_layer epilogue.  x [N,sH,sW,Cin] bf16 -> y [N,H,W,Cout] = relu(conv(x, W, stride s)
 * + b); wd = the dgrad layout segk_pack_deconv_weights produces from the HWIO weights W [k,k,Cin,Cout] (read as the
 * weights of a transposed conv Cout -> Cin).  Its input gradient is segk_deconv2d_fwd(dy, wk, ...) and its weight
 * gradient segk_deconv2d_wgrad with x and dy swapped (x := dy [N,H,W,Cout], dy := x [N,sH,sW,Cin]); k = 4, s = 2. */
int segk_conv2d_strided_fwd(segk_ctx* ctx, const void* x, const void* wd, const float* bias, void* y, int N, int H,
                            int W, int Cin, int Cout, int k, int s, unsigned flags, void* stream);

/* gradient of deconv_layer wrt W[k,k,Cout,Cin] (fp32). k=4, s=2. */
int segk_deconv2d_wgrad(segk_ctx* ctx, const void* x, const void* dy, float* dw, int N, int H,
                        int W, int Cin, int Cout, int k, int s, int accumulate, void* stream);

/* ---- first layer on the tensor cores without a patch tensor --------------------------------
 * conv_layer on the network input (conv1_1, FCN.py:52; U-Net first Conv2D, utils.py:165): 3x3 SAME,
 * Cin in {1,3,4}, Cout in {64,128,256}.  Each pixel's 3x3xCin patch is built in shared memory
 * straight from the image (x_dtype 2 = u8 as fed at FCN.py:312, 0 = bf16) and multiplied on
 * tcgen05; wk is the segk_pack_im2col_weights layout [Cout][64] bf16.  y bf16, bias + optional
 * SEGK_EPI_RELU fused. */
int segk_conv2d_first_fwd(segk_ctx* ctx, const void* x, int x_dtype, const void* wk,
                          const float* bias, void* y, uint32_t* relu_bits, int N, int H, int W, int Cin,
                          int Cout, int kh, int kw, unsigned flags, void* stream);

/* Conv2DBackpropFilter (+ BiasAddGrad when dbias != NULL) of that layer (FCN.py:340): reads the
 * image and dy once.  dw fp32 [kh,kw,Cin,Cout] (overwritten), dbias fp32 [Cout] (overwritten). */
int segk_conv2d_first_wgrad(segk_ctx* ctx, const void* x, int x_dtype, const void* dy, float* dw,
                            float* dbias, int N, int H, int W, int Cin, int Cout, int kh, int kw,
                            void* stream);

/* ---- small-channel layers on CUDA cores (conv1_1 Cin=3/4, conv8 Cout=2, conv_t1 Cin=2,
 *      conv_t3 Cout=2) ----------------------------------------------------------------- */
/* conv_layer forward for ragged channel counts.  x_dtype: 0 = bf16, 2 = u8 (raw image,
 * FCN.py:312).  w fp32 HWIO, output bf16.  Supported: K = kh*kw*Cin <= 64 (any Cout); 1x1 with
 * Cout in {2,4,8} and Cin % 8 == 0; odd k <= 5 with Cout in {2,4} and Cin % 8 == 0 (the 3x3 head
 * of the reference's SegNet, SegNet.py:80) -- the last two also with SEGK_EPI_OUT_F32: fp32 logits. */
int segk_conv2d_small_fwd(segk_ctx* ctx, const void* x, int x_dtype, const float* w,
                          const float* bias, void* y, int N, int H, int W, int Cin, int Cout,
                          int kh, int kw, unsigned flags, void* stream);
/* 1x1 with Cout in {2,4,8}, or odd k <= 5 with Cout in {2,4}: dx = ((dy * rot180(W)^T) masked by
 * relu_mask > 0) * scale */
int segk_conv2d_small_dgrad(segk_ctx* ctx, const void* dy, const float* w, const void* relu_mask,
                            void* dx, float scale, int N, int H, int W, int Cin, int Cout,
                            int kh, int kw, void* stream);
int segk_conv2d_small_wgrad(segk_ctx* ctx, const void* x, int x_dtype, const void* dy, float* dw,
                            int N, int H, int W, int Cin, int Cout, int kh, int kw,
                            void* stream);
/* deconv_layer (any Cin/Cout, k = 2s) on CUDA cores: fwd / dgrad / wgrad; w fp32
 * [k,k,Cout,Cin]; dy may be fp32 (the logits gradient) or bf16. */
int segk_deconv2d_small_fwd(segk_ctx* ctx, const void* x, const float* w, const float* bias,
                            const void* residual, void* y, int N, int H, int W, int Cin,
                            int Cout, int k, int s, unsigned flags, void* stream);
int segk_deconv2d_small_dgrad(segk_ctx* ctx, const void* dy, int dy_is_f32, const float* w,
                              const void* relu_mask, void* dx, int N, int H, int W, int Cin,
                              int Cout, int k, int s, void* stream);
int segk_deconv2d_small_wgrad(segk_ctx* ctx, const void* x, const void* dy, int dy_is_f32,
                              float* dw, int N, int H, int W, int Cin, int Cout, int k, int s,
                              void* stream);

/* ---- ragged-channel layers as tensor-core GEMMs via a patch tensor ---------------------- */
/* conv1_1 (Cin 3/4, FCN.py:52): P[n,y,x,kk] bf16, kk = (ky*kw+kx)*Cin+ci zero-padded to 64, so
 * the layer becomes a 1x1 conv over P (segk_conv2d_fwd / segk_conv2d_wgrad with Cin = 64).
 * x_dtype 2 = u8 image, 0 = bf16. */
int segk_im2col_k64(segk_ctx* ctx, const void* x, int x_dtype, void* P, int N, int H, int W,
                    int Cin, int kh, int kw, void* stream);
/* w[K][Cout] fp32 (HWIO flattened) -> wk[Cout][64] bf16, zero-padded, for the 1x1 conv above */
int segk_pack_im2col_weights(segk_ctx* ctx, const float* w, void* wk, int K, int Cout,
                             void* stream);
/* conv_t3 (16x16 s8, Cout = 2, FCN.py:98-107) backward: P[n,i,j,(ky,kx,co)] bf16 =
 * dy[n, s*i-p+ky, s*j-p+kx, co] (0 outside), so that dx = P . W^T and dW = P^T . x are 1x1
 * GEMMs (segk_conv2d_fwd / segk_conv2d_wgrad with Cin = k*k*Cout). */
int segk_deconv_patch_gather(segk_ctx* ctx, const void* dy, int dy_is_f32, void* P, int N, int H,
                             int W, int Cout, int k, int s, void* stream);
/* conv_t3 forward: yp[n,i,j,(ky,kx,co)] fp32 = x . W (a 1x1 GEMM) -> y[n,oy,ox,co] = b[co] + the
 * 4 overlapping patch entries (+ residual).  y fp32 (logits) or bf16. */
int segk_deconv_col2im(segk_ctx* ctx, const float* yp, const float* bias, const void* residual,
                       void* y, int out_f32, int N, int H, int W, int Cout, int k, int s,
                       void* stream);
/* ---- phase-packed transposed conv for tiny Cout (conv_t3: 16x16 s8, 256 -> 2, FCN.py:98-107) ---------------
 * k = 2s, SAME: the s x s output pixels [r*s-p, (r+1)*s-p) x [c*s-p, (c+1)*s-p) (p = s/2) depend only on the
 * 2 x 2 input neighbourhood (r-1+dy, c-1+dx), so the layer is ONE 4-tap stride-1 implicit GEMM over the
 * (H+1) x (W+1) block grid with R = s*s*Cout columns (a,b,co) whose epilogue writes each row's block of fp32
 * logits (+ bias) directly -- no patch-space tensor, no col2im pass.  Needs Cin % 64 == 0, R in {64,128,256}.
 *   segk_pack_deconv_packed : w[k,k,Cout,Cin] fp32 -> bf[t][Cin/64][R][64] (fwd) and bt[t][R/64][Cin][64] (dgrad),
 *                             t = dy*2+dx, value W[a+s(1-dy)][b+s(1-dx)][co][ci]
 *   segk_deconv_pack_dy     : dy [N,sH,sW,Cout] (fp32 or bf16) -> dyb[N,H+1,W+1,R] bf16 blocks (0 outside the image)
 *   segk_deconv2d_packed_dgrad : dx = 4-tap igemm over dyb (ReLU mask / dx_colsum as segk_conv2d_dgrad)
 *   segk_deconv2d_packed_wgrad : dwt[t][Cin][R] fp32 (overwritten); segk_deconv_unpack_dw -> dw[k,k,Cout,Cin] */
int segk_pack_deconv_packed(segk_ctx* ctx, const float* w, void* bf, void* bt, int k, int s, int Cin, int Cout,
                            void* stream);
int segk_deconv2d_packed_fwd(segk_ctx* ctx, const void* x, const void* bf, const float* bias, float* y, int N,
                             int H, int W, int Cin, int Cout, int k, int s, void* stream);
int segk_deconv_pack_dy(segk_ctx* ctx, const void* dy, int dy_is_f32, void* dyb, int N, int H, int W, int Cout,
                        int s, void* stream);
int segk_deconv2d_packed_dgrad(segk_ctx* ctx, const void* dyb, const void* bt, const void* relu_mask, void* dx,
                               float* dx_colsum, int N, int H, int W, int Cin, int Cout, int k, int s,
                               void* stream);
int segk_deconv2d_packed_wgrad(segk_ctx* ctx, const void* x, const void* dyb, float* dwt, int N, int H, int W,
                               int Cin, int Cout, int k, int s, void* stream);
int segk_deconv_unpack_dw(segk_ctx* ctx, const float* dwt, float* dw, int k, int s, int Cin, int Cout,
                          int accumulate, void* stream);
/* generic matrix pack: w fp32 [T][A][B] -> cp bf16 [T][ceil(B/64)][A][64] (rows A, k = B) and/or
 * tr bf16 [T][ceil(A/64)][B][64] (rows B, k = A) */
int segk_pack_matrix(segk_ctx* ctx, const float* w, void* cp, void* tr, int T, int A, int B,
                     void* stream);

/* ---- shared-helper layers of the secondary builders (Network/utils/utils.py) -------------- */
/* Batch_Normalization (utils.py:300-301) is tf.layers.batch_normalization with training=False and
 * never-updated moving stats: y = gamma * x / sqrt(1 + 1e-3) + beta.  The scale is folded into the
 * packed conv weights (out = in * scale[c] * mult, e.g. W * gamma / sqrt(1+eps)) and back out of the
 * weight gradient; beta is the conv epilogue's bias.  in/out fp32 [rows][C], out may alias in. */
int segk_scale_columns(segk_ctx* ctx, const float* in, const float* scale, float mult, float* out,
                       int64_t rows, int C, void* stream);
/* Batch_Normalization gradients without a pass over activations (utils.py:300-301, inference-mode affine folded
 * into the conv weights W' = W * gamma * mult): given the weight gradient of the folded conv in gw [rows][C] (HWIO
 * flattened, C = Cout) and the unfolded weights w, writes dgamma[c] = mult * sum_k w[k][c] * gw[k][c] and rescales
 * gw in place to the gradient of the unfolded weights (gw[k][c] *= gamma[c] * mult).  d(beta) is the BiasAddGrad
 * of dz (segk_bias_grad).  workspace: >= 4 * C * min(ceil(rows / 64), ceil(2 * SMs / ceil(C / 32))) bytes of
 * device scratch for the per-block partial rows (deterministic two-stage sum). */
int segk_bn_unfold_grads(segk_ctx* ctx, float* gw, const float* w, const float* gamma, float mult,
                         float* dgamma, void* workspace, size_t workspace_bytes, int64_t rows, int C,
                         void* stream);

/* dgamma[c] = sum_r dz[r][c] * (y[r][c] - beta[c]) / gamma[c]  (dz, y bf16 [rows][C]; dz already
 * carries the ReLU mask, so (y - beta)/gamma' is the raw conv output wherever dz != 0).
 * dbeta (nullable): also dbeta[c] = sum_r dz[r][c] (the BiasAddGrad of the same dz) from the same pass.
 * workspace: >= 8 * C * min(rows/4+1, 4*SMs) bytes (8 MB always suffices for C <= 4096). */
int segk_bn_gamma_grad(segk_ctx* ctx, const void* dz, const void* y, const float* beta,
                       const float* gamma, float* dgamma, float* dbeta, void* workspace,
                       size_t workspace_bytes, int64_t rows, int C, void* stream);
/* The same two gradients for an fp32 head with C in {2,4,8} channels (dz, y fp32 [rows][C]): the
 * Batch_Normalization the reference's SegNet applies to its logits (SegNet.py:80-81). */
int segk_bn_grads_f32(segk_ctx* ctx, const float* dz, const float* y, const float* beta, const float* gamma,
                      float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                      int64_t rows, int C, void* stream);
/* ---- dense-block builders (Network/model/FCDenseNet.py:23-60) ---------------------------------------------
 * Tensors of these builders are stored with every channel segment padded to a multiple of 8 and the total to a
 * multiple of 64 (pads are zero), so each conv is a tcgen05 GEMM over 64-wide channel chunks; parameters are
 * remapped logical <-> physical with `map[physical] = logical index or -1`.  Rows are `ld` elements apart, so a
 * kernel can work on a channel PREFIX of a concat buffer (the zero-copy view of Concat(layers_concat), :55).
 *
 * Pre-activation Batch_Normalization + ReLU (utils.py:300-303; BN is the inference-mode affine):
 *   y[r][c] = act(x[r][c] * scale[c] + shift[c]), c < C, with scale = gamma / sqrt(1 + 1e-3), shift = beta
 *   (physical arrays); relu != 0 applies max(., 0). */
int segk_bn_act_fwd(segk_ctx* ctx, const void* x, int ldx, void* y, int ldy, const float* scale, const float* shift,
                    int64_t rows, int C, int relu, const uint8_t* drop_mask, float drop_keep, uint64_t drop_seed,
                    void* stream);
/* its gradient: g = dy * [y > 0]; dx (=, or += when accumulate) g * scale; dscale[c] = sum_r g * x, dshift[c] = sum_r g
 * (two-stage, deterministic).  dy / y have row pitch ldy, x / dx row pitch ldx.  workspace >= the _bytes() value. */
size_t segk_bn_act_bwd_workspace_bytes(segk_ctx* ctx, int C);
int segk_bn_act_bwd(segk_ctx* ctx, const void* dy, const void* y, int ldy, const void* x, void* dx, int ldx,
                    const float* scale, float* dscale, float* dshift, void* workspace, size_t workspace_bytes,
                    int64_t rows, int C, int relu, int accumulate, const uint8_t* drop_mask, float drop_keep,
                    uint64_t drop_seed, void* stream);
/* drop_keep in (0,1) (both calls): the Dropout in front of the BN (FCDenseNet.py:27-28: conv -> Dropout -> BN -> ReLU) is
 * applied to x on the fly, with segk_dropout's keep pattern over the [rows][ldx] tensor x (drop_mask: injected u8 mask
 * or NULL = Philox(drop_seed)); the backward also applies the DropoutGrad to dx (no accumulate).  drop_keep >= 1: none. */
/* Avg_Pooling 2x2 / stride 2 / VALID (utils.py:309) and AvgPoolGrad (dx = dy / 4 on every window element) */
int segk_avgpool2x2_fwd(segk_ctx* ctx, const void* x, int ldx, void* y, int ldy, int N, int H, int W, int C,
                        void* stream);
int segk_avgpool2x2_bwd(segk_ctx* ctx, const void* dy, int lddy, void* dx, int lddx, int N, int H, int W, int C,
                        void* stream);
/*
 * Remaining op families of Network/model, HBM-bound kernels on NHWC bf16 with C % 8 == 0 (opfam.cu).
 * Avg_Pooling(x, kh, kw, stride_h, stride_w, padding VALID) (utils.py:309; the pyramid-pooling windows of
 * PSPNet.py:147-165,546-567): y [N,(H-kh)/sh+1,(W-kw)/sw+1,C]; AvgPoolGrad: dx = sum over the windows containing the pixel
 * of dy / (kh kw). */
int segk_avgpool_fwd(segk_ctx* ctx, const void* x, void* y, int N, int H, int W, int C, int kh, int kw, int sh, int sw,
                     void* stream);
int segk_avgpool_bwd(segk_ctx* ctx, const void* dy, void* dx, int N, int H, int W, int C, int kh, int kw, int sh, int sw,
                     void* stream);
/* Max_Pooling(x, kh, kw, stride, padding) (utils.py:306; 3x3 / stride 2 VALID at PSPNet.py:34, SAME at PSPNet.py:190),
 * same != 0: TF SAME geometry (out = ceil(in / stride), padding ignored by the max).  idx [N,OH,OW,C] u8 = position
 * (ky * kw + kx) of the first maximum of every window; MaxPoolGrad routes dy there (overlapping windows add up). */
int segk_maxpool_fwd(segk_ctx* ctx, const void* x, void* y, uint8_t* idx, int N, int H, int W, int C, int kh, int kw,
                     int stride, int same, void* stream);
int segk_maxpool_bwd(segk_ctx* ctx, const void* dy, const uint8_t* idx, void* dx, int N, int H, int W, int C, int kh,
                     int kw, int stride, int same, void* stream);
/* tf.nn.depthwise_conv2d(x, filter [kh,kw,C,1], strides, 'SAME', rate) (DeepLabv3Plus.py:49,117; EfficientNet.py:173,453;
 * Generative_Segmentation_net.py:53): channel multiplier 1, w fp32 [kh][kw][C], y [N,ceil(H/s),ceil(W/s),C] bf16
 * (+ bias, ReLU by flags); rate > 1 needs stride 1 as in TF.  dgrad: dx [N,H,W,C]; wgrad: dw fp32 [kh][kw][C]
 * (two-stage fixed-order sum, += when accumulate). */
int segk_depthwise_conv2d_fwd(segk_ctx* ctx, const void* x, const float* w, const float* bias, void* y, int N, int H,
                              int W, int C, int kh, int kw, int stride, int rate, unsigned flags, void* stream);
int segk_depthwise_conv2d_dgrad(segk_ctx* ctx, const void* dy, const float* w, void* dx, int N, int H, int W, int C,
                                int kh, int kw, int stride, int rate, void* stream);
int segk_depthwise_conv2d_wgrad(segk_ctx* ctx, const void* x, const void* dy, float* dw, int N, int H, int W, int C,
                                int kh, int kw, int stride, int rate, int accumulate, void* stream);
/* kind 0: tf.sigmoid, kind 1: swish x * sigmoid(x) (EfficientNet.py's activation) on n bf16 elements (n % 8 == 0);
 * backward from the INPUT x: dx = dy * f'(x). */
int segk_activation_fwd(segk_ctx* ctx, const void* x, void* y, int64_t n, int kind, void* stream);
int segk_activation_bwd(segk_ctx* ctx, const void* x, const void* dy, void* dx, int64_t n, int kind, void* stream);
/* squeeze-excite multiply (EfficientNet.py SE block: x * sigmoid(se) broadcast over H, W): y[n,p,c] = x[n,p,c] * s[n,c],
 * x [N,HW,C], s [N,C] bf16; backward: dx = dy * s, ds[n,c] = sum_p dy x (fp32 [N,C]). */
int segk_channel_scale_fwd(segk_ctx* ctx, const void* x, const void* s, void* y, int N, int64_t HW, int C, void* stream);
int segk_channel_scale_bwd(segk_ctx* ctx, const void* x, const void* s, const void* dy, void* dx, float* ds, int N,
                           int64_t HW, int C, void* stream);
/* logical w[T][A][B] fp32 <-> physical wp[T][Ap][Bp]: to_phys != 0: wp = mapped ? w : 0; else w[mapped] = wp.
 * amap[Ap] / bmap[Bp] device int32 arrays (NULL = identity on the first A / B entries). */
int segk_remap_weights(segk_ctx* ctx, float* w, float* wp, int T, int A, int B, int Ap, int Bp, const int* amap,
                       const int* bmap, int to_phys, void* stream);
/* per-channel parameters: dst[i] = map[i] >= 0 ? src[map[i]] * mul + add : 0; and dst[map[i]] = src[i] * mul */
int segk_gather_f32(segk_ctx* ctx, const float* src, const int* map, float* dst, int n, float mul, float add,
                    void* stream);
int segk_scatter_f32(segk_ctx* ctx, const float* src, const int* map, float* dst, int n, float mul, void* stream);

/* ---- remaining op families of Network/utils/utils.py -----------------------------------------
 * Atrous_Conv2D_Layer (utils.py:210-231: tf.nn.atrous_conv2d(x, W, rate, SAME)) on the same tcgen05 implicit GEMM:
 * the taps of the tap table sit `rate` pixels apart; same operands / layouts / epilogue as segk_conv2d_*. */
int segk_atrous_conv2d_fwd(segk_ctx* ctx, const void* x, const void* wk, const float* bias, const void* residual,
                           void* y, int N, int H, int W, int Cin, int Cout, int kh, int kw, int rate,
                           unsigned flags, void* stream);
int segk_atrous_conv2d_dgrad(segk_ctx* ctx, const void* dy, const void* wd, const void* relu_mask,
                             const void* residual, void* dx, float* dx_colsum, float scale, int N, int H, int W,
                             int Cin, int Cout, int kh, int kw, int rate, void* stream);
int segk_atrous_conv2d_wgrad(segk_ctx* ctx, const void* x, const void* dy, float* dw, int N, int H, int W, int Cin,
                             int Cout, int kh, int kw, int rate, int accumulate, void* stream);
/* Resize_Bilinear (utils.py:329-330: tf.image.resize_bilinear(x, size, align_corners=True)) x [N,H,W,C] -> y [N,OH,OW,C]
 * bf16, fp32 interpolation in TF's evaluation order; its gradient in gather form (deterministic). */
int segk_resize_bilinear_fwd(segk_ctx* ctx, const void* x, void* y, int N, int H, int W, int OH, int OW, int C,
                             void* stream);
int segk_resize_bilinear_bwd(segk_ctx* ctx, const void* dy, void* dx, int N, int H, int W, int OH, int OW, int C,
                             void* stream);
/* Global_Avg_Pool (utils.py:312-313: tflearn global_avg_pool = mean over H, W): x [N,H,W,C] -> y [N,C]; gradient dx = dy / (H W) */
int segk_global_avgpool_fwd(segk_ctx* ctx, const void* x, void* y, int N, int H, int W, int C, void* stream);
int segk_global_avgpool_bwd(segk_ctx* ctx, const void* dy, void* dx, int N, int H, int W, int C, void* stream);
/* Global_Max_Pool (utils.py:315-316: tflearn global_max_pool = reduce_max over H, W): x [N,H,W,C] bf16 -> y [N,C];
 * count [N,C] int32 = number of maximal elements, for the gradient, which TF's reduce_max spreads equally over ties:
 * dx = (x == y) ? dy / count : 0 */
int segk_global_maxpool_fwd(segk_ctx* ctx, const void* x, void* y, int* count, int N, int H, int W, int C, void* stream);
int segk_global_maxpool_bwd(segk_ctx* ctx, const void* dy, const void* x, const void* y, const int* count, void* dx,
                            int N, int H, int W, int C, void* stream);
/* Zero_Padding (utils.py:325-327: tf.pad with `pad` zeros on each side of H and W): crop == 0: y [N,H+2p,W+2p,C] <-
 * x [N,H,W,C]; crop == 1 (its gradient): y [N,H,W,C] <- the centre of x [N,H+2p,W+2p,C].  bf16, C % 8 == 0. */
int segk_zero_pad(segk_ctx* ctx, const void* x, void* y, int N, int H, int W, int C, int pad, int crop, void* stream);

/* Concat (utils.py:332) and its gradient: dst[r][coff_dst + c] (=, or += when accumulate)
 * src[r][coff_src + c] for c < C, zeroed where mask[r][c] <= 0 (mask dense [rows][C] or NULL =
 * the fused ReluGrad of the producer).  bf16, all channel counts multiples of 8. */
int segk_channel_copy(segk_ctx* ctx, const void* src, int ld_src, int coff_src, void* dst,
                      int ld_dst, int coff_dst, const void* mask, int accumulate, int64_t rows, int C,
                      int drop_side, const uint8_t* drop_mask, float drop_keep, uint64_t drop_seed,
                      void* stream);
/* drop_side 1 / 2 with drop_keep in (0,1): tf.nn.dropout of the copied values on the fly, the keep pattern indexed by
 * the elements of the source (1) or destination (2) tensor -- the Dropout between a dense-block conv and its concat
 * slot (FCDenseNet.py:33-34,55) and its gradient.  0: plain copy. */

/* ---- weight layout packing (fp32 master -> bf16 kernel layouts) ------------------------
 * Kernel layouts are BLOCKED by 64-wide chunks of the contraction dimension k: an operand [T taps][rows][K] is
 * stored as [T][ceil(K/64)][rows][64] (zero-padded in K), so the operand tile of one (tap, k-chunk) is one
 * contiguous run of rows x 128 bytes (a TMA box).  Any Cin / Cout; outputs 16-byte aligned. */
/* w[kh,kw,Cin,Cout] fp32 (HWIO, FCN.py:125) -> wk[tap][Cin/64][Cout][64] bf16 (fwd: rows = Cout, k = Cin) and
 * wd[ntaps-1-tap][Cout/64][Cin][64] bf16 (dgrad: rows = Cin, k = Cout; taps reversed = rot180).
 * Either may be NULL. */
int segk_pack_conv_weights(segk_ctx* ctx, const float* w, void* wk, void* wd, int kh, int kw,
                           int Cin, int Cout, void* stream);
/* w[k,k,Cout,Cin] fp32 (k = 2s, FCN.py:143) ->
 *   wk[(ay*s+ax)*4 + uy*2+ux][Cin/64][Cout][64] bf16 = w[ay+s*(1-uy)][ax+s*(1-ux)]  (deconv fwd phases)
 *   wd[ky*k+kx][Cout/64][Cin][64] bf16                                                (deconv dgrad)   */
int segk_pack_deconv_weights(segk_ctx* ctx, const float* w, void* wk, void* wd, int k, int s,
                             int Cin, int Cout, void* stream);

/* ---- HBM-bound kernels --------------------------------------------------------------- */
/* max_pool 2x2/s2 VALID (FCN.py:161-163): y [N,H/2,W/2,C] bf16 and the in-window index of
 * the first maximal element (0..3, row-major) as u8 — bit-exact vs the oracle. */
int segk_maxpool2x2_fwd(segk_ctx* ctx, const void* x, void* y, uint8_t* idx, int N, int H, int W,
                        int C, void* stream);
/* MaxPoolGrad from the stored index, fused with the ReluGrad of the pooled activation:
 * dx[n,y,x,c] = ((idx==k ? dy : 0) + residual) * [act > 0]  (act may be NULL: no mask; residual,
 * shape of dx or NULL, is a second gradient path into the same tensor, e.g. a skip connection).
 * act_is_pooled != 0: `act` is the POOLED tensor y [N,H/2,W/2,C] instead of the pre-pool one (residual
 * must be NULL): at the routed element the pre-pool activation equals the pooled value, so the result is
 * bit-identical while a quarter of the mask bytes are read.
 * dbias (fp32 [C], pooled mode only, or NULL): BiasAddGrad of the conv_layer in front of the pool from the same
 * pass -- db[c] = sum over dx[..., c] = sum of the masked dy (every dy lands on exactly one window element). */
int segk_maxpool2x2_bwd(segk_ctx* ctx, const void* dy, const uint8_t* idx, const void* act, int act_is_pooled,
                        const void* residual, void* dx, float* dbias, int N, int H, int W, int C, void* stream);

/* tf.nn.dropout (FCN.py:165-167): y = x * keep_mask / keep_prob.  mask (u8 0/1) is used
 * when non-NULL (parity runs); otherwise Philox4x32-10(seed, element index / 8), 16 bits per element. Same call is
 * the backward (pass dy as x). */
int segk_dropout(segk_ctx* ctx, const void* x, void* y, const uint8_t* mask, int64_t n,
                 float keep_prob, uint64_t seed, void* stream);

/* reduce_mean(softmax_cross_entropy_with_logits) + its gradient + argmax (FCN.py:334,111):
 *   logits f32 [npix,C] (2 <= C <= 64), labels u8 class ids [npix]; a label >= C is ignored (no loss, zero
 *   gradient, not counted; the mean still divides by npix);
 *   dlogits (f32, may be NULL) = (softmax - onehot) * grad_scale   (grad_scale =
 *   1/(world*N*H*W)); pred (u8, may be NULL) = argmax, ties -> 0;
 *   loss_sum (f32[2]): [0] = sum over pixels of the per-pixel loss (deterministic two-stage),
 *   [1] = that sum / npix, i.e. the reduce_mean of FCN.py:334;
 *   cm (int64[4], may be NULL) += confusion counts cm[gt*2+pred].
 *   workspace: >= segk_xent_workspace_bytes(npix) bytes. */
size_t segk_xent_workspace_bytes(int64_t npix);
int segk_softmax_xent_fwd_bwd(segk_ctx* ctx, const float* logits, const uint8_t* labels,
                              float* dlogits, uint8_t* pred, float* loss_sum, int64_t* cm,
                              void* workspace, int64_t npix, int C, float grad_scale,
                              void* stream);
/* tf.nn.softmax + road mask (FCN.py:229,204-206) + tf.argmax (FCN.py:111): prob f32 [npix,C],
 * mask u8 [npix] = prob[...,1] > 0.5, argmax u8 [npix] = first maximal class (each may be NULL) */
int segk_softmax_infer(segk_ctx* ctx, const float* logits, float* prob, uint8_t* mask, uint8_t* argmax,
                       int64_t npix, int C, void* stream);
/* the annotation placeholder as fed by the reference (one-hot [npix,C], FCN.py:313,195-201; dtype 1 = f32,
 * 2 = u8 / bool) -> u8 class ids [npix] (first maximal channel) for segk_softmax_xent_fwd_bwd */
int segk_onehot_to_ids(segk_ctx* ctx, const void* onehot, int dtype, uint8_t* ids, int64_t npix, int C,
                       void* stream);
/* paste_mask (FCN.py:203-211): the road overlay of gen_test_output — where mask != 0 the RGBA colour
 * ([0,255,0,127] in the reference) is alpha-blended over the u8 image (C = 3 or 4 channels; for
 * C = 4 the alpha channel is blended too, as PIL's Image.paste does), bit-exact with PIL's
 * integer blend; elsewhere the image is copied. */
int segk_overlay_mask(segk_ctx* ctx, const uint8_t* image, const uint8_t* mask, uint8_t* out,
                      int64_t npix, int C, int r, int g, int b, int a, void* stream);
/* road / non-road confusion matrix (new): cm[gt*2+pred] += counts */
int segk_confusion_matrix(segk_ctx* ctx, const uint8_t* gt, const uint8_t* pred, int64_t* cm,
                          int64_t npix, void* stream);

/* tf.train.AdamOptimizer ApplyAdam over a flat fp32 arena (FCN.py:338-340), TF formula:
 *   m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t m / (sqrt(v) + eps),
 *   lr_t = lr sqrt(1-b2^t)/(1-b1^t) computed by the caller.  g is multiplied by
 *   grad_scale first (1.0, or 1/world when the allreduce was a plain sum of means). */
int segk_adam_step(segk_ctx* ctx, float* p, float* m, float* v, const float* g, int64_t n,
                   float lr_t, float beta1, float beta2, float eps, float grad_scale,
                   void* stream);
/* The same update on `nranges` element ranges [offsets[i], offsets[i] + lengths[i]) of the arenas in one launch
 * (offsets / lengths: HOST int64 arrays) -- the small variables between the conv weights of a gradient bucket. */
int segk_adam_step_ranges(segk_ctx* ctx, float* p, float* m, float* v, const float* g, const int64_t* offsets,
                          const int64_t* lengths, int nranges, float lr_t, float beta1, float beta2, float eps,
                          float grad_scale, void* stream);
/* The same update for ONE conv_layer weight tensor [kh,kw,Cin,Cout] (Cin, Cout multiples of 64) fused with
 * segk_pack_conv_weights: the fresh parameters are written back and, from the same pass, as bf16 into wk / wd.
 * col_scale (fp32 [Cout] or NULL), col_mult: the packed copies hold p[.., co] * col_scale[co] * col_mult -- the
 * inference-mode Batch_Normalization scale gamma / sqrt(1 + eps) of Conv2D_Block (utils.py:186-208,300-301) folded into
 * the kernel-layout weights, i.e. segk_scale_columns + segk_pack_conv_weights in the same pass (the fp32 master
 * weights stay unscaled; col_scale must already hold this step's updated gamma). */
int segk_adam_pack_conv_weights(segk_ctx* ctx, float* p, float* m, float* v, const float* g, void* wk, void* wd,
                                const float* col_scale, float col_mult, int kh, int kw, int Cin, int Cout, float lr_t,
                                float beta1, float beta2, float eps, float grad_scale, void* stream);
/* tf.train.MomentumOptimizer: a = mu a + g; p -= lr a (new) */
int segk_momentum_step(segk_ctx* ctx, float* p, float* a, const float* g, int64_t n, float lr,
                       float mu, float grad_scale, void* stream);

/* ---- input pipeline (get_batches_fn, FCN.py:242-305, minus PNG decode) ------- */
/* scipy.misc.imresize(..., interp='bilinear') = PIL Image.resize(BILINEAR), two passes of
 *   out = clip8((2^21 + sum_i in[i] * k[i]) >> 22)
 * with PIL's own coefficient tables (int32 [out_size][ksize]) and bounds (int32 [out_size][2] =
 * first tap, tap count), computed on the host (pipeline.py: pil_bilinear_coeffs).  Bit-exact with PIL.
 * C = 1, 3 or 4; C = 4 is RGBA (the reference's "merge" PNGs, FCN.py:225,312) with PIL's premultiplied-alpha
 * resize: RGBA -> RGBa before the horizontal pass, RGBa -> RGBA after the vertical one.
 * Horizontal pass: src u8 [H][W][C] cropped to [y0,y0+crop_h) x [x0,x0+crop_w) (crop_image,
 * FCN.py:176-182), optionally flipped horizontally (flip_image, :184-185) -> dst u8 [crop_h][out_w][C]. */
int segk_resize_h_u8(segk_ctx* ctx, const uint8_t* src, uint8_t* dst, const int* coeffs,
                     const int* bounds, int ksize, int W, int C, int x0, int y0, int crop_w,
                     int crop_h, int out_w, int flip, void* stream);
/* Vertical pass tmp u8 [in_h][out_w][C] -> out [out_h][out_w][C] with the per-image epilogue:
 * mode 0 none; mode 1 bc_img (FCN.py:186-192): uint8(clip(v*contrast + brightness, 0, 255));
 * mode 2 process_gt_image (FCN.py:194-201): out u8 [out_h][out_w] class id, 0 where the resized
 * pixel == (255,0,0) (background) else 1 (road). */
int segk_resize_v_u8(segk_ctx* ctx, const uint8_t* tmp, uint8_t* out, const int* coeffs,
                     const int* bounds, int ksize, int C, int out_w, int out_h, int mode,
                     double contrast, double brightness, void* stream);

/* ---- misc device utilities ------------------------------------------------------------ */
/* u8/f32 NHWC image -> bf16 NHWC (the feed_dict cast of FCN.py:312,395) */
int segk_cast_to_bf16(segk_ctx* ctx, const void* x, int x_dtype, void* y, int64_t n, void* stream);
/* db[c] = sum over rows of dy[rows, C] (BiasAddGrad); dy bf16 or f32 */
int segk_bias_grad(segk_ctx* ctx, const void* dy, int dy_is_f32, float* db, int64_t rows, int C,
                   void* stream);

/* ---- gradient exchange over NVLink (data parallelism; replaces ncclAllReduce) ----------
 * In-place SUM all-reduce of the fp32 elements [offset, offset + n) of a SYMMETRIC buffer (same layout on
 * every rank; offset and n multiples of 4).  Rank r reduces the r-th 1/world share and broadcasts it:
 *   multicast_ptr != NULL : NVLS -- multimem.ld_reduce / multimem.st through the NVSwitch multicast
 *                           address of the buffer (the sum is formed inside the switch);
 *   else                  : peer loads / stores; peer_ptrs = HOST array of `world` device addresses of
 *                           the buffer on each rank (rank order), summed in rank order.
 * Every rank must call it with the same (offset, n); the caller brackets the call with cross-rank
 * barriers on the same stream (inputs complete on all ranks before; all shares stored after). */
int segk_allreduce_f32(segk_ctx* ctx, void* multicast_ptr, const uint64_t* peer_ptrs, int64_t offset,
                       int64_t n, int rank, int world, void* stream);
/* Exchange fused with the optimizer (gradient all-reduce + tf.train.AdamOptimizer's ApplyAdam, FCN.py:338-340,
 * in ONE kernel over NVLS): for its 1/world share of [offset, offset + n) rank r reads the gradient sum
 * over all ranks through the multicast address of the gradient arena, updates its local m / v and the
 * parameter values (TF formula, lr_t = lr*sqrt(1-b2^t)/(1-b1^t)), and multicasts the new parameters into
 * every rank's parameter arena.  Both arenas symmetric; param_local = this rank's own mapping of the
 * parameter arena.  m / v are only maintained for the rank's own shares.  Same barrier contract. */
int segk_allreduce_adam_f32(segk_ctx* ctx, const void* grad_multicast_ptr, void* param_multicast_ptr,
                            const float* param_local, float* m, float* v, int64_t offset, int64_t n,
                            int rank, int world, float lr_t, float beta1, float beta2, float eps,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGK_H */
