/* stc_b200.h — C ABI of libstc_b200.so (sm_100a / B200 only).
 *
 * Drop-in boundary for the STC-UNet hot path.  The reference has no
 * native code: every entry point below replaces a PyTorch call made by the reference's
 * Python modules (file:line under /root/reference cited per function).  Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch); the library never
 *     allocates or frees caller tensors.  Scratch comes from `ws`/`ws_bytes` arguments.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no host sync.
 *   - activations are NHWC ("channels_last"), element type given by `dtype`
 *     (STC_F32 or STC_BF16); statistics / parameters / gradients of parameters are fp32,
 *     BatchNorm sums are fp64.
 *   - return value: 0 = ok, otherwise an STC_ERR_* code; stc_last_error() gives the text
 *     (thread local).  There is NO CPU fallback: on a device that is not cc 10.x every
 *     compute entry point returns STC_ERR_ARCH.
 */
#ifndef STC_B200_H_
#define STC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STC_OK 0
#define STC_ERR_INVALID 1
#define STC_ERR_CUDA 2
#define STC_ERR_ARCH 3
#define STC_ERR_WORKSPACE 4

#define STC_F32 0
#define STC_BF16 1

/* activation codes used by stc_bn_apply / stc_act_* */
#define STC_ACT_NONE 0
#define STC_ACT_RELU 1
#define STC_ACT_HSWISH 2
#define STC_ACT_SIGMOID 3

/* conv engine selection (stc_conv_*): AUTO = tcgen05 when dtype is bf16 and the shape is
 * eligible (Cin % 64 == 0, Cout % 16 == 0), else the fp32-accumulate SIMT kernel. */
#define STC_ENGINE_AUTO 0
#define STC_ENGINE_SIMT 1
#define STC_ENGINE_TCGEN05 2

const char* stc_last_error(void);
int stc_version(void);
/* 0 if the current device is cc 10.x, else STC_ERR_ARCH (also fills last error). */
int stc_check_device(void);
int stc_num_sms(void);

/* ---------------------------------------------------------------- layout / packing */
/* image (N,C,H,W) fp32 NCHW -> NHWC `dtype` with channels zero-padded to Cpad (input of
 * UnetBackbone.forward, unet_backbone.py:36). */
int stc_nchw_to_nhwc(const float* src, void* dst, int N, int C, int H, int W, int Cpad, int dtype, void* stream);
/* Input pipeline on the device (replaces Normalize + DefaultFormatBundle, mmseg/datasets/pipelines/formatting.py:179-217):
 * src = decoded 8-bit HWC pixels (P = N*H*W pixels, C <= 4 channels), dst = NHWC `dtype` with Cpad channels,
 * dst[p][c] = (src[p][swap_rb ? C-1-c : c] - mean[c]) * inv_std[c].  stc_widen_u8_i64: 8-bit label maps -> int64. */
int stc_image_u8_to_nhwc(const uint8_t* src, void* dst, const float* mean, const float* inv_std, long long P, int C, int Cpad, int swap_rb,
                         int dtype, void* stream);
int stc_widen_u8_i64(const uint8_t* src, int64_t* dst, long long n, void* stream);
/* The geometric part of the training pipeline on the device as well (my_config/STC-UNet.py:31-37: RandomCrop -> RandomFlip(horizontal) ->
 * Normalize -> Pad; transforms.py RandomCrop :599-614, RandomFlip :347-380): src = N decoded 8-bit images (Hs x Ws x C), geom = N x {y0, x0,
 * flip} (int32, device; drawn on the host exactly like RandomCrop.get_crop_bbox / RandomFlip), dst = N x H x W x Cpad normalised
 * activations; output pixels beyond the cropped source get pad_val (images) / seg_pad_val (labels). */
int stc_image_u8_crop_flip_to_nhwc(const uint8_t* src, void* dst, const int* geom, const float* mean, const float* inv_std, int N, int Hs, int Ws,
                                   int H, int W, int C, int Cpad, int swap_rb, float pad_val, int dtype, void* stream);
int stc_label_u8_crop_flip_i64(const uint8_t* src, int64_t* dst, const int* geom, int N, int Hs, int Ws, int H, int W, int seg_pad_val,
                               void* stream);
/* Conv2d.weight (Cout,Cin,R,S) fp32 -> packed [R*S][Cout][CinPad] `dtype` (K-major per tap).
 * transpose_flip != 0 packs the dgrad operand: [(R-1-r)*S+(S-1-s)][Cin][CoutPad] = W[co][ci][r][s]. */
int stc_pack_conv_weight(const float* w, void* dst, int Cout, int Cin, int R, int S, int inner_pad,
                         int transpose_flip, int dtype, void* stream);
/* Every weight pack of a step in one launch.  table: n rows of 8 int64 {src device pointer (fp32 OIHW), dst element offset,
 * Cout, Cin, R, S, inner_pad, mode (0 fprop, 1 dgrad, 2 im2col)}; prefix: n+1 cumulative WORK-ITEM counts, an item being one (co,ci) pair
 * (Cout*Cin per row) for modes 0/1 - which must have inner_pad == Cin resp. Cout - and one output element (Cout*inner_pad) for mode 2;
 * total = prefix[n]. */
int stc_pack_conv_weights_batched(const int64_t* table, const int64_t* prefix, int n, void* dst, long long total, int dtype,
                                  void* stream);
/* wgrad workspace [R*S][Cin][Cout] fp32 -> Conv2d.weight.grad (Cout,Cin,R,S) fp32
 * (accumulate != 0 adds into dst). */
int stc_unpack_conv_wgrad(const float* ws, float* dw, int Cout, int Cin, int R, int S, int accumulate, void* stream);
/* Small-Cin convolutions (the 3-channel image conv, unet_backbone.py:120 via InConv) run as a K = Kpad 1x1 conv on the tensor
 * cores: out[p][k] = x[p + tap][ci] with k = tap*Cin + ci (zero padded to Kpad, a multiple of 64 for the tcgen05 engine).
 * The matching weight pack is stc_pack_conv_weight(..., inner_pad = Kpad, transpose_flip = 2) -> [Cout][Kpad], and
 * stc_unpack_im2col_wgrad maps the wgrad workspace [Kpad][Cout] back to Conv2d.weight.grad (Cout,Cin,R,S). */
int stc_im2col(const void* x, void* out, int N, int H, int W, int Cin, int R, int S, int Kpad, int dtype, void* stream);
int stc_unpack_im2col_wgrad(const float* ws, float* dw, int Cout, int Cin, int R, int S, void* stream);

/* ---------------------------------------------------------------- convolution (K1/K2/K3)
 * Replaces nn.Conv2d(k, padding=k//2) in DoubleConv (unet_backbone.py:120,123; unet_head.py:67,70),
 * KernelSelectAttention (unet_backbone.py:60-67), the 1x1 convs of CoordAtt (unet_head.py:123-129)
 * and nn.Linear of TransformerLayer/KSA (unet_backbone.py:69-72,199-204) when R=S=1.
 * x: NHWC (N,H,W,Cin), wp: packed [R*S][Cout][Cin], bias fp32[Cout] or NULL, y: NHWC (N,H,W,Cout).
 * residual (same shape/dtype as y) may be NULL; act is an STC_ACT_* applied last.
 * The same entry point computes dgrad when given dy and the transpose_flip-packed weight. */
int stc_conv_fprop(const void* x, const void* wp, const float* bias, const void* residual, void* y,
                   int N, int H, int W, int Cin, int Cout, int R, int S, int act, int dtype, int engine,
                   void* stream);
/* Conv2d followed by a train-mode BatchNorm (every DoubleConv / KSA branch, unet_backbone.py:120-125,62-66): y = conv(x) + bias
 * AND the batch statistics sums = [sum_p y | sum_p y^2] (fp64, 2*Cout) of the stored outputs in one pass - out of the halo
 * kernel's epilogue when it takes the shape, else conv + stc_bn_reduce.  ws: stc_bn_ws_bytes(N*H*W, Cout) bytes of scratch. */
int stc_conv_bnstats_fused_ok(int W, int Cin, int Cout, int R, int S, int dtype, int engine);
int stc_conv_fprop_bnstats(const void* x, const void* wp, const float* bias, void* y, int N, int H, int W, int Cin, int Cout,
                           int R, int S, int dtype, int engine, double* sums, void* ws, long long ws_bytes, void* stream);
/* dW[(r,s)][ci][co] (fp32 workspace, must be zeroed by the caller when accumulate==0 is wanted)
 * += sum_pixels x[p+(r,s)][ci] * dy[p][co].  Then stc_unpack_conv_wgrad -> OIHW. */
int stc_conv_wgrad(const void* x, const void* dy, float* dw_ws, int N, int H, int W, int Cin, int Cout,
                   int R, int S, int dtype, int engine, void* stream);
/* Virtual channel concat: the conv that consumes `torch.cat([x2, x1], dim=1)` (Up.forward, unet_head.py:54-55;
 * UpConvBlock.forward, up_conv_block.py:99; the smp UNet++ decoder blocks behind unetpp_head.py:16) reads its up-to-5 NHWC sources
 * directly - the K loop walks the sources' 64-channel chunks through one TMA descriptor each - so the concatenated tensor is never
 * written; its dgrad stores every 64-channel chunk straight into the gradient tensor of the source that owns it, and its wgrad
 * reads the sources the same way.  Parts: pointers x0.. with channel counts c0.. (multiples of 64, unused = NULL / 0, packed to the
 * front), all (N,H,W,c_i) bf16; wp / the wgrad workspace use the CONCATENATED channel order.  tcgen05 engine only
 * (stc_conv_cat_ok tells; otherwise the caller materialises the concat with stc_upcat_fwd / stc_concat_channels / stc_catn_fwd). */
int stc_conv_cat_ok(int c0, int c1, int c2, int c3, int c4, int Cother, int dtype, int engine);
int stc_conv_fprop_cat(const void* x0, const void* x1, const void* x2, const void* x3, const void* x4, int c0, int c1, int c2, int c3, int c4,
                       const void* wp, const float* bias, void* y, int N, int H, int W, int Cout, int R, int S, int act, int dtype,
                       int engine, void* stream);
/* dx_i = the i-th channel block of conv_transpose(dy): dy (N,H,W,Cdy), wpt = the transpose_flip pack [taps][sum c_i][Cdy]. */
int stc_conv_dgrad_split(const void* dy, const void* wpt, void* dx0, void* dx1, void* dx2, void* dx3, void* dx4, int c0, int c1, int c2, int c3,
                         int c4, int N, int H, int W, int Cdy, int R, int S, int dtype, int engine, void* stream);
int stc_conv_wgrad_cat(const void* x0, const void* x1, const void* x2, const void* x3, const void* x4, int c0, int c1, int c2, int c3, int c4,
                       const void* dy, float* dw_ws, int N, int H, int W, int Cout, int R, int S, int dtype, int engine, void* stream);
/* column sums: out[c] (+)= sum_p x[p][c]  (Conv2d/Linear bias gradients). */
int stc_colsum(const void* x, float* out, long long P, int C, int accumulate, int dtype, void* stream);

/* ---------------------------------------------------------------- batched GEMM (K3/K4)
 * C[b] = alpha * op(A[b]) * op(B[b]) (+ beta*C) ; op by element strides, so any of NN/NT/TN/TT.
 * A(m,k) at A + b1*sA1 + b2*sA2 + m*sAm + k*sAk (elements), same for B(k,n), C(m,n) (C unit stride in n).
 * Replaces the bmm/mm inside nn.MultiheadAttention (unet_backbone.py:202,207). */
typedef struct {
    int M, N, K;
    int batch1, batch2;
    long long sA1, sA2, sAm, sAk;
    long long sB1, sB2, sBk, sBn;
    long long sC1, sC2, sCm;
    float alpha, beta;
} stc_gemm_desc;
int stc_gemm(const void* A, const void* B, void* C, const stc_gemm_desc* d, int dtype, int engine, void* stream);
/* Same contraction with bf16 operands and an fp32 result (tcgen05 engine only; errors if the descriptor is not eligible).
 * Used for the E x E products of a FOLDED pair of Linear layers: TransformerLayer applies q/k/v then MHA's in-projection, and
 * fc1 then fc2, with nothing in between (unet_backbone.py:199-208), so y = x (W2 W1)^T is ONE token GEMM; the parameter
 * gradients follow from G = dy^T x as dW2 = G W1^T and dW1 = W2^T G. */
int stc_gemm_f32out(const void* A, const void* B, float* C, const stc_gemm_desc* d, void* stream);
/* Which kernel the last stc_conv_fprop / stc_conv_wgrad / stc_gemm call on this thread actually launched (bench.py attributes
 * FLOPs and launch counts with it): STC_ENGINE_SIMT, STC_ENGINE_TCGEN05 (stc::umma_kernel), STC_KERNEL_CONVH
 * (stc::umma_convh_kernel) or STC_KERNEL_WGRADH (stc::umma_wgradh_kernel). */
#define STC_KERNEL_CONVH 3
#define STC_KERNEL_WGRADH 4
int stc_dense_last_engine(void);
/* Diagnostics: a device buffer of at least 148 * 16 int64 into which the tcgen05 kernels of the NEXT launches write per-CTA clock
 * counters (time the MMA issuer / epilogue / producers spent waiting on each barrier; layout in tools/convh_prof.py); NULL = off
 * (the default; the kernels then only test one pointer). */
int stc_debug_profile(void* counters);

/* row softmax over the last dim: P = softmax(scale * S) ; rows x L (MHA, L = H*W tokens). */
int stc_softmax_rows_fwd(const void* S, void* P, long long rows, int L, float scale, int dtype, void* stream);
/* dS = scale * P * (dP - sum_j dP_j P_j) */
int stc_softmax_rows_bwd(const void* P, const void* dP, void* dS, long long rows, int L, float scale, int dtype,
                         void* stream);
/* out[n][l][:] = x[n][l][:] - mean_l x[n][l][:]  (x: N x L x E tokens; mean_ws: N*E floats of scratch).  Used by the fp32 parity
 * path of nn.MultiheadAttention (unet_backbone.py:195-209): softmax(QK^T) and dS are invariant under a common shift of the keys /
 * values, and the centred products lose no digits to the common component. */
int stc_center_tokens(const void* x, void* out, float* mean_ws, int N, int L, int E, int dtype, void* stream);

/* ---------------------------------------------------------------- BatchNorm (K5/K6, C1/C2)
 * nn.SyncBatchNorm / BatchNorm2d train+eval (unet_backbone.py:64,121,124; unet_head.py:68,71,125). */
/* sums[0:C] = sum_p y, sums[C:2C] = sum_p y^2 (fp64).  ws >= stc_bn_ws_bytes(P,C). */
long long stc_bn_ws_bytes(long long P, int C);
int stc_bn_reduce(const void* y, double* sums, long long P, int C, void* ws, long long ws_bytes, int dtype, void* stream);
/* mean/invstd from (possibly all-reduced) sums over `count` samples; updates running stats
 * (momentum, unbiased var) when running_mean != NULL. */
int stc_bn_finalize(const double* sums, double count, float* mean, float* invstd, float* running_mean,
                    float* running_var, int64_t* num_batches_tracked, float momentum, float eps, int C, void* stream);
/* eval mode: mean = running_mean, invstd = rsqrt(running_var+eps) */
int stc_bn_eval_stats(const float* running_mean, const float* running_var, float* mean, float* invstd, float eps,
                      int C, void* stream);
/* a = act(gamma*(y-mean)*invstd + beta) */
int stc_bn_apply(const void* y, const float* mean, const float* invstd, const float* gamma, const float* beta,
                 void* a, long long P, int C, int act, int dtype, void* stream);
/* g = dout * act'(z); sums[0:C] = sum g, sums[C:2C] = sum g*xhat (fp64) */
int stc_bn_bwd_reduce(const void* y, const void* dout, const float* mean, const float* invstd, const float* gamma,
                      const float* beta, double* sums, long long P, int C, int act, void* ws, long long ws_bytes,
                      int dtype, void* stream);
/* dgamma = sums[C:2C], dbeta = sums[0:C] (taken from the LOCAL sums, before any SyncBN all-reduce, as
 * torch's SyncBatchNorm backward does). */
int stc_bn_param_grads(const double* sums, float* dgamma, float* dbeta, int C, void* stream);
/* The same backward with an IMPLICIT upstream gradient  d' = up_scale[n][c] * dout + up_shift[n][c] * shift_scale  (n = image,
 * rows_per_image rows each; ReLU, train mode, C/8 a power of two <= 256 - stc_bn_bwd_aff_ok).  KernelSelectAttention
 * (unet_backbone.py:86-98) hands each branch  df_k = softmax-weight_k[n,c] * dout + dS[n,c]/HW ; these entry points consume
 * dout directly so the three df tensors are never written. */
/* Inference (eval-mode BN, no gradient): Conv2d + BatchNorm folded into one conv, w' = w * s per output channel, b' = (b - running_mean) * s + beta
 * with s = gamma / sqrt(running_var + eps); b may be NULL.  The activation then runs in the conv epilogue (no BN-apply pass). */
int stc_bn_fold_conv(const float* w, const float* b, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                     float eps, float* w_folded, float* b_folded, int Cout, long long per_out, void* stream);
int stc_bn_bwd_aff_ok(int C);
int stc_bn_bwd_reduce_aff(const void* y, const void* dout, const float* up_scale, const float* up_shift, float shift_scale,
                          long long rows_per_image, int N, const float* mean, const float* invstd, const float* gamma,
                          const float* beta, double* sums, int C, void* ws, long long ws_bytes, int dtype, void* stream);
int stc_bn_bwd_apply_aff(const void* y, const void* dout, const float* up_scale, const float* up_shift, float shift_scale,
                         long long rows_per_image, int N, const float* mean, const float* invstd, const float* gamma,
                         const float* beta, const double* sums, double count, void* dy, int C, int dtype, void* stream);

/* dy = gamma*invstd*(g - sum_g/count - xhat*sum_gx/count) (train); eval != 0: dy = gamma*invstd*g. */
int stc_bn_bwd_apply(const void* y, const void* dout, const float* mean, const float* invstd, const float* gamma,
                     const float* beta, const double* sums, double count, void* dy,
                     long long P, int C, int act, int eval, int dtype, void* stream);

/* ---------------------------------------------------------------- pooling / upsampling (K7/K8/K9) */
/* nn.MaxPool2d(2) (unet_backbone.py:107); H, W are the INPUT sizes (even). */
int stc_maxpool2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream);
/* dx gets dy at the first arg-max of each 2x2 window (PyTorch tie rule), 0 elsewhere. */
int stc_maxpool2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, int dtype, void* stream);
/* Up.forward's upsample+pad+cat (unet_head.py:51-55): out[...,0:Cs] = skip ;
 * out[...,Cs:Cs+Cu] = pad(bilinear_x2(low, align_corners)) ; low is (N,h,w,Cu), out is (N,H,W,Cs+Cu). */
int stc_upcat_fwd(const void* skip, const void* low, void* out, int N, int H, int W, int Cs, int h, int w, int Cu,
                  int align_corners, int dtype, void* stream);
int stc_upcat_bwd(const void* dout, void* dskip, void* dlow, int N, int H, int W, int Cs, int h, int w, int Cu,
                  int align_corners, int dtype, void* stream);

/* Up.forward fused (decode_heads/unet_head.py:50-60 with CoordAtt, :131-146): the concatenated tensor cat = [skip, pad(up(low))] is
 * never written.  stc_upcat_pool: y (N,H+W,Cs+Cu) = row means (rows 0..H-1) and column means (rows H..) of cat.
 * stc_upcat_apply_fwd: out = cat + a_h*a_w with a (N,H+W,Cs+Cu); a == NULL gives out = cat.
 * stc_upcat_apply_bwd: g = dout + dy_h/W + dy_w/H (dyhw (N,H+W,Cs+Cu) = gradient of the pooled descriptor, may be NULL);
 * dskip = g[..., :Cs], dlow = adjoint of the bilinear x2 applied to g[..., Cs:]; either output may be NULL.
 * stc_upcat_pool needs a caller-provided workspace of stc_upcat_pool_ws_bytes (fp32 partial reductions of low: the means of the up-sampled
 * half are linear in low, so low is reduced with the adjoint interpolation weights and the result interpolated).
 * stc_upcat_fused_ok returns 1 when these row-structured kernels support the shape ((Cs+Cu)/8 <= 256 lanes). */
int stc_upcat_fused_ok(int N, int H, int W, int Cs, int h, int w, int Cu);
long long stc_upcat_pool_ws_bytes(int N, int h, int w, int Cu);
int stc_upcat_pool(const void* skip, const void* low, void* y, int N, int H, int W, int Cs, int h, int w, int Cu, int align_corners, void* ws,
                   long long ws_bytes, int dtype, void* stream);
int stc_upcat_apply_fwd(const void* skip, const void* low, const void* a, void* out, int N, int H, int W, int Cs, int h, int w, int Cu,
                        int align_corners, int dtype, void* stream);
int stc_upcat_apply_bwd(const void* dout, const void* dyhw, void* dskip, void* dlow, int N, int H, int W, int Cs, int h, int w, int Cu,
                        int align_corners, int dtype, void* stream);

/* torch.cat([a, b], dim=1) on NHWC rows (UpConvBlock.forward, mmseg/models/utils/up_conv_block.py:99; FCNHead concat_input,
 * fcn_head.py:81) and its adjoint (a or b may be NULL to drop that half). */
int stc_concat_channels(const void* a, const void* b, void* out, long long P, int Ca, int Cb, int dtype, void* stream);
/* y[p][0:Cdst) = x[p][0:min(Csrc,Cdst)), zeros beyond Csrc (both multiples of 8): widens the 16 / 32-channel layers of UNet++'s decoder
 * (smp DecoderBlock, decoder_channels (256,128,64,32,16)) to the tensor-core kernels' 64-channel K chunks, and narrows the result back. */
int stc_resize_channels(const void* x, void* y, long long P, int Csrc, int Cdst, int dtype, void* stream);
int stc_split_channels(const void* cat, void* a, void* b, long long P, int Ca, int Cb, int dtype, void* stream);

/* UNet++ DecoderBlock input (segmentation_models_pytorch 0.2.0, used by decode_heads/unetpp_head.py:16): out = cat([in0', in1, .., in4])
 * along channels, in0' = nearest x2 upsample of in0 when up0 != 0 (in0 is then (N,H/2,W/2,c0)).  Unused inputs: NULL / 0 channels.
 * The adjoint writes d_k for every non-NULL pointer (2x2 block sums for the upsampled input). */
int stc_catn_fwd(const void* in0, const void* in1, const void* in2, const void* in3, const void* in4, int c0, int c1, int c2, int c3,
                 int c4, void* out, int N, int H, int W, int up0, int dtype, void* stream);
int stc_catn_bwd(const void* dout, void* d0, void* d1, void* d2, void* d3, void* d4, int c0, int c1, int c2, int c3, int c4, int N, int H,
                 int W, int up0, int dtype, void* stream);

/* DeconvModule (mmseg/models/backbones/unet.py:89-147: ConvTranspose2d(k=4, s=2, p=1) -> BN -> ReLU): the transposed conv runs as a 3x3
 * conv to 4*C sub-pixel channels (host builds that weight) followed by this pixel shuffle,
 *   hi[n, 2h+py, 2w+px, c] = lo[n, h, w, (py*2+px)*C + c];  inverse != 0 moves hi -> lo (the adjoint).  lo is (N,H,W,4C), hi (N,2H,2W,C). */
int stc_depth_to_space2(const void* src, void* dst, int N, int H, int W, int C, int inverse, int dtype, void* stream);

/* ---------------------------------------------------------------- CoordAtt (K10; unet_head.py:131-146,57) */
/* y[n, 0:H, c] = mean_w x ; y[n, H:H+W, c] = mean_h x  ; y is (N, H+W, C) */
int stc_rowcol_mean(const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream);
/* out = x + ah[n,h,c]*aw[n,w,c]; a is (N, H+W, C) with ah first. */
int stc_coordatt_apply(const void* x, const void* a, void* out, int N, int H, int W, int C, int dtype, void* stream);
/* da[n,h,c] = sum_w dout*aw ; da[n,H+w,c] = sum_h dout*ah */
int stc_coordatt_apply_bwd(const void* dout, const void* a, void* da, int N, int H, int W, int C, int dtype, void* stream);
/* dx = dout + dy[n,h,c]/W + dy[n,H+w,c]/H  (grad through x + the two mean-pools) */
int stc_coordatt_dx(const void* dout, const void* dy, void* dx, int N, int H, int W, int C, int dtype, void* stream);

/* ---------------------------------------------------------------- KernelSelectAttention fuse (K11; unet_backbone.py:80-99,46-48) */
/* S[n,c] = mean_hw(f0+f1+f2) (fp32) */
int stc_ksa_pool(const void* f0, const void* f1, const void* f2, float* S, int N, long long HW, int C, int dtype, void* stream);
/* w = softmax over the 3 branches of a[3][N*C] (fp32) */
int stc_softmax3_fwd(const float* a, float* w, long long NC, void* stream);
int stc_softmax3_bwd(const float* w, const float* dw, float* da, long long NC, void* stream);
/* out = x + sum_k w[k][n][c] * f_k */
int stc_ksa_combine(const void* x, const void* f0, const void* f1, const void* f2, const float* w, void* out, int N,
                    long long HW, int C, int dtype, void* stream);
/* dw[k][n][c] = sum_hw dout * f_k (fp32) */
int stc_ksa_dw(const void* dout, const void* f0, const void* f1, const void* f2, float* dw, int N, long long HW, int C,
               int dtype, void* stream);
/* df_k = w_k*dout + dS[n][c]/HW */
int stc_ksa_df(const void* dout, const float* w, const float* dS, void* df0, void* df1, void* df2, int N, long long HW,
               int C, int dtype, void* stream);

/* ---------------------------------------------------------------- small fp32 dense layers (KSA fc/fcs; tiny)
 * y[r][o] = act(sum_i x[r][i] W[o][i] + b[o]) ; fp32 row-major. */
int stc_linear_f32_fwd(const float* x, const float* W, const float* b, float* y, int rows, int in, int out, void* stream);
/* dx = dy W ; dW (+)= dy^T x ; db (+)= colsum dy  (any may be NULL) */
int stc_linear_f32_bwd(const float* x, const float* W, const float* dy, float* dx, float* dW, float* db, int rows,
                       int in, int out, void* stream);

/* ---------------------------------------------------------------- elementwise */
int stc_add(const void* a, const void* b, void* out, long long n, int dtype, void* stream);           /* out = a + b */
int stc_add_n(const void* a, const void* b, const void* c, const void* d, void* out, long long n, int dtype, void* stream); /* out = a+b(+c)(+d); c, d may be NULL */
int stc_axpy_f32(const float* x, float* y, float alpha, long long n, void* stream);                    /* y += alpha x */
int stc_cast(const void* src, void* dst, long long n, int src_dtype, int dst_dtype, void* stream);
int stc_act_bwd(const void* y_out, const void* dy, void* dx, long long n, int act, int dtype, void* stream); /* sigmoid: uses output */
/* dst[n, dst_off + r, :] = src[n, src_off + r, :] for r < count, on (N, rows, C) tensors (CoordAtt split / cat). */
int stc_copy_rows(const void* src, void* dst, int N, int src_rows, int dst_rows, int C, int src_off, int dst_off, int count,
                  int dtype, void* stream);
int stc_scale_channels(const void* x, const float* m, void* y, int N, long long HW, int C, int dtype, void* stream); /* Dropout2d mask: y = x*m[n][c] */

/* ---------------------------------------------------------------- classifier + loss (K12-K15)
 * BaseDecodeHead.cls_seg (decode_head.py:254-259): logits NCHW fp32 = conv1x1(x NHWC) + b. W is (Ccls,Cin) fp32. */
/* mask: optional Dropout2d factors, fp32 (N, Cin), already scaled by 1/(1-p); NULL = no dropout. */
int stc_cls_fwd(const void* x, const float* W, const float* b, const float* mask, float* logits, int N, long long HW, int Cin,
                int Ccls, int dtype, void* stream);
/* dx NHWC = dlogits^T W ; dW (+)= ..., db (+)= ... ; ws >= stc_cls_bwd_ws_bytes */
long long stc_cls_bwd_ws_bytes(int N, long long HW, int Cin, int Ccls);
int stc_cls_bwd(const float* dlogits, const void* x, const float* W, const float* mask, void* dx, float* dW, float* db, int N,
                long long HW, int Cin, int Ccls, void* ws, long long ws_bytes, int dtype, void* stream);
/* BaseDecodeHead.losses (decode_head.py:261-296) = CrossEntropyLoss(avg_non_ignore=False)
 * (cross_entropy_loss.py:45-61) + DiceLoss (dice_loss.py:13-47,92-123) + accuracy (accuracy.py:6-61).
 * logits NCHW fp32, label int64 (N,H,W).  stats (fp64, 3*N*C + 4): per (n,c) [sum p*t*m, sum p^2, sum t],
 * then [ce_sum, n_correct, n_valid, 0].  out3 = {loss_bce, loss_dice, acc_seg} fp32. */
long long stc_seg_loss_stats_len(int N, int C);
int stc_seg_loss_fwd(const float* logits, const int64_t* label, double* stats, float* out3, int N, long long HW, int C,
                     int ignore_index, float smooth, void* stream);
/* dlogits = g_ce * dCE/dlogits + g_dice * dDice/dlogits */
int stc_seg_loss_bwd(const float* logits, const int64_t* label, const double* stats, const float* g_ce, const float* g_dice,
                     float* dlogits, int N, long long HW, int C, int ignore_index, float smooth, void* stream);

/* Row softmax of nn.MultiheadAttention (unet_backbone.py:202,207: softmax(q k^T / sqrt(hd)) and its backward) inside the score products,
 * without the L x L score / dP tensors: a CTA owns a 128-row block over all its N tiles and visits them twice - sweep 0 keeps the row
 * statistics (max and sum of exponentials; or sum(P dP) / sum(P)) in the epilogue thread that owns the row, sweep 1 recomputes the tiles
 * and stores the final bf16 values.  P = softmax_rows(scale * bf16(A B^T));  dS = scale * P * (bf16(A B^T) - sum_j P dP / sum_j P) with
 * P laid out like the output.  Arithmetic of stc_gemm + stc_softmax_rows_fwd / _bwd up to the order of the fp32 row sums.
 * tcgen05 engine, bf16, M % 128 == 0, N a multiple of its tile (stc_gemm_softmax_ok; P may be NULL for the forward form). */
int stc_gemm_softmax_ok(const stc_gemm_desc* d, const void* C, const void* P, int dtype, int engine);
int stc_gemm_softmax(const void* A, const void* B, void* P, const stc_gemm_desc* d, float scale, int dtype, int engine, void* stream);
int stc_gemm_softmax_bwd(const void* A, const void* B, const void* P, void* dS, const stc_gemm_desc* d, float scale, int dtype, int engine,
                         void* stream);

/* Softmax backward inside the dP product of nn.MultiheadAttention's backward (unet_backbone.py:202,207): C = alpha * P .* (A * B - D[row])
 * with A = dO, B = V^T, P = the stored probabilities (bf16, laid out like C) and D[b1, b2, m] = rowsum(dO * O) (fp32, stc_rowdot_heads):
 * dS = scale * P * (dP - D) leaves the GEMM's epilogue directly - the L x L dP tensor and the separate softmax-backward pass
 * (read P, read dP, write dS) disappear.  tcgen05 engine only; stc_gemm_dsoftmax_ok tells. */
int stc_gemm_dsoftmax_ok(const stc_gemm_desc* d, int dtype, int engine);
int stc_gemm_dsoftmax(const void* A, const void* B, const void* P, const float* D, void* C, const stc_gemm_desc* d, int dtype, int engine,
                      void* stream);
/* out[n, h, i] = sum_d a[n, i, h*hd + d] * b[n, i, h*hd + d] for (N, L, heads*hd) token tensors (fp32 out, (N, heads, L)). */
int stc_rowdot_heads(const void* a, const void* b, float* out, int N, int L, int heads, int hd, int dtype, void* stream);

/* ---------------------------------------------------------------- inference post-processing (K17; encoder_decoder.py:157-203,253,272) */
/* preds[:, :, y1:y1+hc, x1:x1+wc] += crop ; count[:, y1.., x1..] += 1  (NCHW fp32; count (N,H,W)) */
int stc_slide_accum(const float* crop, float* preds, float* count, int N, int C, int H, int W, int hc, int wc, int y1,
                    int x1, void* stream);
/* pred[n,h,w] = argmax_c softmax(preds/count) (first max wins); count may be NULL (whole mode) */
int stc_argmax(const float* preds, const float* count, int64_t* pred, int N, int C, long long HW, void* stream);

/* ---------------------------------------------------------------- integer confusion matrix (K18; metrics.py:75-87)
 * cm[label*C + pred] += 1 for label != ignore, label,pred in [0,C).  pred int64, label uint8 or int64
 * (label_is_u8).  areas (int64[4*C]: intersect, union, pred, label) follow metrics.py exactly
 * (pred and label histogrammed independently).  Both are accumulated into (caller zeroes). */
int stc_confusion_hist(const int64_t* pred, const void* label, int label_is_u8, long long n, int C, int ignore_index,
                       int64_t* cm, int64_t* areas, void* stream);

/* ---------------------------------------------------------------- optimizer (f-2; my_config/STC-UNet.py:87) */
/* torch.optim.Adam semantics on a flat fp32 buffer; step is 1-based. */
int stc_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, void* stream);
/* Same update with the learning rate and the step counter in DEVICE memory (dyn = {lr, beta1^t, beta2^t} fp32, step int32): a captured
 * CUDA graph of the whole training step can be replayed while bias corrections and LR schedules keep advancing. */
int stc_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float* dyn, int* step, float beta1, float beta2, float eps,
                      float weight_decay, void* stream);

/* ---------------------------------------------------------------- data-parallel exchanges over NVLink peer memory (C1-C3)
 * Replace the NCCL calls of the reference's data-parallel step (nn.SyncBatchNorm's statistics all-reduces and
 * MMDistributedDataParallel's gradient reduction, mmseg/apis/train.py:104-113) by kernels that read / write the peers' buffers
 * directly, so the whole step can be one CUDA graph.  *_ptrs: `world` device pointers (as integers, HOST array) to every rank's copy of
 * a symmetric buffer, own rank included.  Control block: stc_peer_ctrl_bytes(max_n, ctas) bytes, zeroed once before the first use.
 * seq: DEVICE counters (1 for the small exchange, `ctas` for the arena exchange), zero-initialised, private to this rank.
 * All ranks must issue the same sequence of calls.  Every wait for a peer is bounded in wall-clock time (stc_peer_configure; default
 * 10 minutes, like NCCL's watchdog): when it runs out the kernel writes 1 to the configured error flag and returns (it does not trap, so
 * the CUDA context survives and the host can raise). */
#define STC_PEER_MAX 16
long long stc_peer_ctrl_bytes(int max_n, int ctas);
/* timeout_ms: bound of one cross-rank wait.  err_flag: int the kernels can write and the host can read without synchronising (pinned,
 * device-mapped host memory), or NULL.  Process-wide (one process drives one GPU, SURVEY 8b threading model). */
int stc_peer_configure(long long timeout_ms, int* err_flag);
/* out[i] = sum over ranks of in[i] (fp64, n <= max_n), summed in rank order: identical on every rank. */
int stc_peer_allreduce_small_f64(const unsigned long long* ctrl_ptrs, int rank, int world, int max_n, const double* in, double* out, int n,
                                 unsigned long long* seq, void* stream);
/* In place on the symmetric fp32 arena: arena[elem_off : elem_off + n) = scale * sum over ranks (two-shot reduce-scatter + all-gather,
 * `ctas` CTAs, each pairing with the same CTA of the peers). */
int stc_peer_allreduce_arena_f32(const unsigned long long* arena_ptrs, const unsigned long long* ctrl_ptrs, int rank, int world, int max_n,
                                 long long elem_off, long long n, float scale, unsigned long long* seq, int ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STC_B200_H_ */
