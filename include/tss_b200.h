/* libtss_b200 -- C ABI of the B200 (sm_100a) kernels behind the Fast-SCNN / ContextNet
 * hot path of torch_semantic_segmentation.
 *
 * The reference (bernardomig/torch_semantic_segmentation) is pure Python: it has no FFI
 * of its own.  Each entry point below replaces the stock torch op that a reference
 * call site dispatches to; the call site is cited on every declaration (paths relative
 * to the reference root).  INTEGRATION.md shows the ctypes / torch.library binding a
 * maintainer of the reference would add.
 *
 * Conventions:
 *  - plain pointers and sizes only; every buffer is device memory owned by the caller
 *    (PyTorch's caching allocator); the library never allocates, frees or keeps them.
 *  - activations are dense NHWC ("channels_last"): x[n][h][w][c]; `dtype` selects the
 *    activation storage type (TSS_F32 or TSS_BF16); parameters, statistics and weight
 *    gradients are always fp32; accumulation is always fp32.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *    synchronises, is re-entrant and thread-safe (autograd calls from its own thread).
 *  - return value 0 = success; otherwise an error code, and tss_last_error() returns a
 *    thread-local message.  Unsupported shapes are hard errors: there is no fallback.
 *  - vector paths need C % 8 == 0 and 16-byte aligned base pointers; the 3-channel
 *    input and the 19-class logits have their own entry points.
 */
#ifndef TSS_B200_H_
#define TSS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSS_VERSION 100

#define TSS_OK 0
#define TSS_ERR_ARG 1
#define TSS_ERR_CUDA 2

#define TSS_F32 0
#define TSS_BF16 1

/* epilogue flags shared by the conv entry points */
#define TSS_EPI_RELU 1      /* max(.,0) after affine (+residual) */

int tss_version(void);
const char* tss_last_error(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
uint64_t tss_launch_count(void);
/* Programmatic dependent launch between consecutive kernels of a stream / graph (on by default;
 * env TSS_PDL=0 or tss_set_pdl(0) selects plain serialized launches).  Process-wide. */
int tss_set_pdl(int enabled);
/* Debugging aid: cudaStreamIsCapturing of a stream handle: 0 = not capturing, 1 = capture active, 2 = capture
 * invalidated by an earlier call, < 0 = -(cudaError_t).  Launches nothing. */
int64_t tss_capture_status(int64_t stream_handle);

/* ---- depthwise 3x3 convolution, padding == dilation ("same" for stride 1) --------------
 * replaces nn.Conv2d(C, C, 3, stride, padding=dilation, dilation, groups=C, bias=False)
 * at fastscnn.py:179-180,191-192 and contextnet.py:157-160.
 * x[N][Hi][Wi][C], w[C][3][3] fp32 (= the (C,1,3,3) parameter), y[N][Ho][Wo][C],
 * Ho = (Hi-1)/stride+1.
 * Epilogue: if scale/shift != NULL: y = y*scale[c]+shift[c] (folded eval-mode BatchNorm,
 * fastscnn.py:181), then ReLU if flags&TSS_EPI_RELU.  If stats != NULL (training):
 * stats[c] += sum(y), stats[C+c] += sum(y*y) over all pixels of the raw conv output
 * (BatchNorm batch statistics: fp32 partial sums per CTA, accumulated across CTAs with fp64
 * atomics so that E[y^2] - mean^2 keeps fp32-reference accuracy). */
int tss_dwconv3x3_fwd(const void* x, const float* w, void* y, int N, int Hi, int Wi, int C,
                      int stride, int dilation, const float* scale, const float* shift,
                      int flags, double* stats, int dtype, void* stream);
/* grad wrt input: dx[N][Hi][Wi][C] from dy[N][Ho][Wo][C] (autograd of the above). */
int tss_dwconv3x3_dgrad(const void* dy, const float* w, void* dx, int N, int Hi, int Wi, int C,
                        int stride, int dilation, int dtype, void* stream);
/* Training, stride 1 / dilation 1: the dgrad above with the BatchNorm-backward reduction of the producing
 * layer fused into its epilogue (see tss_pwconv_dgrad_bnred): g[N][H][W][C] = dz * mask(yp), sums += ... */
int tss_dwconv3x3_dgrad_bnred(const void* dy, const float* w, void* g, int N, int H, int W, int C,
                              const void* yp, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, int flags, float* sums, int dtype, void* stream);
/* the same for stride 2 (g and yp have the conv's INPUT geometry (Hi, Wi), dy its output geometry) */
int tss_dwconv3x3_dgrad_s2_bnred(const void* dy, const float* w, void* g, int N, int Hi, int Wi, int C,
                                 const void* yp, const float* mean, const float* rstd, const float* gamma,
                                 const float* beta, int flags, float* sums, int dtype, void* stream);
/* Training: depthwise forward and weight gradient whose input is the RAW conv output x of the producing layer;
 * its BatchNorm (+ReLU with in_flags&TSS_EPI_RELU) z = act(x*in_scale + in_shift) is applied while the input tile is
 * read, so the activated tensor is never materialised (zero padding applies to z).  dilation 1, stride 1 or 2;
 * fwd accumulates the output statistics like tss_dwconv3x3_fwd; wgrad: dw += sum z * dy. */
int tss_dwconv3x3_fwd_bnin(const void* x, const float* in_scale, const float* in_shift, int in_flags, const float* w,
                           void* y, int N, int Hi, int Wi, int C, int stride, double* stats, int dtype, void* stream);
int tss_dwconv3x3_wgrad_bnin(const void* x, const float* in_scale, const float* in_shift, int in_flags, const void* dy,
                             float* dw, int N, int Hi, int Wi, int C, int stride, int dtype, void* stream);
/* Training, stride 1 / dilation 1: BatchNorm-backward APPLY of this depthwise layer + its dgrad + the producer's
 * BatchNorm-backward reduction in ONE kernel (the depthwise counterpart of tss_pwconv_bwd_fused).  dz, y
 * [N][H][W][C]: gradient after this layer's BN/ReLU and its raw conv output (two TMA halo tiles); sums[2C]: the
 * finished reduction of this layer (flags&TSS_EPI_RELU: mask recomputed from y; 0: dz already masked).  dy
 * (may be NULL): the BatchNorm-backward output, stored for the weight gradient.  g: masked gradient for the producer
 * (yp, pmean, ..., psums as in tss_dwconv3x3_dgrad_bnred; yp == NULL: no producer, g is the plain input gradient).
 * dgamma += sums[C+c], dbeta += sums[c]. */
int tss_dwconv3x3_bwd_fused(const void* dz, const void* y, const float* w, const float* mean, const float* rstd,
                            const float* gamma, const float* beta, const float* sums, int flags, int64_t count,
                            void* dy, float* dgamma, float* dbeta, void* g, int N, int H, int W, int C,
                            const void* yp, const float* pmean, const float* prstd, const float* pgamma,
                            const float* pbeta, int pflags, float* psums, int dtype, void* stream);
/* grad wrt weight: dw[C][3][3] (fp32) += sum_{n,ho,wo} x * dy  (warp-shuffle + block
 * reduction, fp32 atomics).  dw must be zeroed (or hold a gradient to accumulate into). */
int tss_dwconv3x3_wgrad(const void* x, const void* dy, float* dw, int N, int Hi, int Wi, int C,
                        int stride, int dilation, int dtype, void* stream);

/* ---- pointwise (1x1) convolution = GEMM  y[M][Nc] = x[M][K] . w[Nc][K]^T -----------------
 * replaces nn.Conv2d(K, Nc, 1[, bias]) at fastscnn.py:167-168,194,97,109-110,115 and
 * contextnet.py:172-175,86.  x has row pitch ldx elements (>= K), y row pitch ldy.
 * w is the fp32 (Nc,K,1,1) parameter.  Epilogue as for the depthwise conv, plus an
 * optional residual res[M][Nc] (pitch ldr) added before the ReLU (fastscnn.py:158-161).
 * `shift` alone (scale == NULL) is a bias (fastscnn.py:97).
 * impl: 0 = SIMT fp32-accumulate reference-precision kernel (any dtype);
 *       1 = tcgen05/TMEM/TMA tensor-core kernel (bf16 only, needs wp = bf16 packed weights
 *           from tss_pack_weights_bf16). */
int tss_pwconv_fwd(const void* x, const float* w, const void* wp, void* y, int64_t M, int K, int Nc,
                   int64_t ldx, int64_t ldy, const float* scale, const float* shift,
                   const void* res, int64_t ldr, int flags, double* stats, int impl,
                   int dtype, void* stream);
/* dx[M][K] = dy[M][Nc] . w[Nc][K]   (wpT = bf16 packed transpose for impl 1) */
int tss_pwconv_dgrad(const void* dy, const float* w, const void* wpT, void* dx, int64_t M, int K,
                     int Nc, int64_t lddy, int64_t lddx, int impl, int dtype, void* stream);
/* dw[Nc][K] (fp32) += dy^T . x ; db[Nc] (fp32, may be NULL) += column sums of dy */
int tss_pwconv_wgrad(const void* x, const void* dy, float* dw, float* db, int64_t M, int K, int Nc,
                     int64_t ldx, int64_t lddy, int impl, int dtype, void* stream);
/* Training: the dgrad above (impl 1) with the BatchNorm-backward reduction of the PRODUCING layer fused
 * into its epilogue.  The input of this 1x1 conv was z = act(BN(yp)) with yp[M][K] (pitch ldyp) the raw
 * conv output of the previous layer; instead of dz the kernel stores g = dz * (z > 0) (mask recomputed
 * from yp; all-ones without TSS_EPI_RELU) and accumulates sums[k] += sum_m g, sums[K+k] += sum_m g*xhat,
 * xhat = (yp-mean)*rstd -- what tss_bn_bwd_reduce would compute in a pass of its own.  The producer's
 * BatchNorm backward then only runs tss_bn_bwd_apply on (g, yp) with flags = 0.  sums zeroed by the caller. */
int tss_pwconv_dgrad_bnred(const void* dy, const void* wpT, void* g, int64_t M, int K, int Nc, int64_t lddy,
                           int64_t ldg, const void* yp, int64_t ldyp, const float* mean, const float* rstd,
                           const float* gamma, const float* beta, int flags, float* sums, void* stream);
/* Training, bf16: 1x1 forward whose input x[M][K] (pitch ldx) is the RAW conv output of the producing layer; its
 * BatchNorm (+ReLU with in_flags&TSS_EPI_RELU) z = act(x*in_scale + in_shift) is applied by the threads that build the
 * GEMM's A operand.  z (pitch ldz, may be NULL) is stored for the weight gradient; y[M][Nc] = z . wp^T (raw output),
 * stats[2*Nc] accumulates its BatchNorm statistics like tss_pwconv_fwd. */
int tss_pwconv_fwd_bnin(const void* x, int64_t ldx, const float* in_scale, const float* in_shift, int in_flags, void* z,
                        int64_t ldz, const void* wp, void* y, int64_t ldy, int64_t M, int K, int Nc, double* stats,
                        void* stream);
/* Training, bf16: BatchNorm-backward APPLY of this 1x1 layer + its dgrad (+ optionally the producer's
 * BatchNorm-backward reduction, as tss_pwconv_dgrad_bnred) in ONE tcgen05 kernel.  dz[M][Nc] is the gradient
 * after the layer's BN/ReLU, y[M][Nc] its raw conv output, sums[2*Nc] the finished reduction (sum g, sum g*xhat
 * with g = dz*mask; flags&TSS_EPI_RELU: the mask is recomputed from y, 0: dz already is g).  The kernel forms
 * dy = gamma*rstd*(g - sums[c]/count - xhat*sums[Nc+c]/count) on the fly as the GEMM's A operand, stores it to
 * dy (pitch lddy; for the weight gradient; may be NULL), accumulates dgamma += sums[Nc+c], dbeta += sums[c]
 * (either may be NULL) and writes dx[M][K] = dy . w (wpT = bf16 [K][Nc]).  With yp != NULL, dx is masked by the
 * producer's ReLU (pflags) and psums[2*K] accumulates the producer's two sums.  No residual, per-rank statistics. */
int tss_pwconv_bwd_fused(const void* dz, const void* y, int64_t lddz, int64_t ldy, const float* mean,
                         const float* rstd, const float* gamma, const float* beta, const float* sums, int flags,
                         int64_t count, void* dy, int64_t lddy, float* dgamma, float* dbeta, const void* wpT,
                         void* dx, int64_t M, int K, int Nc, int64_t lddx, const void* yp, int64_t ldyp,
                         const float* pmean, const float* prstd, const float* pgamma, const float* pbeta,
                         int pflags, float* psums, void* stream);
/* Class-score convolution nn.Conv2d(K, Nc, 1) WITH bias and Nc not a multiple of 16 (fastscnn.py:97, contextnet.py:86)
 * on the tensor-core kernels: zero-padded operands wp[Np][K] (Np % 16 == 0), wpT[K][Npt] (Npt % 8 == 0 = the channel
 * pitch of the incoming gradient) and bias_pad[Np] (fp32).  Then tss_pwconv_fwd(impl 1, Nc = Np, shift = bias_pad),
 * tss_pwconv_dgrad(impl 1, Nc = Npt) and tss_pwconv_wgrad(impl 1, Nc, db) (dw rows < Nc only, db += column sums). */
int tss_class_scores_pack(const float* w, const float* bias, void* wp, void* wpT, float* bias_pad, int Nc,
                          int K, int Np, int Npt, void* stream);
/* bf16 copies of a (Nc,K) fp32 weight: wp[Nc][K] and wpT[K][Nc] (either may be NULL) */
int tss_pack_weights_bf16(const float* w, void* wp, void* wpT, int Nc, int K, void* stream);
/* the same for n_entries weights living in one fp32 parameter arena, in ONE launch (after the optimizer
 * step): table (device int64 [n_entries][5]) = {element offset of the (Nc,K) weight in the arena, Nc, K,
 * device address of wp, device address of wpT}; max_elems = max Nc*K over the table. */
int tss_pack_weights_multi(const float* arena, const int64_t* table, int n_entries, int64_t max_elems,
                           void* stream);

/* ---- fused depthwise 3x3 -> pointwise 1x1, inference --------------------------------------------
 * The tail of the inverted-residual bottleneck (conv2 + conv3, fastscnn.py:149-161,
 * contextnet.py:138-147) and the DS-conv block (fastscnn.py:188-199) with eval-mode BatchNorm, as ONE
 * kernel: t = act1(dw3x3(x)*scale1+shift1) is produced tile by tile straight into the shared-memory
 * operand of the tcgen05 GEMM y = act2(t . wp^T * scale2 + shift2 [+ res]); the depthwise output
 * (6x the block's width in a bottleneck) never goes to HBM.  bf16 only; stride 1, dilation 1,
 * padding 1; C % 64 == 0; Nc % 16 == 0, Nc <= 256; x[N][H][W][C] dense, wp = bf16 (Nc, C) from
 * tss_pack_weights_bf16, w_dw the fp32 (C,1,3,3) parameter. */
int tss_dwpw_fwd(const void* x, const float* w_dw, const float* scale1, const float* shift1, int flags1,
                 const void* wp, void* y, int N, int H, int W, int C, int Nc, int64_t ldy,
                 const float* scale2, const float* shift2, const void* res, int64_t ldr, int flags2,
                 void* stream);

/* ---- dense 3x3 convolution ---------------------------------------------------------------
 * Stem: replaces nn.Conv2d(3, 32, 3, stride=2, padding=1, bias=False) at fastscnn.py:30 /
 * contextnet.py:38,48.  x is the user's NCHW fp32 image batch (N,3,H,W) read in place
 * (no layout/precision pre-pass); y[N][H/2][W/2][Cout] NHWC in `dtype`; w (Cout,3,3,3).
 * Epilogue as above. */
int tss_stem3x3s2_fwd(const float* x, const float* w, void* y, int N, int H, int W, int Cout,
                      const float* scale, const float* shift, int flags, double* stats,
                      int dtype, void* stream);
/* the same on tcgen05 (bf16 output only): the 27-tap patches of 128 consecutive output pixels are built by the
 * threads as the swizzled A operand of an implicit GEMM (image and weights rounded to bf16, fp32 accumulate). */
int tss_stem3x3s2_fwd_tc(const float* x, const float* w, void* y, int N, int H, int W, int Cout,
                         const float* scale, const float* shift, int flags, double* stats, void* stream);
/* weight gradient of the stem on tcgen05 (bf16 dy): dw (Cout,3,3,3) fp32 += sum over pixels; both GEMM operands
 * (dy transposed, image patches) are built K-major in shared memory by the threads, the pixel index is the
 * reduction dimension. */
int tss_stem3x3s2_wgrad_tc(const float* x, const void* dy, float* dw, int N, int H, int W, int Cout,
                           void* stream);
/* the same through a patch matrix: patches (bf16 [N*Ho*Wo][32], workspace) receives the 27 taps of every output pixel
 * (coalesced), then dw32 (fp32 [32][32], workspace) = dy^T . patches on the TMA-fed tensor-core weight-gradient GEMM of the
 * pointwise convs, and dw (Cout,3,3,3) += dw32[:, :27]. */
int tss_stem3x3s2_wgrad_patches(const float* x, const void* dy, void* patches, float* dw32, float* dw, int N,
                                int H, int W, int Cout, void* stream);
/* its two halves: the patch matrix depends on the image only (the training forward launches it on a side stream, off the
 * backward pass's tail), the GEMM + unpack on the gradient */
int tss_stem3x3s2_patches(const float* x, void* patches, int N, int H, int W, void* stream);
int tss_stem3x3s2_wgrad_from_patches(const void* patches, const void* dy, float* dw32, float* dw, int64_t M, int Cout,
                                     void* stream);
/* the same with the stem's BatchNorm-backward APPLY folded into the operand producer: dz is the gradient after the
 * stem's BN/ReLU, y its raw conv output, sums[2*Cout] the finished reduction (flags&TSS_EPI_RELU: mask recomputed from
 * y).  dy is never materialised (the stem has no input gradient, the weight gradient is its only reader);
 * dgamma += sums[Cout+c], dbeta += sums[c]. */
int tss_stem3x3s2_wgrad_tc_bn(const float* x, const void* dz, const void* y, const float* mean, const float* rstd,
                              const float* gamma, const float* beta, const float* sums, int flags, int64_t count,
                              float* dw, float* dgamma, float* dbeta, int N, int H, int W, int Cout, void* stream);
int tss_stem3x3s2_wgrad(const float* x, const void* dy, float* dw, int N, int H, int W, int Cout,
                        int dtype, void* stream);
/* Dense 3x3, stride 1, padding 1, C -> Cout between equal-sized NHWC maps: replaces
 * nn.Conv2d(128, 128, 3, padding=1, bias=False) at contextnet.py:55.  It runs as the pointwise
 * GEMM above over a tap-major patch matrix:
 *   col[N*H*W][9*C], col[m][tap*C + c] = x[n][h+ky-1][w+kx-1][c] (0 outside), tap = ky*3+kx
 *   wk[Cout][9*C],   wk[co][tap*C + c] = w[co][c][ky][kx]
 * im2col3x3 builds col; col2im3x3 is its transpose (dx from the GEMM dgrad's dcol);
 * permute_weights3x3 converts the (Cout,C,3,3) parameter to wk (backward = 0) or accumulates a
 * tap-major gradient dwk into the (Cout,C,3,3) gradient (backward = 1: dst += permuted src). */
int tss_im2col3x3(const void* x, void* col, int N, int H, int W, int C, int dtype, void* stream);
int tss_col2im3x3(const void* dcol, void* dx, int N, int H, int W, int C, int dtype, void* stream);
int tss_permute_weights3x3(const float* src, float* dst, int Cout, int Cin, int backward, void* stream);

/* ---- BatchNorm (training) -----------------------------------------------------------------
 * replaces nn.BatchNorm2d (+ nn.ReLU / F.relu / residual add) at fastscnn.py:169-172,
 * 181-184,193-198,158-161,89.
 * finalize: from stats[2C] (sum, sum of squares over `count` values per channel) compute
 * mean[c], rstd[c] = 1/sqrt(var_biased+eps), scale = gamma*rstd, shift = beta-mean*scale and
 * update running_mean/var (momentum, unbiased var) and num_batches_tracked (int64, may be NULL).
 * clear_n > 0: "consume and clear" -- stats[0..clear_n) (doubles) is zeroed after it has been read,
 * so a per-layer scratch [stats: 2C doubles | backward sums: 2C floats = C doubles] is zero again
 * for the next conv epilogue / bn_bwd_reduce without a memset launch per layer. */
int tss_bn_finalize(double* stats, int64_t count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, int64_t* num_batches_tracked,
                    float momentum, float eps, float* scale, float* shift, float* mean,
                    float* rstd, int C, int64_t clear_n, void* stream);
/* finalize + apply in ONE launch (training): z = act(BN(y) [+ res]) with scale / shift derived from stats inside
 * the kernel; mean / rstd written for the backward pass, running statistics updated, stats[0..clear_n) zeroed by
 * the last CTA to have read it.  ticket: device int, zero on entry and zero again on exit. */
int tss_bn_finalize_apply(double* stats, int64_t count, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, int64_t* num_batches_tracked,
                          float momentum, float eps, float* mean, float* rstd, int* ticket, int64_t clear_n,
                          const void* y, const void* res, void* z, int64_t M, int C, int64_t ldy, int64_t ldr,
                          int64_t ldz, int flags, int dtype, void* stream);
/* eval mode: scale = gamma/sqrt(running_var+eps), shift = beta - running_mean*scale */
int tss_bn_fold(const float* gamma, const float* beta, const float* running_mean,
                const float* running_var, float eps, float* scale, float* shift, int C,
                void* stream);
/* z = act( y*scale+shift [+ y2*scale2+shift2] [+ res] ), rows of C channels, pitches in elements */
int tss_bn_apply(const void* y, const float* scale, const float* shift, const void* y2,
                 const float* scale2, const float* shift2, const void* res, void* z, int64_t M,
                 int C, int64_t ldy, int64_t ldy2, int64_t ldr, int64_t ldz, int flags, int dtype,
                 void* stream);
/* backward, pass 1: g = dz * (z > 0 if relu); sums[c] += sum g ; sums[C+c] += sum g*xhat,
 * xhat = (y-mean)*rstd.  z == NULL with TSS_EPI_RELU: the mask is recomputed from y as
 * fma(y, gamma*rstd, beta - mean*gamma*rstd) > 0 -- the forward's own arithmetic, so the mask is
 * identical and z need not be read (valid when no residual was added before the ReLU). */
int tss_bn_bwd_reduce(const void* dz, const void* z, const void* y, const float* mean,
                      const float* rstd, const float* gamma, const float* beta, float* sums, int64_t M,
                      int C, int64_t lddz, int64_t ldz, int64_t ldy, int flags, int dtype, void* stream);
/* backward, pass 2: dy = gamma*rstd*(g - sums[c]/count - xhat*sums[C+c]/count); optional dres = g.
 * dgamma[c] += sums[C+c], dbeta[c] += sums[c] (done by block 0; may be NULL).  count = values per
 * channel behind the statistics: 0 = the M rows of this call; M*world_size when the statistics and
 * `sums` were all-reduced over a process group (SyncBN). */
int tss_bn_bwd_apply(const void* dz, const void* z, const void* y, const float* mean,
                     const float* rstd, const float* gamma, const float* beta, const float* sums, void* dy, void* dres,
                     float* dgamma, float* dbeta, int64_t M, int64_t count, int C, int64_t lddz, int64_t ldz,
                     int64_t ldy, int64_t lddy, int64_t lddres, int flags, int dtype, void* stream);
/* backward, both passes in ONE launch (replaces the reduce + apply pair behind nn.BatchNorm2d's autograd,
 * fastscnn.py:169, for tensors small enough to stay in L2 between the passes): pass 1 accumulates sums[2*C] (zeroed by
 * the caller) exactly like tss_bn_bwd_reduce (z == NULL under TSS_EPI_RELU: mask recomputed from y), a grid-wide
 * barrier, pass 2 writes dy (and dres = g when given) like tss_bn_bwd_apply with count = M, dgamma / dbeta accumulated
 * by block 0.  sync: DEVICE int32 [4], zeroed once by the caller and owned by ONE stream: {arrivals, generation, sticky
 * time-out flag, unused}; launches that share it must be stream-ordered.  All CTAs are resident at once (the grid is
 * capped at two per SM); a CTA that waits for more than about a second sets sync[2] and nobody waits any more. */
int tss_bn_bwd_onepass(const void* dz, const void* z, const void* y, const float* mean, const float* rstd,
                       const float* gamma, const float* beta, float* sums, void* dy, void* dres, float* dgamma,
                       float* dbeta, int64_t M, int C, int64_t lddz, int64_t ldz, int64_t ldy, int64_t lddy,
                       int64_t lddres, int flags, int* sync, int dtype, void* stream);
/* g = dz * (z > 0): ReLU backward alone (fusion add, eval-free paths) */
int tss_relu_bwd(const void* dz, const void* z, void* g, int64_t M, int C, int64_t lddz, int64_t ldz,
                 int64_t ldg, int dtype, void* stream);
/* out = a + b (residual gradient accumulation), rows of C channels */
int tss_add(const void* a, const void* b, void* out, int64_t M, int C, int64_t lda, int64_t ldb,
            int64_t ldo, int dtype, void* stream);
/* strided row copy: dst[m][0..C) = src[m][0..C) */
int tss_copy_rows(const void* src, void* dst, int64_t M, int C, int64_t lds, int64_t ldd, int dtype,
                  void* stream);

/* ---- pyramid pooling ------------------------------------------------------------------------
 * replaces the four nn.AdaptiveAvgPool2d(bin) of fastscnn.py:108 in ONE pass: windows
 * [floor(i*H/b), ceil((i+1)*H/b)).  out is the concatenation over bins of [N][b*b][C]
 * blocks (bin i starts at element N*C*sum_{j<i} b_j^2), so each bin is a plain row matrix
 * for the 1x1 convolution that follows (fastscnn.py:109-110). */
int tss_adaptive_pool_fwd(const void* x, void* out, int N, int H, int W, int C, const int* bins,
                          int nbins, int dtype, void* stream);
int tss_adaptive_pool_bwd(const void* dout, void* dx, int N, int H, int W, int C, const int* bins,
                          int nbins, int accumulate, int dtype, void* stream);

/* ---- bilinear resize, align_corners=True -------------------------------------------------------
 * replaces F.interpolate(..., mode='bilinear', align_corners=True) / nn.UpsamplingBilinear2d at
 * fastscnn.py:74,119-120 and contextnet.py:119-121.  NHWC, C % 8 == 0, channel pitches ldx/ldy
 * so that the result can be written straight into the concat buffer of fastscnn.py:122. */
int tss_bilinear_fwd(const void* x, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C,
                     int64_t ldx, int64_t ldy, int dtype, void* stream);
/* backward (transpose).  workspace: fp32 [N*Hi*Wo*C] scratch for the separable two-pass form
 * (rows, then columns); NULL selects a single gather pass (small maps only). */
int tss_bilinear_bwd(const void* dy, void* dx, float* workspace, int N, int Hi, int Wi, int Ho, int Wo,
                     int C, int64_t lddy, int64_t lddx, int dtype, void* stream);
/* Final x8 up-sampling of the class scores (fastscnn.py:63-64, contextnet.py:74-76):
 * in NHWC x[N][Hi][Wi][ldx] (C = 19 classes in a pitch of ldx >= C, any C <= 64) -> out NCHW
 * y[N][C][Ho][Wo] (the layout the reference returns), Wo % 8 == 0. */
int tss_upsample_logits_fwd(const void* x, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C,
                            int64_t ldx, int dtype, void* stream);
/* backward: dx32[N][Hi][Wi][lddx] (fp32, zeroed by the caller) += U^T dy (dy NCHW) */
int tss_upsample_logits_bwd(const void* dy, float* dx32, int N, int Hi, int Wi, int Ho, int Wo,
                            int C, int64_t lddx, int dtype, void* stream);
/* ContextNet input shrink (contextnet.py:65-67): NCHW fp32 (N,3,H,W) -> NCHW fp32 (N,3,Ho,Wo) */
int tss_bilinear_nchw_f32(const float* x, float* y, int NC, int Hi, int Wi, int Ho, int Wo,
                          void* stream);
/* dst (dtype) = src (fp32), n elements (n % 8 == 0) */
int tss_cast_from_f32(const float* src, void* dst, int64_t n, int dtype, void* stream);

/* ---- softmax cross-entropy with ignore_index ------------------------------------------------------
 * replaces nn.CrossEntropyLoss(ignore_index=255) (scripts/train_fastscnn.py:132) and the
 * F.cross_entropy(reduction='none') inside losses/ohem_loss.py:11-12.
 * logits NCHW [N][C][HW] (C <= 32, HW % 4 == 0), target int64 [N][HW].
 * count_valid: nvalid[0] = #(target != ignore_index and 0 <= target < num_classes) -- the pixels that carry a loss
 * in fwd (torch raises on a label outside [0, C); here it is treated as ignored, consistently in both kernels).
 * fwd: ONE pass computes per-pixel loss, the loss sum (loss_sum[0] += , fp64) and, if
 * dlogits != NULL, the gradient of the MEAN loss: (softmax - onehot)/nvalid[0] (0 at ignored
 * pixels).  pixel_loss (fp32 [N][HW], may be NULL) receives the reduction='none' values.
 * ohem != NULL (device float[4] = {cut, w_above, tie, w_tie} from tss_ohem_select): the gradient of
 * pixel i is weighted by (loss_i > cut ? w_above : loss_i == tie ? w_tie : 0) instead of 1/nvalid. */
int tss_ce_count_valid(const int64_t* target, int64_t n, int64_t ignore_index, int64_t num_classes,
                       int64_t* nvalid, void* stream);
int tss_ce_fwd(const void* logits, const int64_t* target, int N, int C, int64_t HW,
               int64_t ignore_index, const int64_t* nvalid, double* loss_sum, float* pixel_loss,
               void* dlogits, const float* ohem, int dtype, void* stream);
/* loss[0] = loss_sum[0] / nvalid[0]  (NaN if no valid pixel, like the reference) */
int tss_ce_finalize(const double* loss_sum, const int64_t* nvalid, float* loss, void* stream);
/* x *= s[0] (device scalar), n elements: non-unit upstream gradient of the loss */
int tss_scale_inplace(void* x, const float* s, int64_t n, int dtype, void* stream);

/* Fused head: the final x8 up-sampling (fastscnn.py:63-64) + the cross-entropy above + the gradient
 * w.r.t. the LOW-resolution class scores, without reading the full-resolution logits and without
 * ever forming their gradient.  x: NHWC scores [N][Hi][Wi][ldx] (C classes), target int64
 * [N][Ho][Wo].  Accumulates (+=) loss_sum[0] (fp64), nvalid[0] (non-ignored pixels) and, if
 * dx32 != NULL, the UNSCALED gradient sum_{pixels} U^T (softmax - onehot) into the fp32 buffer
 * dx32 [N][Hi][Wi][lddx]; all three must be zeroed by the caller.  pixel_loss (fp32 [N][Ho][Wo],
 * may be NULL) receives the reduction='none' values (losses/ohem_loss.py:11-12).
 * With ohem != NULL the gradient of each pixel carries its OHEM weight (see tss_ce_fwd) and the
 * caller passes nvalid[0] = 1 to finalize.
 * finalize: loss[0] = loss_sum/nvalid (NaN if no valid pixel, like the reference) and
 * dx[i] = dx32[i] / nvalid in the activation dtype (n = N*Hi*Wi*lddx elements, n % 8 == 0). */
int tss_upsample_ce_fwd(const void* x, const int64_t* target, int N, int C, int Hi, int Wi, int Ho, int Wo,
                        int64_t ldx, int64_t ignore_index, double* loss_sum, int64_t* nvalid,
                        float* pixel_loss, float* dx32, int64_t lddx, const float* ohem, int dtype, void* stream);

/* ---- online hard example mining (losses/ohem_loss.py:10-21) -----------------------------------------
 * pixel_loss: the n reduction='none' cross-entropy values (0 at ignored pixels, which still count in
 * n, as in the reference).  n_keep = int(n * numel_frac).  v = the (n_keep+1)-th largest value, by a
 * 3-pass radix select (no sort, no host synchronisation).  loss[0] = mean of the values > thresh if
 * v > thresh, else the mean of the n_keep largest.  weights (device float[4]) receives the per-pixel
 * gradient weights rule {cut, w_above, tie, w_tie} for the `ohem` argument of the CE kernels (values
 * equal to the n_keep-th largest share the remaining slots evenly).  workspace:
 * tss_ohem_workspace_bytes() bytes, ZERO on first use (the call leaves it reusable). */
int64_t tss_ohem_workspace_bytes(void);
int tss_ohem_select(const float* pixel_loss, int64_t n, int64_t n_keep, float thresh, void* workspace,
                    float* loss, float* weights, void* stream);
int tss_upsample_ce_finalize(const double* loss_sum, const int64_t* nvalid, float* loss, const float* dx32,
                             void* dx, int64_t n, int dtype, void* stream);

/* ---- confusion matrix --------------------------------------------------------------------------
 * replaces ignite ConfusionMatrix.update x4 (engine.py:65-72): cm[t][p] += 1 for 0<=t<C.
 * cm is int64 [C][C].  Warp-aggregated shared-memory atomics, one global atomic per bin per CTA. */
int tss_confusion_from_labels(const int64_t* pred, const int64_t* target, int64_t n, int C,
                              int64_t* cm, void* stream);
/* same, with the argmax over the class planes of NCHW logits fused (ties -> lowest index,
 * NaN = max, like torch.argmax).  If pred_out != NULL the int64 argmax map is written too. */
int tss_confusion_from_logits(const void* logits, const int64_t* target, int N, int C, int64_t HW,
                              int64_t* cm, int64_t* pred_out, int dtype, void* stream);

/* ---- optimizer ---------------------------------------------------------------------------------
 * torch.optim.AdamW (scripts/train_fastscnn.py:125-129) over ONE flat fp32 parameter arena.
 * hyper (device, fp32[8]): lr, beta1, beta2, eps, weight_decay, step (incremented by the call),
 * bias_correction1, bias_correction2 (written by the call). */
int tss_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float* hyper,
                   float grad_scale, void* stream);

/* ---- dropout -------------------------------------------------------------------------------------------
 * nn.Dropout(p) (fastscnn.py:96, contextnet.py:84) on a dense tensor of n elements (n % 8 == 0) without a mask
 * tensor: y = x * keep / (1 - p) with keep from Philox4x32-10 keyed by (seed, offset, element group).
 * rng: DEVICE int64[3] = {seed, offset, ticket (0)}; the forward kernel stores the offset it used in used[0] and
 * advances rng[1] (CUDA-graph replays draw fresh masks); the backward kernel regenerates the mask from
 * (rng[0], used[0]): dx = dy * keep / (1 - p). */
int tss_dropout_fwd(const void* x, void* y, int64_t n, float p, int64_t* rng, int64_t* used, int dtype, void* stream);
int tss_dropout_bwd(const void* dy, void* dx, int64_t n, float p, int64_t* rng, int64_t* used, int dtype, void* stream);

/* ---- pyramid pooling branches, grouped (training) ------------------------------------------------------
 * PyramidPoolingModule (fastscnn.py:101-123): nbins x [AdaptiveAvgPool2d(b) -> Conv2d(C, Cb, 1) -> BatchNorm2d ->
 * ReLU] -> bilinear (align_corners=True) -> cat([x, branches]).  pool = the output of tss_adaptive_pool_fwd:
 * [N*sum(b*b)][C], branch-major.  y / z / dz / dy: [N*sum(b*b)][Cb] in the same row order (raw conv output /
 * activated / gradients).  table: DEVICE int64 [nbins][10] of addresses per branch = {weight (Cb,C) fp32, gamma,
 * beta, running_mean, running_var, num_batches_tracked (int64), dweight, dgamma, dbeta, 0}; 0 = absent.
 * branches_fwd: conv + batch statistics + finalize (mean/rstd out: [nbins][Cb]; running statistics updated with
 *   momentum, unbiased variance) + affine + ReLU, one launch.  A branch with a single value per channel is an error.
 * concat_fwd: cat[N][H][W][C + nbins*Cb] = [x, up(z_0), ..., up(z_{nbins-1})].
 * concat_bwd: dz = transpose of the up-samplings applied to dcat[..., C:] (dcat pitch lddcat).
 * branches_bwd: ReLU mask (from y) + BatchNorm backward -> dy; dgamma += , dbeta += ; dpool[N*sum(b*b)][C] =
 *   dy . w; dweight += dy^T . pool (two launches). */
int tss_ppm_branches_fwd(const void* pool, const int64_t* table, void* y, void* z, float* mean, float* rstd,
                         int N, int C, int Cb, const int* bins, int nbins, float momentum, float eps, int dtype,
                         void* stream);
/* eval mode: z = relu(conv * scale + shift) with the folded BatchNorm; table columns 1 and 2 hold the addresses of
 * the per-branch scale / shift vectors (tss_bn_fold) instead of gamma / beta. */
int tss_ppm_branches_eval(const void* pool, const int64_t* table, void* z, int N, int C, int Cb, const int* bins,
                          int nbins, int dtype, void* stream);
int tss_ppm_concat_fwd(const void* x, const void* z, void* cat, int N, int H, int W, int C, int Cb,
                       const int* bins, int nbins, int dtype, void* stream);
int tss_ppm_concat_bwd(const void* dcat, void* dz, int N, int H, int W, int C, int Cb, int64_t lddcat,
                       const int* bins, int nbins, int dtype, void* stream);
int tss_ppm_branches_bwd(const void* dz, const void* y, const void* pool, const int64_t* table, const float* mean,
                         const float* rstd, void* dy, void* dpool, int N, int C, int Cb, const int* bins, int nbins,
                         int dtype, void* stream);

/* ---- input pipeline on the device --------------------------------------
 * One launch for a batch of decoded samples: label ids -> train ids (TRAIN_MAPPING, data/cityscapes.py:17-20,88),
 * then albu.RandomScale -> RandomCrop -> HorizontalFlip -> Normalize -> ToTensor (scripts/train_fastscnn.py:62-68)
 * with the random draws supplied by the caller.  images: uint8 [N][H][W][3] (RGB), labels: uint8 [N][H][W] (or
 * NULL together with out_label).  geom: DEVICE int32 [N][5] = {scaled height, scaled width, crop y, crop x,
 * flip} per sample (scaled size = int(H*scale), int(W*scale); crop start inside the scaled image).
 * lut: DEVICE int64 [256] (NULL = identity).  norm: HOST float [6] = mean*255 (3), 1/(std*255) (3).
 * out_image: fp32 [N][3][ch][cw] (what the stem reads), out_label: int64 [N][ch][cw]; cw % 4 == 0.
 * Bit-identical to the CPU pipeline: OpenCV's uint8 INTER_LINEAR fixed-point arithmetic (image) and
 * INTER_NEAREST index rule (labels) are reproduced exactly; scale 1 / crop 0 is the evaluation transform. */
int tss_augment_batch(const void* images, const void* labels, const int* geom, const int64_t* lut,
                      const float* norm, float* out_image, int64_t* out_label, int N, int H, int W, int ch,
                      int cw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TSS_B200_H_ */
