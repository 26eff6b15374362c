/* libich_b200.so -- C ABI of the B200-native U-Net hot path.
 *
 * The reference (antoine-spahr/Label-Efficient-Volumetric-Deep-Semantic-Segmentation-of-ICH) is pure Python/PyTorch and
 * has no FFI of its own; the operators below are what its nn.Module / loss-module hot path dispatches to inside
 * PyTorch (cuDNN / ATen).  Each entry point names the reference call site (relative to code/src/) it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch's caching allocator); the library never allocates;
 *   - activations are channel-last rows [voxel][channel] with an explicit channel pitch `*_ld` (elements), so a tensor
 *     may be a channel slab of a wider (concat) buffer; `dtype` selects the activation element type;
 *   - grids are (N, D, H, W); 2-D nets pass D = 1 and KD = 1 / FD = 1;
 *   - `stream` is a cudaStream_t; all work is enqueued asynchronously on it;
 *   - return 0 on success; non-zero on error, message via ich_last_error() (thread-local).
 */
#ifndef ICH_B200_H
#define ICH_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ICH_F32 0
#define ICH_BF16 1

const char* ich_last_error(void);
int ich_abi_version(void);

/* ---- layout: reference tensors are NC(D)HW fp32 (models/optim/UNet2D.py:137); the engine is N(D)HWC ------------- */
int ich_layout_nc_to_nl(const float* src, void* dst, int dtype, int N, int C, long long S, int dst_ld, void* stream);
int ich_layout_nl_to_nc(const void* src, int dtype, int src_ld, float* dst, int N, int C, long long S, void* stream);

/* ---- weight packing (derived caches of the fp32 parameters): dst = permute(flip(src)) of a 5-D tensor [d0..d4], dst dim k runs over
 *      source dim p_k, source dims whose bit is set in flipmask are reversed; dst dtype fp32 or bf16.                                  */
int ich_permute5(const float* src, void* dst, int dtype, int d0, int d1, int d2, int d3, int d4, int p0, int p1, int p2, int p3, int p4,
                 int flipmask, void* stream);
/* batched form: n_jobs packs in one launch; src / dst are HOST arrays of device pointers, dtype / flipmask host arrays [n_jobs],
 * dims / perm host arrays [n_jobs][5] (used to re-derive every stale weight pack of a network after an optimizer step) */
int ich_permute5_batch(int n_jobs, const void* const* src, void* const* dst, const int* dtype, const int* dims, const int* perm,
                       const int* flipmask, void* stream);

/* ---- convolution, "same" padding, stride 1: nn.Conv3d/Conv2d k3 p1 (models/networks/UNet.py:153,155,158,160) and the 1x1
 *      heads (:84, :228).  wpack = [taps*Cin][Cout] fp32.  The data-gradient is the same entry point called with the
 *      flipped/transposed pack.  CUDA-core fp32-accumulate path (fp32 verification mode + odd shapes).               */
int ich_conv_fwd(const void* x, int x_ld, const float* wpack, const float* bias, void* y, int y_ld, int dtype, int N, int D, int H,
                 int W, int Cin, int Cout, int KD, int KH, int KW, int relu, void* stream);
/* weight gradient in torch layout [Cout][Cin][KD*KH*KW] fp32 (autograd of the convs above) */
int ich_conv_wgrad(const void* x, int x_ld, const void* dy, int dy_ld, int dtype, float* dw, int N, int D, int H, int W, int Cin,
                   int Cout, int KD, int KH, int KW, void* stream);

/* ---- tcgen05 / TMEM / TMA implicit-GEMM convolution, bf16 operands, fp32 accumulate (same call sites as ich_conv_fwd).
 *      wpack_bf16 = [taps][Cout][Cin] bf16 (K-major per tap).  Requires Cin % 16 == 0, Cout % 16 == 0, Cout <= 256.
 *      ich_conv_tc_supported() returns 1 when the shape is eligible.                                                  */
int ich_conv_tc_supported(int N, int D, int H, int W, int Cin, int Cout, int KD, int KH, int KW);
/* which kernel (and therefore which weight pack) a shape uses: 0 none, 1 slab kernel ([kd][kh][kw][Cout][Cin]),
 * 2 plane-streaming kernel with the depth taps folded into the MMA N dimension ([kh][kw][2-kd][Cout][Cin])            */
int ich_conv_tc_variant(int N, int D, int H, int W, int Cin, int Cout, int KD, int KH, int KW);
/* host-only: the tiling the slab kernel picks for a shape -- out[0..9] = cout block, rows per slab, M tiles per item, accumulator sets,
 * pipeline stages, kd-split, dynamic shared memory bytes, TMEM columns, K chunk width, work items; returns non-zero if unsupported */
int ich_conv_tc_plan_info(int N, int D, int H, int W, int Cin, int Cout, int KD, int KH, int KW, long long* out);
/* host-only: the tiling the plane-streaming kernel (3x3x3) picks -- out[0..9] = cout block, rows per item, M tiles per item, accumulator
 * ring slots per tile, pipeline stages, weights resident in shared memory, dynamic shared memory bytes, TMEM columns, one ring shared by
 * the tiles (1) or one ring per tile (0), work items; returns non-zero if the shape is not one the kernel takes                      */
int ich_conv_tc_stream_plan_info(int N, int D, int H, int W, int Cin, int Cout, long long* out);
int ich_conv_tc_fwd(const void* x, int x_ld, const void* wpack_bf16, const float* bias, void* y, int y_ld, int N, int D, int H, int W,
                    int Cin, int Cout, int KD, int KH, int KW, int relu, void* stream);
/* same, with the BatchNorm batch statistics (fp64 per-channel sum / sum of squares of the stored outputs) fused into the epilogue */
int ich_conv_tc_fwd_stats(const void* x, int x_ld, const void* wpack_bf16, void* y, int y_ld, double* sum, double* sumsq, int N, int D,
                          int H, int W, int Cin, int Cout, int KD, int KH, int KW, void* stream);
/* first layer (Cin = 1, the CT intensity: UNet.py:153 with in_channels = 1) on tcgen05: the 27-tap (2-D: 9-tap) im2col row of a
 * voxel is built in shared memory (K padded to 32), bf16 x bf16 -> fp32.  HBM-bound.  wpack = [taps][Cout] fp32 (the
 * ich_conv_fwd pack).  sum / sumsq (both or neither): fused BatchNorm batch statistics of the stored outputs, zeroed here.
 * Requires Cout in {8, 16, 24, 32}.  dw in torch layout [Cout][1][taps] fp32.                                             */
int ich_conv_cin1_tc_supported(int N, int D, int H, int W, int Cout, int KD);
int ich_conv_cin1_tc_fwd(const void* x, int x_ld, const float* wpack, const float* bias, void* y, int y_ld, double* sum, double* sumsq, int N,
                         int D, int H, int W, int Cout, int KD, int relu, void* stream);
int ich_conv_cin1_tc_wgrad(const void* x, int x_ld, const void* dy, int dy_ld, float* dw, int N, int D, int H, int W, int Cout, int KD,
                           void* stream);
/* transposed conv k2 s2 on tensor cores (same call site as ich_convT2_fwd): 1x1 GEMM + depth-to-space scatter epilogue.
 * wpack_bf16 = [taps*Cout][Cin] bf16.  Grid args = the COARSE grid.                                                   */
int ich_convT2_tc_supported(int N, int D, int H, int W, int Cin, int Cout, int FD);
int ich_convT2_tc_fwd(const void* x, int x_ld, const void* wpack_bf16, const float* bias, void* y, int y_ld, int N, int D, int H, int W,
                      int Cin, int Cout, int FD, void* stream);
/* weight gradient on tensor cores; dw in torch layout [Cout][Cin][taps] fp32 */
int ich_conv_tc_wgrad_supported(int N, int D, int H, int W, int Cin, int Cout, int KD, int KH, int KW);
int ich_conv_tc_wgrad(const void* x, int x_ld, const void* dy, int dy_ld, float* dw, int N, int D, int H, int W, int Cin, int Cout,
                      int KD, int KH, int KW, void* stream);
/* transposed-conv weight gradient on tensor cores: g = ich_space_to_depth2 of the up-sampled gradient ([voxel][tap*Cout+co]);
 * dw in torch layout [Cin][Cout][taps] fp32.  Grid args = the COARSE grid.                                               */
int ich_convT2_tc_wgrad_supported(int N, int D, int H, int W, int Cin, int Cout, int FD);
int ich_convT2_tc_wgrad(const void* x, int x_ld, const void* g, int g_ld, float* dw, int N, int D, int H, int W, int Cin, int Cout, int FD,
                        void* stream);
/* fine grid [2x voxels][C] (pitch src_ld) -> coarse grid [voxel][tap*C + c], tap = (i<<2)|(j<<1)|l; backward of ConvTranspose k2 s2 */
int ich_space_to_depth2(const void* src, int src_ld, void* dst, int dtype, int N, int D, int H, int W, int C, int FD, void* stream);
/* same, also accumulating colsum[c] (fp64, zeroed here) = sum over all fine voxels of src[.][c] = the ConvTranspose bias gradient */
int ich_space_to_depth2_sum(const void* src, int src_ld, void* dst, int dtype, int N, int D, int H, int W, int C, int FD, double* colsum,
                            void* stream);

/* ---- transposed conv k2 s2: nn.ConvTranspose3d/2d (models/networks/UNet.py:75-76). Grid args = the COARSE grid;
 *      FD = depth factor (2 for 3-D, 1 for 2-D). wpack = [Cin][taps*Cout], wpack_d = [taps*Cout][Cin], taps = 4*FD.     */
int ich_convT2_fwd(const void* x, int x_ld, const float* wpack, const float* bias, void* y, int y_ld, int dtype, int N, int D, int H,
                   int W, int Cin, int Cout, int FD, void* stream);
int ich_convT2_dgrad(const void* dy, int dy_ld, const float* wpack_d, void* dx, int dx_ld, int dtype, int N, int D, int H, int W, int Cin,
                     int Cout, int FD, void* stream);
int ich_convT2_wgrad(const void* x, int x_ld, const void* dy, int dy_ld, int dtype, float* dw, int N, int D, int H, int W, int Cin, int Cout,
                     int FD, void* stream);

/* transposed-conv backward WITHOUT the re-pack: the up-sampled gradient `dup` (channel slab of the fine grid [N, FD*D, 2H, 2W, .], pitch
 * dup_ld) is read in place through one strided tensor map per tap.  wpack_bf16 = [Cin][taps*Cout] bf16.  Grid args = the COARSE grid. */
int ich_convT2_tc_dgrad_supported(int N, int D, int H, int W, int Cin, int Cout, int FD);
int ich_convT2_tc_dgrad(const void* dup, int dup_ld, const void* wpack_bf16, void* dx, int dx_ld, int N, int D, int H, int W, int Cin, int Cout,
                        int FD, void* stream);
int ich_convT2_tc_wgrad_direct_supported(int N, int D, int H, int W, int Cin, int Cout, int FD);
int ich_convT2_tc_wgrad_direct(const void* x, int x_ld, const void* dup, int dup_ld, float* dw, int N, int D, int H, int W, int Cin, int Cout, int FD,
                               void* stream);

/* ---- BatchNorm (+ReLU): nn.BatchNorm3d/2d + nn.ReLU (models/networks/UNet.py:149,154,156,159,161,173-174) ----------- */
int ich_colstats(const void* x, int ld, int dtype, long long M, int C, double* sum, double* sumsq, void* stream);
int ich_bn_finalize(const double* sum, const double* sumsq, long long count, int C, const float* gamma, const float* beta,
                    const float* conv_bias, float* running_mean, float* running_var, float momentum, float eps, float* scale, float* shift,
                    float* save_mean, float* save_invstd, int training, void* stream);
int ich_affine_act(const void* y, int y_ld, const float* scale, const float* shift, void* z, int z_ld, int dtype, long long M, int C, int relu,
                   void* stream);
/* `sums` is a caller-provided workspace of ICH_BN_SUM_COPIES * 2 * C doubles (replicated partial sums, zeroed by the call) */
#define ICH_BN_SUM_COPIES 16
int ich_bn_act_bwd(const void* dz, int dz_ld, const void* y, int y_ld, const float* scale, const float* shift, const float* mean,
                   const float* invstd, double* sums, void* dy, int dy_ld, float* dgamma, float* dbeta, int dtype, long long M, int C, int relu,
                   int training, void* stream);

/* same two operators with nn.Dropout (models/networks/UNet.py:150,175-176: after the second ReLU of a ConvBlock) fused in:
 * z = relu(y*scale+shift) * keep/(1-p), keep = Philox4x32-10(seed; row, channel) >= p (16-bit resolution); the backward
 * regenerates the mask from the same (seed, row, channel) -- no mask tensor.  drop_p == 0 is the plain operator.          */
int ich_affine_act_drop(const void* y, int y_ld, const float* scale, const float* shift, void* z, int z_ld, int dtype, long long M, int C,
                        int relu, float drop_p, long long seed, void* stream);
int ich_bn_act_bwd_drop(const void* dz, int dz_ld, const void* y, int y_ld, const float* scale, const float* shift, const float* mean,
                        const float* invstd, double* sums, void* dy, int dy_ld, float* dgamma, float* dbeta, int dtype, long long M, int C,
                        int relu, int training, float drop_p, long long seed, void* stream);

/* SyncBN form of the backward (multi-GPU): phase 1 = reduction pass only (sums zeroed, then filled with this
 * rank's partial sums), the caller all-reduces `sums` over the ranks, phase 2 = apply pass only with global_count = rows behind the
 * statistics over all ranks.  dgamma / dbeta stay per-rank quantities (read them from `sums` between the phases; pass NULL here). */
int ich_bn_act_bwd_sync(const void* dz, int dz_ld, const void* y, int y_ld, const float* scale, const float* shift, const float* mean,
                        const float* invstd, double* sums, void* dy, int dy_ld, float* dgamma, float* dbeta, int dtype, long long M, int C,
                        int relu, int training, float drop_p, long long seed, int phase, long long global_count, void* stream);

/* ---- nn.MaxPool3d/2d(2,2) (models/networks/UNet.py:82,109); grid args = the INPUT grid ------------------------------ */
int ich_maxpool2_fwd(const void* x, int x_ld, void* y, int y_ld, int dtype, int N, int D, int H, int W, int C, int FD, void* stream);
/* same, also copying x into `skip` (a channel slab of the decoder's concat buffer, pitch skip_ld): the encoder block output is both
 * pooled and kept as the skip tensor (UNet.py:107-109,119), so the concat costs no pass of its own */
int ich_maxpool2_fwd_skip(const void* x, int x_ld, void* y, int y_ld, void* skip, int skip_ld, int dtype, int N, int D, int H, int W, int C,
                          int FD, void* stream);
/* dskip (optional): gradient of the SAME tensor arriving through the skip connection (UNet.py:107,119), added in the same pass */
int ich_maxpool2_bwd(const void* x, int x_ld, const void* dy, int dy_ld, void* dx, int dx_ld, int dtype, int N, int D, int H, int W, int C,
                     int FD, const void* dskip, int dskip_ld, void* stream);

/* ---- nn.Upsample(scale_factor=2, trilinear / bilinear, align_corners=True): the bilinear=True decoder (models/networks/UNet.py:69-72,117).
 *      Grid args = the INPUT grid; FD = depth factor (2 for 3-D, 1 for 2-D); y / dy may be a channel slab of the concat buffer.    */
int ich_upsample2_fwd(const void* x, int x_ld, void* y, int y_ld, int dtype, int N, int D, int H, int W, int C, int FD, void* stream);
int ich_upsample2_bwd(const void* dy, int dy_ld, void* dx, int dx_ld, int dtype, int N, int D, int H, int W, int C, int FD, void* stream);

/* ---- torch.cat([res, x], 1) (models/networks/UNet.py:119) as channel-slab copies; AdaptiveAvgPool(1) (:295,318) ------ */
int ich_slab_copy(const void* src, int src_ld, void* dst, int dst_ld, int dtype, long long M, int C, void* stream);
int ich_avgpool_fwd(const void* x, int ld, int dtype, float* out, int N, long long S, int C, void* stream);
int ich_avgpool_bwd(const float* dout, void* dx, int ld, int dtype, int N, long long S, int C, void* stream);

/* ---- final 1x1 conv + Sigmoid / Softmax (models/networks/UNet.py:84-91,122); act: 0 none, 1 sigmoid, 2 softmax -------- */
int ich_head_fwd(const void* x, int x_ld, int dtype, const float* w, const float* b, float* out, int N, long long S, int Cin, int Cout, int act,
                 void* stream);
int ich_head_dlogit(const float* out, const float* dout, void* dl, int dtype, int N, long long S, int Cout, int act, void* stream);
int ich_head1_bwd(const void* x, int x_ld, int dtype, const float* w, const float* out, const float* dout, void* dx, int dx_ld, float* dw,
                  float* db, long long M, int Cin, int act, void* stream);

/* ---- last ConvBlock unit fused with the single-class head (models/networks/UNet.py:173-174 then :122): z = ReLU(BN(y)) is never
 *      materialised.  fwd: out[m] = act(b + sum_c w[c] * relu(y[m][c] * scale[c] + shift[c])); bwd: d(logit) from out / dout, dz = d(logit) * w
 *      rebuilt inside the two BatchNorm-backward passes -> dy, d(gamma), d(beta) and the head's d(w), d(b).  act: 0 none, 1 sigmoid.
 *      sums: workspace of ICH_BN_SUM_COPIES * (3 * C + 1) doubles. ------------------------------------------------------------------- */
int ich_bn_head_supported(int dtype, int C);
int ich_bn_head_fwd(const void* y, int y_ld, int dtype, const float* scale, const float* shift, const float* w, const float* b, float* out,
                    long long M, int C, int relu, int act, void* stream);
int ich_bn_head_bwd(const void* y, int y_ld, int dtype, const float* scale, const float* shift, const float* mean, const float* invstd,
                    const float* w, const float* out, const float* dout, double* sums, void* dy, int dy_ld, float* dgamma, float* dbeta,
                    float* dw, float* db, long long M, int C, int relu, int training, int act, void* stream);

/* ---- BinaryDiceLoss / ComboLoss (models/optim/LossFunctions.py:39-63,143-166) ---------------------------------------- */
int ich_seg_loss_fwd(const float* pred, const float* mask, int B, long long S, float P, float eps, float alpha_empty, float w_bce, float w_dice,
                     float beta, int reduction, double* acc, float* per_sample, float* loss, void* stream);
int ich_seg_loss_bwd(const float* pred, const float* mask, const double* acc, const float* gscale, int B, long long S, float P, float eps,
                     float alpha_empty, float w_bce, float w_dice, float beta, float* dpred, void* stream);

/* ---- TverskyLoss (models/optim/LossFunctions.py:65-114): acc as above (P = 1), TL = 1 - (TP+eps)/(TP + beta FN + gamma FP + eps) -------- */
int ich_tversky_loss_fwd(const float* pred, const float* mask, int B, long long S, float eps, float alpha_empty, float beta, float gamma,
                         int reduction, double* acc, float* per_sample, float* loss, void* stream);
int ich_tversky_loss_bwd(const float* mask, const double* acc, const float* gscale, int B, long long S, float eps, float alpha_empty, float beta,
                         float gamma, float* dpred, void* stream);

/* ---- InfoNCELoss / LocalInfoNCELoss (models/optim/LossFunctions.py:208-230,308-341) ----------------------------------- */
int ich_infonce_fwd(const float* P, int B, int R, int E, float tau, float* Pn, float* invn, float* lse, float* rowloss, float* loss,
                    unsigned int* counter, void* stream);
int ich_infonce_bwd(const float* Pn, const float* invn, const float* lse, int B, int R, int E, float tau, const float* gout, float* dP,
                    void* stream);
int ich_region_gather(const float* f, const int* corners, float* P, int bs, int H, int W, int C, int A, int K, int view, void* stream);
int ich_region_scatter(const float* dP, const int* corners, float* df, int bs, int H, int W, int C, int A, int K, int view, void* stream);

/* ---- batch_binary_confusion_matrix (utils/tensor_utils.py:12-36) with the >= 0.5 threshold of models/optim/UNet2D.py:220 -- */
int ich_confusion(const float* pred, const float* target, int B, long long S, float thr, double* out, void* stream);

/* ---- input staging: raw CT -> window -> clip -> engine dtype in ONE pass.  Replaces the host-side
 *      window_ct (utils/ct_utils.py:13-36), the `input.to(device).float()` of models/optim/UNet2D.py:137-138 / :297 and the layout hop (a
 *      one-channel volume is already channel-last).  src_dtype: 0 fp32, 1 int16, 2 uint16, 3 uint8 (Hounsfield units or raw intensities). */
int ich_stage_ct(const void* src, int src_dtype, void* dst, int dtype, long long M, float win_min, float win_max, float out_lo, float out_hi,
                 void* stream);

/* ---- sliding-window volume inference on the device (models/optim/UNet2D.py:272-314 as a 3-D window driver):
 *      window extraction from a one-channel volume [D][H][W]; stitching of the fp32 predictions [n_win][wd][wh][ww] with the
 *      `pred >= 0.5` mask of UNet2D.py:220,303.  starts = device int [n_win][3] = (d0, h0, w0).  mode 0: windows do not overlap, acc
 *      (optional) and mask (optional, uint8) are written directly; mode 1: acc / cnt accumulate (fp32 atomics), then
 *      ich_blend_threshold divides and thresholds (mean blending of the overlaps). */
int ich_window_gather(const void* vol, int dtype, int D, int H, int W, const int* starts, int n_win, int wd, int wh, int ww, void* out, void* stream);
int ich_window_scatter(const float* pred, int D, int H, int W, const int* starts, int n_win, int wd, int wh, int ww, int mode, float thr,
                       float* acc, float* cnt, unsigned char* mask, void* stream);
int ich_blend_threshold(float* acc, const float* cnt, long long M, float thr, unsigned char* mask, void* stream);

/* ---- MLPHead (models/networks/UNet.py:179-209): Linear (+ReLU) on a [B, K] fp32 matrix, w = [N][K] (nn.Linear layout).
 *      bwd: `out` = the forward output (ReLU mask), any of dx / dw / db may be NULL (db is produced with dw). */
int ich_linear_fwd(const float* x, const float* w, const float* bias, float* out, int B, int K, int N, int relu, void* stream);
int ich_linear_bwd(const float* dout, const float* out, const float* x, const float* w, float* dx, float* dw, float* db, int B, int K, int N,
                   int relu, void* stream);

/* ---- GatedConv output (models/networks/GatedUNet.py:303-322): out = feat * sigmoid(gate) and its two gradients ------------------------- */
int ich_gate_mul_fwd(const void* feat, const void* gate, void* out, int dtype, long long M, void* stream);
int ich_gate_mul_bwd(const void* feat, const void* gate, const void* dout, void* dfeat, void* dgate, int dtype, long long M, void* stream);

#ifdef __cplusplus
}
#endif
#endif
