/*
 * b21.h — C-ABI of libb21.so, the B200 (sm_100a) kernel library behind the BraTS21 segmentation hot path.
 *
 * The reference (Alxaline/BraTS21) has no FFI boundary: its hot path is in-process Python calling torch.nn
 * (SURVEY.md §8b).  This header is therefore the *proposed* boundary: one entry point per kernel family, each
 * citing the reference code it replaces.  All pointers are raw DEVICE pointers owned by the caller (PyTorch's
 * caching allocator in practice); the library allocates nothing persistent.  Every call is asynchronous on
 * `stream` (a cudaStream_t passed as void*), re-entrant, and performs no hidden synchronisation.
 *
 * Return value: 0 on success, negative on error (B21_ERR_*); b21_last_error() returns a thread-local message.
 *
 * Tensor layouts
 *   "ndhwc bf16": activations, channels-last, [N][D][H][W][ld] with `ld` >= C the channel stride in elements
 *                 (so a tensor can be a channel slice of a wider concat buffer). C and ld are multiples of 8.
 *   "ncdhw f32" : network inputs / logits exactly as the reference's torch tensors (contiguous NCDHW fp32).
 */
#ifndef B21_H_
#define B21_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B21_OK 0
#define B21_ERR_BAD_ARG (-1)
#define B21_ERR_UNSUPPORTED (-2)
#define B21_ERR_CUDA (-3)

/* Number of contention-spreading slots in a norm-statistics buffer: double[B21_STAT_SLOTS][N][8][2]. */
#define B21_STAT_SLOTS 32

const char* b21_last_error(void);
int b21_version(void);

/* ------------------------------------------------------------------------------------------------ conv3d
 * Replaces torch.nn.Conv3d as used by conv3x3/conv1x1 (networks/equiunet2020.py:19-41), ConvEvoBlockCorrected /
 * ConvEvo (networks/equiunet2021.py:192-222) and SimpleASPPEVO (networks/equiunet2021.py:165-172):
 * stride 1, "same" padding (= dilation for k=3, 0 for k=1), groups 1.
 */

/* Padded output-channel count of the packed weight for `cout` (tile-rounded). */
int b21_conv_cout_padded(int cout);

/* Repack an fp32 torch weight [cout][cin][k][k][k] (k = 1 or 3) into bf16 [k^3][cout_padded][cin_padded] with
 * zero fill; cin_padded >= cin, multiple of 8.  If `transpose_flip` != 0 the packed weight is the one the
 * data-gradient needs: in/out channels swapped and taps mirrored (w'[co=ci][ci=co][2-kd][2-kh][2-kw]). */
int b21_pack_conv_weight(const float* w, void* packed, int cout, int cin, int cin_padded, int k,
                         int transpose_flip, void* stream);

/* y[n,d,h,w,0:cout] = conv(x)[...] + bias.  x: ndhwc bf16 with cin channels (ldx stride), y: ndhwc bf16 (ldy).
 * taps = 1 (k=1) or 27 (k=3).  If `stats` != NULL it must hold B21_STAT_SLOTS*n*8*2 doubles; it is zeroed and
 * receives per (slot, n, group of cout/8 channels) the sum and sum of squares of the fp32 pre-rounding outputs
 * (what GroupNorm(8, C) / EvoNorm3D group_std need: networks/factory.py:182, networks/equiunet2021.py:48-52). */
int b21_conv3d_fwd(const void* x, int ldx, const void* w_packed, const float* bias, void* y, int ldy,
                   double* stats, int n, int d, int h, int w, int cin, int cout, int taps, int dil,
                   void* stream);

/* Plane-marching variant of b21_conv3d_fwd for k = 3, dilation 1 layers whose weights fit in shared memory
 * (b21_conv_march_supported): every input plane is read once and serves all 27 taps; the three kd taps are folded
 * into the MMA N dimension.  Same semantics (bias, stats, channel-slice ld) as b21_conv3d_fwd with taps = 27, dil = 1.
 * The weight is packed by b21_pack_conv_weight_march into b21_conv_march_weight_bytes(cin, cout) bytes:
 * bf16 [kh*3+kw][ceil16(cin)/8][3*cout/8][8 n][8 k], n = (2-kd)*cout + co; `transpose_flip` as above. */
int b21_conv_march_supported(int cin, int cout);
long long b21_conv_march_weight_bytes(int cin, int cout);
int b21_pack_conv_weight_march(const float* w, void* packed, int cout, int cin, int transpose_flip, void* stream);
int b21_conv3d_march_fwd(const void* x, int ldx, const void* w_march, const float* bias, void* y, int ldy,
                         double* stats, int n, int d, int h, int w, int cin, int cout, void* stream);

/* Sliding-window variant of b21_conv3d_fwd for k = 3, dilation 1 layers with 16 <= cin <= 96 whose weights do not fit
 * in shared memory (b21_conv_slide_supported): 18 x 10 halo planes are shared-memory resident (each read once per
 * d-segment), the 27 weight tiles stream through a TMA ring and each feeds three output planes held in a TMEM ring.
 * Same semantics (bias, stats, channel-slice ld) as b21_conv3d_fwd with taps = 27, dil = 1.  The weight is packed by
 * b21_pack_conv_weight_slide into b21_conv_slide_weight_bytes(cin, cout) bytes:
 * bf16 [cout/NT][kd*9+kh*3+kw][ceil16(cin)/8][NT/8][8 n][8 k]; `transpose_flip` as above. */
int b21_conv_slide_supported(int cin, int cout);
long long b21_conv_slide_weight_bytes(int cin, int cout);
int b21_pack_conv_weight_slide(const float* w, void* packed, int cout, int cin, int transpose_flip, void* stream);
int b21_conv3d_slide_fwd(const void* x, int ldx, const void* w_slide, const float* bias, void* y, int ldy,
                         double* stats, int n, int d, int h, int w, int cin, int cout, void* stream);

/* Batched re-pack: after an optimizer step every conv weight of a network is re-packed (generic, plane-march and
 * sliding-window images, forward and transposed) -- ~90 launches of 2-30 us.  b21_pack_job_{tap,march,slide} fill one
 * HOST job record each (same arguments and layout rules as the matching b21_pack_conv_weight* call).  The caller
 * orders the records so that jobs with the same source `w` are consecutive (a group), gives every job of a group the
 * same block range blk0 / nblk = ceil(cout/16) * ceil(cin/16) (one block per 16 x 16 source tile; groups in ascending
 * blk0), copies the records to the device once and then calls b21_pack_batch: ONE launch of `total_blocks` blocks x 256
 * threads reads every source tile once and rewrites the valid elements of all its images (identical to the single
 * calls; padding rows / channels, zeroed by the single calls, are not touched). */
typedef struct b21_pack_job {
  const void* w;    /* fp32 weight [cout][cin][k^3] (device) */
  void* out;        /* packed bf16 image (device) */
  long long total;  /* elements of the image */
  int kind, cout, cin, tf;
  int p0, p1, p2, p3;
  int blk0, nblk;   /* block range of the job's group, set by the caller */
} b21_pack_job;
int b21_pack_job_tap(const float* w, void* packed, int cout, int cin, int cin_padded, int k, int transpose_flip,
                     b21_pack_job* job);
int b21_pack_job_march(const float* w, void* packed, int cout, int cin, int transpose_flip, b21_pack_job* job);
int b21_pack_job_slide(const float* w, void* packed, int cout, int cin, int transpose_flip, b21_pack_job* job);
int b21_pack_batch(const b21_pack_job* jobs_dev, int njobs, int total_blocks, void* stream);

/* First convolution of the networks (k = 3, dilation 1, at most 4 real input channels: networks/equiunet2020.py:19-25
 * encoder1.ConvBnRelu1, networks/equiunet2021.py:198 encoder1.conv_evo_1).  With 8 bytes per voxel the plane-marching
 * kernel is bound by its per-plane hand-shakes, so this variant builds the im2col of each input plane in shared memory
 * (8-byte cp.async, zero fill = padding; k = (kh*3+kw)*4 + c, 36 -> 48), folds the three kd taps into N and groups the
 * output planes in quads: 18 tcgen05.mma per four planes, every barrier / commit once per two or four planes.  x holds the real channels in the first
 * 8 bytes of each dense 16-byte voxel record (ldx = 8); same bias / stats semantics as b21_conv3d_fwd;
 * act: 0 none, 1 swish (x * sigmoid x) before the store, 2 the same through tanh.approx.  The weight is packed by
 * b21_pack_conv_weight_input into b21_conv_input_weight_bytes(cout) bytes: bf16 [6][3*cout/8][8 n][8 k], n = (2-kd)*cout + co. */
int b21_conv_input_supported(int cin_true, int cout);
long long b21_conv_input_weight_bytes(int cout);
int b21_pack_conv_weight_input(const float* w, void* packed, int cout, int cin, void* stream);
int b21_pack_job_input(const float* w, void* packed, int cout, int cin, b21_pack_job* job);
int b21_conv3d_input_fwd(const void* x, int ldx, const void* w_input, const float* bias, void* y, int ldy,
                         double* stats, int act, int n, int d, int h, int w, int cout, void* stream);

/* Persistent 1x1x1 variant of b21_conv3d_fwd (taps = 1) for the HBM-bound ConvEvo bridges / up-convs
 * (networks/equiunet2021.py:214-222,262-269): weights resident in shared memory, activation tiles streamed through a
 * TMA ring, double-buffered TMEM accumulator.  x / y are [n][nvox][ld] bf16; `w_packed` is the k = 1 packing of
 * b21_pack_conv_weight.  Same bias / stats semantics as b21_conv3d_fwd. */
int b21_conv_point_supported(int cin, int cout);
int b21_conv1x1_fwd(const void* x, int ldx, const void* w_packed, const float* bias, void* y, int ldy,
                    double* stats, int n, long long nvox, int cin, int cout, void* stream);

/* ------------------------------------------------------------------------------ folded EvoNorm (inference)
 * EvoNorm-S0 is y = x*sigmoid(x) * a_c + b_c where only the affine (a, b) depends on the global group statistics
 * (networks/equiunet2021.py:48-52,95-105).  The *_fold conv entry points store S = swish(conv + bias) directly
 * (act = 1) with the statistics, and the affine (A, B) [n][c] — including the ResidualSE gate (equiunet2021.py:204-205)
 * — is folded into the consumer: per-sample packed weights W*A[n][ci] (b21_pack_conv_weight_*_fold) and a bias table
 * T[n][border class][co] = bias + sum_ci B[n][ci] * (sum of the taps of W[co][ci] that stay inside the volume)
 * (b21_border_weight_sums once per weight, b21_bias_table per forward; class = cd*9+ch*3+cw, c = 0 first voxel /
 * 1 interior / 2 last voxel of the axis; one class for k = 1).  This removes every normalisation pass over HBM.
 * chan_sum (fp32 [n][cout], zeroed by the caller) receives the channel sums of the stored outputs (SE squeeze). */
int b21_evo_se_affine(const double* stats, const float* gamma, const float* beta, const float* chan_sum,
                      const float* w1, const float* b1, const float* w2, const float* b2, float* a_out, float* b_out,
                      int ldab, int n, int c, int hidden, long long nvox, float eps, void* stream);
int b21_border_weight_sums(const float* w, float* ws, int cout, int cin, int taps, void* stream);
int b21_bias_table(const float* ws, const float* bias, const float* b_in, int ldab, float* table, int n, int cout,
                   int cin, int ncls, void* stream);
/* The per-sample packing kernels also emit the bias table in the same launch when `table` != NULL
 * (ws / bias / b_in as for b21_bias_table; B rows share `ldscale`).  They read the fp32 weight in coalesced
 * 16 x 16 x taps tiles and rewrite the VALID elements of every per-sample image only: the caller zero-fills `packed`
 * once (padding rows / channels must be 0). */
int b21_pack_conv_weight_fold(const float* w, void* packed, int cout, int cin, int cin_padded, int k,
                              const float* scale, int ldscale, int nsamples, const float* ws, const float* bias,
                              const float* b_in, float* table, void* stream);
int b21_pack_conv_weight_march_fold(const float* w, void* packed, int cout, int cin, const float* scale, int ldscale,
                                    int nsamples, const float* ws, const float* bias, const float* b_in, float* table,
                                    void* stream);
int b21_pack_conv_weight_slide_fold(const float* w, void* packed, int cout, int cin, const float* scale, int ldscale,
                                    int nsamples, const float* ws, const float* bias, const float* b_in, float* table,
                                    void* stream);
int b21_conv3d_march_fwd_fold(const void* x, int ldx, const void* w_march, long long wstride_n, const float* bias,
                              const float* bias_table, void* y, int ldy, double* stats, float* chan_sum, int act,
                              int n, int d, int h, int w, int cin, int cout, void* stream);
/* b21_conv3d_march_fwd_fold with the input channels split over two tensors (x: [0, cin1), x2: [cin1, cin)): the
 * consumer of a channel concat (torch.cat at equiunet2021.py:310,315,319) reads both producers' dense outputs. */
int b21_conv3d_march_fwd_fold2(const void* x, int ldx, int cin1, const void* x2, int ldx2, const void* w_march,
                               long long wstride_n, const float* bias, const float* bias_table, void* y, int ldy,
                               double* stats, float* chan_sum, int act, int n, int d, int h, int w, int cin, int cout,
                               void* stream);
int b21_conv3d_slide_fwd_fold(const void* x, int ldx, const void* w_slide, long long wstride_n, const float* bias,
                              const float* bias_table, void* y, int ldy, double* stats, float* chan_sum, int act,
                              int n, int d, int h, int w, int cin, int cout, void* stream);
int b21_conv1x1_fwd_fold(const void* x, int ldx, const void* w_packed, int per_sample, const float* bias,
                         const float* bias_n, void* y, int ldy, double* stats, int act, int n, long long nvox, int cin,
                         int cout, void* stream);
/* MONAI MaxAvgPool (mode 2: [max | mean], equiunet2021.py:261) or max-pool (mode 1) of A[n][c] * x + B[n][c]. */
int b21_affine_pool(const void* x, int ldx, const float* a_in, const float* b_in, int ldab, void* pooled, int ldpool,
                    int mode, int n, int d, int h, int w, int c, void* stream);

/* ------------------------------------------------------------------------------------- normalisation / SE
 * norm_apply: y = GroupNorm(8,C)(x) -> ReLU (mode 0; networks/factory.py:182 + equiunet2020.py:60-61) or
 * EvoNorm3D-S0 (mode 1; networks/equiunet2021.py:48-52,95-105: x*sigmoid(x)/sqrt(var_unbiased+eps)*gamma+beta)
 * from the statistics buffer b21_conv3d_fwd filled.  x may alias y.  If chan_sum != NULL (fp32 [chan_slots][n][c],
 * zeroed by the caller) the per-channel sums of the outputs are accumulated into it (squeeze-excite mean): block b adds
 * to copy b % chan_slots (one copy makes ~1200 blocks serialise on c addresses); the caller sums the copies. */
int b21_norm_apply(const void* x, int ldx, void* y, int ldy, const double* stats, const float* gamma,
                   const float* beta, float* chan_sum, int chan_slots, int mode, int n, long long nvox, int c,
                   float eps, void* stream);

/* The rest of the norm / activation factory of EquiUnet (networks/factory.py:179-200: get_norm_layer "instance" =
 * nn.InstanceNorm3d(affine), "batch" = nn.BatchNorm3d(affine), "none"; get_act "relu" | "leakyrelu" | "elu" through
 * MONAI's Act), as three passes: b21_channel_stats (per (n, c) sum / sum of squares over the voxels, double
 * out[n][c][2]), b21_norm_coeffs (statistics -> per-(n, c) affine a, b with y = a*x + b the normalisation; kind 0 =
 * GroupNorm(8) from the conv-epilogue statistics, 1 = instance, 2 = batch norm in training mode (batch statistics over
 * n and voxels; running_mean / running_var updated with `momentum`, unbiased variance, as nn.BatchNorm3d), 3 = batch
 * norm in eval mode (running statistics), 4 = none (a = 1, b = 0)), and b21_affine_act (y = act(a[n][c]*x + b[n][c]),
 * act 0 identity / 1 ReLU / 2 LeakyReLU(slope) / 3 ELU(alpha = 1); x may alias y). */
int b21_channel_stats(const void* x, int ldx, double* out, int n, long long nvox, int c, void* stream);
int b21_norm_coeffs(int kind, const double* stats, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float* a_out, float* b_out, int n, int c, long long nvox,
                    float eps, void* stream);
int b21_affine_act(const void* x, int ldx, void* y, int ldy, const float* a, const float* b, int act, float slope, int n,
                   long long nvox, int c, void* stream);

/* MONAI ResidualSELayer(r=2, relu, sigmoid) gate used at networks/equiunet2021.py:204-205:
 * scale[n][c] = 1 + sigmoid(W2 relu(W1 (chan_sum[n] * inv_count) + b1) + b2), so that x + x*s == x*scale. */
int b21_se_gate(const float* chan_sum, const float* w1, const float* b1, const float* w2, const float* b2,
                float* scale, int n, int c, int hidden, float inv_count, void* stream);

/* x*scale[n][c] written full-res (`full`, may alias x, may be NULL) and/or pooled 2x2x2 (`pooled`):
 * mode 0 = scale only, 1 = nn.MaxPool3d(2,2) (equiunet2020.py:433), 2 = MONAI MaxAvgPool -> [max | avg] channel
 * concat, 2c channels (equiunet2021.py:261).  scale may be NULL (= 1). */
int b21_scale_pool(const void* x, int ldx, const float* scale, void* full, int ldfull, void* pooled, int ldpool,
                   int mode, int n, int d, int h, int w, int c, void* stream);

/* nn.Upsample(scale_factor=2, mode="trilinear", align_corners=True) (equiunet2020.py:439, equiunet2021.py:270),
 * ndhwc bf16 [n,d,h,w,c] -> [n,2d,2h,2w,c] written into a channel slice (ldy). */
int b21_upsample2x(const void* x, int ldx, void* y, int ldy, int n, int d, int h, int w, int c, void* stream);

/* Same interpolation by an integer factor s on ncdhw fp32 planes (deep-supervision heads,
 * equiunet2020.py:444-458, equiunet2021.py:274-280). */
int b21_upsample_f32(const float* x, float* y, int planes, int d, int h, int w, int s, void* stream);

/* conv1x1 to k<=4 classes (outconv / out_conv / deep heads: equiunet2020.py:441-458, equiunet2021.py:271-280):
 * out (ncdhw fp32 [n][k][nvox]) = b + W (x * scale[n] + offset[n]); scale / offset ([n][ldso] fp32, the folded affine of
 * the input: SE gate or EvoNorm (A, B)) may be NULL (= 1 / 0). */
int b21_head_conv(const void* x, int ldx, const float* scale, const float* offset, int ldso, const float* w,
                  const float* b, float* out, int n, long long nvox, int c, int k, void* stream);

/* ------------------------------------------------------------------------- sliding window / TTA / labels
 * A TTA variant is (perm[3], flip[3]): augmented[a0,a1,a2] = volume[s0,s1,s2] with
 * s_j = flip[j] ? dim_j-1-a_{perm[j]} : a_{perm[j]}  (tta/transforms.py:16-98,149-173 are all of this form).
 *
 * pack_windows: crop `nwin` (<=16) windows of size d,h,w at `origins` ([nwin][3], augmented frame, may be
 * negative = zero padding as F.pad in utils/inferers.py:101-109) out of ncdhw fp32 volumes and write ndhwc bf16
 * [nwin][d][h][w][cpad] (channels >= vc are zero).  Replaces augment_image + the slice/cat at inferers.py:126-132. */
int b21_pack_windows(const float* vol, int vc, int vd, int vh, int vw, void* out, int cpad, int nwin, int d, int h,
                     int w, const int* origins, const int* vol_index, const int* perm, const int* flip, void* stream);

/* acc[k][ad][ah][aw] (+)= max(prof_d[z]*prof_h[y]*prof_w[x], wfloor) * logits[win][k][z][y][x] for each of the nwin
 * (<= 16) windows, added in window order (utils/inferers.py:149-151; importance map = outer product of the three 1-D
 * profiles, clamped below at its smallest non-zero value `wfloor` as MONAI's compute_importance_map does; 0 = no
 * clamp).  logits == NULL accumulates the weights themselves (the count map, k = 1). */
int b21_blend_accumulate(const float* logits, float* acc, const float* prof_d, const float* prof_h,
                         const float* prof_w, int nwin, int k, int d, int h, int w, int ad, int ah, int aw,
                         const int* origins, float wfloor, void* stream);

/* prob_sum[k][vd][vh][vw] (+)= sigmoid(acc/cnt) de-augmented (inferers.py:154-162 + deaugment_mask +
 * engine.py:239-249).  cnt may be NULL; pad_before = the F.pad offsets of the augmented frame (or NULL). */
int b21_tta_accumulate(const float* acc, const float* cnt, float* prob_sum, int k, int ad, int ah, int aw,
                       const int* pad_before, int vd, int vh, int vw, const int* perm, const int* flip,
                       int apply_sigmoid, int overwrite, void* stream);

/* (prob_sum / count >= thresh) -> onehot uint8 [3][nvox] (TC, WT, ET), zeroed where every image channel is 0
 * (remove_background_voxels, utils/transforms.py:536-550), and the BraTS label map uint8 [nvox]
 * (ConvertToBratsClassesBasedOnMultiChannel + ChangeLabel3To4, utils/transforms.py:169-206). */
int b21_labels_finalize(const float* prob_sum, float count, float thresh, const float* image, int image_channels,
                        uint8_t* onehot, uint8_t* label, long long nvox, int et_label, void* stream);

/* remove_background_voxels (utils/transforms.py:536-550) as its own pass, for the flows where the reference runs the
 * post transforms between the threshold and the background mask (learning/engine.py:249-256): label planes
 * uint8 [planes][nvox] are zeroed where every channel of image fp32 [image_channels][nvox] is exactly 0. */
int b21_mask_background(uint8_t* label, int planes, const float* image, int image_channels, long long nvox,
                        void* stream);

/* ------------------------------------------------------------------------------------------- training step
 * Backward of the network body and the criterion (learning/engine.py:88-130: forward, Dice over the heads,
 * scaler.scale(loss).backward(), optimizer step).  Gradients of activations are ndhwc bf16; gradients of
 * parameters are fp32 and ACCUMULATED into the buffers passed (zero them at the start of a step).
 */

/* dW[cout][cin][taps] (fp32, torch Conv3d weight layout) += sum_v dz[v][co] * x[v + off(tap)][ci]: weight gradient
 * of b21_conv3d_fwd (same x / taps / dil; dz is the gradient of its output).  The data gradient is b21_conv3d_fwd
 * itself on dz with the `transpose_flip` packing of the weight. */
int b21_conv3d_wgrad(const void* x, int ldx, const void* dz, int lddz, float* dw, int n, int d, int h, int w, int cin,
                     int cout, int taps, int dil, void* stream);

/* Plane-marching variant of b21_conv3d_wgrad for k = 3, dilation 1, cin <= 96, cout <= 128, h, w >= 8
 * (b21_conv_wgrad_march_supported): x halo planes and dz planes are read once into shared memory, taps are descriptor
 * start addresses, the three kd taps are folded into the MMA N dimension.  `cin` is the channel count of x (multiple
 * of 8, may include zero padding), `cin_true` <= cin the input-channel count of dw [cout][cin_true][27]. */
int b21_conv_wgrad_march_supported(int cin, int cout);
int b21_conv3d_wgrad_march(const void* x, int ldx, const void* dz, int lddz, float* dw, int n, int d, int h, int w,
                           int cin, int cin_true, int cout, void* stream);

/* Backward of b21_norm_apply (mode 0 GroupNorm(8)+ReLU, mode 1 EvoNorm3D-S0) from the pre-norm tensor z and the
 * forward statistics; dy is the gradient of the layer output.  With se_w1 != NULL the layer output was additionally
 * multiplied by the MONAI ResidualSELayer gate (b21_se_gate: se_scale, from the channel means se_mean) and dy is the
 * gradient of that product; the gate's MLP gradients are accumulated into d_w1/d_b1/d_w2/d_b2.  dgamma/dbeta are
 * accumulated; colsum (optional, fp32 [c]) accumulates sum_v dz (= bias gradient of the producing conv).
 * workspace: b21_norm_bwd_workspace_bytes(n, c) bytes, 16-byte aligned (reduction tables spread over 32 / 16 copies so
 * that the ~1200 blocks of a pass do not serialise on a few L2 atomic addresses).  dz may alias dy. */
long long b21_norm_bwd_workspace_bytes(int n, int c);
int b21_norm_bwd(const void* dy, int lddy, const void* z, int ldz, void* dz, int lddz, const double* stats,
                 const float* gamma, const float* beta, float* dgamma, float* dbeta, float* colsum,
                 const float* se_scale, const float* se_mean, const float* se_w1, const float* se_b1,
                 const float* se_w2, const float* se_b2, float* d_w1, float* d_b1, float* d_w2, float* d_b2,
                 int hidden, void* workspace, long long workspace_bytes, int mode, int n, long long nvox, int c,
                 float eps, void* stream);

/* Backward of b21_scale_pool modes 1/2: dy = add (optional) + max-routed dpool[:c] (+ dpool[c:2c]/8 for mode 2);
 * y is the tensor that was pooled. */
int b21_pool_bwd(const void* y, int ldy, const void* dpool, int ldp, const void* add, int ldadd, void* dy, int lddy,
                 int mode, int n, int d, int h, int w, int c, void* stream);

/* Adjoints of b21_upsample2x / b21_upsample_f32 (d, h, w are the LOW-resolution dims). */
int b21_upsample2x_bwd(const void* dy, int lddy, void* dx, int lddx, int n, int d, int h, int w, int c, void* stream);
int b21_upsample_f32_bwd(const float* dy, float* dx, int planes, int d, int h, int w, int s, void* stream);

/* Backward of b21_head_conv: dx (bf16, += if accumulate) = scale * W^T dl; dws[slot][n][k][c] += sum_v dl x (so that
 * dW = scale * dws and dscale = sum_k W dws); db[slot][k] += sum_v dl.  Block b adds to copy b % slots of both tables
 * (the caller zeroes them and sums the copies): one copy makes ~1200 blocks serialise on k*c L2 atomic addresses. */
int b21_head_conv_bwd(const void* x, int ldx, const float* scale, const float* w, const float* dl, void* dx, int lddx,
                      int accumulate, float* dws, float* db, int slots, int n, long long nvox, int c, int k,
                      void* stream);

/* dst += src on ndhwc bf16 (gradient fan-in). */
int b21_add_inplace(void* dst, int ldd, const void* src, int lds, long long nvox_total, int c, void* stream);

/* monai.losses.DiceLoss(include_background, sigmoid, squared_pred, batch=True, jaccard?) as the reference builds it
 * (src/definer.py:184-203).  dice_fwd: sums (double [k][3] scratch), loss[0] += weight * mean_k f_k, coef (float
 * [k][2]) for the backward.  dice_bwd: dlogits = gscale * gout[0] * dloss/dlogits (gout may be NULL = 1). */
int b21_dice_fwd(const float* logits, const float* target, double* sums, float* loss, float* coef, int n, int k,
                 long long nvox, int jaccard, float smooth_nr, float smooth_dr, float weight, void* stream);
int b21_dice_bwd(const float* logits, const float* target, const float* coef, const float* gout, float gscale,
                 float ce_weight, float* dlogits, int n, int k, long long nvox, void* stream);
/* Cross-entropy half of DiceCELoss (learning/losses.py:470-595, `--criterion dice_ce`, src/definer.py:204-212):
 * loss[0] += weight * mean over the n * nvox voxels of CrossEntropy(logits[:, :, v], argmax_k target[:, k, v])
 * (first maximum, as torch.argmax at losses.py:572); scratch = one double.  Its gradient is folded into b21_dice_bwd:
 * ce_weight = lambda_ce adds lambda_ce * (softmax_k - [k == y]) / (n * nvox) to the Dice term (0 = plain Dice). */
int b21_ce_fwd(const float* logits, const float* target, double* scratch, float* loss, int n, int k, long long nvox,
               float weight, void* stream);

/* Fused multi-tensor Ranger2020 step (learning/optimizer.py:136-255).  table: int64 [ntensors][6] = {param, grad,
 * exp_avg, exp_avg_sq, slow_buffer (device pointers, fp32), numel}; chunks: int32 [nchunks][2] = {tensor, offset}
 * with b21_ranger_chunk() elements per chunk; gscale multiplies every gradient (loss-scale / data-parallel mean).
 * dyn (optional, device float[4] = {lr * step_size, rectified, lookahead, gscale}) overrides the step-dependent scalar
 * arguments at run time, so that one captured launch serves every step of a CUDA-graphed training loop. */
int b21_ranger_chunk(void);
int b21_ranger_step(const long long* table, const int* chunks, int nchunks, float gscale, float lr, float step_size,
                    float beta1, float beta2, float eps, float weight_decay, int rectified, int lookahead, float alpha,
                    const float* dyn, void* stream);

/* Gradient centralisation (centralized_gradient, learning/optimizer.py:11-20, applied at optimizer.py:187-188 when
 * use_gc): every dim-0 slice of a qualifying gradient tensor has its mean subtracted in place, before
 * b21_ranger_step.  rows: int64 [nrows][2] = {device pointer to the fp32 row, row length}. */
int b21_grad_centralize(const long long* rows, int nrows, void* stream);

/* ------------------------------------------------------------------------------------- input side (pre.cu)
 * CropForegroundd + NormalizeIntensityd(nonzero, channel_wise[, remove_outliers]) + shape_to_divisible(k)
 * (src/definer.py:561-567, utils/transforms.py:328-406,483-512) on an fp32 [c][d][h][w] device volume.
 * bbox (device int[6]) = {min d, min h, min w, max d+1, max h+1, max w+1} of the voxels where any channel > 0
 * ({d, h, w, 0, 0, 0} when there is none).  stats (device double [c][3]) = {count, sum, sum of squares} of the
 * voxels != 0 inside bbox.  b21_normalize_crop_pad writes out[c][od][oh][ow]: the box, shifted by (pad_d, pad_h,
 * pad_w), zero elsewhere; non-zero voxels become (x - mean) / std (population std, 0 -> 1), clipped to +-clip when
 * clip > 0 (remove_outliers); zeros stay zero. */
int b21_foreground_bbox(const float* img, int c, int d, int h, int w, int* bbox, void* stream);
int b21_nonzero_stats(const float* img, int c, int d, int h, int w, const int* bbox, double* stats, void* stream);
int b21_normalize_crop_pad(const float* img, float* out, int c, int d, int h, int w, const int* bbox,
                           const double* stats, int od, int oh, int ow, int pad_d, int pad_h, int pad_w, float clip,
                           void* stream);

/* --------------------------------------------------------------------------- label post-processing (post.cu)
 * b21_keep_components: KeepLargestConnectedComponent / get_largest_component (utils/transforms.py:209-230,579-600)
 * on a uint8 label map [d][h][w], in place: components of (label != 0) under full 26-connectivity
 * (skimage.morphology.label default) with at most `threshold` voxels are zeroed; threshold < 0 keeps only the
 * largest component (the first in raster order on ties).  work: b21_keep_components_workspace_bytes(nvox) bytes.
 * b21_replace_rare_labels: ReplaceWithClosestValue (utils/transforms.py:233-268,603-647), in place: every label
 * value carried by at most `thresh` voxels is replaced — provided at least one NON-ZERO value is that rare — by the
 * value of the nearest voxel (euclidean, inside the 2-D slice orthogonal to `axis`) with a kept label; equidistant
 * candidates: the smallest row-major slice index.  dims (n0, n1, n2) are those of the squeezed label map.
 * b21_labels_to_channels: MONAI ConvertToMultiChannelBasedOnBratsClasses: uint8 [3][nvox] = (TC = 1|4, WT = 1|2|4,
 * ET = 4) (definer.py:691). */
long long b21_keep_components_workspace_bytes(long long nvox);
int b21_keep_components(uint8_t* label, void* work, int d, int h, int w, int threshold, void* stream);
long long b21_replace_rare_workspace_bytes(int thresh);
int b21_replace_rare_labels(uint8_t* label, void* work, int n0, int n1, int n2, int thresh, int axis, void* stream);
int b21_labels_to_channels(const uint8_t* label, uint8_t* onehot, long long nvox, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B21_H_ */
