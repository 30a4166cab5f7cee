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

#ifdef __cplusplus
}
#endif
#endif /* B21_H_ */
