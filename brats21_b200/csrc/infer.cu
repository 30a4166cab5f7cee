// infer.cu — kernels of the sliding-window + TTA inference wrappers (utils/inferers.py, tta/*, learning/engine.py
// :424-440,236-252, utils/transforms.py:169-206,536-550), all HBM-bound, fp32 accumulators.
//
// A TTA variant is a signed axis permutation of the volume.  It is described by (perm[3], flip[3]) such that the
// augmented volume A relates to the source volume V by
//        A[a0,a1,a2] = V[s0,s1,s2],   s_j = flip[j] ? Vdim_j - 1 - a_{perm[j]} : a_{perm[j]}
// (source axis j is read from augmented axis perm[j], mirrored when flip[j]).  All 16 reference variants
// (OnAxes x HorizontalFlip x Rotate90) and all 8 axis-flip subsets are of this form.
//
//   pack_windows     crop windows of A (never materialised) from V (NCDHW fp32) -> NDHWC bf16, zero padded
//   blend_accumulate acc[k, A-frame] += w_d*w_h*w_w * logits;  separable importance profile (constant/gaussian)
//   blend_count      cnt[A-frame]   += w_d*w_h*w_w              (once per window grid)
//   tta_accumulate   prob_sum[k, V-frame] += sigmoid(acc/cnt) gathered through the inverse of the variant
//   labels_finalize  mean >= thresh -> {TC,WT,ET} bits, background removal, BraTS label map (uint8)
#include "ptx.cuh"
#include "host_common.h"

namespace b21 {

constexpr int kMaxWin = 16;
struct WinList {
  int org[kMaxWin][3];  // window origin in the (padded) augmented frame
  int vol[kMaxWin];     // which volume of the batch
};
struct Variant {
  int perm[3];
  int flip[3];
};

__global__ void __launch_bounds__(256) pack_windows_kernel(const float* __restrict__ vol, int VC, int VD, int VH,
                                                           int VW, __nv_bfloat16* __restrict__ out, int cpad, int B,
                                                           int d, int h, int w, WinList wl, Variant tv) {
  const long long total = (long long)B * d * h * w;
  const int Vdim[3] = {VD, VH, VW};
  const size_t plane = size_t(VD) * VH * VW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long v = i;
    const int x = int(v % w); v /= w;
    const int y = int(v % h); v /= h;
    const int z = int(v % d);
    const int b = int(v / d);
    const int a[3] = {wl.org[b][0] + z, wl.org[b][1] + y, wl.org[b][2] + x};
    int s[3];
    bool inside = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int aj = a[tv.perm[j]];
      s[j] = tv.flip[j] ? Vdim[j] - 1 - aj : aj;
      inside = inside && (aj >= 0) && (aj < Vdim[j]);
    }
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = 0.f;
    if (inside) {
      const float* src = vol + size_t(wl.vol[b]) * VC * plane + (size_t(s[0]) * VH + s[1]) * VW + s[2];
      for (int c = 0; c < VC && c < 8; ++c) f[c] = __ldg(src + c * plane);
    }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
    __nv_bfloat16* dst = out + size_t(i) * cpad;
    *reinterpret_cast<uint4*>(dst) = o;
    for (int c = 8; c < cpad; c += 8) *reinterpret_cast<uint4*>(dst + c) = make_uint4(0, 0, 0, 0);
  }
}

// one launch per window: deterministic accumulation order (windows of a batch may overlap)
__global__ void __launch_bounds__(256) blend_accumulate_kernel(const float* __restrict__ logits, float* acc,
                                                               const float* __restrict__ pd,
                                                               const float* __restrict__ ph,
                                                               const float* __restrict__ pw, int K, int d, int h,
                                                               int w, int AD, int AH, int AW, int o0, int o1, int o2) {
  const long long total = (long long)K * d * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long v = i;
    const int x = int(v % w); v /= w;
    const int y = int(v % h); v /= h;
    const int z = int(v % d);
    const int k = int(v / d);
    const float wt = __fmul_rn(__fmul_rn(__ldg(pd + z), __ldg(ph + y)), __ldg(pw + x));
    const size_t ai = ((size_t(k) * AD + (o0 + z)) * AH + (o1 + y)) * AW + (o2 + x);
    // separate multiply and add (no FMA contraction): the reference does `out += importance_map * seg_prob`
    acc[ai] = __fadd_rn(acc[ai], logits ? __fmul_rn(wt, __ldg(logits + i)) : wt);
  }
}

__global__ void __launch_bounds__(256) tta_accumulate_kernel(const float* __restrict__ acc,
                                                             const float* __restrict__ cnt, float* prob_sum, int K,
                                                             int AD, int AH, int AW, int p0, int p1, int p2, int VD,
                                                             int VH, int VW, Variant tv, int apply_sigmoid,
                                                             int overwrite) {
  const long long nv = (long long)VD * VH * VW;
  const long long total = (long long)K * nv;
  const int Vdim[3] = {VD, VH, VW};
  const int pad[3] = {p0, p1, p2};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long v = i;
    int s[3];
    s[2] = int(v % VW); v /= VW;
    s[1] = int(v % VH); v /= VH;
    s[0] = int(v % VD);
    const int k = int(v / VD);
    int a[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) a[tv.perm[j]] = (tv.flip[j] ? Vdim[j] - 1 - s[j] : s[j]);
    const size_t ai = (size_t(a[0] + pad[0]) * AH + (a[1] + pad[1])) * AW + (a[2] + pad[2]);
    float val = __ldg(acc + size_t(k) * AD * AH * AW + ai);
    if (cnt) val = val / __ldg(cnt + ai);
    if (apply_sigmoid) val = 1.f / (1.f + expf(-val));
    prob_sum[i] = overwrite ? val : prob_sum[i] + val;
  }
}

__global__ void __launch_bounds__(256) labels_finalize_kernel(const float* __restrict__ prob_sum, float count,
                                                              float thresh, const float* __restrict__ image, int IC,
                                                              uint8_t* __restrict__ onehot,
                                                              uint8_t* __restrict__ label, long long nvox,
                                                              int et_label) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox;
       i += (long long)gridDim.x * blockDim.x) {
    bool fg = true;
    if (image) {
      fg = false;
      for (int c = 0; c < IC; ++c) fg = fg || (__ldg(image + size_t(c) * nvox + i) != 0.f);
    }
    const bool tc = fg && (__fdiv_rn(__ldg(prob_sum + i), count) >= thresh);
    const bool wt = fg && (__fdiv_rn(__ldg(prob_sum + nvox + i), count) >= thresh);
    const bool et = fg && (__fdiv_rn(__ldg(prob_sum + 2 * nvox + i), count) >= thresh);
    if (onehot) {
      onehot[i] = tc;
      onehot[nvox + i] = wt;
      onehot[2 * nvox + i] = et;
    }
    if (label) {
      // ConvertToBratsClassesBasedOnMultiChannel (utils/transforms.py:185-191): ET first, then NCR/NET, then ED
      uint8_t l = 0;
      if (et) l = uint8_t(et_label);
      if (tc && !et) l = 1;
      if (wt && !tc) l = 2;
      label[i] = l;
    }
  }
}

static inline int grid1d(long long items, int threads) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 8;
  return int(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

static int parse_variant(const int* perm, const int* flip, Variant* tv) {
  int seen = 0;
  for (int j = 0; j < 3; ++j) {
    if (perm[j] < 0 || perm[j] > 2) return -1;
    seen |= 1 << perm[j];
    tv->perm[j] = perm[j];
    tv->flip[j] = flip[j] ? 1 : 0;
  }
  return seen == 7 ? 0 : -1;
}

}  // namespace b21

using namespace b21;

extern "C" int b21_pack_windows(const float* vol, int vc, int vd, int vh, int vw, void* out, int cpad, int nwin,
                                int d, int h, int w, const int* origins, const int* vol_index, const int* perm,
                                const int* flip, void* stream) {
  B21_CHECK_ARG(vol && out && origins && perm && flip, "pack_windows: null pointer");
  B21_CHECK_ARG(nwin >= 1 && nwin <= kMaxWin, "pack_windows: 1..%d windows per call (got %d)", kMaxWin, nwin);
  B21_CHECK_ARG(vc >= 1 && vc <= 8 && cpad % 8 == 0 && cpad >= 8, "pack_windows: <= 8 input channels, cpad multiple of 8");
  WinList wl;
  for (int b = 0; b < nwin; ++b) {
    for (int j = 0; j < 3; ++j) wl.org[b][j] = origins[b * 3 + j];
    wl.vol[b] = vol_index ? vol_index[b] : 0;
  }
  Variant tv;
  B21_CHECK_ARG(parse_variant(perm, flip, &tv) == 0, "pack_windows: perm must be a permutation of 0,1,2");
  const long long items = (long long)nwin * d * h * w;
  pack_windows_kernel<<<grid1d(items, 256), 256, 0, (cudaStream_t)stream>>>(
      vol, vc, vd, vh, vw, reinterpret_cast<__nv_bfloat16*>(out), cpad, nwin, d, h, w, wl, tv);
  B21_LAUNCH_CHECK("pack_windows_kernel");
  return B21_OK;
}

extern "C" int b21_blend_accumulate(const float* logits, float* acc, const float* prof_d, const float* prof_h,
                                    const float* prof_w, int nwin, int k, int d, int h, int w, int ad, int ah, int aw,
                                    const int* origins, void* stream) {
  B21_CHECK_ARG(acc && prof_d && prof_h && prof_w && origins, "blend_accumulate: null pointer");
  B21_CHECK_ARG(nwin >= 1 && k >= 1, "blend_accumulate: bad sizes");
  const long long per = (long long)k * d * h * w;
  for (int b = 0; b < nwin; ++b) {
    const int* o = origins + b * 3;
    B21_CHECK_ARG(o[0] >= 0 && o[1] >= 0 && o[2] >= 0 && o[0] + d <= ad && o[1] + h <= ah && o[2] + w <= aw,
                  "blend_accumulate: window %d [%d,%d,%d] leaves the accumulator", b, o[0], o[1], o[2]);
    blend_accumulate_kernel<<<grid1d(per, 256), 256, 0, (cudaStream_t)stream>>>(
        logits ? logits + size_t(b) * per : nullptr, acc, prof_d, prof_h, prof_w, k, d, h, w, ad, ah, aw, o[0], o[1],
        o[2]);
  }
  B21_LAUNCH_CHECK("blend_accumulate_kernel");
  return B21_OK;
}

extern "C" int b21_tta_accumulate(const float* acc, const float* cnt, float* prob_sum, int k, int ad, int ah, int aw,
                                  const int* pad_before, int vd, int vh, int vw, const int* perm, const int* flip,
                                  int apply_sigmoid, int overwrite, void* stream) {
  B21_CHECK_ARG(acc && prob_sum && perm && flip, "tta_accumulate: null pointer");
  Variant tv;
  B21_CHECK_ARG(parse_variant(perm, flip, &tv) == 0, "tta_accumulate: perm must be a permutation of 0,1,2");
  const int vdim[3] = {vd, vh, vw};
  const int adim[3] = {ad, ah, aw};
  int pad[3] = {0, 0, 0};
  if (pad_before) for (int j = 0; j < 3; ++j) pad[j] = pad_before[j];
  for (int j = 0; j < 3; ++j)
    B21_CHECK_ARG(vdim[j] + pad[tv.perm[j]] <= adim[tv.perm[j]], "tta_accumulate: augmented frame too small on axis %d", j);
  const long long items = (long long)k * vd * vh * vw;
  tta_accumulate_kernel<<<grid1d(items, 256), 256, 0, (cudaStream_t)stream>>>(
      acc, cnt, prob_sum, k, ad, ah, aw, pad[0], pad[1], pad[2], vd, vh, vw, tv, apply_sigmoid, overwrite);
  B21_LAUNCH_CHECK("tta_accumulate_kernel");
  return B21_OK;
}

extern "C" int b21_labels_finalize(const float* prob_sum, float count, float thresh, const float* image,
                                   int image_channels, uint8_t* onehot, uint8_t* label, long long nvox, int et_label,
                                   void* stream) {
  B21_CHECK_ARG(prob_sum && (onehot || label) && nvox > 0, "labels_finalize: null pointer");
  labels_finalize_kernel<<<grid1d(nvox, 256), 256, 0, (cudaStream_t)stream>>>(prob_sum, count, thresh, image,
                                                                              image_channels, onehot, label, nvox,
                                                                              et_label);
  B21_LAUNCH_CHECK("labels_finalize_kernel");
  return B21_OK;
}
