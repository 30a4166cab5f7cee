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

// One launch per window BATCH.  Windows of a batch overlap, and the reference adds them one after the other
// (`out[slc] += w * p`, utils/inferers.py:149-151): each thread owns V consecutive x voxels of the accumulator inside the
// batch's bounding box and walks the windows in order, so the fp32 sum is formed in exactly the reference's order
// while the accumulator is read and written once per batch (not once per window) and the logits are read once.
// wfloor > 0: MONAI clamps the importance map at its smallest non-zero value (only reachable when the truncated
// Gaussian has zeros inside the window, i.e. sigma_scale < 1/8).
struct BlendBox {
  int lo[3];
  int size[3];
};

template <int V>
__global__ void __launch_bounds__(256) blend_accumulate_kernel(const float* __restrict__ logits, float* acc,
                                                               const float* __restrict__ pd,
                                                               const float* __restrict__ ph,
                                                               const float* __restrict__ pw, int nwin, int K, int d,
                                                               int h, int w, int AD, int AH, int AW, WinList wl,
                                                               BlendBox bb, float wfloor) {
  const int bwv = bb.size[2] / V;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= bb.size[1] * bwv) return;
  const int xv = row % bwv, yb = row / bwv;
  const int z = bb.lo[0] + blockIdx.y, y = bb.lo[1] + yb, x = bb.lo[2] + xv * V;
  const int k = blockIdx.z;
  float* ap = acc + ((size_t(k) * AD + z) * AH + y) * AW + x;
  float v[V];
  bool touched = false;
  for (int b = 0; b < nwin; ++b) {
    const int lz = z - wl.org[b][0], ly = y - wl.org[b][1], lx = x - wl.org[b][2];
    if (lz < 0 || lz >= d || ly < 0 || ly >= h || lx < 0 || lx >= w) continue;  // (lx is a multiple of V: all or none)
    if (!touched) {
      if (V == 4) {
        const float4 t = *reinterpret_cast<const float4*>(ap);
        v[0] = t.x; v[1 % V] = t.y; v[2 % V] = t.z; v[3 % V] = t.w;
      } else {
        v[0] = ap[0];
      }
      touched = true;
    }
    const float wzy = __fmul_rn(__ldg(pd + lz), __ldg(ph + ly));
    float lg[V];
    if (logits) {
      const float* lp = logits + (((size_t(b) * K + k) * d + lz) * h + ly) * w + lx;
      if (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(lp));
        lg[0] = t.x; lg[1 % V] = t.y; lg[2 % V] = t.z; lg[3 % V] = t.w;
      } else {
        lg[0] = __ldg(lp);
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float wt = __fmul_rn(wzy, __ldg(pw + lx + i));
      if (wfloor > 0.f) wt = fmaxf(wt, wfloor);
      // separate multiply and add (no FMA contraction): the reference does `out += importance_map * seg_prob`
      v[i] = __fadd_rn(v[i], logits ? __fmul_rn(wt, lg[i]) : wt);
    }
  }
  if (touched) {
    if (V == 4) *reinterpret_cast<float4*>(ap) = make_float4(v[0], v[1 % V], v[2 % V], v[3 % V]);
    else ap[0] = v[0];
  }
}

// flip-only variants (perm = identity) with x extents that are multiples of 4: float4 along x, mirrored reads stay
// coalesced.  prob_sum[k][s] (+)= sigmoid(acc[k][a(s) + pad] / cnt[a(s) + pad]).
__global__ void __launch_bounds__(256) tta_accumulate_vec4_kernel(const float* __restrict__ acc,
                                                                  const float* __restrict__ cnt, float* prob_sum,
                                                                  int AD, int AH, int AW, int p0, int p1, int p2,
                                                                  int VD, int VH, int VW, int f0, int f1, int f2,
                                                                  int apply_sigmoid, int overwrite) {
  const int vw4 = VW >> 2;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= VH * vw4) return;
  const int x4 = row % vw4, s1 = row / vw4, s0 = blockIdx.y, k = blockIdx.z;
  const int a0 = (f0 ? VD - 1 - s0 : s0) + p0, a1 = (f1 ? VH - 1 - s1 : s1) + p1;
  const int a2 = (f2 ? VW - 4 - 4 * x4 : 4 * x4) + p2;
  const size_t ai = (size_t(a0) * AH + a1) * AW + a2;
  float4 a = __ldg(reinterpret_cast<const float4*>(acc + size_t(k) * AD * AH * AW + ai));
  if (cnt) {
    const float4 c = __ldg(reinterpret_cast<const float4*>(cnt + ai));
    a.x = a.x / c.x; a.y = a.y / c.y; a.z = a.z / c.z; a.w = a.w / c.w;
  }
  if (f2) {
    float t = a.x; a.x = a.w; a.w = t;
    t = a.y; a.y = a.z; a.z = t;
  }
  if (apply_sigmoid) {
    a.x = 1.f / (1.f + expf(-a.x)); a.y = 1.f / (1.f + expf(-a.y));
    a.z = 1.f / (1.f + expf(-a.z)); a.w = 1.f / (1.f + expf(-a.w));
  }
  float4* dst = reinterpret_cast<float4*>(prob_sum + ((size_t(k) * VD + s0) * VH + s1) * VW + 4 * x4);
  if (!overwrite) {
    const float4 o = *dst;
    a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
  }
  *dst = a;
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

// remove_background_voxels (utils/transforms.py:536-550) on a label map: the reference applies it AFTER the post
// transforms (learning/engine.py:249-256), so with component cleaning / rare-label replacement it cannot be folded into
// labels_finalize.
__global__ void __launch_bounds__(256) mask_background_kernel(uint8_t* __restrict__ label,
                                                              const float* __restrict__ image, int IC, int planes,
                                                              long long nvox) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox;
       i += (long long)gridDim.x * blockDim.x) {
    bool fg = false;
    for (int c = 0; c < IC; ++c) fg = fg || (__ldg(image + size_t(c) * nvox + i) != 0.f);
    if (!fg)
      for (int p = 0; p < planes; ++p) label[size_t(p) * nvox + i] = 0;
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
                                    const int* origins, float wfloor, void* stream) {
  B21_CHECK_ARG(acc && prof_d && prof_h && prof_w && origins, "blend_accumulate: null pointer");
  B21_CHECK_ARG(nwin >= 1 && nwin <= kMaxWin && k >= 1, "blend_accumulate: 1..%d windows per call (got %d)", kMaxWin, nwin);
  WinList wl;
  BlendBox bb;
  int hi[3] = {0, 0, 0};
  const int ext[3] = {d, h, w};
  bool vec = (w % 4 == 0) && (aw % 4 == 0) && ((reinterpret_cast<uintptr_t>(acc) & 15) == 0) &&
             ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  for (int b = 0; b < nwin; ++b) {
    const int* o = origins + b * 3;
    B21_CHECK_ARG(o[0] >= 0 && o[1] >= 0 && o[2] >= 0 && o[0] + d <= ad && o[1] + h <= ah && o[2] + w <= aw,
                  "blend_accumulate: window %d [%d,%d,%d] leaves the accumulator", b, o[0], o[1], o[2]);
    for (int j = 0; j < 3; ++j) {
      wl.org[b][j] = o[j];
      bb.lo[j] = (b == 0 || o[j] < bb.lo[j]) ? o[j] : bb.lo[j];
      hi[j] = (b == 0 || o[j] + ext[j] > hi[j]) ? o[j] + ext[j] : hi[j];
    }
    wl.vol[b] = 0;
    vec = vec && (o[2] % 4 == 0);
  }
  for (int j = 0; j < 3; ++j) bb.size[j] = hi[j] - bb.lo[j];
  B21_CHECK_ARG(bb.size[0] <= 65535 && k <= 65535, "blend_accumulate: accumulator too deep for the launch grid");
  const int v = vec ? 4 : 1;
  const int rows = bb.size[1] * (bb.size[2] / v);
  dim3 grid((rows + 255) / 256, bb.size[0], k);
  if (vec)
    blend_accumulate_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, acc, prof_d, prof_h, prof_w, nwin, k, d, h,
                                                                       w, ad, ah, aw, wl, bb, wfloor);
  else
    blend_accumulate_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, acc, prof_d, prof_h, prof_w, nwin, k, d, h,
                                                                       w, ad, ah, aw, wl, bb, wfloor);
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
  const bool ident = tv.perm[0] == 0 && tv.perm[1] == 1 && tv.perm[2] == 2;
  if (ident && vw % 4 == 0 && aw % 4 == 0 && pad[2] % 4 == 0 && vd <= 65535 && k <= 65535 &&
      ((reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(cnt) | reinterpret_cast<uintptr_t>(prob_sum)) & 15) == 0) {
    dim3 grid((vh * (vw / 4) + 255) / 256, vd, k);
    tta_accumulate_vec4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(acc, cnt, prob_sum, ad, ah, aw, pad[0], pad[1],
                                                                      pad[2], vd, vh, vw, tv.flip[0], tv.flip[1],
                                                                      tv.flip[2], apply_sigmoid, overwrite);
    B21_LAUNCH_CHECK("tta_accumulate_vec4_kernel");
    return B21_OK;
  }
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

extern "C" int b21_mask_background(uint8_t* label, int planes, const float* image, int image_channels, long long nvox,
                                   void* stream) {
  B21_CHECK_ARG(label && image && planes >= 1 && image_channels >= 1 && nvox > 0, "mask_background: bad arguments");
  mask_background_kernel<<<grid1d(nvox, 256), 256, 0, (cudaStream_t)stream>>>(label, image, image_channels, planes, nvox);
  B21_LAUNCH_CHECK("mask_background_kernel");
  return B21_OK;
}
