// Static description of the NerfMLP GEMM chain (Ref-NeRF configuration of the reference:
// configs/blender_refnerf.gin:34-52, models.py:497-531) and of the packed-weight blob.
//
// GEMM layers, forward order:
//   S0..S7  spatial net   (K: 96->pad128 | 256 | 256 | 256 | 256 | 256+96(pad128) | 256 | 256) -> 256
//   H       heads         256 -> 139 (pad 144): rows 0..127 bottleneck, 128 raw_density, 129..131 grad_pred,
//                          132 raw_roughness, 133..135 raw_rgb_diffuse, 136..138 raw_tint
//   V0..V7  view net      (K: 201->pad256 | 256 x4 | 256+201(pad256) | 256 | 256) -> 256
//   C       rgb head      256 -> 3 (pad 16)
// Skip concat order is [activation, inputs] (models.py:579-580, 693-694).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/refnerf_b200.h"

namespace rn {

constexpr int kNumLayers = 18;
constexpr int kLayerH = 8;
constexpr int kLayerV0 = 9;
constexpr int kLayerC = 17;
constexpr int kWidth = 256;
constexpr int kEncReal = 96, kEncPad = 128;
constexpr int kViewReal = 201, kViewPad = 256;
constexpr int kBottleneck = 128;
constexpr int kHeadsReal = 139, kHeadsPad = 144;
constexpr int kHeadScalars = 11;  // density, grad_pred(3), roughness, diffuse(3), tint(3)

struct LayerDef {
  int n_real, n_pad;    // outputs
  int k1_real, k1_pad;  // first K source
  int k2_real, k2_pad;  // second K source (skip connection), 0 if none
  int nt_pad;           // padded N used as the K extent of the transposed (dgrad) operand, multiple of 64
  int k_tot() const { return k1_pad + k2_pad; }
};

inline LayerDef layer_def(int l) {
  if (l == 0) return {256, 256, kEncReal, kEncPad, 0, 0, 256};
  if (l == 5) return {256, 256, 256, 256, kEncReal, kEncPad, 256};
  if (l < 8) return {256, 256, 256, 256, 0, 0, 256};
  if (l == kLayerH) return {kHeadsReal, kHeadsPad, 256, 256, 0, 0, 192};
  if (l == kLayerV0) return {256, 256, kViewReal, kViewPad, 0, 0, 256};
  if (l == kLayerV0 + 5) return {256, 256, 256, 256, kViewReal, kViewPad, 256};
  if (l < kLayerC) return {256, 256, 256, 256, 0, 0, 256};
  return {3, 16, 256, 256, 0, 0, 64};
}

inline int elem_bytes(int prec) { return prec == RN_PREC_FP32 ? 4 : 2; }
inline int planes(int prec) { return prec == RN_PREC_BF16X3 ? 2 : 1; }
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Offsets into the packed blob.  Per layer: Wf [n_pad, k_tot] (planes), Wt [k_tot, nt_pad] (planes),
// bias f32 [n_pad].  After all layers: wd f32 [256] (raw_density weight row, for the normals pass), wcat (below).
struct PackedLayout {
  size_t wf[kNumLayers], wt[kNumLayers], bias[kNumLayers];
  size_t wf_plane[kNumLayers], wt_plane[kNumLayers];  // byte stride between hi and lo planes
  size_t wd;
  size_t wcat;   // bf16 [256, 512]: [Wt(V0) | Wt(V5)[256:512, :]] -- d v0 = [dY0 | dY5] * wcat^T in ONE dgrad op (fused chains)
  size_t wcat_plane;   // byte stride between its hi and lo planes (split-bf16)
  size_t total;
};

inline PackedLayout packed_layout(int prec) {
  PackedLayout p;
  size_t off = 0;
  const int eb = elem_bytes(prec), pl = planes(prec);
  for (int l = 0; l < kNumLayers; ++l) {
    LayerDef d = layer_def(l);
    p.wf[l] = off;
    p.wf_plane[l] = align256((size_t)d.n_pad * d.k_tot() * eb);
    off += p.wf_plane[l] * pl;
    p.wt[l] = off;
    p.wt_plane[l] = align256((size_t)d.k_tot() * d.nt_pad * eb);
    off += p.wt_plane[l] * pl;
    p.bias[l] = off;
    off += align256((size_t)d.n_pad * 4);
  }
  p.wd = off;
  off += align256(256 * 4);
  p.wcat = off;
  p.wcat_plane = align256((size_t)256 * 512 * 2);
  off += p.wcat_plane * pl;
  p.total = off;
  return p;
}

// Parameter index helpers (order documented in include/refnerf_b200.h)
inline int param_weight_index_spatial(int i) { return 2 * i; }
constexpr int kParamDensity = 16, kParamGradPred = 18, kParamRoughness = 20, kParamDiffuse = 22, kParamTint = 24,
              kParamBottleneck = 26, kParamView0 = 28, kParamRgb = 44;

}  // namespace rn
