// Host-side sequencing of the NerfMLP level pass (models.py:533-750) over row chunks, and the C ABI.
// No allocation, no host synchronisation: everything is enqueued on the caller's stream and all
// scratch comes from the caller-provided workspace.
#include <stdio.h>
#include <string.h>

#include "gemm.cuh"
#include "layers.h"
#include "pointwise.cuh"

static thread_local char g_err[512] = "";

int rn_set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int rn_set_cuda_error(cudaError_t e, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e), file, line);
  return RN_ERR_CUDA;
}

// ---- instrumentation: launch counter and per-class CUDA-event timing --------------------------
#include <atomic>
#include <mutex>
#include <vector>
static std::atomic<long long> g_launches{0};
void rn_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {
struct ProfClass {
  std::vector<cudaEvent_t> ev;  // pairs: start, stop
  size_t used = 0;
  double flops = 0.0;
  double exec_flops = 0.0;
};
bool g_prof_on = false;
ProfClass g_prof[RN_PROF_NUM];
std::mutex g_prof_mu;
}  // namespace

void rn_prof_begin(int cls, cudaStream_t st, double algo_flops, double exec_flops) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfClass& p = g_prof[cls];
  if (p.used + 2 > p.ev.size()) {
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      p.ev.push_back(e);
    }
  }
  cudaEventRecord(p.ev[p.used], st);
  p.flops += algo_flops;
  p.exec_flops += exec_flops;
}
void rn_prof_end(int cls, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfClass& p = g_prof[cls];
  cudaEventRecord(p.ev[p.used + 1], st);
  p.used += 2;
}

extern "C" int64_t rn_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" int rn_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  return RN_OK;
}
extern "C" int rn_prof_summary2(int cls, int64_t* launches, double* total_ms, double* algo_flops, double* exec_flops);
extern "C" int rn_prof_summary(int cls, int64_t* launches, double* total_ms, double* algo_flops) {
  return rn_prof_summary2(cls, launches, total_ms, algo_flops, nullptr);
}
extern "C" int rn_prof_summary2(int cls, int64_t* launches, double* total_ms, double* algo_flops, double* exec_flops) {
  if (cls < 0 || cls >= RN_PROF_NUM) return rn_set_error(RN_ERR_ARG, "rn_prof_summary: bad class");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfClass& p = g_prof[cls];
  double ms = 0.0;
  for (size_t i = 0; i + 1 < p.used; i += 2) {
    cudaError_t e = cudaEventSynchronize(p.ev[i + 1]);
    if (e != cudaSuccess) return rn_set_cuda_error(e, __FILE__, __LINE__);
    float t = 0.f;
    cudaEventElapsedTime(&t, p.ev[i], p.ev[i + 1]);
    ms += t;
  }
  if (launches) *launches = (int64_t)(p.used / 2);
  if (total_ms) *total_ms = ms;
  if (algo_flops) *algo_flops = p.flops;
  if (exec_flops) *exec_flops = p.exec_flops;
  p.used = 0;
  p.flops = 0.0;
  p.exec_flops = 0.0;
  return RN_OK;
}


namespace rn {
namespace {

#define RN_TRY(expr)            \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != RN_OK) return rc__; \
  } while (0)

const char* kParamNames[RN_MLP_NUM_PARAMS] = {
    "spatial_net.0.weight", "spatial_net.0.bias", "spatial_net.1.weight", "spatial_net.1.bias",
    "spatial_net.2.weight", "spatial_net.2.bias", "spatial_net.3.weight", "spatial_net.3.bias",
    "spatial_net.4.weight", "spatial_net.4.bias", "spatial_net.5.weight", "spatial_net.5.bias",
    "spatial_net.6.weight", "spatial_net.6.bias", "spatial_net.7.weight", "spatial_net.7.bias",
    "raw_density.weight", "raw_density.bias", "grad_pred.weight", "grad_pred.bias",
    "raw_roughness.weight", "raw_roughness.bias", "raw_rgb_diffuse.weight", "raw_rgb_diffuse.bias",
    "raw_tint.weight", "raw_tint.bias", "bottleneck.weight", "bottleneck.bias",
    "viewdir_mlp.0.weight", "viewdir_mlp.0.bias", "viewdir_mlp.1.weight", "viewdir_mlp.1.bias",
    "viewdir_mlp.2.weight", "viewdir_mlp.2.bias", "viewdir_mlp.3.weight", "viewdir_mlp.3.bias",
    "viewdir_mlp.4.weight", "viewdir_mlp.4.bias", "viewdir_mlp.5.weight", "viewdir_mlp.5.bias",
    "viewdir_mlp.6.weight", "viewdir_mlp.6.bias", "viewdir_mlp.7.weight", "viewdir_mlp.7.bias",
    "rgb.weight", "rgb.bias"};

// head rows inside the packed H layer: {param index of weight, first packed row, rows}
struct HeadSeg { int param, row0, rows; };
const HeadSeg kHeadSegs[6] = {{kParamBottleneck, 0, 128}, {kParamDensity, 128, 1},  {kParamGradPred, 129, 3},
                              {kParamRoughness, 132, 1},  {kParamDiffuse, 133, 3}, {kParamTint, 136, 3}};

int layer_param(int l) {  // weight param index of a non-head layer
  if (l < 8) return 2 * l;
  if (l >= kLayerV0 && l < kLayerC) return kParamView0 + 2 * (l - kLayerV0);
  return kParamRgb;
}

int64_t param_numel(int i) {
  const bool is_bias = i & 1;
  const int w = i & ~1;
  int n, k;
  if (w < 16) { LayerDef d = layer_def(w / 2); n = d.n_real; k = d.k1_real + d.k2_real; }
  else if (w == kParamDensity || w == kParamRoughness) { n = 1; k = 256; }
  else if (w == kParamGradPred || w == kParamDiffuse || w == kParamTint) { n = 3; k = 256; }
  else if (w == kParamBottleneck) { n = 128; k = 256; }
  else if (w < kParamRgb) { LayerDef d = layer_def(kLayerV0 + (w - kParamView0) / 2); n = d.n_real; k = d.k1_real + d.k2_real; }
  else { n = 3; k = 256; }
  return is_bias ? n : (int64_t)n * k;
}

struct Carver {
  uint8_t* base;
  size_t off = 0;
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += align256(bytes);
    return p;
  }
  ActBuf act(int prec, int64_t rows, int ld) {
    ActBuf b;
    b.ld = ld;
    b.hi = take((size_t)rows * ld * elem_bytes(prec));
    b.lo = prec == RN_PREC_BF16X3 ? take((size_t)rows * ld * 2) : nullptr;
    return b;
  }
};

struct Workspace {
  ActBuf x0, v0, sp[8], vw[8], g[2], gs[8], dheads, d_bott, d_scal, d_rgb_raw;
  uint32_t *msp[8], *mvw[8];   // ReLU bits of the hidden activations (fused chains), valid when nsp / nvw == 8
  float *heads_raw, *rgb_raw, *gx0, *gx0b, *dv0f, *dcolor;
  float *stage16, *scal;   // fp16 mode: f32 staging of the 16-wide chain seeds; amax / scale scalars (pointwise.cu)
  float* gW[kNumLayers];
  float* gB[kNumLayers];
  float* wg_partial;           // deterministic wgrad: per-CTA partial tiles, num_sms x (256 x 257) floats
  int nsp, nvw;
  size_t bytes;
  ActBuf& a(int i) { return sp[(i - 1) % nsp]; }   // spatial activation a_i = relu(y_{i-1}), i = 1..8
  ActBuf& b(int i) { return vw[(i - 1) % nvw]; }   // view activation
  uint32_t* ma(int i) { return nsp == 8 ? msp[i - 1] : nullptr; }
  uint32_t* mb(int i) { return nvw == 8 ? mvw[i - 1] : nullptr; }
};

// Activations a training forward keeps for its backward (no recompute): chain inputs, every hidden activation
// and the raw head outputs, for ALL rows of the call (chunks index into it by row).
struct Saved {
  ActBuf x0, v0, sp[8], vw[8];
  uint32_t *msp[8], *mvw[8];   // ReLU bits of sp / vw, [rows, 8] words (fused chains)
  float *heads_raw, *rgb_raw;
  size_t bytes;
};
// `hprec`: format of the hidden activations (the fused split-bf16 chains keep only their hi plane: wgrad runs on
// bf16 operands and the dgrad chains need the ReLU bits only)
Saved carve_saved(void* base, int prec, int hprec, int64_t rows) {
  Saved s;
  Carver c{reinterpret_cast<uint8_t*>(base)};
  s.x0 = c.act(prec, rows, kEncPad);
  s.v0 = c.act(prec, rows, kViewPad);
  for (int i = 0; i < 8; ++i) s.sp[i] = c.act(hprec, rows, 256);
  for (int i = 0; i < 8; ++i) s.vw[i] = c.act(hprec, rows, 256);
  for (int i = 0; i < 8; ++i) s.msp[i] = (uint32_t*)c.take(relu_bits_bytes(rows));
  for (int i = 0; i < 8; ++i) s.mvw[i] = (uint32_t*)c.take(relu_bits_bytes(rows));
  s.heads_raw = (float*)c.take((size_t)rows * 16 * 4);
  s.rgb_raw = (float*)c.take((size_t)rows * 4 * 4);
  s.bytes = c.off;
  return s;
}
ActBuf rows_from(ActBuf b, int prec, int64_t row0) {
  if (b.hi) b.hi = reinterpret_cast<uint8_t*>(b.hi) + (size_t)row0 * b.ld * elem_bytes(prec);
  if (b.lo) b.lo = reinterpret_cast<uint8_t*>(b.lo) + (size_t)row0 * b.ld * 2;
  return b;
}

// mode 0: eval forward, 1: training forward (+normals pass), 2: backward.  With `external` the chain inputs,
// hidden activations and raw head outputs live in a Saved region (see use_saved) instead of the workspace.
// `x3chain`: fused split-bf16 chains -- chain inputs in two planes, hidden activations and every gradient tile in
// one bf16 plane
Workspace carve(void* base, int prec, int64_t rc, int mode, bool external = false, bool x3chain = false, bool det = false) {
  Workspace w;
  w.wg_partial = nullptr;
  Carver c{reinterpret_cast<uint8_t*>(base)};
  const int hprec = x3chain ? RN_PREC_BF16 : prec;
  w.nsp = (mode == 0 && !external) ? 2 : 8;
  w.nvw = (mode == 2 || external) ? 8 : 2;
  if (!external) {
    w.x0 = c.act(prec, rc, kEncPad);
    w.v0 = c.act(prec, rc, kViewPad);
    for (int i = 0; i < w.nsp; ++i) w.sp[i] = c.act(hprec, rc, 256);
    for (int i = 0; i < w.nvw; ++i) w.vw[i] = c.act(hprec, rc, 256);
    for (int i = 0; i < 8; ++i) w.msp[i] = w.nsp == 8 ? (uint32_t*)c.take(relu_bits_bytes(rc)) : nullptr;
    for (int i = 0; i < 8; ++i) w.mvw[i] = w.nvw == 8 ? (uint32_t*)c.take(relu_bits_bytes(rc)) : nullptr;
    w.heads_raw = (float*)c.take((size_t)rc * 16 * 4);
    w.rgb_raw = (float*)c.take((size_t)rc * 4 * 4);
  }
  if (mode == 1 && !x3chain) {   // (the fused chains keep the running gradient tile on chip)
    w.g[0] = c.act(prec, rc, 256);
    w.g[1] = c.act(prec, rc, 256);
  }
  if (mode >= 1) w.gx0 = (float*)c.take((size_t)rc * 128 * 4);
  w.gx0b = (mode == 1 && x3chain) ? (float*)c.take((size_t)rc * 128 * 4) : nullptr;
  if (mode == 2) {
    for (int i = 0; i < 8; ++i) w.gs[i] = c.act(hprec, rc, 256);   // dY of the 8 layers of the net being back-propagated
    w.g[0] = w.gs[0];
    w.g[1] = w.gs[1];
    w.dv0f = (float*)c.take((size_t)rc * 256 * 4);
    w.dcolor = (float*)c.take((size_t)rc * 8 * 4);
    // d heads = [d bottleneck (128) | d scalar heads (11, padded to 16) | unused]: one 192-wide buffer so the
    // heads dgrad reads it as a single K extent
    w.dheads = c.act(hprec, rc, 192);
    w.d_bott = w.dheads;
    w.d_scal = w.dheads;
    w.d_scal.hi = reinterpret_cast<uint8_t*>(w.dheads.hi) + 128 * elem_bytes(hprec);
    if (w.dheads.lo) w.d_scal.lo = reinterpret_cast<uint8_t*>(w.dheads.lo) + 128 * 2;
    w.d_rgb_raw = c.act(hprec, rc, 16);
    w.stage16 = (float*)c.take((size_t)rc * 16 * 4);
    w.scal = (float*)c.take(256);
    for (int l = 0; l < kNumLayers; ++l) {
      LayerDef d = layer_def(l);
      w.gW[l] = (float*)c.take((size_t)d.n_pad * d.k_tot() * 4);
      w.gB[l] = (float*)c.take((size_t)d.n_pad * 4);
    }
    if (det) w.wg_partial = (float*)c.take((size_t)256 * (256 * 257) * 4);   // up to 256 CTAs
  }
  w.bytes = c.off;
  return w;
}
// point the chunk's activation buffers at rows [row0, ..) of the saved region
void use_saved(Workspace& w, const Saved& s, int prec, int hprec, int64_t row0) {
  w.x0 = rows_from(s.x0, prec, row0);
  w.v0 = rows_from(s.v0, prec, row0);
  for (int i = 0; i < 8; ++i) {
    w.sp[i] = rows_from(s.sp[i], hprec, row0);
    w.vw[i] = rows_from(s.vw[i], hprec, row0);
    w.msp[i] = s.msp[i] + (size_t)row0 * 8;   // row0 is a multiple of 128: whole 32-row groups
    w.mvw[i] = s.mvw[i] + (size_t)row0 * 8;
  }
  w.heads_raw = s.heads_raw + (size_t)row0 * 16;
  w.rgb_raw = s.rgb_raw + (size_t)row0 * 4;
}

inline bool cfg_x3chain(const RnMlpConfig* cfg) { return cfg->prec == RN_PREC_BF16X3 && cfg->gemm_impl == 0; }
inline int cfg_hprec(const RnMlpConfig* cfg) { return cfg_x3chain(cfg) ? RN_PREC_BF16 : cfg->prec; }

struct Packed {
  const uint8_t* base;
  PackedLayout lay;
  int prec;
  const void* wf_hi(int l) const { return base + lay.wf[l]; }
  const void* wf_lo(int l) const { return prec == RN_PREC_BF16X3 ? base + lay.wf[l] + lay.wf_plane[l] : nullptr; }
  // transposed operand, starting at row `row` (rows are input features)
  const void* wt_hi(int l, int row) const {
    return base + lay.wt[l] + (size_t)row * layer_def(l).nt_pad * elem_bytes(prec);
  }
  const void* wt_lo(int l, int row) const {
    return prec == RN_PREC_BF16X3 ? base + lay.wt[l] + lay.wt_plane[l] + (size_t)row * layer_def(l).nt_pad * 2 : nullptr;
  }
  const float* bias(int l) const { return reinterpret_cast<const float*>(base + lay.bias[l]); }
  const float* wd() const { return reinterpret_cast<const float*>(base + lay.wd); }
  const void* wcat() const { return base + lay.wcat; }
  const void* wcat_lo() const { return prec == RN_PREC_BF16X3 ? base + lay.wcat + lay.wcat_plane : nullptr; }
};

struct Ctx {
  const RnMlpConfig* cfg;
  Packed pk;
  const float *tdist, *origins, *dirs, *viewdirs, *radii;
  int s;
  cudaStream_t st;
  MlpScalars sc;
  int impl;
  bool chain = false;  // bf16 / fp16 / split-bf16 with gemm_impl 0: fused chains
  bool f16 = false;    // fp16 mode: weights, activations and (dynamically scaled) gradient tiles are fp16
  bool x3 = false;     // fused split-bf16: forward and normals chains on chain_x3.cu (3 MMAs per K step), loss backward
                       // with one-plane bf16 gradient tiles against split weights (2 MMAs), bf16 wgrad
  int chain_impl = 0;  // ChainArgs::impl of the forward / normals chains
  int hprec() const { return x3 ? RN_PREC_BF16 : cfg->prec; }   // format of hidden activations and gradient tiles
  // per-layer tcgen05 wgrad kernels also produce the bias gradient (no separate column-sum pass over dY)
  bool fused_bias() const { return !chain && impl == 0 && cfg->prec != RN_PREC_FP32; }
  bool algo = true;  // launches carry algorithmic FLOPs (false while recomputing activations in backward)
};

// y = act(W x + b) for chain layer l; input a1 (+ skip input a2)
int fwd_layer(const Ctx& c, int l, int64_t rows, ActBuf a1, ActBuf a2, GemmEpilogue epi) {
  LayerDef d = layer_def(l);
  GemmArgs g;
  g.prec = c.cfg->prec;
  g.impl = c.impl;
  g.m = rows;
  g.n = d.n_pad;
  g.a1 = a1; g.k1 = d.k1_pad; g.a1_valid = d.k1_pad;
  g.a2 = a2; g.k2 = d.k2_pad; g.a2_valid = d.k2_pad;
  g.b_hi = c.pk.wf_hi(l); g.b_lo = c.pk.wf_lo(l); g.b_ld = d.k_tot();
  epi.bias = c.pk.bias(l);
  g.epi = epi;
  g.algo_flops = c.algo ? 2.0 * (double)rows * d.n_real * (d.k1_real + d.k2_real) : 0.0;
  return launch_gemm(g, c.st);
}

// dX[:, row0 : row0+n] = dY W[:, row0 : row0+n]  (dY given as up to two K sources)
int dgrad_layer(const Ctx& c, int l, int64_t rows, ActBuf dy1, int k1, int dy1_valid, ActBuf dy2, int k2, int dy2_valid,
                int in_row0, int n, GemmEpilogue epi) {
  GemmArgs g;
  g.prec = c.cfg->prec;
  g.impl = c.impl;
  g.m = rows;
  g.n = n;
  g.a1 = dy1; g.k1 = k1; g.a1_valid = dy1_valid;
  g.a2 = dy2; g.k2 = k2; g.a2_valid = dy2_valid;
  g.b_hi = c.pk.wt_hi(l, in_row0); g.b_lo = c.pk.wt_lo(l, in_row0); g.b_ld = layer_def(l).nt_pad;
  g.epi = epi;
  {
    LayerDef d = layer_def(l);
    g.algo_flops = c.algo ? 2.0 * (double)rows * d.n_real * (in_row0 == 0 ? d.k1_real : d.k2_real) : 0.0;
  }
  return launch_gemm(g, c.st);
}

GemmEpilogue epi_act(ActBuf out, int cols, int relu) {
  GemmEpilogue e;
  e.out = out; e.out_cols = cols; e.relu = relu;
  return e;
}
GemmEpilogue epi_masked(ActBuf out, ActBuf mask) {
  GemmEpilogue e;
  e.out = out; e.out_cols = 256; e.mask = mask;
  return e;
}
GemmEpilogue epi_f32(float* p, int ld, int cols, int accum) {
  GemmEpilogue e;
  e.f32 = p; e.f32_ld = ld; e.f32_col0 = 0; e.f32_cols = cols; e.f32_accum = accum;
  return e;
}

double op_flops(int64_t rows, int l, bool skip_part, bool both) {
  LayerDef d = layer_def(l);
  const int k = both ? d.k1_real + d.k2_real : (skip_part ? d.k2_real : d.k1_real);
  return 2.0 * (double)rows * d.n_real * k;
}

// layers [l0, l0+8) + the head layer `lh` as one fused launch; `keep` != null saves every hidden activation
int chain_layers(const Ctx& c, int l0, int lh, int64_t rows, ActBuf in, int in_cols, Workspace* keep, bool spatial,
                 GemmEpilogue final_epi) {
  ChainArgs a;
  a.impl = c.chain_impl;
  a.m = rows;
  a.in = in;
  a.in_cols = in_cols;
  // the real input width (96 IPE columns of 128, 201 view-net inputs of 256): the buffer is zero beyond it, the TMA loads
  // zero-fill there and chain_x3t.cu skips the K steps that are padding only
  a.in_valid = layer_def(l0).k1_real < in_cols ? layer_def(l0).k1_real : in_cols;
  a.act_f16 = a.w_f16 = c.f16;
  a.num_ops = 9;
  double flops = 0.0;
  for (int i = 0; i < 9; ++i) {
    const int l = i < 8 ? l0 + i : lh;
    LayerDef d = layer_def(l);
    ChainOpArgs& L = a.op[i];
    L.n = d.n_pad;
    L.kb_act = i == 0 ? 0 : 4;
    L.kb_in = (i == 0 ? d.k1_pad : d.k2_pad) / 64;
    L.kind = i < 8 ? 0 : 1;
    L.mode = 0;
    L.gepi = 0;
    L.w = c.pk.wf_hi(l);
    L.w_lo = c.pk.wf_lo(l);
    L.w_ld = d.k_tot();
    L.bias = c.pk.bias(l);
    L.save_hi = (keep && i < 8) ? (spatial ? keep->a(i + 1).hi : keep->b(i + 1).hi) : nullptr;
    L.save_bits = (keep && i < 8) ? (spatial ? keep->ma(i + 1) : keep->mb(i + 1)) : nullptr;
    flops += op_flops(rows, l, false, true);
  }
  final_epi.bias = c.pk.bias(lh);
  a.gepi[0] = final_epi;
  if (c.x3 && lh == kLayerH) {
    // chain_x3.cu: global ops are at most 128 wide -> the heads layer runs as two ops over the same activation tile:
    // the bottleneck (128 columns, leaves in activation format) and the 11 scalar heads (16 columns, f32)
    // The scalar heads go first: the bottleneck's epilogue stages its output in the activation tile, which the other
    // op still has to read.
    LayerDef d = layer_def(lh);
    ChainOpArgs& S = a.op[8];
    ChainOpArgs& B = a.op[9];
    B = S;
    B.n = kBottleneck;
    S.n = kHeadsPad - kBottleneck;
    S.gepi = 1;
    S.w = reinterpret_cast<const uint8_t*>(c.pk.wf_hi(lh)) + (size_t)kBottleneck * d.k_tot() * 2;
    S.w_lo = reinterpret_cast<const uint8_t*>(c.pk.wf_lo(lh)) + (size_t)kBottleneck * d.k_tot() * 2;
    GemmEpilogue eb = final_epi, es = final_epi;
    eb.f32 = nullptr; eb.f32_cols = 0;
    es.out = ActBuf{nullptr, nullptr, 0}; es.out_cols = 0;
    es.f32_col0 = 0;
    es.bias = c.pk.bias(lh) + kBottleneck;
    a.gepi[0] = eb;
    a.gepi[1] = es;
    a.num_ops = 10;
  }
  a.algo_flops = c.algo ? flops : 0.0;
  return launch_chain(a, c.st);
}

// one backward (dgrad) op of a fused chain for layer l: rows [row0, ..) of the transposed weights
ChainOpArgs bwd_op(const Ctx& c, int l, int in_row0, int n, int kb_act, int kb_in, const uint32_t* mask, void* save) {
  ChainOpArgs L;
  L.n = n;
  L.kb_act = kb_act;
  L.kb_in = kb_in;
  L.kind = mask ? 0 : 1;
  L.mode = 1;
  L.w = c.pk.wt_hi(l, in_row0);
  L.w_lo = c.pk.wt_lo(l, in_row0);
  L.w_ld = layer_def(l).nt_pad;
  L.mask_bits = mask;
  L.save_hi = save;
  return L;
}

// fp16 mode: the seed of the density-gradient chain (|raw_density weights| ~ 1/16, shrinking ~2.4x per layer at
// init) is scaled by an exact power of two so that the fp16 gradient tiles stay in the normal range; the
// normalisation kernel divides it out again.
constexpr float kNormalsSeedScale = 64.f;

// d raw_density / d x0 through the 8 spatial layers, fused.  SS chain: the seed gradient (raw_density weight row where
// a8 > 0) is generated inside the kernel from the ReLU bits; TS chain: it is read from w.g[0].
int normals_chain(const Ctx& c, Workspace& w, int64_t rows) {
  ChainArgs a;
  a.impl = c.chain_impl;
  a.m = rows;
  a.act_f16 = a.w_f16 = c.f16;
  a.seed_scale = c.f16 ? kNormalsSeedScale : 1.f;
  const bool gen_seed = true;   // the seed gradient (raw_density weight row where a8 > 0) is generated in the kernel
  int n = 0;
  if (gen_seed) {
    ChainOpArgs& S = a.op[n++];
    S.n = 256; S.kind = 2; S.mode = 1; S.bias = c.pk.wd(); S.mask_bits = w.ma(8);
  } else {
    a.in = w.g[0];
    a.in_cols = 256;
    a.in_valid = 256;
  }
  double flops = 0.0;
  for (int l = 7; l >= 1; --l) {
    if (l == 5) {
      a.op[n] = bwd_op(c, 5, 256, kEncPad, 4, 0, nullptr, nullptr);
      a.op[n].gepi = 0;
      flops += op_flops(rows, 5, true, false);
      ++n;
    }
    const bool from_in = l == 7 && !gen_seed;
    a.op[n] = bwd_op(c, l, 0, 256, from_in ? 0 : 4, from_in ? 4 : 0, w.ma(l), nullptr);
    flops += op_flops(rows, l, false, false);
    ++n;
  }
  a.op[n] = bwd_op(c, 0, 0, kEncPad, 4, 0, nullptr, nullptr);
  a.op[n].gepi = 1;
  flops += op_flops(rows, 0, false, false);
  ++n;
  a.num_ops = n;
  a.gepi[0] = epi_f32(w.gx0, 128, 128, 0);
  // chain_x3: the layer-0 share goes to its own buffer (ipe_grad_normals adds the two): no read-modify-write in the epilogue
  a.gepi[1] = w.gx0b ? epi_f32(w.gx0b, 128, 128, 0) : epi_f32(w.gx0, 128, 128, 1);
  a.algo_flops = c.algo ? flops : 0.0;
  return launch_chain(a, c.st);
}

// forward for one chunk; normals_out != nullptr runs the in-forward density-gradient pass
int forward_chunk(const Ctx& c, Workspace& w, int64_t row0, int64_t rows, const RnMlpOutputs& o, bool want_normals,
                  bool write_outputs) {
  const int prec = c.cfg->prec;
  const bool keep_spatial = w.nsp == 8;   // training forward / backward recompute: every activation is needed later
  const bool keep_view = w.nvw == 8;
  const ActBuf none = {nullptr, nullptr, 0};
  RN_TRY(launch_encode(prec, c.tdist, c.origins, c.dirs, c.radii, c.s, row0, rows, w.x0, kEncPad, c.st));
  GemmEpilogue heads_epi = epi_act(w.v0, kBottleneck, 0);
  heads_epi.f32 = w.heads_raw; heads_epi.f32_ld = 16; heads_epi.f32_col0 = kBottleneck; heads_epi.f32_cols = 16;
  if (c.chain) {
    RN_TRY(chain_layers(c, 0, kLayerH, rows, w.x0, kEncPad, keep_spatial ? &w : nullptr, true, heads_epi));
  } else {
    for (int l = 0; l < 8; ++l)
      RN_TRY(fwd_layer(c, l, rows, l == 0 ? w.x0 : w.a(l), l == 5 ? w.x0 : none, epi_act(w.a(l + 1), 256, 1)));
    RN_TRY(fwd_layer(c, kLayerH, rows, w.a(8), none, heads_epi));
  }
  if (want_normals) {
    // d raw_density / d x0 through the spatial net (models.py:603-609); result is a constant
    if (!c.chain) RN_TRY(launch_density_grad_seed(prec, w.a(8), c.pk.wd(), w.g[0], rows, c.st));
    if (c.chain) {
      RN_TRY(normals_chain(c, w, rows));
    } else {
      int cur = 0;
      for (int l = 7; l >= 1; --l) {
        if (l == 5)
          RN_TRY(dgrad_layer(c, l, rows, w.g[cur], 256, 256, none, 0, 0, 256, kEncPad, epi_f32(w.gx0, 128, 128, 0)));
        RN_TRY(dgrad_layer(c, l, rows, w.g[cur], 256, 256, none, 0, 0, 0, 256, epi_masked(w.g[cur ^ 1], w.a(l))));
        cur ^= 1;
      }
      RN_TRY(dgrad_layer(c, 0, rows, w.g[cur], 256, 256, none, 0, 0, 0, kEncPad, epi_f32(w.gx0, 128, 128, 1)));
    }
    RN_TRY(launch_ipe_grad_normals(w.gx0, (c.chain && w.gx0b) ? w.gx0b : nullptr, 128, c.tdist, c.origins, c.dirs, c.radii, c.s, row0, rows,
                                   o.normals + row0 * 3, (c.chain && c.f16) ? 1.f / kNormalsSeedScale : 1.f, c.st));
  }
  // outputs of the heads are written even in the recompute pass (cheap); callers pass scratch or real outputs
  RN_TRY(launch_heads_prologue_fwd(prec, w.heads_raw, c.viewdirs, c.s, row0, rows, c.sc, w.v0, o.density + row0,
                                   o.normals_pred + row0 * 3, o.grad_pred + row0 * 3, o.roughness + row0,
                                   o.tint + row0 * 3, c.st));
  if (c.chain) {
    RN_TRY(chain_layers(c, kLayerV0, kLayerC, rows, w.v0, kViewPad, keep_view ? &w : nullptr, false, epi_f32(w.rgb_raw, 4, 4, 0)));
  } else {
    for (int l = 0; l < 8; ++l)
      RN_TRY(fwd_layer(c, kLayerV0 + l, rows, l == 0 ? w.v0 : w.b(l), l == 5 ? w.v0 : none, epi_act(w.b(l + 1), 256, 1)));
    RN_TRY(fwd_layer(c, kLayerC, rows, w.b(8), none, epi_f32(w.rgb_raw, 4, 4, 0)));
  }
  if (write_outputs)
    RN_TRY(launch_color_fwd(w.rgb_raw, w.heads_raw, rows, c.sc, o.rgb + row0 * 3, o.diffuse + row0 * 3,
                            o.specular + row0 * 3, c.st));
  return RN_OK;
}

int wgrad_layer(const Ctx& c, Workspace& w, int l, int64_t rows, ActBuf dy, int dy_valid, int n_real_total, ActBuf x1,
                ActBuf x2) {
  LayerDef d = layer_def(l);
  if (c.chain || c.fused_bias()) {
    // tcgen05 (bf16 / fp16 / split-bf16): one launch per X source covers all of dY (X is read once, not once per
    // 128-column slab); the first one also produces the bias gradient
    WgradArgs g;
    g.prec = c.hprec();
    g.impl = c.impl;
    g.m = rows;
    g.dy = dy; g.dy_valid = dy_valid; g.n0 = 0; g.n_real = n_real_total;
    g.all_slabs = 1;
    g.x_f16 = g.dy_f16 = c.f16;
    g.partial = w.wg_partial;
    g.bias_out = w.gB[l];
    g.x = x1; g.x_valid = d.k1_pad; g.kx = d.k1_pad; g.k_real = d.k1_real;
    g.out = w.gW[l]; g.out_ld = d.k_tot();
    g.algo_flops = 2.0 * (double)rows * n_real_total * d.k1_real;
    RN_TRY(launch_wgrad(g, c.st));
    if (d.k2_pad) {
      g.bias_out = nullptr;
      g.x = x2; g.x_valid = d.k2_pad; g.kx = d.k2_pad; g.k_real = d.k2_real;
      g.out = w.gW[l] + d.k1_pad;
      g.algo_flops = 2.0 * (double)rows * n_real_total * d.k2_real;
      RN_TRY(launch_wgrad(g, c.st));
    }
    return RN_OK;
  }
  for (int n0 = 0; n0 < n_real_total; n0 += 128) {
    WgradArgs g;
    g.prec = c.cfg->prec;
    g.impl = c.impl;
    g.m = rows;
    g.dy = dy; g.dy_valid = dy_valid; g.n0 = n0; g.n_real = n_real_total;
    g.x = x1; g.x_valid = d.k1_pad; g.kx = d.k1_pad; g.k_real = d.k1_real;
    g.out = w.gW[l]; g.out_ld = d.k_tot();
    g.bias_out = c.fused_bias() ? w.gB[l] : nullptr;   // per-layer tcgen05 path: the slab's bias gradient rides along
    const int nslab = n_real_total - n0 < 128 ? n_real_total - n0 : 128;
    g.algo_flops = 2.0 * (double)rows * nslab * d.k1_real;
    RN_TRY(launch_wgrad(g, c.st));
    if (d.k2_pad) {
      g.bias_out = nullptr;
      g.x = x2; g.x_valid = d.k2_pad; g.kx = d.k2_pad; g.k_real = d.k2_real;
      g.out = w.gW[l] + d.k1_pad;
      g.algo_flops = 2.0 * (double)rows * nslab * d.k2_real;
      RN_TRY(launch_wgrad(g, c.st));
    }
  }
  return RN_OK;
}

// backward of one chunk with fused dgrad chains (bf16): the chains save dY of every layer, the wgrads follow
int backward_chunk_chain(const Ctx& c, Workspace& w, int64_t row0, int64_t rows, const RnMlpOutputs& g) {
  const int prec = c.hprec();
  const int wpl = c.x3 ? 2 : 1;
  const ActBuf none = {nullptr, nullptr, 0};
  auto off = [&](const float* p, int per) -> const float* { return p ? p + row0 * per : nullptr; };
  // fp16 mode: the gradient tiles entering the two dgrad chains are scaled by a power of two chosen from their
  // max |x| (S1: view chain, S2: spatial chain), so that the fp16 tiles of all 8 layers stay in the normal range;
  // the scales are divided out when the packed weight gradients are unpacked.
  const ActBuf stage = {w.stage16, nullptr, 16};
  uint32_t* amax = reinterpret_cast<uint32_t*>(w.scal);
  if (c.f16) {
    cudaError_t e = cudaMemsetAsync(w.scal, 0, 256, c.st);
    if (e != cudaSuccess) return rn_set_cuda_error(e, __FILE__, __LINE__);
    RN_TRY(launch_color_bwd(RN_PREC_FP32, w.rgb_raw, w.heads_raw, rows, c.sc, off(g.rgb, 3), off(g.diffuse, 3),
                            off(g.specular, 3), stage, w.dcolor, c.st));
    RN_TRY(launch_amax_f32(w.stage16, rows * 16, amax + 0, c.st));
    RN_TRY(launch_grad_scale(w.scal, 0, c.st));
    RN_TRY(launch_scale_to_f16(w.stage16, 16, rows, 16, w.d_rgb_raw.hi, w.d_rgb_raw.ld, w.scal + 4, c.st));
  } else {
    RN_TRY(launch_color_bwd(prec, w.rgb_raw, w.heads_raw, rows, c.sc, off(g.rgb, 3), off(g.diffuse, 3), off(g.specular, 3),
                            w.d_rgb_raw, w.dcolor, c.st));
  }
  const bool fuse_dv0 = true;   // d v0 = [dY0 | dY5] * wcat^T as ONE op, dY5 re-read from its save
  {  // view net: rgb head, V7..V0
    ChainArgs a;
    a.impl = 0;
    a.w_planes = wpl;
    a.act_f16 = a.w_f16 = c.f16;
    a.m = rows;
    a.in = w.d_rgb_raw;
    a.in_cols = 64;
    a.in_valid = 16;
    int n = 0;
    double flops = op_flops(rows, kLayerC, false, false);
    a.op[n++] = bwd_op(c, kLayerC, 0, 256, 0, 1, w.mb(8), w.gs[7].hi);
    for (int l = 7; l >= 1; --l) {
      const int L = kLayerV0 + l;
      if (l == 5) {
        flops += op_flops(rows, L, true, false);
        if (!fuse_dv0) {
          a.op[n] = bwd_op(c, L, 256, 256, 4, 0, nullptr, nullptr);
          a.op[n].gepi = 0;
          ++n;
        }
      }
      a.op[n++] = bwd_op(c, L, 0, 256, 4, 0, w.mb(l), w.gs[l - 1].hi);
      flops += op_flops(rows, L, false, false);
    }
    flops += op_flops(rows, kLayerV0, false, false);
    if (fuse_dv0) {
      ChainOpArgs& F = a.op[n];
      F.n = 256; F.kb_act = 4; F.kb_in = 4; F.in2 = 1; F.kind = 1; F.mode = 1; F.gepi = 0;
      F.w = c.pk.wcat(); F.w_lo = c.pk.wcat_lo(); F.w_ld = 512;
      ++n;
      a.in2 = w.gs[5];
      a.in2_cols = 256;
      a.in2_valid = 256;
      // columns 0..127 (bottleneck gradient) leave as the bf16 operand of the heads dgrad / wgrad, 128..255 as f32
      GemmEpilogue e;
      e.out = w.d_bott; e.out_cols = 128;
      e.f32 = w.dv0f + 128; e.f32_ld = 256; e.f32_col0 = 128; e.f32_cols = 128; e.f32_accum = 0;
      a.gepi[0] = e;
    } else {
      a.op[n] = bwd_op(c, kLayerV0, 0, 256, 4, 0, nullptr, nullptr);
      a.op[n].gepi = 1;
      ++n;
      a.gepi[0] = epi_f32(w.dv0f, 256, 256, 0);
      a.gepi[1] = epi_f32(w.dv0f, 256, 256, 1);
    }
    a.num_ops = n;
    a.algo_flops = flops;
    RN_TRY(launch_chain(a, c.st));
  }
  RN_TRY(wgrad_layer(c, w, kLayerC, rows, w.d_rgb_raw, 16, 3, w.b(8), none));   // bias gradient fused
  for (int l = 7; l >= 0; --l) {
    const int L = kLayerV0 + l;
    RN_TRY(wgrad_layer(c, w, L, rows, w.gs[l], 256, 256, l == 0 ? w.v0 : w.b(l), l == 5 ? w.v0 : none));
  }
  if (!fuse_dv0) RN_TRY(launch_f32_to_act(prec, w.dv0f, 256, 0, rows, 128, w.d_bott, c.st));
  if (c.f16) {
    // scalar-head gradients in true scale (f32 staging); the bottleneck gradient is already fp16 in the view chain's scale
    RN_TRY(launch_heads_prologue_bwd(RN_PREC_FP32, w.heads_raw, c.viewdirs, c.s, row0, rows, c.sc, w.dv0f, w.dcolor,
                                     off(g.density, 1), off(g.normals_pred, 3), off(g.grad_pred, 3), off(g.roughness, 1),
                                     off(g.tint, 3), stage, w.scal + 5, c.st));
    RN_TRY(launch_amax_f32(w.stage16, rows * 16, amax + 1, c.st));
    RN_TRY(launch_amax_f16(w.d_bott.hi, w.d_bott.ld, 128, rows, amax + 2, c.st));
    RN_TRY(launch_grad_scale(w.scal, 1, c.st));
    RN_TRY(launch_rescale_f16(w.d_bott.hi, w.d_bott.ld, 128, rows, w.scal + 8, c.st));
    RN_TRY(launch_scale_to_f16(w.stage16, 16, rows, 16, w.d_scal.hi, w.d_scal.ld, w.scal + 6, c.st));
  } else {
    RN_TRY(launch_heads_prologue_bwd(prec, w.heads_raw, c.viewdirs, c.s, row0, rows, c.sc, w.dv0f, w.dcolor,
                                     off(g.density, 1), off(g.normals_pred, 3), off(g.grad_pred, 3), off(g.roughness, 1),
                                     off(g.tint, 3), w.d_scal, nullptr, c.st));
  }
  {  // spatial net: heads, S7..S1 (no gradient w.r.t. x0 is needed: sdist is detached)
    ChainArgs a;
    a.impl = 0;
    a.w_planes = wpl;
    a.act_f16 = a.w_f16 = c.f16;
    a.m = rows;
    a.in = w.dheads;
    a.in_cols = 192;
    a.in_valid = 144;
    int n = 0;
    double flops = op_flops(rows, kLayerH, false, false);
    a.op[n++] = bwd_op(c, kLayerH, 0, 256, 0, 3, w.ma(8), w.gs[7].hi);
    for (int l = 7; l >= 1; --l) {
      a.op[n++] = bwd_op(c, l, 0, 256, 4, 0, w.ma(l), w.gs[l - 1].hi);
      flops += op_flops(rows, l, false, false);
    }
    a.num_ops = n;
    a.algo_flops = flops;
    RN_TRY(launch_chain(a, c.st));
  }
  RN_TRY(wgrad_layer(c, w, kLayerH, rows, w.dheads, kHeadsPad, kHeadsReal, w.a(8), none));   // packed head rows 0..138
  for (int l = 7; l >= 0; --l)
    RN_TRY(wgrad_layer(c, w, l, rows, w.gs[l], 256, 256, l == 0 ? w.x0 : w.a(l), l == 5 ? w.x0 : none));
  return RN_OK;
}

int backward_chunk(const Ctx& c, Workspace& w, int64_t row0, int64_t rows, const RnMlpOutputs& g) {
  const int prec = c.cfg->prec;
  const ActBuf none = {nullptr, nullptr, 0};
  auto off = [&](const float* p, int per) -> const float* { return p ? p + row0 * per : nullptr; };
  RN_TRY(launch_color_bwd(prec, w.rgb_raw, w.heads_raw, rows, c.sc, off(g.rgb, 3), off(g.diffuse, 3), off(g.specular, 3),
                          w.d_rgb_raw, w.dcolor, c.st));
  // rgb head
  RN_TRY(wgrad_layer(c, w, kLayerC, rows, w.d_rgb_raw, 16, 3, w.b(8), none));
  if (!c.fused_bias()) RN_TRY(launch_colsum(prec, w.d_rgb_raw, rows, 16, w.gB[kLayerC], c.st));
  RN_TRY(dgrad_layer(c, kLayerC, rows, w.d_rgb_raw, 64, 16, none, 0, 0, 0, 256, epi_masked(w.g[0], w.b(8))));
  int cur = 0;
  for (int l = 7; l >= 0; --l) {
    const int L = kLayerV0 + l;
    RN_TRY(wgrad_layer(c, w, L, rows, w.g[cur], 256, 256, l == 0 ? w.v0 : w.b(l), l == 5 ? w.v0 : none));
    if (!c.fused_bias()) RN_TRY(launch_colsum(prec, w.g[cur], rows, 256, w.gB[L], c.st));
    if (l == 5) RN_TRY(dgrad_layer(c, L, rows, w.g[cur], 256, 256, none, 0, 0, 256, 256, epi_f32(w.dv0f, 256, 256, 0)));
    if (l > 0) {
      RN_TRY(dgrad_layer(c, L, rows, w.g[cur], 256, 256, none, 0, 0, 0, 256, epi_masked(w.g[cur ^ 1], w.b(l))));
      cur ^= 1;
    } else {
      RN_TRY(dgrad_layer(c, L, rows, w.g[cur], 256, 256, none, 0, 0, 0, 256, epi_f32(w.dv0f, 256, 256, 1)));
    }
  }
  RN_TRY(launch_f32_to_act(prec, w.dv0f, 256, 0, rows, 128, w.d_bott, c.st));
  RN_TRY(launch_heads_prologue_bwd(prec, w.heads_raw, c.viewdirs, c.s, row0, rows, c.sc, w.dv0f, w.dcolor,
                                   off(g.density, 1), off(g.normals_pred, 3), off(g.grad_pred, 3), off(g.roughness, 1),
                                   off(g.tint, 3), w.d_scal, nullptr, c.st));
  // heads
  if (c.fused_bias()) {
    // [d bottleneck | d scalar heads] live in one 192-wide buffer: one all-slab launch, bias gradient included
    RN_TRY(wgrad_layer(c, w, kLayerH, rows, w.dheads, kHeadsPad, kHeadsReal, w.a(8), none));
  } else {
    LayerDef d = layer_def(kLayerH);
    WgradArgs a;
    a.prec = prec; a.impl = c.impl; a.m = rows;
    a.x = w.a(8); a.x_valid = 256; a.kx = 256; a.k_real = 256; a.out_ld = d.k_tot();
    a.dy = w.d_bott; a.dy_valid = 128; a.n0 = 0; a.n_real = 128; a.out = w.gW[kLayerH];
    a.algo_flops = 2.0 * (double)rows * 128 * 256;
    RN_TRY(launch_wgrad(a, c.st));
    a.dy = w.d_scal; a.dy_valid = 16; a.n0 = 0; a.n_real = kHeadScalars; a.out = w.gW[kLayerH] + (size_t)128 * d.k_tot();
    a.algo_flops = 2.0 * (double)rows * kHeadScalars * 256;
    RN_TRY(launch_wgrad(a, c.st));
    RN_TRY(launch_colsum(prec, w.d_bott, rows, 128, w.gB[kLayerH], c.st));
    RN_TRY(launch_colsum(prec, w.d_scal, rows, 16, w.gB[kLayerH] + 128, c.st));
  }
  {
    RN_TRY(dgrad_layer(c, kLayerH, rows, w.d_bott, 128, 128, w.d_scal, 64, 16, 0, 256, epi_masked(w.g[0], w.a(8))));
  }
  cur = 0;
  for (int l = 7; l >= 0; --l) {
    RN_TRY(wgrad_layer(c, w, l, rows, w.g[cur], 256, 256, l == 0 ? w.x0 : w.a(l), l == 5 ? w.x0 : none));
    if (!c.fused_bias()) RN_TRY(launch_colsum(prec, w.g[cur], rows, 256, w.gB[l], c.st));
    if (l > 0) {
      RN_TRY(dgrad_layer(c, l, rows, w.g[cur], 256, 256, none, 0, 0, 0, 256, epi_masked(w.g[cur ^ 1], w.a(l))));
      cur ^= 1;
    }
  }
  return RN_OK;
}

int make_ctx(Ctx& c, const RnMlpConfig* cfg, const void* packed, const float* tdist, const float* origins,
             const float* dirs, const float* viewdirs, const float* radii, int s, void* stream) {
  if (!cfg || !packed) return rn_set_error(RN_ERR_ARG, "rn_mlp: null config / packed weights");
  if (cfg->prec < 0 || cfg->prec > 3) return rn_set_error(RN_ERR_ARG, "rn_mlp: bad precision");
  if (cfg->prec == RN_PREC_FP16 && cfg->gemm_impl != 0) return rn_set_error(RN_ERR_UNSUPPORTED, "rn_mlp: the fp16 mode runs on the fused SS chains only (gemm_impl = 0)");
  if (cfg->chunk_rows <= 0 || cfg->chunk_rows % 128) return rn_set_error(RN_ERR_ARG, "rn_mlp: chunk_rows must be a positive multiple of 128");
  if (s < 1) return rn_set_error(RN_ERR_ARG, "rn_mlp: bad sample count");
  c.cfg = cfg;
  c.pk.base = reinterpret_cast<const uint8_t*>(packed);
  c.pk.lay = packed_layout(cfg->prec);
  c.pk.prec = cfg->prec;
  c.tdist = tdist; c.origins = origins; c.dirs = dirs; c.viewdirs = viewdirs; c.radii = radii;
  c.s = s;
  c.st = (cudaStream_t)stream;
  c.sc = {cfg->srgb_mapping, cfg->srgb_normalization, cfg->density_bias, cfg->roughness_bias,
          cfg->rgb_premultiplier, cfg->rgb_bias, cfg->rgb_padding};
  c.impl = cfg->gemm_impl == 1 ? 1 : 0;
  if (cfg->gemm_impl == 3) return rn_set_error(RN_ERR_UNSUPPORTED, "rn_mlp: gemm_impl 3 (TMEM-operand chain experiment) was removed");
  c.chain = (cfg->prec == RN_PREC_BF16 || cfg->prec == RN_PREC_FP16 || cfg->prec == RN_PREC_BF16X3) && cfg->gemm_impl == 0;
  c.f16 = cfg->prec == RN_PREC_FP16;
  c.x3 = c.chain && cfg->prec == RN_PREC_BF16X3;
  c.chain_impl = c.x3 ? 2 : 0;
  return RN_OK;
}

}  // namespace
}  // namespace rn

using namespace rn;

extern "C" const char* rn_last_error(void) { return g_err; }
extern "C" int rn_abi_version(void) { return 3; }
extern "C" const char* rn_mlp_param_name(int i) { return (i >= 0 && i < RN_MLP_NUM_PARAMS) ? kParamNames[i] : nullptr; }
extern "C" int64_t rn_mlp_param_numel(int i) { return (i >= 0 && i < RN_MLP_NUM_PARAMS) ? param_numel(i) : -1; }

extern "C" size_t rn_mlp_packed_bytes(int prec) {
  if (prec < 0 || prec > 3) return 0;
  return packed_layout(prec).total;
}

extern "C" size_t rn_mlp_workspace_bytes(const RnMlpConfig* cfg, int training) {
  if (!cfg || cfg->chunk_rows <= 0) return 0;
  const bool ext = training == 2;
  size_t b = carve(nullptr, cfg->prec, cfg->chunk_rows, training ? 2 : 0, ext, cfg_x3chain(cfg), cfg->deterministic_wgrad != 0).bytes;
  if (training) {
    size_t b1 = carve(nullptr, cfg->prec, cfg->chunk_rows, 1, ext, cfg_x3chain(cfg)).bytes;
    if (b1 > b) b = b1;
    b += (size_t)cfg->chunk_rows * 16 * 4;  // scratch head outputs of the recompute pass
  }
  return b;
}

extern "C" size_t rn_mlp_saved_bytes(const RnMlpConfig* cfg, int64_t n_rows) {
  if (!cfg || n_rows <= 0) return 0;
  return carve_saved(nullptr, cfg->prec, cfg_hprec(cfg), n_rows).bytes;
}

extern "C" int rn_mlp_pack(const float* const* params, void* packed, int prec, void* stream) {
  if (!params || !packed || prec < 0 || prec > 3) return rn_set_error(RN_ERR_ARG, "rn_mlp_pack: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  PackedLayout lay = packed_layout(prec);
  uint8_t* base = reinterpret_cast<uint8_t*>(packed);
  cudaError_t e = cudaMemsetAsync(packed, 0, lay.total, st);
  if (e != cudaSuccess) return rn_set_cuda_error(e, __FILE__, __LINE__);
  // every weight block / bias vector is one segment of a table-driven launch (pointwise.cu)
  PackTable tw, tb;
  tw.n = tb.n = 0;
  int rc_flush = RN_OK;
  auto flush = [&](PackTable& t) {
    if (t.n && rc_flush == RN_OK) rc_flush = launch_pack_batch(prec, t, st);
    t.n = 0;
  };
  auto seg = [&](const float* src, int src_ld, int nr, int nc, int transpose, void* dst_hi, void* dst_lo, int dst_ld, int r0,
                 int c0) {
    if (tw.n == kMaxBatchSegs) flush(tw);
    tw.seg[tw.n++] = PackSeg{src, dst_hi, prec == RN_PREC_BF16X3 ? dst_lo : nullptr, src_ld, nr, nc, transpose, dst_ld, r0, c0, 0};
  };
  auto copy_f32 = [&](const float* src, float* dst, int n) {
    if (tb.n == kMaxBatchSegs) flush(tb);
    tb.seg[tb.n++] = PackSeg{src, dst, nullptr, n, 1, n, 0, n, 0, 0, 1};
  };
  for (int l = 0; l < kNumLayers; ++l) {
    LayerDef d = layer_def(l);
    void* wf_hi = base + lay.wf[l];
    void* wf_lo = base + lay.wf[l] + lay.wf_plane[l];
    void* wt_hi = base + lay.wt[l];
    void* wt_lo = base + lay.wt[l] + lay.wt_plane[l];
    float* bias = reinterpret_cast<float*>(base + lay.bias[l]);
    if (l == kLayerH) {
      for (const HeadSeg& hs : kHeadSegs) {
        seg(params[hs.param], 256, hs.rows, 256, 0, wf_hi, wf_lo, d.k_tot(), hs.row0, 0);
        seg(params[hs.param], 256, hs.rows, 256, 1, wt_hi, wt_lo, d.nt_pad, 0, hs.row0);
        copy_f32(params[hs.param + 1], bias + hs.row0, hs.rows);
      }
      continue;
    }
    const int pi = layer_param(l);
    const int kr = d.k1_real + d.k2_real;
    seg(params[pi], kr, d.n_real, d.k1_real, 0, wf_hi, wf_lo, d.k_tot(), 0, 0);
    seg(params[pi], kr, d.n_real, d.k1_real, 1, wt_hi, wt_lo, d.nt_pad, 0, 0);
    if (d.k2_pad) {
      seg(params[pi] + d.k1_real, kr, d.n_real, d.k2_real, 0, wf_hi, wf_lo, d.k_tot(), 0, d.k1_pad);
      seg(params[pi] + d.k1_real, kr, d.n_real, d.k2_real, 1, wt_hi, wt_lo, d.nt_pad, d.k1_pad, 0);
    }
    copy_f32(params[pi + 1], bias, d.n_real);
  }
  copy_f32(params[kParamDensity], reinterpret_cast<float*>(base + lay.wd), 256);
  if (prec != RN_PREC_FP32) {
    // wcat[j, 0:256] = W_V0[:, j], wcat[j, 256:512] = W_V5[:, 256 + j]  (j = view-net input feature, 201 real)
    void* wc = base + lay.wcat;
    void* wc_lo = base + lay.wcat + lay.wcat_plane;
    const int p0 = layer_param(kLayerV0), p5 = layer_param(kLayerV0 + 5);
    seg(params[p0], kViewReal, 256, kViewReal, 1, wc, wc_lo, 512, 0, 0);
    seg(params[p5] + 256, 256 + kViewReal, 256, kViewReal, 1, wc, wc_lo, 512, 0, 256);
  }
  flush(tw);
  flush(tb);
  if (rc_flush != RN_OK) return rc_flush;
  return RN_OK;
}

extern "C" int rn_mlp_forward(const RnMlpConfig* cfg, const void* packed, const float* tdist, const float* origins,
                              const float* dirs, const float* viewdirs, const float* radii, int64_t n_rays, int s,
                              const RnMlpOutputs* out, void* workspace, size_t workspace_bytes, void* saved,
                              size_t saved_bytes, void* stream) {
  Ctx c;
  RN_TRY(make_ctx(c, cfg, packed, tdist, origins, dirs, viewdirs, radii, s, stream));
  if (!out || !out->density || !out->rgb || !out->normals_pred || !out->grad_pred || !out->tint || !out->diffuse ||
      !out->specular || !out->roughness)
    return rn_set_error(RN_ERR_ARG, "rn_mlp_forward: missing output buffers");
  const bool want_normals = out->normals != nullptr;
  const int64_t rows_total = n_rays * s;
  const int64_t rc = cfg->chunk_rows;
  Workspace w = carve(workspace, cfg->prec, rc, want_normals ? 1 : 0, saved != nullptr, c.x3);
  if (w.bytes > workspace_bytes) return rn_set_error(RN_ERR_ARG, "rn_mlp_forward: workspace too small");
  Saved sv;
  if (saved) {
    sv = carve_saved(saved, cfg->prec, c.hprec(), rows_total);
    if (sv.bytes > saved_bytes) return rn_set_error(RN_ERR_ARG, "rn_mlp_forward: saved region too small");
  }
  for (int64_t row0 = 0; row0 < rows_total; row0 += rc) {
    const int64_t rows = rows_total - row0 < rc ? rows_total - row0 : rc;
    if (saved) use_saved(w, sv, cfg->prec, c.hprec(), row0);
    RN_TRY(forward_chunk(c, w, row0, rows, *out, want_normals, true));
  }
  return RN_OK;
}

extern "C" int rn_mlp_backward(const RnMlpConfig* cfg, const void* packed, const float* tdist, const float* origins,
                               const float* dirs, const float* viewdirs, const float* radii, int64_t n_rays, int s,
                               const RnMlpOutputs* g, float* const* grads, void* workspace, size_t workspace_bytes,
                               const void* saved, size_t saved_bytes, void* stream) {
  Ctx c;
  RN_TRY(make_ctx(c, cfg, packed, tdist, origins, dirs, viewdirs, radii, s, stream));
  if (!g || !grads) return rn_set_error(RN_ERR_ARG, "rn_mlp_backward: null gradients");
  const int64_t rows_total = n_rays * s;
  const int64_t rc = cfg->chunk_rows;
  if (cfg->deterministic_wgrad && !c.chain) return rn_set_error(RN_ERR_UNSUPPORTED, "rn_mlp_backward: deterministic_wgrad needs the fused-chain modes (gemm_impl 0, bf16 / fp16 / bf16x3)");
  Workspace w = carve(workspace, cfg->prec, rc, 2, saved != nullptr, c.x3, cfg->deterministic_wgrad != 0);
  Saved sv;
  if (saved) {
    sv = carve_saved(const_cast<void*>(saved), cfg->prec, c.hprec(), rows_total);
    if (sv.bytes > saved_bytes) return rn_set_error(RN_ERR_ARG, "rn_mlp_backward: saved region too small");
  }
  const size_t scratch_off = w.bytes;
  if (w.bytes + (size_t)rc * 16 * 4 > workspace_bytes) return rn_set_error(RN_ERR_ARG, "rn_mlp_backward: workspace too small");
  // scratch destinations for the per-sample head outputs of the recompute pass (density 1, normals_pred 3,
  // grad_pred 3, roughness 1, tint 3 = 11 floats per row, packed as separate arrays)
  float* scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + scratch_off);
  cudaStream_t st = c.st;
  auto zero_packed = [&]() -> int {
    // gW[0] .. gB[last] are carved back to back: one memset
    uint8_t* lo = reinterpret_cast<uint8_t*>(w.gW[0]);
    uint8_t* hi = reinterpret_cast<uint8_t*>(w.gB[kNumLayers - 1]) + align256((size_t)layer_def(kNumLayers - 1).n_pad * 4);
    cudaError_t e = cudaMemsetAsync(lo, 0, (size_t)(hi - lo), st);
    if (e != cudaSuccess) return rn_set_cuda_error(e, __FILE__, __LINE__);
    return RN_OK;
  };
  // packed-layout gradients -> parameter gradients (+=); fp16 mode: times 1/S1 (view net, rgb head) or 1/S2 (the rest)
  auto unpack_packed = [&](bool scaled) -> int {
    UnpackTable t;
    t.n = 0;
    int rc = RN_OK;
    auto add = [&](const float* src, int src_ld, int r0, int c0, int nr, int nc, float* dst, int dst_ld, const float* sc) {
      if (t.n == kMaxBatchSegs) {
        if (rc == RN_OK) rc = launch_unpack_add_batch(t, st);
        t.n = 0;
      }
      t.seg[t.n++] = UnpackSeg{src, dst, sc, src_ld, r0, c0, nr, nc, dst_ld};
    };
    for (int l = 0; l < kNumLayers; ++l) {
      LayerDef d = layer_def(l);
      const float* sc = scaled ? (l >= kLayerV0 ? w.scal + 5 : w.scal + 7) : nullptr;
      if (l == kLayerH) {
        for (const HeadSeg& hs : kHeadSegs) {
          add(w.gW[l], d.k_tot(), hs.row0, 0, hs.rows, 256, grads[hs.param], 256, sc);
          add(w.gB[l], d.n_pad, 0, hs.row0, 1, hs.rows, grads[hs.param + 1], hs.rows, sc);
        }
        continue;
      }
      const int pi = layer_param(l);
      const int kr = d.k1_real + d.k2_real;
      add(w.gW[l], d.k_tot(), 0, 0, d.n_real, d.k1_real, grads[pi], kr, sc);
      if (d.k2_pad) add(w.gW[l], d.k_tot(), 0, d.k1_pad, d.n_real, d.k2_real, grads[pi] + d.k1_real, kr, sc);
      add(w.gB[l], d.n_pad, 0, 0, 1, d.n_real, grads[pi + 1], d.n_real, sc);
    }
    if (rc == RN_OK) rc = launch_unpack_add_batch(t, st);
    return rc;
  };
  const bool per_chunk = c.f16;   // the scales are chosen per chunk, so every chunk is unpacked on its own
  if (!per_chunk) RN_TRY(zero_packed());
  for (int64_t row0 = 0; row0 < rows_total; row0 += rc) {
    const int64_t rows = rows_total - row0 < rc ? rows_total - row0 : rc;
    RnMlpOutputs tmp;
    memset(&tmp, 0, sizeof(tmp));
    // recompute-pass outputs land in scratch, addressed so that "+ row0*k" inside forward_chunk hits the scratch base
    tmp.density = scratch - row0;
    tmp.roughness = scratch + rc - row0;
    tmp.normals_pred = scratch + 2 * rc - row0 * 3;
    tmp.grad_pred = scratch + 5 * rc - row0 * 3;
    tmp.tint = scratch + 8 * rc - row0 * 3;
    if (saved) {
      use_saved(w, sv, cfg->prec, c.hprec(), row0);
    } else {
      c.algo = false;
      RN_TRY(forward_chunk(c, w, row0, rows, tmp, false, false));
      c.algo = true;
    }
    if (per_chunk) RN_TRY(zero_packed());
    if (c.chain) {
      RN_TRY(backward_chunk_chain(c, w, row0, rows, *g));
    } else {
      RN_TRY(backward_chunk(c, w, row0, rows, *g));
    }
    if (per_chunk) RN_TRY(unpack_packed(true));
  }
  if (!per_chunk) RN_TRY(unpack_packed(false));
  return RN_OK;
}

// ---- unit-level entry points ------------------------------------------------------------------
extern "C" int rn_encode(const float* tdist, const float* origins, const float* dirs, const float* radii, int64_t n_rays,
                         int s, float* feat_out, void* stream) {
  ActBuf out = {feat_out, nullptr, 96};
  return launch_encode(RN_PREC_FP32, tdist, origins, dirs, radii, s, 0, n_rays * s, out, 96, (cudaStream_t)stream);
}

extern "C" int rn_ide(const float* dirs, const float* kappa_inv, int64_t n, float* out, void* stream) {
  return launch_ide(dirs, kappa_inv, n, out, (cudaStream_t)stream);
}

extern "C" size_t rn_gemm_scratch_bytes(int64_t m, int n, int k) {
  size_t rows = (size_t)(m > 256 ? m : 256);
  return 2 * align256(rows * (size_t)(k > 256 ? k : 256) * 4) + 2 * align256((size_t)256 * 512 * 4) + 4096;
}

namespace {
ActBuf scratch_act(Carver& c, int prec, int64_t rows, int ld) { return c.act(prec, rows, ld); }
}  // namespace

extern "C" int rn_gemm_test(const float* a, const float* b, const float* bias, int64_t m, int n, int k, int relu,
                            int prec, int impl, float* cmat, void* scratch, size_t scratch_bytes, void* stream) {
  if (k % 64 || n % 16 || n > 256 || k > 512) return rn_set_error(RN_ERR_ARG, "rn_gemm_test: need k%64==0, k<=512, n%16==0, n<=256");
  if (scratch_bytes < rn_gemm_scratch_bytes(m, n, k)) return rn_set_error(RN_ERR_ARG, "rn_gemm_test: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  Carver c{reinterpret_cast<uint8_t*>(scratch)};
  ActBuf abuf = scratch_act(c, prec, m, k);
  ActBuf bbuf = scratch_act(c, prec, n, k);
  RN_TRY(launch_f32_to_act(prec, a, k, 0, m, k, abuf, st));
  RN_TRY(launch_f32_to_act(prec, b, k, 0, n, k, bbuf, st));
  GemmArgs g;
  g.prec = prec; g.impl = impl; g.m = m; g.n = n;
  // exercise the two-source path when k > 256
  const int k1 = k > 256 ? 256 : k;
  g.a1 = abuf; g.k1 = k1; g.a1_valid = k1;
  if (k > k1) {
    g.a2 = abuf;
    g.a2.hi = reinterpret_cast<uint8_t*>(abuf.hi) + (size_t)k1 * elem_bytes(prec);
    if (abuf.lo) g.a2.lo = reinterpret_cast<uint8_t*>(abuf.lo) + (size_t)k1 * 2;
    g.k2 = k - k1; g.a2_valid = k - k1;
  }
  g.b_hi = bbuf.hi; g.b_lo = bbuf.lo; g.b_ld = k;
  g.epi.bias = bias; g.epi.relu = relu;
  g.epi.f32 = cmat; g.epi.f32_ld = n; g.epi.f32_col0 = 0; g.epi.f32_cols = n;
  return launch_gemm(g, st);
}

extern "C" int rn_wgrad_test(const float* dy, const float* x, int64_t m, int n, int k, int prec, int impl, float* cmat,
                             void* scratch, size_t scratch_bytes, void* stream) {
  if (n % 8 || n > 256 || k % 64 || k > 256) return rn_set_error(RN_ERR_ARG, "rn_wgrad_test: need n%8==0, n<=256, k%64==0, k<=256");
  if (scratch_bytes < rn_gemm_scratch_bytes(m, n, k)) return rn_set_error(RN_ERR_ARG, "rn_wgrad_test: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  Carver c{reinterpret_cast<uint8_t*>(scratch)};
  ActBuf ybuf = scratch_act(c, prec, m, n);
  ActBuf xbuf = scratch_act(c, prec, m, k);
  RN_TRY(launch_f32_to_act(prec, dy, n, 0, m, n, ybuf, st));
  RN_TRY(launch_f32_to_act(prec, x, k, 0, m, k, xbuf, st));
  cudaError_t e = cudaMemsetAsync(cmat, 0, (size_t)n * k * 4, st);
  if (e != cudaSuccess) return rn_set_cuda_error(e, __FILE__, __LINE__);
  WgradArgs g;
  g.prec = prec; g.impl = impl; g.m = m;
  g.dy = ybuf; g.dy_valid = n; g.n0 = 0; g.n_real = n;
  g.x = xbuf; g.x_valid = k; g.kx = k; g.k_real = k;
  g.out = cmat; g.out_ld = k;
  if (n <= 128) return launch_wgrad(g, st);
  // wider dY: the tcgen05 kernels cover both 128-feature slabs in ONE launch (the training path's wgrad2 kernel with
  // its row-coalesced red.v4 epilogue), the SIMT kernel takes one launch per slab
  if (impl == 0 && prec != RN_PREC_FP32) {
    g.all_slabs = 1;
    return launch_wgrad(g, st);
  }
  for (int n0 = 0; n0 < n; n0 += 128) {
    g.n0 = n0;
    RN_TRY(launch_wgrad(g, st));
  }
  return RN_OK;
}

extern "C" int rn_gemm_bench(int64_t m, int prec, int impl, int iters, float* ms_out, void* scratch, size_t scratch_bytes,
                             void* stream) {
  if (scratch_bytes < rn_gemm_scratch_bytes(m, 256, 256) + align256((size_t)m * 256 * 4) * 2)
    return rn_set_error(RN_ERR_ARG, "rn_gemm_bench: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  Carver c{reinterpret_cast<uint8_t*>(scratch)};
  ActBuf abuf = scratch_act(c, prec, m, 256);
  ActBuf bbuf = scratch_act(c, prec, 256, 256);
  ActBuf obuf = scratch_act(c, prec, m, 256);
  // operands: whatever bytes are in scratch are fine for timing, but avoid NaN-pattern slowdowns: zero them
  cudaMemsetAsync(scratch, 0, c.off, st);
  GemmArgs g;
  g.prec = prec; g.impl = impl; g.m = m; g.n = 256;
  g.a1 = abuf; g.k1 = 256; g.a1_valid = 256;
  g.b_hi = bbuf.hi; g.b_lo = bbuf.lo; g.b_ld = 256;
  g.epi.relu = 1; g.epi.out = obuf; g.epi.out_cols = 256;
  RN_TRY(launch_gemm(g, st));  // warm-up
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int i = 0; i < iters; ++i) RN_TRY(launch_gemm(g, st));
  cudaEventRecord(e1, st);
  cudaError_t e = cudaEventSynchronize(e1);
  if (e != cudaSuccess) return rn_set_cuda_error(e, __FILE__, __LINE__);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = ms / iters;
  return RN_OK;
}
