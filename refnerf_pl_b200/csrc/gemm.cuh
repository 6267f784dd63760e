// GEMM building blocks of the NerfMLP chain: argument structs and launchers (definitions in
// gemm_simt.cu and gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace rn {

// Epilogue shared by the SIMT and tcgen05 kernels; applied to 8 consecutive columns of one row.
struct GemmEpilogue {
  const float* bias = nullptr;  // [n]
  int relu = 0;
  ActBuf mask = {nullptr, nullptr, 0};  // if set: v = (mask[row, col] > 0) ? v : 0   (ReLU backward)
  ActBuf out = {nullptr, nullptr, 0};   // columns [0, out_cols) in activation format
  int out_cols = 0;
  float* f32 = nullptr;                 // columns [f32_col0, f32_col0 + f32_cols) -> f32[row*f32_ld + col - f32_col0]
  int f32_ld = 0, f32_col0 = 0, f32_cols = 0, f32_accum = 0;
};

struct GemmArgs {
  int prec = RN_PREC_FP32;
  int impl = 0;        // 0 = tcgen05 (bf16 modes) / SIMT (fp32); 1 = force SIMT
  int64_t m = 0;       // rows
  int n = 0;           // output columns, multiple of 16, <= 256
  ActBuf a1 = {nullptr, nullptr, 0};
  int k1 = 0;          // multiple of 64
  int a1_valid = 0;    // valid columns of a1 (columns beyond read as zero), multiple of 8
  ActBuf a2 = {nullptr, nullptr, 0};
  int k2 = 0, a2_valid = 0;
  const void* b_hi = nullptr;  // weights [n, k1+k2] K-major in `prec` format
  const void* b_lo = nullptr;
  int b_ld = 0;
  GemmEpilogue epi;
  double algo_flops = 0.0;  // algorithmic (unpadded, non-recompute) FLOPs carried by this launch (profiling)
};

// dW[n0 + i, j] += sum_r dY[r, n0 + i] * X[r, j]   for i < 128, j < kx  (guards: n0+i < n_real, j < k_real)
struct WgradArgs {
  int prec = RN_PREC_FP32;
  int impl = 0;
  int64_t m = 0;
  ActBuf dy = {nullptr, nullptr, 0};
  int dy_valid = 0;   // valid columns of dy
  int n0 = 0;         // first dy column of this 128-wide slab
  int n_real = 0;     // rows of dW that exist
  ActBuf x = {nullptr, nullptr, 0};
  int x_valid = 0;
  int kx = 0;         // multiple of 64, <= 256
  int k_real = 0;
  float* out = nullptr;  // dW (fp32), row-major [.., out_ld], element (n, j) at out[n*out_ld + j]
  int out_ld = 0;
  int all_slabs = 0;          // bf16 tcgen05 only: cover every 128-wide slab of dY (n_real <= 256) in one launch
  float* bias_out = nullptr;  // tcgen05 kernels: also accumulate the bias gradient here, bias_out[c] += sum_r dY[r, c] for the
                              // dY columns this launch covers (all of them with all_slabs, else n0 .. n0+127)
  float* partial = nullptr;   // with all_slabs: deterministic reduction -- every CTA writes its [256 x 256] partial tile (and
                              // 256 bias partial sums) to partial + cta * (256 * 257) and a second kernel adds them in CTA order
  int x_f16 = 0, dy_f16 = 0;  // with all_slabs: the operands are fp16 instead of bf16 (both or neither: the MMA
                              // rejects mixed a/b formats with an illegal-instruction fault)
  double algo_flops = 0.0;
};

// Fused chain (bf16, tcgen05): a sequence of GEMM "ops" runs back to back on 128-row tiles with the
// running activation resident in shared memory.  Hidden ops turn the TMEM accumulator into the next A
// operand (forward: bias + ReLU; backward: 1-bit ReLU mask written by the forward chain) and may also save it
// to global memory; global ops run an ordinary GemmEpilogue.  Weights stream through the TMA ring.
struct ChainOpArgs {
  int n = 0;                 // output columns; 256 for hidden ops
  int kb_act = 0;            // 64-wide K blocks from the running activation (0 or 4)
  int kb_in = 0;             // K blocks from the chain input tile (0 or all of them)
  int in2 = 0;               // take the input K blocks from ChainArgs::in2 instead (chain_pair.cu only)
  int kind = 0;              // 0 hidden, 1 global epilogue, 2 seed (backward chains, chain_pair.cu): the activation tile is
                             // generated as mask_bits ? bias[col] : 0 -- no GEMM, no input tensor
  int mode = 0;              // hidden: 0 bias + ReLU, 1 ReLU mask
  int gepi = 0;              // global: epilogue index (0/1)
  const void* w = nullptr;   // bf16 weights [n, (kb_act + kb_in) * 64] K-major
  const void* w_lo = nullptr;   // lo plane of split-bf16 weights (chain_x3.cu; chain_pair.cu with ChainArgs::w_planes = 2)
  int w_ld = 0;
  const float* bias = nullptr;
  const uint32_t* mask_bits = nullptr;  // mode 1: ReLU bits of the matching forward activation (layout: relu_bits_index)
  void* save_hi = nullptr;              // optional bf16 [m,256] copy of the hidden result (TMA store)
  uint32_t* save_bits = nullptr;        // mode 0: optional ReLU bits of the result (8 words per row, relu_bits_index)
};
struct ChainArgs {
  int64_t m = 0;
  ActBuf in = {nullptr, nullptr, 0};
  int in_cols = 0;           // K extent of the input tile: multiple of 64, <= 256
  int in_valid = 0;          // valid columns of the input buffer (beyond: zero)
  ActBuf in2 = {nullptr, nullptr, 0};   // optional second input tensor [m, 64*k] (may be a save_hi buffer of an earlier op
  int in2_cols = 0, in2_valid = 0;      // of the same launch: the kernel orders the TMA store before the TMA load)
  int impl = 0;              // 0: chain_pair.cu (one-plane activations, two row tiles), 2: chain_x3.cu (split-bf16 activations
                             // AND weights, 3 MMAs per K step, one row tile, K-block-granular epilogue hand-off)
  int w_planes = 1;          // chain_pair.cu: 2 = split-bf16 weights against one-plane activations (2 MMAs per K step)
  int act_f16 = 0;           // chain_pair.cu: inputs, activation tile and activation-format outputs are fp16 (else bf16)
  int w_f16 = 0;             // chain_pair.cu: weights are fp16 (else bf16); must equal act_f16 (mixed a/b formats fault)
  float seed_scale = 1.f;    // chain_pair.cu: factor applied to the vector of a seed op
  int num_ops = 0;
  ChainOpArgs op[12];
  GemmEpilogue gepi[2];
  double algo_flops = 0.0;
  double exec_flops = 0.0;   // computed by the launchers: 2 * rows * n * k_padded * MMAs per K step, summed over the ops
};
int launch_chain(const ChainArgs& a, cudaStream_t st);      // chain_pair.cu: CTA pairs, cta_group::2, two row tiles in flight
int launch_chain_x3t(const ChainArgs& a, cudaStream_t st);  // chain_x3t.cu: the same chain with the activation in tensor memory (TS-form MMAs)
int launch_chain_x3(const ChainArgs& a, cudaStream_t st);   // chain_x3.cu: CTA pairs, split-bf16 operands, one row tile
// ReLU bit masks: 8 words per row (word w = columns [32w, 32w+32)), stored so that the 32 rows a warp owns are
// contiguous per word: word index of (row, w) = ((row / 32) * 8 + w) * 32 + row % 32.  Buffers are sized for rows
// rounded up to a multiple of 32.
inline size_t relu_bits_bytes(int64_t rows) { return (size_t)((rows + 31) / 32) * 32 * 32; }

int launch_gemm(const GemmArgs& g, cudaStream_t st);
int launch_wgrad(const WgradArgs& g, cudaStream_t st);

int launch_gemm_simt(const GemmArgs& g, cudaStream_t st);
int launch_wgrad_simt(const WgradArgs& g, cudaStream_t st);
int launch_gemm_tc(const GemmArgs& g, cudaStream_t st);
int launch_wgrad_tc(const WgradArgs& g, cudaStream_t st);

// ---- epilogue device code (shared) ------------------------------------------------------------
template <int PREC>
__device__ __forceinline__ void epi_store8(const GemmEpilogue& e, size_t row, int col, float* v) {
  if (e.bias) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += e.bias[col + i];
  }
  if (e.relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (e.mask.hi) {
    float mk[8];
    act_load8_hi<PREC>(e.mask, row, col, mk);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = mk[i] > 0.f ? v[i] : 0.f;
  }
  if (e.out.hi && col < e.out_cols) act_store8<PREC>(e.out, row, col, v);
  if (e.f32 && col + 8 > e.f32_col0 && col < e.f32_col0 + e.f32_cols) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = col + i - e.f32_col0;
      if (c >= 0 && c < e.f32_cols) {
        float* p = e.f32 + row * e.f32_ld + c;
        *p = e.f32_accum ? (*p + v[i]) : v[i];
      }
    }
  }
}

}  // namespace rn
