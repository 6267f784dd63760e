// fp32-accumulate FFMA GEMMs: the exact-parity anchor (RN_PREC_FP32) and the on-device checker for
// the tcgen05 kernels (same operands, same epilogue, plain CUDA-core arithmetic).
#include "gemm.cuh"

namespace rn {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, LDS = 68;

template <int PREC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(GemmArgs g) {
  __shared__ float As[BK][LDS];
  __shared__ float Bs[BK][LDS];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int ty = tid >> 3, tx = tid & 7;  // rows ty*2..+1, cols tx*8..+7
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int ktot = g.k1 + g.k2;
  const ActBuf bbuf = {const_cast<void*>(g.b_hi), const_cast<void*>(g.b_lo), g.b_ld};
  for (int k0 = 0; k0 < ktot; k0 += BK) {
    {
      const int half = tid >> 7;          // 0: A tile, 1: B tile
      const int t = tid & 127;
      const int r = t >> 1, kg = (t & 1) * 8;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (half == 0) {
        const int64_t row = m0 + r;
        if (row < g.m) {
          const int k = k0 + kg;
          if (k < g.k1) {
            if (k < g.a1_valid) act_load8<PREC>(g.a1, (size_t)row, k, v);
          } else {
            if (k - g.k1 < g.a2_valid) act_load8<PREC>(g.a2, (size_t)row, k - g.k1, v);
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) As[kg + e][r] = v[e];
      } else {
        const int nr = n0 + r;
        if (nr < g.n) act_load8<PREC>(bbuf, (size_t)nr, k0 + kg, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) Bs[kg + e][r] = v[e];
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float a0 = As[kk][ty * 2], a1 = As[kk][ty * 2 + 1];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float b = Bs[kk][tx * 8 + j];
        acc[0][j] = fmaf(a0, b, acc[0][j]);
        acc[1][j] = fmaf(a1, b, acc[1][j]);
      }
    }
    __syncthreads();
  }
  const int col = n0 + tx * 8;
  if (col < g.n) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t row = m0 + ty * 2 + i;
      if (row < g.m) epi_store8<PREC>(g.epi, (size_t)row, col, acc[i]);
    }
  }
}

// dW[n, j] += sum_r dY[r, n] X[r, j]; grid (n tiles of 64, j tiles of 64, row splits)
template <int PREC>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(WgradArgs g, int64_t rows_per_split) {
  __shared__ float Ys[BK][LDS];
  __shared__ float Xs[BK][LDS];
  const int tid = threadIdx.x;
  const int n0 = g.n0 + blockIdx.x * BN;
  const int j0 = blockIdx.y * BN;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = min(g.m, r_begin + rows_per_split);
  const int ty = tid >> 3, tx = tid & 7;  // dW rows (n) ty*2..+1, cols (j) tx*8..+7
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += BK) {
    {
      const int half = tid >> 7;
      const int t = tid & 127;
      const int r = t >> 3, cg = (t & 7) * 8;  // 16 rows x 8 groups
      const int64_t row = r0 + r;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (row < r_end) {
        if (half == 0) {
          if (n0 + cg < g.dy_valid) act_load8<PREC>(g.dy, (size_t)row, n0 + cg, v);
        } else {
          if (j0 + cg < g.x_valid) act_load8<PREC>(g.x, (size_t)row, j0 + cg, v);
        }
      }
      float* dst = half == 0 ? &Ys[r][cg] : &Xs[r][cg];
#pragma unroll
      for (int e = 0; e < 8; ++e) dst[e] = v[e];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float a0 = Ys[kk][ty * 2], a1 = Ys[kk][ty * 2 + 1];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float b = Xs[kk][tx * 8 + j];
        acc[0][j] = fmaf(a0, b, acc[0][j]);
        acc[1][j] = fmaf(a1, b, acc[1][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int n = n0 + ty * 2 + i;
    if (n >= g.n_real || n >= g.n0 + 128) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int jj = j0 + tx * 8 + j;
      if (jj < g.k_real) atomicAdd(g.out + (size_t)n * g.out_ld + jj, acc[i][j]);
    }
  }
}

}  // namespace

int launch_gemm_simt(const GemmArgs& g, cudaStream_t st) {
  if (g.m <= 0) return RN_OK;
  dim3 grid((unsigned)((g.m + BM - 1) / BM), (unsigned)((g.n + BN - 1) / BN));
  rn_prof_begin(RN_PROF_GEMM_SIMT, st, g.algo_flops);
  switch (g.prec) {
    case RN_PREC_FP32: gemm_simt_kernel<RN_PREC_FP32><<<grid, 256, 0, st>>>(g); break;
    case RN_PREC_BF16: gemm_simt_kernel<RN_PREC_BF16><<<grid, 256, 0, st>>>(g); break;
    case RN_PREC_BF16X3: gemm_simt_kernel<RN_PREC_BF16X3><<<grid, 256, 0, st>>>(g); break;
    default: return rn_set_error(RN_ERR_ARG, "gemm: bad precision");
  }
  rn_prof_end(RN_PROF_GEMM_SIMT, st);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_wgrad_simt(const WgradArgs& g, cudaStream_t st) {
  if (g.m <= 0) return RN_OK;
  const int n_tiles = 2;  // one 128-wide slab of dY features
  const int j_tiles = (g.kx + BN - 1) / BN;
  int splits = (int)((g.m + 4095) / 4096);
  if (splits > 512) splits = 512;
  if (splits < 1) splits = 1;
  int64_t rows_per = (g.m + splits - 1) / splits;
  rows_per = (rows_per + BK - 1) / BK * BK;
  dim3 grid(n_tiles, j_tiles, (unsigned)((g.m + rows_per - 1) / rows_per));
  switch (g.prec) {
    case RN_PREC_FP32: wgrad_simt_kernel<RN_PREC_FP32><<<grid, 256, 0, st>>>(g, rows_per); break;
    case RN_PREC_BF16: wgrad_simt_kernel<RN_PREC_BF16><<<grid, 256, 0, st>>>(g, rows_per); break;
    case RN_PREC_BF16X3: wgrad_simt_kernel<RN_PREC_BF16X3><<<grid, 256, 0, st>>>(g, rows_per); break;
    default: return rn_set_error(RN_ERR_ARG, "wgrad: bad precision");
  }
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_gemm(const GemmArgs& g, cudaStream_t st) {
  if (g.n % 16 || g.n > 256 || g.k1 % 64 || g.k2 % 64) return rn_set_error(RN_ERR_ARG, "gemm: n%16, n<=256, k%64 required");
  if (g.prec == RN_PREC_FP32 || g.impl == 1) return launch_gemm_simt(g, st);
  return launch_gemm_tc(g, st);
}

int launch_wgrad(const WgradArgs& g, cudaStream_t st) {
  if (g.kx % 64 || g.kx > 256) return rn_set_error(RN_ERR_ARG, "wgrad: kx%64, kx<=256 required");
  if (g.prec == RN_PREC_FP32 || g.impl == 1) return launch_wgrad_simt(g, st);
  return launch_wgrad_tc(g, st);
}

}  // namespace rn
