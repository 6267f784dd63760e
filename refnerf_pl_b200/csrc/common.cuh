// Shared device helpers for the refnerf_pl_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/refnerf_b200.h"

#define RN_EPS32 1.1920928955078125e-07f
#define RN_FULL 0xffffffffu

#define RN_CUDA_CHECK_LAUNCH()                       \
  do {                                               \
    rn_count_launch();                               \
    cudaError_t e__ = cudaGetLastError();            \
    if (e__ != cudaSuccess) return rn_set_cuda_error(e__, __FILE__, __LINE__); \
  } while (0)

int rn_set_cuda_error(cudaError_t e, const char* file, int line);
int rn_set_error(int code, const char* msg);
void rn_count_launch();
// optional CUDA-event timing of kernel classes (bench.py roofline); no-ops unless rn_prof_enable(1)
enum { RN_PROF_GEMM_TC = 0, RN_PROF_WGRAD_TC = 1, RN_PROF_GEMM_SIMT = 2, RN_PROF_CHAIN_TC = 3,
       // the non-GEMM kernels of a step (time only, no FLOP accounting)
       RN_PROF_ENCODE = 4, RN_PROF_HEADS_FWD = 5, RN_PROF_HEADS_BWD = 6, RN_PROF_IPE_GRAD = 7, RN_PROF_COLOR = 8,
       RN_PROF_GLUE = 9 /* pack / unpack / column sums / format conversions */, RN_PROF_RAY = 10 /* raymarch.cu entry points */,
       RN_PROF_NUM = 11 };
// algo_flops: algorithmic (dense, unpadded, 2*MAC) FLOPs of the launch; exec_flops: FLOPs the tensor pipe actually executes
// for it (padded shapes x the number of MMAs per K step of the arithmetic mode)
void rn_prof_begin(int cls, cudaStream_t st, double algo_flops, double exec_flops = 0.0);
void rn_prof_end(int cls, cudaStream_t st);
struct RnProfScope {   // brackets everything launched on `st` during its lifetime
  int cls;
  cudaStream_t st;
  RnProfScope(int c, cudaStream_t s) : cls(c), st(s) { rn_prof_begin(c, s, 0.0); }
  ~RnProfScope() { rn_prof_end(cls, st); }
};

// ------------------------------------------------------------------------------------------
// Activation buffers: [rows, ld] matrices that feed / leave the GEMMs.
//   RN_PREC_FP32   : hi = float*            (exact-parity SIMT path)
//   RN_PREC_BF16   : hi = bf16*             (throughput tensor path)
//   RN_PREC_BF16X3 : hi, lo = bf16*         (split-bf16 tensor path: x ~= hi + lo, 16-bit mantissa)
//   RN_PREC_FP16   : hi = fp16*             (weights, activations and scaled gradient tiles of the fp16 mode)
// ------------------------------------------------------------------------------------------
struct ActBuf {
  void* hi;
  void* lo;
  int ld;  // elements
};

__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }

__device__ __forceinline__ uint16_t float_to_bf16_bits(float x) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}

// two floats -> packed bf16x2 (one F2FP.PACK instruction instead of two scalar F2F conversions)
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// two floats -> packed fp16x2 (saturating: a finite activation never becomes inf)
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ uint16_t float_to_f16_bits(float x) { return (uint16_t)(pack_f16x2(x, 0.f) & 0xffffu); }

// pack two floats into hi (and lo residual) bf16 pairs
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(a, b);
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
  lo = pack_bf16x2(ra, rb);
}

// store 8 consecutive columns (col % 8 == 0) of one row
template <int PREC>
__device__ __forceinline__ void act_store8(const ActBuf& b, size_t row, int col, const float* v) {
  if (PREC == RN_PREC_FP32) {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(b.hi) + row * b.ld + col);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else if (PREC == RN_PREC_FP16) {
    uint32_t h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = pack_f16x2(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(b.hi) + row * b.ld + col) = make_uint4(h[0], h[1], h[2], h[3]);
  } else {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(b.hi) + row * b.ld + col) = make_uint4(h[0], h[1], h[2], h[3]);
    if (PREC == RN_PREC_BF16X3)
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(b.lo) + row * b.ld + col) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// load 8 consecutive columns of one row as floats (hi + lo)
template <int PREC>
__device__ __forceinline__ void act_load8(const ActBuf& b, size_t row, int col, float* v) {
  if (PREC == RN_PREC_FP32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(b.hi) + row * b.ld + col);
    float4 a = p[0], c = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
  } else if (PREC == RN_PREC_FP16) {
    uint4 h = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(b.hi) + row * b.ld + col);
    uint32_t hh[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack_f16x2(hh[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  } else {
    uint4 h = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(b.hi) + row * b.ld + col);
    uint32_t hh[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(hh[i] << 16);
      v[2 * i + 1] = __uint_as_float(hh[i] & 0xffff0000u);
    }
    if (PREC == RN_PREC_BF16X3) {
      uint4 l = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(b.lo) + row * b.ld + col);
      uint32_t ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[2 * i] += __uint_as_float(ll[i] << 16);
        v[2 * i + 1] += __uint_as_float(ll[i] & 0xffff0000u);
      }
    }
  }
}

// sign bits only (relu mask): positive iff hi > 0 (lo never flips the sign of a nonzero hi)
template <int PREC>
__device__ __forceinline__ void act_load8_hi(const ActBuf& b, size_t row, int col, float* v) {
  if (PREC == RN_PREC_FP32) {
    act_load8<RN_PREC_FP32>(b, row, col, v);
  } else if (PREC == RN_PREC_FP16) {
    act_load8<RN_PREC_FP16>(b, row, col, v);
  } else {
    act_load8<RN_PREC_BF16>(b, row, col, v);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RN_FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(RN_FULL, v, o));
  return v;
}
// inclusive prefix sum across the warp
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(RN_FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// inclusive running max across the warp (exact, so any evaluation order gives the same result)
__device__ __forceinline__ float warp_scan_max(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(RN_FULL, v, o);
    if (lane >= o) v = fmaxf(v, t);
  }
  return v;
}

__device__ __forceinline__ float softplus_f(float x) {
  // torch.nn.functional.softplus, beta=1, threshold=20
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
