// K1 (ray casting + integrated positional encoding), K2 (heads activations, reflection, integrated
// directional encoding, n.v), colour combine, and the small glue kernels of the NerfMLP pipeline.
// Reference semantics: render.py:22-129, coord.py:102-133, math.py:22-34, ref_utils.py:22-161,
// models.py:582-729, image.py:51-59.  fp32 op order follows the reference where it matters for
// bit-level agreement of the IPE phase (mean * 2^15 amplifies one ulp of `mean` to ~0.03 rad).
#include "pointwise.cuh"

#include "ide_coeffs.inc"

namespace rn {
namespace {

#define DISPATCH_PREC(prec, ...)                                  \
  switch (prec) {                                                 \
    case RN_PREC_FP32: { constexpr int PREC = RN_PREC_FP32; __VA_ARGS__; } break;     \
    case RN_PREC_BF16: { constexpr int PREC = RN_PREC_BF16; __VA_ARGS__; } break;     \
    case RN_PREC_BF16X3: { constexpr int PREC = RN_PREC_BF16X3; __VA_ARGS__; } break; \
    case RN_PREC_FP16: { constexpr int PREC = RN_PREC_FP16; __VA_ARGS__; } break;     \
    default: return rn_set_error(RN_ERR_ARG, "bad precision");    \
  }

// ------------------------------------------------------------------------------------------
// conical frustum -> lifted mean / variance (render.py:64-80, 22-43; coord.py:129-133 with the
// octahedron-1 basis = anti-diagonal -1: lifted mean = (-z,-y,-x), lifted var = (cov_zz,cov_yy,cov_xx))
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void frustum_lift(float t0, float t1, const float o[3], const float d[3], float radius,
                                             float lm[3], float lv[3]) {
  const float mu = __fdiv_rn(__fadd_rn(t0, t1), 2.f);
  const float hw = __fdiv_rn(__fsub_rn(t1, t0), 2.f);
  const float mu2 = __fmul_rn(mu, mu), hw2 = __fmul_rn(hw, hw);
  const float denom = fmaxf(RN_EPS32, __fadd_rn(__fmul_rn(3.f, mu2), hw2));
  const float t_mean = __fadd_rn(mu, __fdiv_rn(__fmul_rn(__fmul_rn(2.f, mu), hw2), denom));
  const float hw4 = __fmul_rn(hw2, hw2);
  const float c415 = (float)(4.0 / 15.0), c512 = (float)(5.0 / 12.0);
  const float t_var = __fsub_rn(
      __fdiv_rn(hw2, 3.f),
      __fdiv_rn(__fmul_rn(__fmul_rn(c415, hw4), __fsub_rn(__fmul_rn(12.f, mu2), hw2)), __fmul_rn(denom, denom)));
  float r_var = __fsub_rn(__fadd_rn(__fdiv_rn(mu2, 4.f), __fmul_rn(c512, hw2)), __fdiv_rn(__fmul_rn(c415, hw4), denom));
  r_var = __fmul_rn(r_var, __fmul_rn(radius, radius));
  const float dmag = fmaxf(1e-10f, __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float mean = __fadd_rn(__fmul_rn(d[c], t_mean), o[c]);
    const float cov = __fadd_rn(__fmul_rn(t_var, __fmul_rn(d[c], d[c])),
                                __fmul_rn(r_var, __fsub_rn(1.f, __fmul_rn(d[c], __fdiv_rn(d[c], dmag)))));
    lm[2 - c] = -mean;
    lv[2 - c] = cov;
  }
}

// math.py:22-34: x if |x| < 100pi else python-sign remainder(x, 100pi)
__device__ __forceinline__ float safe_arg(float x) {
  const float t = 314.159271f;  // float32(100*pi)
  if (fabsf(x) < t) return x;
  float r = fmodf(x, t);
  if (r != 0.f && r < 0.f) r += t;
  return r;
}

// sin / cos of an argument already reduced by safe_arg (|x| <= 100 pi): two-constant Cody-Waite reduction to [-pi, pi]
// (error < 1e-7 for |x| < 320) followed by the hardware approximation, whose absolute error on [-pi, pi] is 2^-21.4
// (CUDA math API) -- together ~5e-7, inside the 2e-6 the IPE features are gated at, at a third of sinf's instructions.
__device__ __forceinline__ float reduce_2pi(float x) {
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(-k, 6.2831854820251465f, x);         // float32(2 pi)
  return fmaf(-k, -1.7484555e-07f, r);                 // 2 pi - float32(2 pi)
}
__device__ __forceinline__ float sin_reduced(float x) { return __sinf(reduce_2pi(x)); }
__device__ __forceinline__ float cos_reduced(float x) { return __cosf(reduce_2pi(x)); }
// exp of a non-positive argument (the IPE attenuation): ex2.approx, relative error ~2^-21 near 0, absolute error
// negligible for large |x|
__device__ __forceinline__ float exp_neg(float x) { return __expf(x); }

constexpr int kEncRows = 64;       // rows per block (4 threads per row)
constexpr int kEncLd = 129;

template <int PREC>
__global__ void __launch_bounds__(256)
encode_kernel(const float* __restrict__ tdist, const float* __restrict__ origins, const float* __restrict__ dirs,
              const float* __restrict__ radii, int s, int64_t row0, int64_t rows, ActBuf out, int ncols) {
  __shared__ float tile[kEncRows * kEncLd];
  const int tid = threadIdx.x;
  const int r = tid >> 2, q = tid & 3;
  const int64_t lrow = (int64_t)blockIdx.x * kEncRows + r;  // row within [0, rows)
  if (lrow < rows) {
    const int64_t row = row0 + lrow;
    const int64_t ray = (row >> 31) ? row / s : (int64_t)((uint32_t)row / (uint32_t)s);   // 32-bit division when it fits
    const int smp = (int)(row - ray * s);
    const float t0 = tdist[ray * (s + 1) + smp], t1 = tdist[ray * (s + 1) + smp + 1];
    float o[3], d[3], lm[3], lv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] = origins[ray * 3 + c];
      d[c] = dirs[ray * 3 + c];
    }
    frustum_lift(t0, t1, o, d, radii[ray], lm, lv);
    float* trow = tile + r * kEncLd;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = q * 4 + kk;
      const float sc = (float)(1u << k);
      const float sc2 = sc * sc;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float a = __fmul_rn(lm[c], sc);
        const float e = exp_neg(__fmul_rn(-0.5f, __fmul_rn(lv[c], sc2)));
        trow[k * 3 + c] = __fmul_rn(e, sin_reduced(safe_arg(a)));
        trow[48 + k * 3 + c] = __fmul_rn(e, sin_reduced(safe_arg(__fadd_rn(a, 1.57079637f))));
      }
    }
  }
  __syncthreads();
  const int ng = ncols >> 3;
  for (int g = tid; g < kEncRows * ng; g += 256) {
    const int rr = g / ng, cg = g - rr * ng;
    const int64_t lr = (int64_t)blockIdx.x * kEncRows + rr;
    if (lr >= rows) continue;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = cg * 8 + e;
      v[e] = (col < 96) ? tile[rr * kEncLd + col] : 0.f;
    }
    act_store8<PREC>(out, (size_t)lr, cg * 8, v);
  }
}

// d raw_density / d means from d raw_density / d features (closed form), then
// normals = -l2_normalize(.)  (models.py:603-609, ref_utils.py:40-42)
__global__ void __launch_bounds__(256)
ipe_grad_normals_kernel(const float* __restrict__ gx0, const float* __restrict__ gx0b, int ld, const float* __restrict__ tdist,
                        const float* __restrict__ origins, const float* __restrict__ dirs,
                        const float* __restrict__ radii, int s, int64_t row0, int64_t rows,
                        float* __restrict__ normals_out, float gscale) {
  __shared__ float tile[kEncRows * 97];
  const int tid = threadIdx.x;
  for (int g = tid; g < kEncRows * 24; g += 256) {  // 24 float4 per row
    const int rr = g / 24, c4 = g - rr * 24;
    const int64_t lr = (int64_t)blockIdx.x * kEncRows + rr;
    if (lr >= rows) continue;
    float4 v = *reinterpret_cast<const float4*>(gx0 + (size_t)lr * ld + c4 * 4);
    if (gx0b) {   // second partial sum (the skip layer's share of the gradient), kept apart to avoid a read-modify-write
      const float4 u = *reinterpret_cast<const float4*>(gx0b + (size_t)lr * ld + c4 * 4);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    float* t = tile + rr * 97 + c4 * 4;
    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
  }
  __syncthreads();
  const int r = tid >> 2, q = tid & 3;
  const int64_t lrow = (int64_t)blockIdx.x * kEncRows + r;
  float dl[3] = {0.f, 0.f, 0.f};
  if (lrow < rows) {
    const int64_t row = row0 + lrow;
    const int64_t ray = row / s;
    const int smp = (int)(row - ray * s);
    const float t0 = tdist[ray * (s + 1) + smp], t1 = tdist[ray * (s + 1) + smp + 1];
    float o[3], d[3], lm[3], lv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] = origins[ray * 3 + c];
      d[c] = dirs[ray * 3 + c];
    }
    frustum_lift(t0, t1, o, d, radii[ray], lm, lv);
    const float* trow = tile + r * 97;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = q * 4 + kk;
      const float sc = (float)(1u << k);
      const float sc2 = sc * sc;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float a = __fmul_rn(lm[c], sc);
        const float e = exp_neg(__fmul_rn(-0.5f, __fmul_rn(lv[c], sc2)));
        const float gs = trow[k * 3 + c], gc = trow[48 + k * 3 + c];
        dl[c] += sc * e * (gs * cos_reduced(safe_arg(a)) + gc * cos_reduced(safe_arg(__fadd_rn(a, 1.57079637f))));
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    dl[c] += __shfl_xor_sync(RN_FULL, dl[c], 1);
    dl[c] += __shfl_xor_sync(RN_FULL, dl[c], 2);
  }
  if (lrow < rows && q == 0) {
    // d mean = basis * d lifted_mean = (-dl[2], -dl[1], -dl[0]);  normals = -g / sqrt(max(|g|^2, eps))
    const float gx = -dl[2] * gscale, gy = -dl[1] * gscale, gz = -dl[0] * gscale;   // gscale: exact power of two
    const float inv = 1.f / sqrtf(fmaxf(gx * gx + gy * gy + gz * gz, RN_EPS32));
    float* o = normals_out + (size_t)lrow * 3;
    o[0] = -gx * inv;
    o[1] = -gy * inv;
    o[2] = -gz * inv;
  }
}

template <int PREC>
__global__ void __launch_bounds__(256)
density_grad_seed_kernel(ActBuf a8, const float* __restrict__ wd, ActBuf out, int64_t rows) {
  const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;  // group of 8 columns
  const int64_t row = g >> 5;
  const int col = (int)(g & 31) * 8;
  if (row >= rows) return;
  float a[8], v[8];
  act_load8_hi<PREC>(a8, (size_t)row, col, a);
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = a[e] > 0.f ? wd[col + e] : 0.f;
  act_store8<PREC>(out, (size_t)row, col, v);
}

// ------------------------------------------------------------------------------------------
// integrated directional encoding (ref_utils.py:98-161), deg_view = 5.
// Polynomials in z are evaluated by Horner in fp64 from the reference's fp32-rounded coefficients:
// the l=16 band has coefficients up to 9e4 with cancellation, and the reference's own fp32
// Vandermonde-then-dot evaluation deviates from exact by up to 1e-2 there.
// ------------------------------------------------------------------------------------------
template <bool GRAD>
__device__ __forceinline__ void ide_core(double x, double y, double z, double kinv, float* out, int out_stride,
                                         const float* g, int g_stride, double& dx, double& dy, double& dz, double& dk) {
  int pair = 0;
#pragma unroll
  for (int li = 0; li < 5; ++li) {
    const int l = 1 << li;
    const double sigma = 0.5 * l * (l + 1);
    const double att = exp(-sigma * kinv);
    double pr = 1.0, pi = 0.0;    // (x+iy)^m
    double qr = 0.0, qi = 0.0;    // m (x+iy)^(m-1)
#pragma unroll
    for (int m = 0; m <= l; ++m) {
      double P = kIdeCoef[pair][l - m];
      double dP = 0.0;
#pragma unroll
      for (int k = l - m - 1; k >= 0; --k) {
        if (GRAD) dP = dP * z + P;
        P = P * z + kIdeCoef[pair][k];
      }
      const double pa = P * att;
      if (!GRAD) {
        out[pair * out_stride] = (float)(pr * pa);
        out[(36 + pair) * out_stride] = (float)(pi * pa);
      } else {
        const double gr = g[pair * g_stride], gi = g[(36 + pair) * g_stride];
        const double s = gr * pr + gi * pi;
        dk -= sigma * s * pa;
        dz += s * dP * att;
        dx += (gr * qr + gi * qi) * pa;
        dy += (gi * qr - gr * qi) * pa;
      }
      // advance powers: q_{m+1} = (m+1) p_m ; p_{m+1} = p_m (x+iy)
      qr = (m + 1) * pr;
      qi = (m + 1) * pi;
      const double nr = pr * x - pi * y;
      pi = pr * y + pi * x;
      pr = nr;
      ++pair;
    }
  }
}

constexpr int kProRows = 128;
constexpr int kProLd = 77;  // 73 values per row, odd stride

struct HeadsFwd {
  float density, rough, tint[3], np[3], nrm, g[3], dot, refd[3];
};

__device__ __forceinline__ void heads_forward(const float* hr, const float vd[3], const MlpScalars& sc, HeadsFwd& h) {
  h.density = softplus_f(hr[0] + sc.density_bias);                  // models.py:623
  h.g[0] = hr[1]; h.g[1] = hr[2]; h.g[2] = hr[3];
  h.nrm = sqrtf(fmaxf(h.g[0] * h.g[0] + h.g[1] * h.g[1] + h.g[2] * h.g[2], RN_EPS32));
#pragma unroll
  for (int c = 0; c < 3; ++c) h.np[c] = -(h.g[c] / h.nrm);          // models.py:615
  h.rough = softplus_f(hr[4] + sc.roughness_bias);                  // models.py:640-641
#pragma unroll
  for (int c = 0; c < 3; ++c) h.tint[c] = sigmoid_f(hr[8 + c]);     // models.py:637
  // reflect(-viewdirs, normals_pred) (ref_utils.py:22-37, models.py:662-663)
  const float ndv = -(h.np[0] * vd[0] + h.np[1] * vd[1] + h.np[2] * vd[2]);  // n . (-v)
#pragma unroll
  for (int c = 0; c < 3; ++c) h.refd[c] = 2.f * ndv * h.np[c] + vd[c];
  h.dot = h.np[0] * vd[0] + h.np[1] * vd[1] + h.np[2] * vd[2];      // models.py:680-682
}

template <int PREC>
__global__ void __launch_bounds__(kProRows)
heads_prologue_fwd_kernel(const float* __restrict__ heads_raw, const float* __restrict__ viewdirs, int s, int64_t row0,
                          int64_t rows, MlpScalars sc, ActBuf v0, float* __restrict__ density,
                          float* __restrict__ normals_pred, float* __restrict__ grad_pred,
                          float* __restrict__ roughness, float* __restrict__ tint) {
  __shared__ float tile[kProRows * kProLd];
  const int tid = threadIdx.x;
  const int64_t lrow = (int64_t)blockIdx.x * kProRows + tid;
  if (lrow < rows) {
    const int64_t ray = (row0 + lrow) / s;
    float hr[16];
    const float4* hp = reinterpret_cast<const float4*>(heads_raw + (size_t)lrow * 16);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float4 v = hp[i];
      hr[4 * i] = v.x; hr[4 * i + 1] = v.y; hr[4 * i + 2] = v.z; hr[4 * i + 3] = v.w;
    }
    const float vd[3] = {viewdirs[ray * 3], viewdirs[ray * 3 + 1], viewdirs[ray * 3 + 2]};
    HeadsFwd h;
    heads_forward(hr, vd, sc, h);
    density[lrow] = h.density;
    roughness[lrow] = h.rough;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      normals_pred[(size_t)lrow * 3 + c] = h.np[c];
      grad_pred[(size_t)lrow * 3 + c] = h.g[c];
      tint[(size_t)lrow * 3 + c] = h.tint[c];
    }
    double dx, dy, dz, dk;
    float* trow = tile + tid * kProLd;
    ide_core<false>(h.refd[0], h.refd[1], h.refd[2], h.rough, trow, 1, nullptr, 0, dx, dy, dz, dk);
    trow[72] = h.dot;
  }
  __syncthreads();
  // v0 columns 128..255: [ide(72) | n.v | zeros]
  for (int g = tid; g < kProRows * 16; g += kProRows) {
    const int rr = g >> 4, cg = g & 15;
    const int64_t lr = (int64_t)blockIdx.x * kProRows + rr;
    if (lr >= rows) continue;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = cg * 8 + e;
      v[e] = col < 73 ? tile[rr * kProLd + col] : 0.f;
    }
    act_store8<PREC>(v0, (size_t)lr, 128 + cg * 8, v);
  }
}

template <int PREC>
__global__ void __launch_bounds__(kProRows)
heads_prologue_bwd_kernel(const float* __restrict__ heads_raw, const float* __restrict__ viewdirs, int s, int64_t row0,
                          int64_t rows, MlpScalars sc, const float* __restrict__ dv0f, const float* __restrict__ dcolor,
                          const float* __restrict__ g_density, const float* __restrict__ g_normals_pred,
                          const float* __restrict__ g_grad_pred, const float* __restrict__ g_roughness,
                          const float* __restrict__ g_tint, ActBuf d_scal, const float* __restrict__ dv0_unscale) {
  // the 73 gradient values of a row (d IDE, d n.v) are staged through shared memory with coalesced reads: one
  // thread per row reading its own 1 KB-strided row would touch 32 different rows per load instruction
  __shared__ float tile[kProRows * kProLd];
  for (int g = threadIdx.x; g < kProRows * 19; g += kProRows) {   // 19 float4 = columns 128..203 of a row
    const int rr = g / 19, c4 = g - rr * 19;
    const int64_t lr = (int64_t)blockIdx.x * kProRows + rr;
    if (lr >= rows) continue;
    const float4 v = __ldg(reinterpret_cast<const float4*>(dv0f + (size_t)lr * 256 + 128) + c4);
    float* t = tile + rr * kProLd + c4 * 4;
    t[0] = v.x;
    if (c4 < 18) { t[1] = v.y; t[2] = v.z; t[3] = v.w; }          // c4 == 18: only column 200 (= value 72) is used
  }
  __syncthreads();
  const int64_t lrow = (int64_t)blockIdx.x * kProRows + threadIdx.x;
  if (lrow >= rows) return;
  const float us = dv0_unscale ? *dv0_unscale : 1.f;   // fp16 mode: dv0f carries the view chain's power-of-two scale
  const int64_t ray = (row0 + lrow) / s;
  float hr[16];
  const float4* hp = reinterpret_cast<const float4*>(heads_raw + (size_t)lrow * 16);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float4 v = hp[i];
    hr[4 * i] = v.x; hr[4 * i + 1] = v.y; hr[4 * i + 2] = v.z; hr[4 * i + 3] = v.w;
  }
  const float vd[3] = {viewdirs[ray * 3], viewdirs[ray * 3 + 1], viewdirs[ray * 3 + 2]};
  HeadsFwd h;
  heads_forward(hr, vd, sc, h);
  // IDE backward
  const float* gide = tile + threadIdx.x * kProLd;
  double dx = 0, dy = 0, dz = 0, dk = 0;
  ide_core<true>(h.refd[0], h.refd[1], h.refd[2], h.rough, nullptr, 0, gide, 1, dx, dy, dz, dk);
  const float dref[3] = {(float)dx * us, (float)dy * us, (float)dz * us};
  const float ddot = gide[72] * us;
  dk *= (double)us;
  // d normals_pred: upstream + n.v feature + reflection (r = 2 (n.v') n - v', v' = -viewdirs)
  const float ndv = -(h.np[0] * vd[0] + h.np[1] * vd[1] + h.np[2] * vd[2]);
  const float dr_dot_n = dref[0] * h.np[0] + dref[1] * h.np[1] + dref[2] * h.np[2];
  float dnp[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float up = g_normals_pred ? g_normals_pred[(size_t)lrow * 3 + c] : 0.f;
    dnp[c] = up + ddot * vd[c] + 2.f * (dr_dot_n * (-vd[c]) + ndv * dref[c]);
  }
  // np = -g / nrm, nrm = sqrt(max(|g|^2, eps))
  float dg[3];
  const float gg = h.g[0] * h.g[0] + h.g[1] * h.g[1] + h.g[2] * h.g[2];
  if (gg > RN_EPS32) {
    const float pd = h.np[0] * dnp[0] + h.np[1] * dnp[1] + h.np[2] * dnp[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) dg[c] = -(dnp[c] - h.np[c] * pd) / h.nrm;
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) dg[c] = -dnp[c] / h.nrm;
  }
  float out[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) out[i] = 0.f;
  {
    const float x = hr[0] + sc.density_bias;
    out[0] = (g_density ? g_density[lrow] : 0.f) * (x > 20.f ? 1.f : sigmoid_f(x));
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) out[1 + c] = dg[c] + (g_grad_pred ? g_grad_pred[(size_t)lrow * 3 + c] : 0.f);
  {
    const float x = hr[4] + sc.roughness_bias;
    out[4] = ((g_roughness ? g_roughness[lrow] : 0.f) + (float)dk) * (x > 20.f ? 1.f : sigmoid_f(x));
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    out[5 + c] = dcolor[(size_t)lrow * 8 + c];
    const float dt = dcolor[(size_t)lrow * 8 + 3 + c] + (g_tint ? g_tint[(size_t)lrow * 3 + c] : 0.f);
    out[8 + c] = dt * h.tint[c] * (1.f - h.tint[c]);
  }
  act_store8<PREC>(d_scal, (size_t)lrow, 0, out);
  act_store8<PREC>(d_scal, (size_t)lrow, 8, out + 8);
}

// ------------------------------------------------------------------------------------------
// colour combine (models.py:699-729, image.py:51-59)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float srgb_f(float x) {
  return x <= 0.0031308f ? (float)(323.0 / 25.0) * x : (211.f * powf(fmaxf(RN_EPS32, x), (float)(5.0 / 12.0)) - 11.f) / 200.f;
}
__device__ __forceinline__ float srgb_grad(float x) {
  if (x <= 0.0031308f) return (float)(323.0 / 25.0);
  if (x <= RN_EPS32) return 0.f;
  return (211.f / 200.f) * (float)(5.0 / 12.0) * powf(x, (float)(5.0 / 12.0 - 1.0));
}
__device__ __forceinline__ float clip01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
__device__ __forceinline__ float clip01_grad(float x) { return (x >= 0.f && x <= 1.f) ? 1.f : 0.f; }

struct ColorFwd {
  float rgbs[3], dl[3], tint[3], spec[3], lin[3], nrm;
  int amax;
};

__device__ __forceinline__ void color_forward(const float* rr, const float* hr, const MlpScalars& sc, ColorFwd& c) {
  c.nrm = 1.f;
  c.amax = 0;
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    c.rgbs[k] = sigmoid_f(sc.rgb_premultiplier * rr[k] + sc.rgb_bias);  // models.py:699-700
    c.dl[k] = sigmoid_f(hr[5 + k] - 1.09861229f);                       // models.py:705-706 (log 3 in fp32)
    c.tint[k] = sigmoid_f(hr[8 + k]);
    c.spec[k] = c.tint[k] * c.rgbs[k];                                  // models.py:708
    c.lin[k] = c.spec[k] + c.dl[k];
    if (c.lin[k] > mx) { mx = c.lin[k]; c.amax = k; }
  }
  if (sc.srgb_mapping && sc.srgb_normalization) c.nrm = fmaxf(mx, 1.f);  // models.py:718-720
}

__global__ void __launch_bounds__(256)
color_fwd_kernel(const float* __restrict__ rgb_raw, const float* __restrict__ heads_raw, int64_t rows, MlpScalars sc,
                 float* __restrict__ rgb, float* __restrict__ diffuse, float* __restrict__ specular) {
  const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (row >= rows) return;
  const float4 r4 = *reinterpret_cast<const float4*>(rgb_raw + (size_t)row * 4);
  const float rr[3] = {r4.x, r4.y, r4.z};
  float hr[16];
  const float4* hp = reinterpret_cast<const float4*>(heads_raw + (size_t)row * 16);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float4 v = hp[i];
    hr[4 * i] = v.x; hr[4 * i + 1] = v.y; hr[4 * i + 2] = v.z; hr[4 * i + 3] = v.w;
  }
  ColorFwd c;
  color_forward(rr, hr, sc, c);
  const float pad = sc.rgb_padding;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float o, d, sp;
    if (sc.srgb_mapping) {
      o = clip01(srgb_f(c.lin[k] / c.nrm));
      d = clip01(srgb_f(c.dl[k]));
      sp = clip01(srgb_f(c.spec[k]));
    } else {
      o = c.lin[k];
      d = c.dl[k];
      sp = c.spec[k];
    }
    rgb[(size_t)row * 3 + k] = o * (1.f + 2.f * pad) - pad;  // models.py:729
    diffuse[(size_t)row * 3 + k] = d;
    specular[(size_t)row * 3 + k] = sp;
  }
}

template <int PREC>
__global__ void __launch_bounds__(256)
color_bwd_kernel(const float* __restrict__ rgb_raw, const float* __restrict__ heads_raw, int64_t rows, MlpScalars sc,
                 const float* __restrict__ g_rgb, const float* __restrict__ g_diffuse,
                 const float* __restrict__ g_specular, ActBuf d_rgb_raw, float* __restrict__ dcolor) {
  const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (row >= rows) return;
  const float4 r4 = *reinterpret_cast<const float4*>(rgb_raw + (size_t)row * 4);
  const float rr[3] = {r4.x, r4.y, r4.z};
  float hr[16];
  const float4* hp = reinterpret_cast<const float4*>(heads_raw + (size_t)row * 16);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float4 v = hp[i];
    hr[4 * i] = v.x; hr[4 * i + 1] = v.y; hr[4 * i + 2] = v.z; hr[4 * i + 3] = v.w;
  }
  ColorFwd c;
  color_forward(rr, hr, sc, c);
  const float pad = sc.rgb_padding;
  float dlin[3], ddl[3], dspec[3];
  float dy[3];
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float go = (g_rgb ? g_rgb[(size_t)row * 3 + k] : 0.f) * (1.f + 2.f * pad);
    const float gd = g_diffuse ? g_diffuse[(size_t)row * 3 + k] : 0.f;
    const float gs = g_specular ? g_specular[(size_t)row * 3 + k] : 0.f;
    if (sc.srgb_mapping) {
      const float y = c.lin[k] / c.nrm;
      dy[k] = go * clip01_grad(srgb_f(y)) * srgb_grad(y);
      dot += dy[k] * c.lin[k];
      ddl[k] = gd * clip01_grad(srgb_f(c.dl[k])) * srgb_grad(c.dl[k]);
      dspec[k] = gs * clip01_grad(srgb_f(c.spec[k])) * srgb_grad(c.spec[k]);
    } else {
      dy[k] = go;
      ddl[k] = gd;
      dspec[k] = gs;
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    dlin[k] = dy[k] / c.nrm;
    if (sc.srgb_mapping && sc.srgb_normalization && c.nrm > 1.f && k == c.amax) dlin[k] -= dot / (c.nrm * c.nrm);
  }
  float out[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float draw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float dsp = dlin[k] + dspec[k];
    const float ddf = dlin[k] + ddl[k];
    out[k] = ddf * c.dl[k] * (1.f - c.dl[k]);                                    // d raw_rgb_diffuse
    out[3 + k] = dsp * c.rgbs[k];                                                // d tint (post-sigmoid)
    draw[k] = dsp * c.tint[k] * c.rgbs[k] * (1.f - c.rgbs[k]) * sc.rgb_premultiplier;  // d rgb head output
  }
  float4* dc = reinterpret_cast<float4*>(dcolor + (size_t)row * 8);
  dc[0] = make_float4(out[0], out[1], out[2], out[3]);
  dc[1] = make_float4(out[4], out[5], out[6], out[7]);
  float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  act_store8<PREC>(d_rgb_raw, (size_t)row, 0, draw);
  act_store8<PREC>(d_rgb_raw, (size_t)row, 8, z);
}

// ------------------------------------------------------------------------------------------
// glue
// ------------------------------------------------------------------------------------------
template <int PREC>
__global__ void __launch_bounds__(256)
colsum_kernel(ActBuf a, int64_t rows, int ncols, float* __restrict__ out) {
  // block: 128 rows; thread -> (column group of 8, row lane); warps read 32 adjacent groups of one row
  __shared__ float acc[256];
  const int ng = ncols >> 3;                 // <= 32
  const int cg = threadIdx.x % ng;
  const int rl = threadIdx.x / ng;
  const int nrl = 256 / ng;
  for (int i = threadIdx.x; i < ncols; i += 256) acc[i] = 0.f;
  __syncthreads();
  float sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int64_t base = (int64_t)blockIdx.x * 128;
  if (rl < nrl)
    for (int r = rl; r < 128; r += nrl) {
      const int64_t row = base + r;
      if (row >= rows) break;
      float v[8];
      act_load8<PREC>(a, (size_t)row, cg * 8, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum[e] += v[e];
    }
  if (rl < nrl) {
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&acc[cg * 8 + e], sum[e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ncols; i += 256) atomicAdd(&out[i], acc[i]);
}

template <int PREC>
__global__ void __launch_bounds__(256)
f32_to_act_kernel(const float* __restrict__ src, int ld, int c0, int64_t rows, int ncols, ActBuf dst) {
  const int ng = ncols >> 3;
  const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t row = g / ng;
  const int cg = (int)(g - row * ng);
  if (row >= rows) return;
  const float4* p = reinterpret_cast<const float4*>(src + (size_t)row * ld + c0 + cg * 8);
  const float4 a = p[0], b = p[1];
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  act_store8<PREC>(dst, (size_t)row, cg * 8, v);
}

template <int PREC>
__global__ void pack_segment_kernel(const float* __restrict__ src, int src_ld, int nr, int nc, int transpose,
                                    void* dst_hi, void* dst_lo, int dst_ld, int r0, int c0) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= nr * nc) return;
  const int i = idx / nc, j = idx - i * nc;
  const float v = src[(size_t)i * src_ld + j];
  const size_t d = transpose ? (size_t)(r0 + j) * dst_ld + (c0 + i) : (size_t)(r0 + i) * dst_ld + (c0 + j);
  if (PREC == RN_PREC_FP32) {
    reinterpret_cast<float*>(dst_hi)[d] = v;
  } else if (PREC == RN_PREC_FP16) {
    reinterpret_cast<uint16_t*>(dst_hi)[d] = float_to_f16_bits(v);
  } else {
    const uint16_t h = float_to_bf16_bits(v);
    reinterpret_cast<uint16_t*>(dst_hi)[d] = h;
    if (PREC == RN_PREC_BF16X3) reinterpret_cast<uint16_t*>(dst_lo)[d] = float_to_bf16_bits(v - bf16_bits_to_float(h));
  }
}

__global__ void unpack_add_kernel(const float* __restrict__ src, int src_ld, int r0, int c0, int nr, int nc,
                                  float* __restrict__ dst, int dst_ld, const float* __restrict__ scale) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= nr * nc) return;
  const int i = idx / nc, j = idx - i * nc;
  const float v = src[(size_t)(r0 + i) * src_ld + c0 + j];
  dst[(size_t)i * dst_ld + j] += scale ? v * *scale : v;
}

// Table-driven variants: ONE launch packs / unpacks every segment of a NerfMLP (blockIdx.y = segment) instead of one
// tiny launch per weight block -- ~90 launches per re-pack and ~50 per backward otherwise.
template <int PREC>
__global__ void __launch_bounds__(256)
pack_batch_kernel(const __grid_constant__ PackTable t) {
  const PackSeg& g = t.seg[blockIdx.y];
  const int total = g.nr * g.nc;
  for (int idx = blockIdx.x * 256 + threadIdx.x; idx < total; idx += gridDim.x * 256) {
    const int i = idx / g.nc, j = idx - i * g.nc;
    const float v = g.src[(size_t)i * g.src_ld + j];
    const size_t d = g.transpose ? (size_t)(g.r0 + j) * g.dst_ld + (g.c0 + i) : (size_t)(g.r0 + i) * g.dst_ld + (g.c0 + j);
    if (PREC == RN_PREC_FP32 || g.f32) {
      reinterpret_cast<float*>(g.dst_hi)[d] = v;
    } else if (PREC == RN_PREC_FP16) {
      reinterpret_cast<uint16_t*>(g.dst_hi)[d] = float_to_f16_bits(v);
    } else {
      const uint16_t h = float_to_bf16_bits(v);
      reinterpret_cast<uint16_t*>(g.dst_hi)[d] = h;
      if (PREC == RN_PREC_BF16X3 && g.dst_lo) reinterpret_cast<uint16_t*>(g.dst_lo)[d] = float_to_bf16_bits(v - bf16_bits_to_float(h));
    }
  }
}

__global__ void __launch_bounds__(256)
unpack_add_batch_kernel(const __grid_constant__ UnpackTable t) {
  const UnpackSeg& g = t.seg[blockIdx.y];
  const int total = g.nr * g.nc;
  const float sc = g.scale ? *g.scale : 1.f;
  for (int idx = blockIdx.x * 256 + threadIdx.x; idx < total; idx += gridDim.x * 256) {
    const int i = idx / g.nc, j = idx - i * g.nc;
    const float v = g.src[(size_t)(g.r0 + i) * g.src_ld + g.c0 + j];
    g.dst[(size_t)i * g.dst_ld + j] += g.scale ? v * sc : v;
  }
}

// ------------------------------------------------------------------------------------------
// fp16 mode: dynamic power-of-two scaling of the gradient tiles entering the dgrad chains.
// scal layout (floats / uint bit patterns): [0] amax view-chain seed, [1] amax scalar-head seed, [2] amax bottleneck
// gradient (view-chain scale) | [4] S1, [5] 1/S1, [6] S2, [7] 1/S2, [8] S2/S1
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_amax_commit(float m, uint32_t* dst) {
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(dst, __float_as_uint(m));   // non-negative floats order like uints
}
__global__ void __launch_bounds__(256)
amax_f32_kernel(const float* __restrict__ src, int64_t n4, uint32_t* __restrict__ dst) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  block_amax_commit(m, dst);
}
// fp16 [rows, ld] matrix, first ncols columns (ncols % 8 == 0)
__global__ void __launch_bounds__(256)
amax_f16_kernel(const uint16_t* __restrict__ src, int ld, int ncols, int64_t rows, uint32_t* __restrict__ dst) {
  const int ng = ncols >> 3;
  const int64_t total = rows * ng;
  float m = 0.f;
  for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < total; g += (int64_t)gridDim.x * 256) {
    const int64_t row = g / ng;
    const int cg = (int)(g - row * ng);
    const uint4 h = *reinterpret_cast<const uint4*>(src + (size_t)row * ld + cg * 8);
    const uint32_t hh[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack_f16x2(hh[i]);
      m = fmaxf(m, fmaxf(fabsf(f.x), fabsf(f.y)));
    }
  }
  block_amax_commit(m, dst);
}
// S = 2^e such that amax * S lies in [2^11, 2^12): 16x headroom below the fp16 maximum, 2^25 above its smallest normal
__device__ __forceinline__ float pow2_scale_for(float amax) {
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e;
  frexpf(amax, &e);   // amax = m * 2^e, m in [0.5, 1)
  int k = 12 - e;
  k = k < -60 ? -60 : (k > 60 ? 60 : k);
  return ldexpf(1.f, k);
}
__global__ void grad_scale_kernel(float* __restrict__ scal, int stage) {
  const uint32_t* a = reinterpret_cast<const uint32_t*>(scal);
  if (stage == 0) {
    const float s1 = pow2_scale_for(__uint_as_float(a[0]));
    scal[4] = s1;
    scal[5] = 1.f / s1;
  } else {
    const float s1inv = scal[5];
    const float s2 = pow2_scale_for(fmaxf(__uint_as_float(a[1]), __uint_as_float(a[2]) * s1inv));
    scal[6] = s2;
    scal[7] = 1.f / s2;
    scal[8] = s2 * s1inv;
  }
}
// dst fp16 [rows, ld_dst][:, 0:ncols] = src f32 [rows, ld_src][:, 0:ncols] * *scale
__global__ void __launch_bounds__(256)
scale_to_f16_kernel(const float* __restrict__ src, int ld_src, int64_t rows, int ncols, uint16_t* __restrict__ dst, int ld_dst,
                    const float* __restrict__ scale) {
  const int ng = ncols >> 3;
  const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t row = g / ng;
  const int cg = (int)(g - row * ng);
  if (row >= rows) return;
  const float sc = *scale;
  const float4* p = reinterpret_cast<const float4*>(src + (size_t)row * ld_src + cg * 8);
  const float4 a = p[0], b = p[1];
  *reinterpret_cast<uint4*>(dst + (size_t)row * ld_dst + cg * 8) =
      make_uint4(pack_f16x2(a.x * sc, a.y * sc), pack_f16x2(a.z * sc, a.w * sc), pack_f16x2(b.x * sc, b.y * sc),
                 pack_f16x2(b.z * sc, b.w * sc));
}
// in place: fp16 [rows, ld][:, 0:ncols] *= *ratio (a power of two: exact unless the result leaves the normal range)
__global__ void __launch_bounds__(256)
rescale_f16_kernel(uint16_t* __restrict__ buf, int ld, int ncols, int64_t rows, const float* __restrict__ ratio) {
  const int ng = ncols >> 3;
  const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t row = g / ng;
  const int cg = (int)(g - row * ng);
  if (row >= rows) return;
  const float r = *ratio;
  uint4* p = reinterpret_cast<uint4*>(buf + (size_t)row * ld + cg * 8);
  const uint4 h = *p;
  const uint32_t hh[4] = {h.x, h.y, h.z, h.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = unpack_f16x2(hh[i]);
    o[i] = pack_f16x2(f.x * r, f.y * r);
  }
  *p = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(128)
ide_kernel(const float* __restrict__ dirs, const float* __restrict__ kinv, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  double dx, dy, dz, dk;
  ide_core<false>(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2], kinv[i], out + i * 72, 1, nullptr, 0, dx, dy, dz, dk);
}

inline unsigned nblk(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

int launch_encode(int prec, const float* tdist, const float* origins, const float* dirs, const float* radii, int s,
                  int64_t row0, int64_t rows, ActBuf out, int ncols, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_ENCODE, st);
  if (rows <= 0) return RN_OK;
  DISPATCH_PREC(prec, (encode_kernel<PREC><<<nblk(rows, kEncRows), 256, 0, st>>>(tdist, origins, dirs, radii, s, row0, rows, out, ncols)));
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_ipe_grad_normals(const float* gx0, const float* gx0b, int ld, const float* tdist, const float* origins, const float* dirs,
                            const float* radii, int s, int64_t row0, int64_t rows, float* normals_out, float gscale,
                            cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_IPE_GRAD, st);
  if (rows <= 0) return RN_OK;
  ipe_grad_normals_kernel<<<nblk(rows, kEncRows), 256, 0, st>>>(gx0, gx0b, ld, tdist, origins, dirs, radii, s, row0, rows, normals_out, gscale);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_density_grad_seed(int prec, ActBuf a8, const float* wd, ActBuf out, int64_t rows, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_GLUE, st);
  if (rows <= 0) return RN_OK;
  DISPATCH_PREC(prec, (density_grad_seed_kernel<PREC><<<nblk(rows * 32, 256), 256, 0, st>>>(a8, wd, out, rows)));
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_heads_prologue_fwd(int prec, const float* heads_raw, const float* viewdirs, int s, int64_t row0, int64_t rows,
                              MlpScalars sc, ActBuf v0, float* density, float* normals_pred, float* grad_pred,
                              float* roughness, float* tint, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_HEADS_FWD, st);
  if (rows <= 0) return RN_OK;
  DISPATCH_PREC(prec, (heads_prologue_fwd_kernel<PREC><<<nblk(rows, kProRows), kProRows, 0, st>>>(
                          heads_raw, viewdirs, s, row0, rows, sc, v0, density, normals_pred, grad_pred, roughness, tint)));
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_heads_prologue_bwd(int prec, const float* heads_raw, const float* viewdirs, int s, int64_t row0, int64_t rows,
                              MlpScalars sc, const float* dv0f, const float* dcolor, const float* g_density,
                              const float* g_normals_pred, const float* g_grad_pred, const float* g_roughness,
                              const float* g_tint, ActBuf d_scal, const float* dv0_unscale, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_HEADS_BWD, st);
  if (rows <= 0) return RN_OK;
  DISPATCH_PREC(prec, (heads_prologue_bwd_kernel<PREC><<<nblk(rows, kProRows), kProRows, 0, st>>>(
                          heads_raw, viewdirs, s, row0, rows, sc, dv0f, dcolor, g_density, g_normals_pred, g_grad_pred,
                          g_roughness, g_tint, d_scal, dv0_unscale)));
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_color_fwd(const float* rgb_raw, const float* heads_raw, int64_t rows, MlpScalars sc, float* rgb,
                     float* diffuse, float* specular, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_COLOR, st);
  if (rows <= 0) return RN_OK;
  color_fwd_kernel<<<nblk(rows, 256), 256, 0, st>>>(rgb_raw, heads_raw, rows, sc, rgb, diffuse, specular);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_color_bwd(int prec, const float* rgb_raw, const float* heads_raw, int64_t rows, MlpScalars sc,
                     const float* g_rgb, const float* g_diffuse, const float* g_specular, ActBuf d_rgb_raw,
                     float* dcolor, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_COLOR, st);
  if (rows <= 0) return RN_OK;
  DISPATCH_PREC(prec, (color_bwd_kernel<PREC><<<nblk(rows, 256), 256, 0, st>>>(rgb_raw, heads_raw, rows, sc, g_rgb, g_diffuse,
                                                                          g_specular, d_rgb_raw, dcolor)));
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_colsum(int prec, ActBuf a, int64_t rows, int ncols, float* out, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_GLUE, st);
  if (rows <= 0) return RN_OK;
  if (ncols % 8 || ncols > 256) return rn_set_error(RN_ERR_ARG, "colsum: ncols must be a multiple of 8, <= 256");
  DISPATCH_PREC(prec, (colsum_kernel<PREC><<<nblk(rows, 128), 256, 0, st>>>(a, rows, ncols, out)));
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_f32_to_act(int prec, const float* src, int ld, int c0, int64_t rows, int ncols, ActBuf dst, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_GLUE, st);
  if (rows <= 0) return RN_OK;
  DISPATCH_PREC(prec, (f32_to_act_kernel<PREC><<<nblk(rows * (ncols >> 3), 256), 256, 0, st>>>(src, ld, c0, rows, ncols, dst)));
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_pack_segment(int prec, const float* src, int src_ld, int nr, int nc, int transpose, void* dst_hi, void* dst_lo,
                        int dst_ld, int r0, int c0, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_GLUE, st);
  DISPATCH_PREC(prec, (pack_segment_kernel<PREC><<<nblk((int64_t)nr * nc, 256), 256, 0, st>>>(src, src_ld, nr, nc, transpose, dst_hi,
                                                                                       dst_lo, dst_ld, r0, c0)));
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_unpack_add(const float* src, int src_ld, int r0, int c0, int nr, int nc, float* dst, int dst_ld,
                      const float* scale, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_GLUE, st);
  unpack_add_kernel<<<nblk((int64_t)nr * nc, 256), 256, 0, st>>>(src, src_ld, r0, c0, nr, nc, dst, dst_ld, scale);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_pack_batch(int prec, const PackTable& t, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_GLUE, st);
  if (t.n <= 0) return RN_OK;
  if (t.n > kMaxBatchSegs) return rn_set_error(RN_ERR_ARG, "pack_batch: too many segments");
  const dim3 grid(32, (unsigned)t.n);
  DISPATCH_PREC(prec, (pack_batch_kernel<PREC><<<grid, 256, 0, st>>>(t)));
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}
int launch_unpack_add_batch(const UnpackTable& t, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_GLUE, st);
  if (t.n <= 0) return RN_OK;
  if (t.n > kMaxBatchSegs) return rn_set_error(RN_ERR_ARG, "unpack_add_batch: too many segments");
  const dim3 grid(32, (unsigned)t.n);
  unpack_add_batch_kernel<<<grid, 256, 0, st>>>(t);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

static unsigned reduce_grid(int64_t items) {
  const int64_t b = (items + 255) / 256;
  return (unsigned)(b < 1 ? 1 : (b > 1184 ? 1184 : b));   // 8 CTAs per SM
}
int launch_amax_f32(const float* src, int64_t n, uint32_t* dst, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_GLUE, st);
  if (n <= 0) return RN_OK;
  if (n & 3) return rn_set_error(RN_ERR_ARG, "amax_f32: element count must be a multiple of 4");
  amax_f32_kernel<<<reduce_grid(n / 4), 256, 0, st>>>(src, n / 4, dst);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}
int launch_amax_f16(const void* src, int ld, int ncols, int64_t rows, uint32_t* dst, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_GLUE, st);
  if (rows <= 0) return RN_OK;
  amax_f16_kernel<<<reduce_grid(rows * (ncols >> 3)), 256, 0, st>>>(reinterpret_cast<const uint16_t*>(src), ld, ncols, rows, dst);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}
int launch_grad_scale(float* scal, int stage, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_GLUE, st);
  grad_scale_kernel<<<1, 1, 0, st>>>(scal, stage);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}
int launch_scale_to_f16(const float* src, int ld_src, int64_t rows, int ncols, void* dst, int ld_dst, const float* scale,
                        cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_GLUE, st);
  if (rows <= 0) return RN_OK;
  scale_to_f16_kernel<<<nblk(rows * (ncols >> 3), 256), 256, 0, st>>>(src, ld_src, rows, ncols, reinterpret_cast<uint16_t*>(dst),
                                                                     ld_dst, scale);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}
int launch_rescale_f16(void* buf, int ld, int ncols, int64_t rows, const float* ratio, cudaStream_t st) {
  RnProfScope prof_scope(RN_PROF_GLUE, st);
  if (rows <= 0) return RN_OK;
  rescale_f16_kernel<<<nblk(rows * (ncols >> 3), 256), 256, 0, st>>>(reinterpret_cast<uint16_t*>(buf), ld, ncols, rows, ratio);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_ide(const float* dirs, const float* kappa_inv, int64_t n, float* out, cudaStream_t st) {
  if (n <= 0) return RN_OK;
  ide_kernel<<<nblk(n, 128), 128, 0, st>>>(dirs, kappa_inv, n, out);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

}  // namespace rn
