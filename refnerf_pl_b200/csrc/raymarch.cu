// K4: warp-per-ray interval resampling, alpha compositing (fwd/bwd) and the step-function losses.
//
// One warp owns one ray.  The ray's fenceposts / CDF / weights are staged in shared memory,
// prefix sums and products run as warp-shuffle scans, the inverse-CDF lookup is a per-lane binary
// search over the staged CDF.  All global traffic is 128-byte coalesced row segments.
// Reference semantics: internal/stepfun.py, internal/math.py:88-111, internal/render.py:132-254.
#include "common.cuh"

namespace {

constexpr int kWarps = 4;  // warps (= rays) per block

__device__ __forceinline__ float nan_to_zero_clip01(float r) {
  if (r != r) r = 0.f;                  // torch.nan_to_num(x, 0)
  return fminf(fmaxf(r, 0.f), 1.f);     // +-inf -> clip
}

// inclusive prefix sum across the warp in fp64 (the CDF accumulator, see below)
__device__ __forceinline__ double warp_scan_incl_f64(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(RN_FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// ---------------------------------------------------------------------------------------------
// resample: models.py:200-203 + stepfun.py:134-258 + math.py:88-111 + coord.py:98
//
// CDF arithmetic: the reference's `torch.cumsum` (stepfun.py:149-154) accumulates a float row in DOUBLE on the CPU
// (ATen's acc_type<float, is_cuda=false>) and rounds every prefix to float.  The kernels do the same -- fp64 prefix
// sums of the fp32 probabilities, rounded once -- so that the CDF differs from the reference's only through the
// softmax (exp and the order of its normalising sum), and the fp64 sums make it monotone without a running max.
// `cw_in` (optional) replaces the computed CDF: the search + interpolation stage can then be checked bit for bit
// against the reference's own CDF.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32)
resample_kernel(const float* __restrict__ sdist_in, const float* __restrict__ w_in, const float* __restrict__ cw_in,
                const float* __restrict__ u,
                const float* __restrict__ near_, const float* __restrict__ far_, int64_t n_rays, int s_in, int s_out,
                float padding, float anneal, float dom_lo, float dom_hi, float* __restrict__ sdist_out,
                float* __restrict__ tdist_out, float* __restrict__ cw_out, int32_t* __restrict__ idx_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 2 * (s_in + 1) + s_out;
  float* ts = smem + warp * per_warp;  // fenceposts of the input step function
  float* cw = ts + (s_in + 1);         // CDF (also scratch for logits / exp)
  float* cs = cw + (s_in + 1);         // sampled centres
  const int64_t ray = (int64_t)blockIdx.x * kWarps + warp;
  if (ray >= n_rays) return;

  const float* tin = sdist_in + ray * (s_in + 1);
  for (int i = lane; i <= s_in; i += 32) ts[i] = tin[i];
  __syncwarp();

  if (cw_in) {
    const float* ci = cw_in + ray * (s_in + 1);
    for (int i = lane; i <= s_in; i += 32) cw[i] = ci[i];
  } else {
  // logits (models.py:200-203) and their max
  const float* win = w_in + ray * s_in;
  float m = -INFINITY;
  for (int i = lane; i < s_in; i += 32) {
    float l = (ts[i + 1] > ts[i]) ? __fmul_rn(anneal, logf(__fadd_rn(win[i], padding))) : -INFINITY;
    cw[i + 1] = l;
    m = fmaxf(m, l);
  }
  m = warp_max(m);
  // softmax numerators and their sum (stepfun.py:160)
  float sum = 0.f;
  for (int i = lane; i < s_in; i += 32) {
    float e = expf(__fsub_rn(cw[i + 1], m));
    cw[i + 1] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  // CDF: cw[j] = min(1, sum_{i<j} w_i), cw[0] = 0, cw[s_in] = 1 (stepfun.py:149-154), fp64 prefix sums
  double carry = 0.0;
  for (int base = 0; base < s_in; base += 32) {
    int i = base + lane;
    const double w = (i < s_in) ? (double)__fdiv_rn(cw[i + 1], sum) : 0.0;
    const double c = warp_scan_incl_f64(w, lane) + carry;
    carry = __shfl_sync(RN_FULL, c, 31);
    if (i < s_in - 1) cw[i + 1] = fminf((float)c, 1.f);
  }
  if (lane == 0) {
    cw[0] = 0.f;
    cw[s_in] = 1.f;
  }
  }
  __syncwarp();
  if (cw_out) {
    float* o = cw_out + ray * (s_in + 1);
    for (int i = lane; i <= s_in; i += 32) o[i] = cw[i];
  }

  // inverse CDF at the fixed grid u (math.py:88-111 in index form: idx = #{cw <= u} - 1)
  for (int j = lane; j < s_out; j += 32) {
    const float uj = u[j];
    int lo = 0, hi = s_in + 1;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cw[mid] <= uj) lo = mid + 1; else hi = mid;
    }
    const int idx = lo - 1;
    const int i0 = min(max(idx, 0), s_in), i1 = min(max(idx + 1, 0), s_in);
    const float x0 = cw[i0], x1 = cw[i1], f0 = ts[i0], f1 = ts[i1];
    const float off = nan_to_zero_clip01(__fdiv_rn(__fsub_rn(uj, x0), __fsub_rn(x1, x0)));
    cs[j] = __fadd_rn(f0, __fmul_rn(off, __fsub_rn(f1, f0)));
    if (idx_out) idx_out[ray * s_out + j] = idx;
  }
  __syncwarp();

  // midpoints + reflected end fenceposts (stepfun.py:246-258), then t = s*far + (1-s)*near
  const float nr = near_[ray], fr = far_[ray];
  float* so = sdist_out + ray * (s_out + 1);
  float* to = tdist_out ? tdist_out + ray * (s_out + 1) : nullptr;
  for (int j = lane; j <= s_out; j += 32) {
    float s;
    if (j == 0) {
      float mid0 = __fdiv_rn(__fadd_rn(cs[1], cs[0]), 2.f);
      s = fmaxf(dom_lo, __fsub_rn(__fmul_rn(2.f, cs[0]), mid0));
    } else if (j == s_out) {
      float midl = __fdiv_rn(__fadd_rn(cs[s_out - 1], cs[s_out - 2]), 2.f);
      s = fminf(dom_hi, __fsub_rn(__fmul_rn(2.f, cs[s_out - 1]), midl));
    } else {
      s = __fdiv_rn(__fadd_rn(cs[j], cs[j - 1]), 2.f);
    }
    so[j] = s;
    if (to) to[j] = __fadd_rn(__fmul_rn(s, fr), __fmul_rn(__fsub_rn(1.f, s), nr));
  }
}

// ---- 128 -> 128 fast path of the resampler: lane l owns input bins 4l..4l+3 and output samples 4l..4l+3 ------------
// Same arithmetic as resample_kernel (no-FMA interpolation, running-max monotone CDF, idx = #{cw <= u} - 1); the
// prefix sums are 3 local adds + one warp scan instead of four 32-wide scans, loads / stores are vectorised or
// staged, and the four binary searches of a lane run interleaved.
__global__ void __launch_bounds__(kWarps * 32)
resample128_kernel(const float* __restrict__ sdist_in, const float* __restrict__ w_in, const float* __restrict__ cw_in,
                   const float* __restrict__ u,
                   const float* __restrict__ near_, const float* __restrict__ far_, int64_t n_rays, float padding,
                   float anneal, float dom_lo, float dom_hi, float* __restrict__ sdist_out, float* __restrict__ tdist_out,
                   float* __restrict__ cw_out, int32_t* __restrict__ idx_out) {
  constexpr int s = 128;
  __shared__ float sm_t[kWarps][s + 4];
  __shared__ float sm_cw[kWarps][s + 4];
  __shared__ float sm_c[kWarps][s + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kWarps + warp;
  if (ray >= n_rays) return;
  float* ts = sm_t[warp];
  float* cw = sm_cw[warp];
  float* cs = sm_c[warp];
  const float4 w4 = cw_in ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(reinterpret_cast<const float4*>(w_in + ray * s) + lane);
  const float4 u4 = __ldg(reinterpret_cast<const float4*>(u) + lane);
  const float* tin = sdist_in + ray * (s + 1);
  for (int i = lane; i <= s; i += 32) ts[i] = __ldg(tin + i);
  if (cw_in) {
    const float* ci = cw_in + ray * (s + 1);
    for (int i = lane; i <= s; i += 32) cw[i] = __ldg(ci + i);
  }
  __syncwarp();
  if (!cw_in) {
  float t[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) t[k] = ts[4 * lane + k];
  const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
  // logits (models.py:200-203), softmax (stepfun.py:160)
  float lg[4], m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    lg[k] = (t[k + 1] > t[k]) ? __fmul_rn(anneal, logf(__fadd_rn(wv[k], padding))) : -INFINITY;
    m = fmaxf(m, lg[k]);
  }
  m = warp_max(m);
  float e[4], sum = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    e[k] = expf(__fsub_rn(lg[k], m));
    sum += e[k];
  }
  sum = warp_sum(sum);
  // CDF: cw[j] = min(1, float(sum_{i<j} p_i)) with the sum in fp64 (see the header), cw[0] = 0, cw[128] = 1
  double c[4], run = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    run += (double)__fdiv_rn(e[k], sum);
    c[k] = run;
  }
  double excl = __shfl_up_sync(RN_FULL, warp_scan_incl_f64(run, lane), 1);
  if (lane == 0) excl = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int j = 4 * lane + k + 1;          // cw index
    if (j < s) cw[j] = fminf((float)(excl + c[k]), 1.f);
  }
  if (lane == 0) cw[0] = 0.f;
  if (lane == 31) cw[s] = 1.f;
  }
  __syncwarp();
  if (cw_out) {
    float* o = cw_out + ray * (s + 1);
    for (int i = lane; i <= s; i += 32) o[i] = cw[i];
  }
  // inverse CDF at u[4l..4l+3]: four interleaved binary searches over cw[0..128]
  const float uj[4] = {u4.x, u4.y, u4.z, u4.w};
  int lo[4] = {0, 0, 0, 0}, hi[4] = {s + 1, s + 1, s + 1, s + 1};
#pragma unroll
  for (int it = 0; it < 8; ++it) {       // ceil(log2(130)) = 8 halvings of [0, 129)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (lo[k] < hi[k]) {
        const int mid = (lo[k] + hi[k]) >> 1;
        if (cw[mid] <= uj[k]) lo[k] = mid + 1; else hi[k] = mid;
      }
    }
  }
  float cv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = lo[k] - 1;
    const int i0 = min(max(idx, 0), s), i1 = min(max(idx + 1, 0), s);
    const float x0 = cw[i0], x1 = cw[i1], f0 = ts[i0], f1 = ts[i1];
    const float offk = nan_to_zero_clip01(__fdiv_rn(__fsub_rn(uj[k], x0), __fsub_rn(x1, x0)));
    cv[k] = __fadd_rn(f0, __fmul_rn(offk, __fsub_rn(f1, f0)));
    if (idx_out) idx_out[ray * s + 4 * lane + k] = idx;
  }
  *(reinterpret_cast<float4*>(cs) + lane) = make_float4(cv[0], cv[1], cv[2], cv[3]);
  __syncwarp();
  // midpoints + reflected end fenceposts (stepfun.py:246-258), then t = s*far + (1-s)*near
  const float nr = near_[ray], fr = far_[ray];
  float* so = sdist_out + ray * (s + 1);
  float* to = tdist_out ? tdist_out + ray * (s + 1) : nullptr;
  for (int j = lane; j <= s; j += 32) {
    float v;
    if (j == 0) {
      const float mid0 = __fdiv_rn(__fadd_rn(cs[1], cs[0]), 2.f);
      v = fmaxf(dom_lo, __fsub_rn(__fmul_rn(2.f, cs[0]), mid0));
    } else if (j == s) {
      const float midl = __fdiv_rn(__fadd_rn(cs[s - 1], cs[s - 2]), 2.f);
      v = fminf(dom_hi, __fsub_rn(__fmul_rn(2.f, cs[s - 1]), midl));
    } else {
      v = __fdiv_rn(__fadd_rn(cs[j], cs[j - 1]), 2.f);
    }
    so[j] = v;
    if (to) to[j] = __fadd_rn(__fmul_rn(v, fr), __fmul_rn(__fsub_rn(1.f, v), nr));
  }
}

// ---------------------------------------------------------------------------------------------
// compositing forward: render.py:132-149 + render.py:152-254
// ---------------------------------------------------------------------------------------------
// accumulate sum_i w[i] * v[i, c] for a [S,3] array read as 3*S contiguous floats
__device__ __forceinline__ void accum_vec3(const float* __restrict__ v, const float* ws, int s, int lane, float acc[3]) {
  const int n = 3 * s;
  for (int f = lane; f < n; f += 32) {
    const float x = v[f];
    const int smp = f / 3, ch = f - 3 * smp;
    const float p = ws[smp] * x;
    acc[0] += (ch == 0) ? p : 0.f;
    acc[1] += (ch == 1) ? p : 0.f;
    acc[2] += (ch == 2) ? p : 0.f;
  }
}

// ---- S = 128 fast path --------------------------------------------------------------------------------------
// Lane l owns the 4 consecutive samples 4l..4l+3: density / weights move as one float4 per lane, a [128,3] array
// as three float4 per lane (12 floats = 4 samples x 3 channels, all indices compile-time), the transmittance scan
// is 3 local adds + one warp scan, and the 22 per-ray sums are reduced through a shared-memory transpose (every
// lane sums one value over the 32 lanes) instead of 22 shuffle trees.  No integer division, no shared-memory weights.
constexpr int kFastVals = 22;   // rgb 0-2, diffuse 3-5, specular 6-8, dist 9, acc 10, logd 11, normals 12-14,
                                // normals_pred 15-17, tint 18-20, roughness 21
__device__ __forceinline__ void acc_samples4(const float* __restrict__ v, int lane, const float w[4], float acc[3]) {
  const float4* p = reinterpret_cast<const float4*>(v) + 3 * lane;
  const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  acc[0] = w[0] * a.x + w[1] * a.w + w[2] * b.z + w[3] * c.y;
  acc[1] = w[0] * a.y + w[1] * b.x + w[2] * b.w + w[3] * c.z;
  acc[2] = w[0] * a.z + w[1] * b.y + w[2] * c.x + w[3] * c.w;
}
// the same in two steps, so that the forward kernel can put ALL of a ray's loads in flight before the first use
// (the kernel is bound by load latency: ncu long-scoreboard stalls, profiles/r01_composite_fwd_ncu_full.txt)
struct Vec12 { float4 a, b, c; };
__device__ __forceinline__ Vec12 load_samples4(const float* __restrict__ v, int lane) {
  const float4* p = reinterpret_cast<const float4*>(v) + 3 * lane;
  Vec12 r;
  r.a = __ldg(p); r.b = __ldg(p + 1); r.c = __ldg(p + 2);
  return r;
}
__device__ __forceinline__ void acc_loaded4(const Vec12& x, const float w[4], float acc[3]) {
  acc[0] = w[0] * x.a.x + w[1] * x.a.w + w[2] * x.b.z + w[3] * x.c.y;
  acc[1] = w[0] * x.a.y + w[1] * x.b.x + w[2] * x.b.w + w[3] * x.c.z;
  acc[2] = w[0] * x.a.z + w[1] * x.b.y + w[2] * x.c.x + w[3] * x.c.w;
}

__global__ void __launch_bounds__(kWarps * 32, 5)
composite_fwd128_kernel(const float* __restrict__ density, const float* __restrict__ tdist, const float* __restrict__ dirs,
                        const float* __restrict__ far_, const float* __restrict__ rgb, const float* __restrict__ diffuse,
                        const float* __restrict__ specular, const float* __restrict__ normals,
                        const float* __restrict__ normals_pred, const float* __restrict__ roughness,
                        const float* __restrict__ tint, int64_t n_rays, float bg, float* __restrict__ weights_out,
                        float* __restrict__ comp_out, float* __restrict__ extras_out, double* __restrict__ pct_out) {
  constexpr int s = 128;
  __shared__ float sm_t[kWarps][s + 4];
  __shared__ float sm_cw[kWarps][s + 4];
  __shared__ float sm_red[kWarps][kFastVals][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kWarps + warp;
  if (ray >= n_rays) return;
  float* ts = sm_t[warp];
  float* cw = sm_cw[warp];

  const float4 d4 = __ldg(reinterpret_cast<const float4*>(density + ray * s) + lane);
  const float* tin = tdist + ray * (s + 1);
  float tstage[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) tstage[i] = (lane + 32 * i <= s) ? __ldg(tin + lane + 32 * i) : 0.f;
  // every other load of the ray is issued now, before anything waits on the first one
  const Vec12 c_rgb = load_samples4(rgb + ray * 384, lane);
  const Vec12 c_dif = load_samples4(diffuse + ray * 384, lane);
  const Vec12 c_spe = load_samples4(specular + ray * 384, lane);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  Vec12 c_nrm = {z4, z4, z4}, c_npr = {z4, z4, z4}, c_tin = {z4, z4, z4};
  float4 r4 = z4;
  if (extras_out) {
    if (normals) c_nrm = load_samples4(normals + ray * 384, lane);
    if (normals_pred) c_npr = load_samples4(normals_pred + ray * 384, lane);
    if (tint) c_tin = load_samples4(tint + ray * 384, lane);
    if (roughness) r4 = __ldg(reinterpret_cast<const float4*>(roughness + ray * s) + lane);
  }
#pragma unroll
  for (int i = 0; i < 5; ++i)
    if (lane + 32 * i <= s) ts[lane + 32 * i] = tstage[i];
  const float dx = dirs[ray * 3 + 0], dy = dirs[ray * 3 + 1], dz = dirs[ray * 3 + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  __syncwarp();
  float t[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) t[k] = ts[4 * lane + k];
  const float dens[4] = {d4.x, d4.y, d4.z, d4.w};
  float dd[4], pre[4];
  float run = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    dd[k] = dens[k] * ((t[k + 1] - t[k]) * dnorm);
    pre[k] = run;          // sum of this lane's earlier samples
    run += dd[k];
  }
  const float off = warp_scan_incl(run, lane) - run;   // sum over earlier lanes
  float w[4], vals[kFastVals];
  float acc = 0.f, dist = 0.f, logd = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    w[k] = (1.f - expf(-dd[k])) * expf(-(off + pre[k]));
    const float tmid = 0.5f * (t[k] + t[k + 1]);
    acc += w[k];
    dist += w[k] * tmid;
    logd += w[k] * logf(tmid);
  }
  *(reinterpret_cast<float4*>(weights_out + ray * s) + lane) = make_float4(w[0], w[1], w[2], w[3]);
  acc_loaded4(c_rgb, w, vals + 0);
  acc_loaded4(c_dif, w, vals + 3);
  acc_loaded4(c_spe, w, vals + 6);
  vals[9] = dist;
  vals[10] = acc;
  vals[11] = logd;
  int nvals = 12;
  if (extras_out) {
    nvals = kFastVals;
    acc_loaded4(c_nrm, w, vals + 12);      // absent arrays were loaded as zeros
    acc_loaded4(c_npr, w, vals + 15);
    acc_loaded4(c_tin, w, vals + 18);
    vals[21] = w[0] * r4.x + w[1] * r4.y + w[2] * r4.z + w[3] * r4.w;
  }
  // transpose reduction: value v of lane l -> red[v][l]; lane v then sums row v
#pragma unroll
  for (int v = 0; v < kFastVals; ++v)
    if (v < nvals) sm_red[warp][v][lane] = vals[v];
  __syncwarp();
  float total = 0.f;
  if (lane < nvals) {
#pragma unroll 8
    for (int l = 0; l < 32; ++l) total += sm_red[warp][lane][l];
  }
  const float acc_t = __shfl_sync(RN_FULL, total, 10);
  const float bg_w = fmaxf(0.f, 1.f - acc_t);
  {
    float o;
    if (lane < 9) o = total + bg_w * bg;
    else if (lane < 11) o = total;                 // 9 distance, 10 acc
    else if (lane == 11) {
      // distance_mean = clip(nan_to_num(exp(sum w log t_mid / max(eps, acc)), inf), t0, tS)  (render.py:232-238)
      float dm = expf(total / fmaxf(RN_EPS32, acc_t));
      if (dm != dm) dm = INFINITY;
      o = fminf(fmaxf(dm, ts[0]), ts[s]);
    } else o = (lane == 12) ? bg_w : 0.f;
    if (lane < 16) comp_out[ray * 16 + lane] = o;
  }
  if (extras_out && lane >= 12 && lane < 24) extras_out[ray * 12 + (lane - 12)] = lane < kFastVals ? total : 0.f;
  if (pct_out) {
    // weighted_percentile (stepfun.py:294-307) on t_aug=[tdist, far], w_aug=[weights, bg_w]: see the generic kernel
    // CDF of the weights with a running max (monotone to the last ulp): the lane owns samples 4l..4l+3, so this is
    // 3 local adds + one warp sum scan + one warp max scan (as in resample128_kernel)
    {
      float c[4], r = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        r += w[k];
        c[k] = r;
      }
      const float offc = warp_scan_incl(r, lane) - r;
      float pm = warp_scan_max(offc + c[3], lane);
      pm = __shfl_up_sync(RN_FULL, pm, 1);
      if (lane == 0) pm = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        pm = fmaxf(pm, offc + c[k]);
        cw[4 * lane + k + 1] = fminf(pm, 1.f);
      }
    }
    if (lane == 0) {
      cw[0] = 0.f;
      cw[s + 1] = 1.f;
      ts[s + 1] = far_[ray];
    }
    __syncwarp();
    const float ps[3] = {0.05f, 0.5f, 0.95f};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int cnt = 0;
      for (int i = lane; i < s + 2; i += 32) cnt += (ps[k] >= cw[i]) ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(RN_FULL, cnt, o);
      if (lane == k) {
        int idx = min(max(cnt - 1, 0), s);
        const double x0 = cw[idx], x1 = cw[idx + 1], f0 = ts[idx], f1 = ts[idx + 1];
        const double mm = (f1 - f0) / (x1 - x0);
        const double bb = f0 - mm * x0;
        pct_out[ray * 3 + k] = mm * (double)ps[k] + bb;
      }
    }
  }
}

__global__ void __launch_bounds__(kWarps * 32)
composite_fwd_kernel(const float* __restrict__ density, const float* __restrict__ tdist, const float* __restrict__ dirs,
                     const float* __restrict__ far_, const float* __restrict__ rgb, const float* __restrict__ diffuse,
                     const float* __restrict__ specular, const float* __restrict__ normals,
                     const float* __restrict__ normals_pred, const float* __restrict__ roughness,
                     const float* __restrict__ tint, int64_t n_rays, int s, float bg, float* __restrict__ weights_out,
                     float* __restrict__ comp_out, float* __restrict__ extras_out, double* __restrict__ pct_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 3 * (s + 2);
  float* ts = smem + warp * per_warp;  // tdist (+far) [s+2]
  float* ws = ts + (s + 2);            // weights [s]
  float* cw = ws + (s + 2);            // CDF for percentiles [s+2]
  const int64_t ray = (int64_t)blockIdx.x * kWarps + warp;
  if (ray >= n_rays) return;

  const float* tin = tdist + ray * (s + 1);
  for (int i = lane; i <= s; i += 32) ts[i] = tin[i];
  const float dx = dirs[ray * 3 + 0], dy = dirs[ray * 3 + 1], dz = dirs[ray * 3 + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  __syncwarp();

  // w_i = (1 - exp(-dd_i)) * exp(-sum_{j<i} dd_j)
  const float* din = density + ray * s;
  float carry = 0.f, acc = 0.f, dist = 0.f, logd = 0.f;
  for (int base = 0; base < s; base += 32) {
    const int i = base + lane;
    float dd = 0.f;
    if (i < s) dd = din[i] * ((ts[i + 1] - ts[i]) * dnorm);
    const float incl = warp_scan_incl(dd, lane) + carry;
    carry = __shfl_sync(RN_FULL, incl, 31);
    if (i < s) {
      const float w = (1.f - expf(-dd)) * expf(-(incl - dd));
      ws[i] = w;
      weights_out[ray * s + i] = w;
      acc += w;
      const float tmid = 0.5f * (ts[i] + ts[i + 1]);
      dist += w * tmid;
      logd += w * logf(tmid);
    }
  }
  acc = warp_sum(acc);
  dist = warp_sum(dist);
  logd = warp_sum(logd);
  __syncwarp();

  const float bg_w = fmaxf(0.f, 1.f - acc);
  float c0[3] = {0, 0, 0}, c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
  accum_vec3(rgb + ray * 3 * s, ws, s, lane, c0);
  accum_vec3(diffuse + ray * 3 * s, ws, s, lane, c1);
  accum_vec3(specular + ray * 3 * s, ws, s, lane, c2);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    c0[c] = warp_sum(c0[c]);
    c1[c] = warp_sum(c1[c]);
    c2[c] = warp_sum(c2[c]);
  }
  if (lane == 0) {
    float* o = comp_out + ray * 16;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] = c0[c] + bg_w * bg;
      o[3 + c] = c1[c] + bg_w * bg;
      o[6 + c] = c2[c] + bg_w * bg;
    }
    o[9] = dist;
    o[10] = acc;
    // distance_mean = clip(nan_to_num(exp(sum w log t_mid / max(eps, acc)), inf), t0, tS)  (render.py:232-238)
    float dm = expf(logd / fmaxf(RN_EPS32, acc));
    if (dm != dm) dm = INFINITY;        // torch.nan_to_num(x, nan=inf)
    dm = fminf(fmaxf(dm, ts[0]), ts[s]);
    o[11] = dm;
    o[12] = bg_w;
    o[13] = o[14] = o[15] = 0.f;
  }
  if (extras_out) {
    float e0[3] = {0, 0, 0}, e1[3] = {0, 0, 0}, e2[3] = {0, 0, 0};
    float er = 0.f;
    if (normals) accum_vec3(normals + ray * 3 * s, ws, s, lane, e0);
    if (normals_pred) accum_vec3(normals_pred + ray * 3 * s, ws, s, lane, e1);
    if (tint) accum_vec3(tint + ray * 3 * s, ws, s, lane, e2);
    if (roughness)
      for (int i = lane; i < s; i += 32) er += ws[i] * roughness[ray * s + i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      e0[c] = warp_sum(e0[c]);
      e1[c] = warp_sum(e1[c]);
      e2[c] = warp_sum(e2[c]);
    }
    er = warp_sum(er);
    if (lane == 0) {
      float* o = extras_out + ray * 12;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        o[c] = e0[c];
        o[3 + c] = e1[c];
        o[6 + c] = e2[c];
      }
      o[9] = er;
      o[10] = o[11] = 0.f;
    }
  }
  if (pct_out) {
    // weighted_percentile (stepfun.py:294-307) on t_aug=[tdist, far], w_aug=[weights, bg_w]:
    // cw = [0, min(1, cumsum(weights)), 1]  (bg_w is the dropped last weight); fp64 interp (math.py:114-142)
    float cr = 0.f, crmax = 0.f;
    for (int base = 0; base < s; base += 32) {
      const int i = base + lane;
      const float w = (i < s) ? ws[i] : 0.f;
      float c = warp_scan_incl(w, lane) + cr;
      cr = __shfl_sync(RN_FULL, c, 31);
      c = fmaxf(warp_scan_max(c, lane), crmax);
      crmax = __shfl_sync(RN_FULL, c, 31);
      if (i < s) cw[i + 1] = fminf(c, 1.f);
    }
    if (lane == 0) {
      cw[0] = 0.f;
      cw[s + 1] = 1.f;
      ts[s + 1] = far_[ray];
    }
    __syncwarp();
    const float ps[3] = {0.05f, 0.5f, 0.95f};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int cnt = 0;
      for (int i = lane; i < s + 2; i += 32) cnt += (ps[k] >= cw[i]) ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(RN_FULL, cnt, o);
      if (lane == 0) {
        int idx = min(max(cnt - 1, 0), s);
        const double x0 = cw[idx], x1 = cw[idx + 1], f0 = ts[idx], f1 = ts[idx + 1];
        const double mm = (f1 - f0) / (x1 - x0);
        const double bb = f0 - mm * x0;
        pct_out[ray * 3 + k] = mm * (double)ps[k] + bb;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// compositing backward
// ---------------------------------------------------------------------------------------------
// gw[smp] += sum_c g[c] * v[smp,c]  and  dv[smp,c] = w[smp] * g[c]; v read as flat 3*S floats
__device__ __forceinline__ void bwd_vec3(const float* __restrict__ v, float* __restrict__ dv, const float* ws, float* gw,
                                         const float g[3], int s, int lane) {
  const int n = 3 * s;
  for (int f = lane; f < n; f += 32) {
    const int smp = f / 3, ch = f - 3 * smp;
    const float gc = (ch == 0) ? g[0] : ((ch == 1) ? g[1] : g[2]);
    if (v) atomicAdd(&gw[smp], gc * v[f]);   // shared-memory atomic: 3 lanes hit one sample
    if (dv) dv[f] = ws[smp] * gc;
  }
}

// ---- S = 128 fast path of the backward: lane l owns samples 4l..4l+3 (see composite_fwd128_kernel) -------------
// gw[k] += sum_c g[c] * v[k][c];  dv[k][c] = w[k] * g[c]   for one [128,3] array (v / dv may be null)
__device__ __forceinline__ void bwd_samples4(const float* __restrict__ v, float* __restrict__ dv, int lane, const float w[4],
                                             const float g[3], float gw[4]) {
  if (v) {
    const float4* p = reinterpret_cast<const float4*>(v) + 3 * lane;
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    gw[0] += g[0] * a.x + g[1] * a.y + g[2] * a.z;
    gw[1] += g[0] * a.w + g[1] * b.x + g[2] * b.y;
    gw[2] += g[0] * b.z + g[1] * b.w + g[2] * c.x;
    gw[3] += g[0] * c.y + g[1] * c.z + g[2] * c.w;
  }
  if (dv) {
    float4* q = reinterpret_cast<float4*>(dv) + 3 * lane;
    q[0] = make_float4(w[0] * g[0], w[0] * g[1], w[0] * g[2], w[1] * g[0]);
    q[1] = make_float4(w[1] * g[1], w[1] * g[2], w[2] * g[0], w[2] * g[1]);
    q[2] = make_float4(w[2] * g[2], w[3] * g[0], w[3] * g[1], w[3] * g[2]);
  }
}

__global__ void __launch_bounds__(kWarps * 32)
composite_bwd128_kernel(const float* __restrict__ density, const float* __restrict__ tdist, const float* __restrict__ dirs,
                        const float* __restrict__ rgb, const float* __restrict__ diffuse, const float* __restrict__ specular,
                        const float* __restrict__ normals, const float* __restrict__ normals_pred,
                        const float* __restrict__ roughness, const float* __restrict__ tint,
                        const float* __restrict__ weights, const float* __restrict__ comp,
                        const float* __restrict__ g_weights, const float* __restrict__ g_comp,
                        const float* __restrict__ g_extras, int64_t n_rays, float bg, float* __restrict__ d_density,
                        float* __restrict__ d_rgb, float* __restrict__ d_diffuse, float* __restrict__ d_specular,
                        float* __restrict__ d_normals_pred, float* __restrict__ d_roughness, float* __restrict__ d_tint) {
  constexpr int s = 128;
  __shared__ float sm_t[kWarps][s + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kWarps + warp;
  if (ray >= n_rays) return;
  float* ts = sm_t[warp];
  const float4 w4 = __ldg(reinterpret_cast<const float4*>(weights + ray * s) + lane);
  const float4 d4 = __ldg(reinterpret_cast<const float4*>(density + ray * s) + lane);
  float4 gw4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g_weights) gw4 = __ldg(reinterpret_cast<const float4*>(g_weights + ray * s) + lane);
  for (int i = lane; i <= s; i += 32) ts[i] = __ldg(tdist + ray * (s + 1) + i);
  const float dx = dirs[ray * 3 + 0], dy = dirs[ray * 3 + 1], dz = dirs[ray * 3 + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  const float acc = comp[ray * 16 + 10];
  const float bgmask = (1.f - acc > 0.f) ? bg : 0.f;  // d max(0, 1-acc)/d acc
  float g0[3], g1[3], g2[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    g0[c] = g_comp[ray * 16 + c];
    g1[c] = g_comp[ray * 16 + 3 + c];
    g2[c] = g_comp[ray * 16 + 6 + c];
  }
  const float g_dist = g_comp[ray * 16 + 9], g_acc = g_comp[ray * 16 + 10];
  const float gbg = -bgmask * (g0[0] + g0[1] + g0[2] + g1[0] + g1[1] + g1[2] + g2[0] + g2[1] + g2[2]);
  __syncwarp();
  float t[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) t[k] = ts[4 * lane + k];
  const float w[4] = {w4.x, w4.y, w4.z, w4.w};
  const float dens[4] = {d4.x, d4.y, d4.z, d4.w};
  float gw[4] = {gw4.x, gw4.y, gw4.z, gw4.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) gw[k] += g_acc + gbg + g_dist * 0.5f * (t[k] + t[k + 1]);
  bwd_samples4(rgb + ray * 384, d_rgb + ray * 384, lane, w, g0, gw);
  bwd_samples4(diffuse + ray * 384, d_diffuse + ray * 384, lane, w, g1, gw);
  bwd_samples4(specular + ray * 384, d_specular + ray * 384, lane, w, g2, gw);
  if (g_extras) {
    float e0[3], e1[3], e2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      e0[c] = g_extras[ray * 12 + c];
      e1[c] = g_extras[ray * 12 + 3 + c];
      e2[c] = g_extras[ray * 12 + 6 + c];
    }
    const float er = g_extras[ray * 12 + 9];
    if (normals) bwd_samples4(normals + ray * 384, nullptr, lane, w, e0, gw);
    if (normals_pred) bwd_samples4(normals_pred + ray * 384, d_normals_pred ? d_normals_pred + ray * 384 : nullptr, lane, w, e1, gw);
    if (tint) bwd_samples4(tint + ray * 384, d_tint ? d_tint + ray * 384 : nullptr, lane, w, e2, gw);
    if (roughness) {
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(roughness + ray * s) + lane);
      gw[0] += er * r4.x; gw[1] += er * r4.y; gw[2] += er * r4.z; gw[3] += er * r4.w;
      if (d_roughness) *(reinterpret_cast<float4*>(d_roughness + ray * s) + lane) = make_float4(w[0] * er, w[1] * er, w[2] * er, w[3] * er);
    }
  } else {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d_normals_pred) for (int j = 0; j < 3; ++j) *(reinterpret_cast<float4*>(d_normals_pred + ray * 384) + 3 * lane + j) = z;
    if (d_tint) for (int j = 0; j < 3; ++j) *(reinterpret_cast<float4*>(d_tint + ray * 384) + 3 * lane + j) = z;
    if (d_roughness) *(reinterpret_cast<float4*>(d_roughness + ray * s) + lane) = z;
  }
  // w_i = (1-e^{-dd_i}) T_i,  T_i = exp(-sum_{j<i} dd_j);  dL/ddd_i = gw_i T_i e^{-dd_i} - sum_{j>i} gw_j w_j
  float delta[4], incl_dd[4], incl_p[4];
  float run_dd = 0.f, run_p = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    delta[k] = (t[k + 1] - t[k]) * dnorm;
    run_dd += dens[k] * delta[k];
    run_p += gw[k] * w[k];
    incl_dd[k] = run_dd;
    incl_p[k] = run_p;
  }
  const float scan_dd = warp_scan_incl(run_dd, lane), scan_p = warp_scan_incl(run_p, lane);
  const float off_dd = scan_dd - run_dd, off_p = scan_p - run_p;
  const float total = __shfl_sync(RN_FULL, scan_p, 31);
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    o[k] = (gw[k] * expf(-(off_dd + incl_dd[k])) - (total - (off_p + incl_p[k]))) * delta[k];
  *(reinterpret_cast<float4*>(d_density + ray * s) + lane) = make_float4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(kWarps * 32)
composite_bwd_kernel(const float* __restrict__ density, const float* __restrict__ tdist, const float* __restrict__ dirs,
                     const float* __restrict__ rgb, const float* __restrict__ diffuse, const float* __restrict__ specular,
                     const float* __restrict__ normals, const float* __restrict__ normals_pred,
                     const float* __restrict__ roughness, const float* __restrict__ tint,
                     const float* __restrict__ weights, const float* __restrict__ comp,
                     const float* __restrict__ g_weights, const float* __restrict__ g_comp,
                     const float* __restrict__ g_extras, int64_t n_rays, int s, float bg, float* __restrict__ d_density,
                     float* __restrict__ d_rgb, float* __restrict__ d_diffuse, float* __restrict__ d_specular,
                     float* __restrict__ d_normals_pred, float* __restrict__ d_roughness, float* __restrict__ d_tint) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 3 * (s + 2);
  float* ts = smem + warp * per_warp;
  float* ws = ts + (s + 2);
  float* gw = ws + (s + 2);  // dL/dw per sample
  const int64_t ray = (int64_t)blockIdx.x * kWarps + warp;
  if (ray >= n_rays) return;

  for (int i = lane; i <= s; i += 32) ts[i] = tdist[ray * (s + 1) + i];
  for (int i = lane; i < s; i += 32) ws[i] = weights[ray * s + i];
  const float dx = dirs[ray * 3 + 0], dy = dirs[ray * 3 + 1], dz = dirs[ray * 3 + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  const float acc = comp[ray * 16 + 10];
  const float bgmask = (1.f - acc > 0.f) ? bg : 0.f;  // d max(0, 1-acc)/d acc
  float g0[3], g1[3], g2[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    g0[c] = g_comp[ray * 16 + c];
    g1[c] = g_comp[ray * 16 + 3 + c];
    g2[c] = g_comp[ray * 16 + 6 + c];
  }
  const float g_dist = g_comp[ray * 16 + 9], g_acc = g_comp[ray * 16 + 10];
  const float gbg = -bgmask * (g0[0] + g0[1] + g0[2] + g1[0] + g1[1] + g1[2] + g2[0] + g2[1] + g2[2]);
  __syncwarp();
  for (int i = lane; i < s; i += 32) {
    float g = g_weights ? g_weights[ray * s + i] : 0.f;
    g += g_acc + gbg + g_dist * 0.5f * (ts[i] + ts[i + 1]);
    gw[i] = g;
  }
  __syncwarp();
  bwd_vec3(rgb + ray * 3 * s, d_rgb + ray * 3 * s, ws, gw, g0, s, lane);
  bwd_vec3(diffuse + ray * 3 * s, d_diffuse + ray * 3 * s, ws, gw, g1, s, lane);
  bwd_vec3(specular + ray * 3 * s, d_specular + ray * 3 * s, ws, gw, g2, s, lane);
  if (g_extras) {
    float e0[3], e1[3], e2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      e0[c] = g_extras[ray * 12 + c];
      e1[c] = g_extras[ray * 12 + 3 + c];
      e2[c] = g_extras[ray * 12 + 6 + c];
    }
    const float er = g_extras[ray * 12 + 9];
    if (normals) bwd_vec3(normals + ray * 3 * s, nullptr, ws, gw, e0, s, lane);
    if (normals_pred) bwd_vec3(normals_pred + ray * 3 * s, d_normals_pred ? d_normals_pred + ray * 3 * s : nullptr, ws, gw, e1, s, lane);
    if (tint) bwd_vec3(tint + ray * 3 * s, d_tint ? d_tint + ray * 3 * s : nullptr, ws, gw, e2, s, lane);
    if (roughness)
      for (int i = lane; i < s; i += 32) {
        atomicAdd(&gw[i], er * roughness[ray * s + i]);
        if (d_roughness) d_roughness[ray * s + i] = ws[i] * er;
      }
  } else {
    if (d_normals_pred) for (int f = lane; f < 3 * s; f += 32) d_normals_pred[ray * 3 * s + f] = 0.f;
    if (d_tint) for (int f = lane; f < 3 * s; f += 32) d_tint[ray * 3 * s + f] = 0.f;
    if (d_roughness) for (int i = lane; i < s; i += 32) d_roughness[ray * s + i] = 0.f;
  }
  __syncwarp();

  // w_i = (1-e^{-dd_i}) T_i,  T_i = exp(-sum_{j<i} dd_j)
  // dL/ddd_i = gw_i * T_i * e^{-dd_i} - sum_{j>i} gw_j w_j
  float total = 0.f;
  for (int i = lane; i < s; i += 32) total += gw[i] * ws[i];
  total = warp_sum(total);
  float carry_dd = 0.f, carry_gw = 0.f;
  for (int base = 0; base < s; base += 32) {
    const int i = base + lane;
    float dd = 0.f, p = 0.f, delta = 0.f;
    if (i < s) {
      delta = (ts[i + 1] - ts[i]) * dnorm;
      dd = density[ray * s + i] * delta;
      p = gw[i] * ws[i];
    }
    const float incl_dd = warp_scan_incl(dd, lane) + carry_dd;
    const float incl_p = warp_scan_incl(p, lane) + carry_gw;
    carry_dd = __shfl_sync(RN_FULL, incl_dd, 31);
    carry_gw = __shfl_sync(RN_FULL, incl_p, 31);
    if (i < s) {
      const float trans_next = expf(-incl_dd);  // T_i * e^{-dd_i}
      const float suffix = total - incl_p;      // sum_{j>i} gw_j w_j
      d_density[ray * s + i] = (gw[i] * trans_next - suffix) * delta;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// lossfun_outer (stepfun.py:31-89) and lossfun_distortion (stepfun.py:261-272)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int count_le(const float* a, int n, float v) {  // #{a <= v}, a sorted
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <bool BWD>
__global__ void __launch_bounds__(kWarps * 32)
lossfun_outer_kernel(const float* __restrict__ t, const float* __restrict__ w, const float* __restrict__ t_env,
                     const float* __restrict__ w_env, const float* __restrict__ g_loss, int64_t n_rays, int s, int se,
                     float* __restrict__ out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 3 * (se + 1);
  float* te = smem + warp * per_warp;  // envelope fenceposts
  float* cy = te + (se + 1);           // [0, cumsum(w_env)]
  float* gcy = cy + (se + 1);          // grad wrt cy (bwd)
  const int64_t ray = (int64_t)blockIdx.x * kWarps + warp;
  if (ray >= n_rays) return;
  for (int i = lane; i <= se; i += 32) {
    te[i] = t_env[ray * (se + 1) + i];
    gcy[i] = 0.f;
  }
  float carry = 0.f;
  for (int base = 0; base < se; base += 32) {
    const int i = base + lane;
    const float v = (i < se) ? w_env[ray * se + i] : 0.f;
    const float c = warp_scan_incl(v, lane) + carry;
    carry = __shfl_sync(RN_FULL, c, 31);
    if (i < se) cy[i + 1] = c;
  }
  if (lane == 0) cy[0] = 0.f;
  __syncwarp();
  for (int i = lane; i < s; i += 32) {
    const float ta = t[ray * (s + 1) + i], tb = t[ray * (s + 1) + i + 1];
    const int lo = max(count_le(te, se + 1, ta) - 1, 0);   // idx_lo of t[i]
    const int hi = min(count_le(te, se + 1, tb), se);      // idx_hi of t[i+1]
    const float w_outer = cy[hi] - cy[lo];
    const float wi = w[ray * s + i];
    const float r = fmaxf(0.f, wi - w_outer);
    if (!BWD) {
      out[ray * s + i] = r * r / (wi + RN_EPS32);
    } else {
      // d loss / d w_outer = -2 r / (w + eps)
      const float g = -2.f * r / (wi + RN_EPS32) * g_loss[ray * s + i];
      if (g != 0.f) {
        atomicAdd(&gcy[hi], g);
        atomicAdd(&gcy[lo], -g);
      }
    }
  }
  if (BWD) {
    __syncwarp();
    // cy[j] = sum_{i<j} w_env[i]  =>  d w_env[i] = sum_{j>i} gcy[j]  (suffix sum)
    float tot = 0.f;
    for (int i = lane; i <= se; i += 32) tot += gcy[i];
    tot = warp_sum(tot);
    float cr = 0.f;
    for (int base = 0; base <= se; base += 32) {
      const int i = base + lane;
      const float v = (i <= se) ? gcy[i] : 0.f;
      const float c = warp_scan_incl(v, lane) + cr;
      cr = __shfl_sync(RN_FULL, c, 31);
      if (i < se) out[ray * se + i] = tot - c;
    }
  }
}

template <bool BWD>
__global__ void __launch_bounds__(kWarps * 32)
distortion_kernel(const float* __restrict__ t, const float* __restrict__ w, const float* __restrict__ g_loss,
                  int64_t n_rays, int s, float* __restrict__ out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ut = smem + warp * 2 * s;
  float* ws = ut + s;
  const int64_t ray = (int64_t)blockIdx.x * kWarps + warp;
  if (ray >= n_rays) return;
  for (int i = lane; i < s; i += 32) {
    ut[i] = 0.5f * (t[ray * (s + 1) + i] + t[ray * (s + 1) + i + 1]);
    ws[i] = w[ray * s + i];
  }
  __syncwarp();
  float acc = 0.f;
  const float g = BWD ? g_loss[ray] : 0.f;
  for (int i = lane; i < s; i += 32) {
    float inner = 0.f;
    const float ui = ut[i];
    for (int j = 0; j < s; ++j) inner += ws[j] * fabsf(ui - ut[j]);
    const float width = t[ray * (s + 1) + i + 1] - t[ray * (s + 1) + i];
    if (!BWD) {
      acc += ws[i] * inner + ws[i] * ws[i] * width * (1.f / 3.f);
    } else {
      out[ray * s + i] = g * (2.f * inner + 2.f * ws[i] * width * (1.f / 3.f));
    }
  }
  if (!BWD) {
    acc = warp_sum(acc);
    if (lane == 0) out[ray] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// orientation + predicted-normal losses (train_utils.py:165-204), per ray, in one pass over the per-sample
// normals:  ori[r] = sum_s w (min(0, n_t . (-v)))^2   (n_t = normals_pred or the density-gradient normals),
//           pred[r] = sum_s w (1 - n . n_pred).
// One warp per ray; the [s,3] rows are staged through shared memory with coalesced accesses, each lane then
// owns samples lane, lane+32, ...  Backward: d w and d normals_pred (the density-gradient normals are constants).
// ---------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(kWarps * 32)
normal_losses_kernel(const float* __restrict__ w, const float* __restrict__ normals, const float* __restrict__ npred,
                     const float* __restrict__ viewdirs, const float* __restrict__ g_ori, const float* __restrict__ g_pred,
                     int64_t n_rays, int s, int ori_is_pred, float* __restrict__ ori_out, float* __restrict__ pred_out,
                     float* __restrict__ d_w, float* __restrict__ d_npred) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sn = smem + warp * 6 * s;   // density-gradient normals [s,3] (zeros when absent)
  float* sp = sn + 3 * s;            // predicted normals [s,3]; backward: overwritten by d normals_pred
  const int64_t ray = (int64_t)blockIdx.x * kWarps + warp;
  if (ray >= n_rays) return;
  const size_t base3 = (size_t)ray * s * 3;
  for (int i = lane; i < 3 * s; i += 32) {
    sn[i] = normals ? normals[base3 + i] : 0.f;
    sp[i] = npred[base3 + i];
  }
  __syncwarp();
  const float vx = -viewdirs[ray * 3], vy = -viewdirs[ray * 3 + 1], vz = -viewdirs[ray * 3 + 2];
  const float go = BWD ? (g_ori ? g_ori[ray] : 0.f) : 0.f;
  const float gp = BWD ? ((g_pred && normals) ? g_pred[ray] : 0.f) : 0.f;
  float ori = 0.f, pred = 0.f;
  for (int i = lane; i < s; i += 32) {
    const float wi = w[(size_t)ray * s + i];
    const float nx = sn[3 * i], ny = sn[3 * i + 1], nz = sn[3 * i + 2];
    const float px = sp[3 * i], py = sp[3 * i + 1], pz = sp[3 * i + 2];
    const float tx = ori_is_pred ? px : nx, ty = ori_is_pred ? py : ny, tz = ori_is_pred ? pz : nz;
    const float ndv = fminf(tx * vx + ty * vy + tz * vz, 0.f);
    const float one_m = 1.f - (nx * px + ny * py + nz * pz);
    if (!BWD) {
      ori += wi * ndv * ndv;
      pred += wi * one_m;
    } else {
      d_w[(size_t)ray * s + i] = go * ndv * ndv + gp * one_m;
      const float c = ori_is_pred ? go * wi * 2.f * ndv : 0.f;
      const float d = -gp * wi;
      sp[3 * i] = c * vx + d * nx;
      sp[3 * i + 1] = c * vy + d * ny;
      sp[3 * i + 2] = c * vz + d * nz;
    }
  }
  if (!BWD) {
    ori = warp_sum(ori);
    pred = warp_sum(pred);
    if (lane == 0) {
      ori_out[ray] = ori;
      pred_out[ray] = normals ? pred : 0.f;
    }
  } else {
    __syncwarp();
    for (int i = lane; i < 3 * s; i += 32) d_npred[base3 + i] = sp[i];
  }
}

// ---------------------------------------------------------------------------------------------
// max_dilate_weights: stepfun.py:92-131 (+ the [1:-1] trim of its caller, models.py:186-187)
//
// The reference sorts the 3S+1 fenceposts t U (t[:-1] - eps) U (t[1:] + eps) and evaluates, for every new fencepost,
// a dense [3S+1, S] mask "which dilated interval contains me" followed by a masked max.  Here one warp owns a ray:
//   * the three source lists are each sorted already, so the sort is a 3-way merge: every source element finds its
//     rank with two binary searches in the other two lists (ties: t before t0 before t1, any stable order gives the
//     same sorted VALUES) and scatters its domain-clipped value;
//   * t0 and t1 are monotone, so the intervals [t0_j, t1_j) containing a fencepost tau form one index window
//     [#{t1 <= tau}, #{t0 <= tau} - 1]: two more binary searches and a range-max query on a sparse table of the pdf
//     (log2(S) levels in shared memory) replace the mask;
//   * weights = pdf * width, optional renormalisation by the warp-reduced sum.
// fp32 operations are the reference's (explicit rn intrinsics), so the fenceposts are bit-exact; the weights differ
// only through the order of the renormalising sum.
// ---------------------------------------------------------------------------------------------
// (count_le: defined with the interlevel loss above)
__device__ __forceinline__ int count_lt(const float* a, int n, float x) {   // #{i < n : a[i] < x}
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kWarps * 32)
max_dilate_kernel(const float* __restrict__ t_in, const float* __restrict__ w_in, int64_t n_rays, int s, int levels,
                  float dilation, float dom_lo, float dom_hi, int renormalize, int trim, float* __restrict__ t_out,
                  float* __restrict__ w_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = 3 * s + 1;                         // dilated fenceposts
  const int per_warp = (s + 1) + 2 * s + m + levels * s;
  float* ts = smem + warp * per_warp;              // t        [s+1]
  float* t0 = ts + (s + 1);                        // t[:-1] - dilation   [s]
  float* t1 = t0 + s;                              // t[1:]  + dilation   [s]
  float* td = t1 + s;                              // sorted + clipped    [3s+1]
  float* tab = td + m;                             // sparse table of the pdf: tab[k*s + j] = max p[j .. j + 2^k - 1]
  const int64_t ray = (int64_t)blockIdx.x * kWarps + warp;
  if (ray >= n_rays) return;
  const float eps2 = RN_EPS32 * RN_EPS32;
  const float* tin = t_in + ray * (s + 1);
  const float* win = w_in + ray * s;
  for (int i = lane; i <= s; i += 32) ts[i] = tin[i];
  __syncwarp();
  for (int j = lane; j < s; j += 32) {
    t0[j] = __fsub_rn(ts[j], dilation);
    t1[j] = __fadd_rn(ts[j + 1], dilation);
    tab[j] = __fdiv_rn(win[j], fmaxf(eps2, __fsub_rn(ts[j + 1], ts[j])));   // weight_to_pdf, stepfun.py:92-94
  }
  __syncwarp();
  for (int k = 1; k < levels; ++k) {
    const int half = 1 << (k - 1);
    for (int j = lane; j < s; j += 32) {
      const float a = tab[(k - 1) * s + j];
      const float b = (j + half < s) ? tab[(k - 1) * s + j + half] : a;
      tab[k * s + j] = fmaxf(a, b);
    }
    __syncwarp();
  }
  // 3-way merge by rank, values clipped to the domain (stepfun.py:113-116)
  auto clipd = [&](float x) { return fminf(fmaxf(x, dom_lo), dom_hi); };
  for (int i = lane; i <= s; i += 32) {
    const float x = ts[i];
    td[i + count_lt(t0, s, x) + count_lt(t1, s, x)] = clipd(x);
  }
  for (int j = lane; j < s; j += 32) {
    const float x0 = t0[j], x1 = t1[j];
    td[j + count_le(ts, s + 1, x0) + count_lt(t1, s, x0)] = clipd(x0);
    td[j + count_le(ts, s + 1, x1) + count_le(t0, s, x1)] = clipd(x1);
  }
  __syncwarp();
  // pdf at every new fencepost but the last: max over the window of dilated intervals that contain it (:118-126);
  // weights = pdf * width (:127), optionally renormalised (:129-130).  With `trim` the first and last fencepost /
  // weight are dropped on the way out (models.py:186-187).
  float* wo = w_out + ray * (trim ? (m - 3) : (m - 1));
  auto out_index = [&](int i) -> int { return trim ? ((i >= 1 && i < m - 2) ? i - 1 : -1) : i; };
  float part = 0.f;
  for (int i = lane; i < m - 1; i += 32) {
    const float tau = td[i];
    const int jlo = count_le(t1, s, tau);          // intervals with t1_j <= tau are over
    const int jhi = count_le(t0, s, tau) - 1;      // intervals with t0_j <= tau have begun
    float pd = 0.f;                                // (the reference's masked max runs over zeros as well)
    if (jhi >= jlo) {
      const int k = 31 - __clz(jhi - jlo + 1);
      pd = fmaxf(pd, fmaxf(tab[k * s + jlo], tab[k * s + jhi - (1 << k) + 1]));
    }
    const float wv = __fmul_rn(pd, __fsub_rn(td[i + 1], tau));
    part += wv;                                    // the sum runs over ALL weights, trimmed or not
    const int o = out_index(i);
    if (o >= 0) wo[o] = wv;
  }
  if (renormalize) {
    const float denom = fmaxf(eps2, warp_sum(part));
    for (int i = lane; i < m - 1; i += 32) {       // every lane rescales exactly the elements it wrote
      const int o = out_index(i);
      if (o >= 0) wo[o] = __fdiv_rn(wo[o], denom);
    }
  }
  float* to = t_out + ray * (trim ? (m - 2) : m);
  if (trim) {
    for (int i = lane; i < m - 2; i += 32) to[i] = td[i + 1];
  } else {
    for (int i = lane; i < m; i += 32) to[i] = td[i];
  }
}

inline unsigned blocks_for(int64_t n_rays) { return (unsigned)((n_rays + kWarps - 1) / kWarps); }

template <typename K>
int ensure_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    if (bytes > 200 * 1024) return rn_set_error(RN_ERR_UNSUPPORTED, "too many samples per ray for the shared-memory staging");
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return rn_set_cuda_error(e, __FILE__, __LINE__);
  }
  return RN_OK;
}

}  // namespace

extern "C" int rn_resample(const float* sdist_in, const float* weights_in, const float* cw_in, const float* u, const float* near_,
                           const float* far_, int64_t n_rays, int s_in, int s_out, float padding, float anneal,
                           float dom_lo, float dom_hi, float* sdist_out, float* tdist_out, float* cw_out,
                           int32_t* idx_out, void* stream) {
  RnProfScope prof_scope(RN_PROF_RAY, (cudaStream_t)stream);
  if (n_rays < 0 || s_in < 1 || s_out < 2) return rn_set_error(RN_ERR_ARG, "rn_resample: need s_in >= 1 and s_out >= 2");
  if (!weights_in && !cw_in) return rn_set_error(RN_ERR_ARG, "rn_resample: need weights or a CDF");
  if (n_rays == 0) return RN_OK;
  if (s_in == 128 && s_out == 128) {
    resample128_kernel<<<blocks_for(n_rays), kWarps * 32, 0, (cudaStream_t)stream>>>(
        sdist_in, weights_in, cw_in, u, near_, far_, n_rays, padding, anneal, dom_lo, dom_hi, sdist_out, tdist_out, cw_out, idx_out);
    RN_CUDA_CHECK_LAUNCH();
    return RN_OK;
  }
  size_t smem = (size_t)kWarps * (2 * (s_in + 1) + s_out) * sizeof(float);
  if (int rc = ensure_smem(resample_kernel, smem)) return rc;
  resample_kernel<<<blocks_for(n_rays), kWarps * 32, smem, (cudaStream_t)stream>>>(
      sdist_in, weights_in, cw_in, u, near_, far_, n_rays, s_in, s_out, padding, anneal, dom_lo, dom_hi, sdist_out, tdist_out,
      cw_out, idx_out);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

extern "C" int rn_max_dilate_weights(const float* t, const float* w, int64_t n_rays, int s, float dilation, float dom_lo,
                                     float dom_hi, int renormalize, int trim, float* t_out, float* w_out, void* stream) {
  RnProfScope prof_scope(RN_PROF_RAY, (cudaStream_t)stream);
  if (n_rays < 0 || s < 1 || s > 1024) return rn_set_error(RN_ERR_ARG, "rn_max_dilate_weights: need 1 <= s <= 1024");
  if (n_rays == 0) return RN_OK;
  int levels = 1;
  while ((1 << levels) <= s) ++levels;            // floor(log2 s) + 1 sparse-table levels
  const size_t smem = (size_t)kWarps * ((s + 1) + 2 * s + (3 * s + 1) + (size_t)levels * s) * sizeof(float);
  if (smem > 200 * 1024) return rn_set_error(RN_ERR_UNSUPPORTED, "rn_max_dilate_weights: too many samples for the shared-memory tables");
  if (int rc = ensure_smem(max_dilate_kernel, smem)) return rc;
  max_dilate_kernel<<<blocks_for(n_rays), kWarps * 32, smem, (cudaStream_t)stream>>>(t, w, n_rays, s, levels, dilation, dom_lo, dom_hi,
                                                                                     renormalize, trim, t_out, w_out);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

extern "C" int rn_composite_fwd(const float* density, const float* tdist, const float* dirs, const float* far_,
                                const float* rgb, const float* diffuse, const float* specular, const float* normals,
                                const float* normals_pred, const float* roughness, const float* tint, int64_t n_rays,
                                int s, float bg, float* weights_out, float* comp_out, float* extras_out,
                                double* pct_out, void* stream) {
  RnProfScope prof_scope(RN_PROF_RAY, (cudaStream_t)stream);
  if (n_rays < 0 || s < 1) return rn_set_error(RN_ERR_ARG, "rn_composite_fwd: bad sizes");
  if (n_rays == 0) return RN_OK;
  if (s == 128) {
    composite_fwd128_kernel<<<blocks_for(n_rays), kWarps * 32, 0, (cudaStream_t)stream>>>(
        density, tdist, dirs, far_, rgb, diffuse, specular, normals, normals_pred, roughness, tint, n_rays, bg, weights_out,
        comp_out, extras_out, pct_out);
    RN_CUDA_CHECK_LAUNCH();
    return RN_OK;
  }
  size_t smem = (size_t)kWarps * 3 * (s + 2) * sizeof(float);
  if (int rc = ensure_smem(composite_fwd_kernel, smem)) return rc;
  composite_fwd_kernel<<<blocks_for(n_rays), kWarps * 32, smem, (cudaStream_t)stream>>>(
      density, tdist, dirs, far_, rgb, diffuse, specular, normals, normals_pred, roughness, tint, n_rays, s, bg,
      weights_out, comp_out, extras_out, pct_out);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

extern "C" int rn_composite_bwd(const float* density, const float* tdist, const float* dirs, const float* rgb,
                                const float* diffuse, const float* specular, const float* normals,
                                const float* normals_pred, const float* roughness, const float* tint,
                                const float* weights, const float* comp, const float* g_weights, const float* g_comp,
                                const float* g_extras, int64_t n_rays, int s, float bg, float* d_density, float* d_rgb,
                                float* d_diffuse, float* d_specular, float* d_normals_pred, float* d_roughness,
                                float* d_tint, void* stream) {
  RnProfScope prof_scope(RN_PROF_RAY, (cudaStream_t)stream);
  if (n_rays < 0 || s < 1) return rn_set_error(RN_ERR_ARG, "rn_composite_bwd: bad sizes");
  if (n_rays == 0) return RN_OK;
  if (s == 128) {
    composite_bwd128_kernel<<<blocks_for(n_rays), kWarps * 32, 0, (cudaStream_t)stream>>>(
        density, tdist, dirs, rgb, diffuse, specular, normals, normals_pred, roughness, tint, weights, comp, g_weights, g_comp,
        g_extras, n_rays, bg, d_density, d_rgb, d_diffuse, d_specular, d_normals_pred, d_roughness, d_tint);
    RN_CUDA_CHECK_LAUNCH();
    return RN_OK;
  }
  size_t smem = (size_t)kWarps * 3 * (s + 2) * sizeof(float);
  if (int rc = ensure_smem(composite_bwd_kernel, smem)) return rc;
  composite_bwd_kernel<<<blocks_for(n_rays), kWarps * 32, smem, (cudaStream_t)stream>>>(
      density, tdist, dirs, rgb, diffuse, specular, normals, normals_pred, roughness, tint, weights, comp, g_weights,
      g_comp, g_extras, n_rays, s, bg, d_density, d_rgb, d_diffuse, d_specular, d_normals_pred, d_roughness, d_tint);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

extern "C" int rn_lossfun_outer_fwd(const float* t, const float* w, const float* t_env, const float* w_env,
                                    int64_t n_rays, int s, int se, float* loss_out, void* stream) {
  RnProfScope prof_scope(RN_PROF_RAY, (cudaStream_t)stream);
  if (n_rays == 0) return RN_OK;
  size_t smem = (size_t)kWarps * 3 * (se + 1) * sizeof(float);
  if (int rc = ensure_smem(lossfun_outer_kernel<false>, smem)) return rc;
  lossfun_outer_kernel<false><<<blocks_for(n_rays), kWarps * 32, smem, (cudaStream_t)stream>>>(
      t, w, t_env, w_env, nullptr, n_rays, s, se, loss_out);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

extern "C" int rn_lossfun_outer_bwd(const float* t, const float* w, const float* t_env, const float* w_env,
                                    const float* g_loss, int64_t n_rays, int s, int se, float* d_w_env, void* stream) {
  RnProfScope prof_scope(RN_PROF_RAY, (cudaStream_t)stream);
  if (n_rays == 0) return RN_OK;
  size_t smem = (size_t)kWarps * 3 * (se + 1) * sizeof(float);
  if (int rc = ensure_smem(lossfun_outer_kernel<true>, smem)) return rc;
  lossfun_outer_kernel<true><<<blocks_for(n_rays), kWarps * 32, smem, (cudaStream_t)stream>>>(
      t, w, t_env, w_env, g_loss, n_rays, s, se, d_w_env);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

extern "C" int rn_distortion_fwd(const float* t, const float* w, int64_t n_rays, int s, float* loss_out, void* stream) {
  RnProfScope prof_scope(RN_PROF_RAY, (cudaStream_t)stream);
  if (n_rays == 0) return RN_OK;
  size_t smem = (size_t)kWarps * 2 * s * sizeof(float);
  if (int rc = ensure_smem(distortion_kernel<false>, smem)) return rc;
  distortion_kernel<false><<<blocks_for(n_rays), kWarps * 32, smem, (cudaStream_t)stream>>>(t, w, nullptr, n_rays, s, loss_out);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

extern "C" int rn_distortion_bwd(const float* t, const float* w, const float* g_loss, int64_t n_rays, int s, float* d_w,
                                 void* stream) {
  RnProfScope prof_scope(RN_PROF_RAY, (cudaStream_t)stream);
  if (n_rays == 0) return RN_OK;
  size_t smem = (size_t)kWarps * 2 * s * sizeof(float);
  if (int rc = ensure_smem(distortion_kernel<true>, smem)) return rc;
  distortion_kernel<true><<<blocks_for(n_rays), kWarps * 32, smem, (cudaStream_t)stream>>>(t, w, g_loss, n_rays, s, d_w);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

// ---------------------------------------------------------------------------------------------
// train_utils.compute_data_loss (train_utils.py:33-88) of ONE level in one launch: the three sums the reference takes with
// ~10 elementwise / reduction launches over [N,3] tensors.  One block, fp64 partial sums combined in a fixed order, so the
// result is bit-reproducible (and closer to exact than the reference's own fp32 sum).
//   out[0] = sum lm (rgb - gt)^2     (the 'mses' statistic)
//   out[1] = sum lm term(rgb - gt)   term = r^2 ('mse') or sqrt(r^2 + pad^2) ('charb')
//   out[2] = sum lm                  (lm broadcast over the 3 channels; lm == NULL: 1, disable_multiscale_loss)
// backward: d_rgb = lm (g0 2 r + g1 dterm/dr).
// ---------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ double block_sum_1024(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  v = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(1024)
data_loss_fwd_kernel(const float* __restrict__ rgb, const float* __restrict__ gt, const float* __restrict__ lm, int64_t n,
                     int charb, float pad2, float* __restrict__ out) {
  __shared__ double sh[32];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int64_t i = threadIdx.x; i < 3 * n; i += blockDim.x) {
    const float m = lm ? lm[i / 3] : 1.f;
    const float r = rgb[i] - gt[i];
    const float r2 = r * r;
    s0 += (double)(m * r2);
    s1 += (double)(m * (charb ? sqrtf(r2 + pad2) : r2));
    s2 += (double)m;
  }
  s0 = block_sum_1024(s0, sh);
  s1 = block_sum_1024(s1, sh);
  s2 = block_sum_1024(s2, sh);
  if (threadIdx.x == 0) {
    out[0] = (float)s0;
    out[1] = (float)s1;
    out[2] = (float)s2;
  }
}

__global__ void __launch_bounds__(256)
data_loss_bwd_kernel(const float* __restrict__ rgb, const float* __restrict__ gt, const float* __restrict__ lm,
                     const float* __restrict__ g, int64_t n, int charb, float pad2, float* __restrict__ d_rgb) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= 3 * n) return;
  const float m = lm ? lm[i / 3] : 1.f;
  const float r = rgb[i] - gt[i];
  const float dterm = charb ? r / sqrtf(r * r + pad2) : 2.f * r;
  d_rgb[i] = m * (g[0] * 2.f * r + g[1] * dterm);
}
}  // namespace

extern "C" int rn_data_loss_fwd(const float* rgb, const float* gt, const float* lossmult, int64_t n_rays, int charb,
                                float charb_padding, float* sums_out, void* stream) {
  RnProfScope prof_scope(RN_PROF_RAY, (cudaStream_t)stream);
  if (!rgb || !gt || !sums_out || n_rays < 0) return rn_set_error(RN_ERR_ARG, "rn_data_loss_fwd: bad arguments");
  data_loss_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rgb, gt, lossmult, n_rays, charb, charb_padding * charb_padding, sums_out);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

extern "C" int rn_data_loss_bwd(const float* rgb, const float* gt, const float* lossmult, const float* g_sums, int64_t n_rays,
                                int charb, float charb_padding, float* d_rgb, void* stream) {
  RnProfScope prof_scope(RN_PROF_RAY, (cudaStream_t)stream);
  if (n_rays == 0) return RN_OK;
  if (!rgb || !gt || !g_sums || !d_rgb || n_rays < 0) return rn_set_error(RN_ERR_ARG, "rn_data_loss_bwd: bad arguments");
  data_loss_bwd_kernel<<<(unsigned)((3 * n_rays + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rgb, gt, lossmult, g_sums, n_rays, charb, charb_padding * charb_padding, d_rgb);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

extern "C" int rn_normal_losses_fwd(const float* weights, const float* normals, const float* normals_pred,
                                    const float* viewdirs, int64_t n_rays, int s, int ori_target_is_pred, float* ori_out,
                                    float* pred_out, void* stream) {
  RnProfScope prof_scope(RN_PROF_RAY, (cudaStream_t)stream);
  if (n_rays == 0) return RN_OK;
  if (!weights || !normals_pred || !viewdirs || !ori_out || !pred_out || s < 1) return rn_set_error(RN_ERR_ARG, "rn_normal_losses_fwd: bad arguments");
  if (!ori_target_is_pred && !normals) return rn_set_error(RN_ERR_ARG, "rn_normal_losses_fwd: orientation target 'normals' needs the density-gradient normals");
  size_t smem = (size_t)kWarps * 6 * s * sizeof(float);
  if (int rc = ensure_smem(normal_losses_kernel<false>, smem)) return rc;
  normal_losses_kernel<false><<<blocks_for(n_rays), kWarps * 32, smem, (cudaStream_t)stream>>>(
      weights, normals, normals_pred, viewdirs, nullptr, nullptr, n_rays, s, ori_target_is_pred, ori_out, pred_out, nullptr, nullptr);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

extern "C" int rn_normal_losses_bwd(const float* weights, const float* normals, const float* normals_pred,
                                    const float* viewdirs, const float* g_ori, const float* g_pred, int64_t n_rays, int s,
                                    int ori_target_is_pred, float* d_weights, float* d_normals_pred, void* stream) {
  RnProfScope prof_scope(RN_PROF_RAY, (cudaStream_t)stream);
  if (n_rays == 0) return RN_OK;
  if (!weights || !normals_pred || !viewdirs || !d_weights || !d_normals_pred || s < 1) return rn_set_error(RN_ERR_ARG, "rn_normal_losses_bwd: bad arguments");
  if (!ori_target_is_pred && !normals) return rn_set_error(RN_ERR_ARG, "rn_normal_losses_bwd: orientation target 'normals' needs the density-gradient normals");
  size_t smem = (size_t)kWarps * 6 * s * sizeof(float);
  if (int rc = ensure_smem(normal_losses_kernel<true>, smem)) return rc;
  normal_losses_kernel<true><<<blocks_for(n_rays), kWarps * 32, smem, (cudaStream_t)stream>>>(
      weights, normals, normals_pred, viewdirs, g_ori, g_pred, n_rays, s, ori_target_is_pred, nullptr, nullptr, d_weights, d_normals_pred);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}
