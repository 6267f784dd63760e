// On-device ray generation: camera_utils.pixels_to_rays (camera_utils.py:502-614) for perspective
// cameras without lens distortion, with the optional NDC conversion (camera_utils.py:31-97, near = 1).
//
// One thread per pixel.  The reference's default numpy path promotes "integer pixel + 0.5" to float64 and keeps
// float64 through the matrix products, norms and the NDC projection; the kernel does the same arithmetic in
// double (no FMA contraction: explicit __dmul_rn / __dadd_rn) and rounds the results to the fp32 fields utils.Rays
// carries.  The work is a few hundred DP operations and 60 bytes per ray: bound by launch latency / HBM writes.
#include "common.cuh"

namespace {

struct D3 { double x, y, z; };

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
// row-major 3x3 (row stride `ld`) times vector, summed left to right like numpy's matmul on a length-3 axis
__device__ __forceinline__ D3 mat3_vec(const float* __restrict__ m, int ld, D3 v) {
  D3 r;
  r.x = add(add(mul((double)m[0], v.x), mul((double)m[1], v.y)), mul((double)m[2], v.z));
  r.y = add(add(mul((double)m[ld], v.x), mul((double)m[ld + 1], v.y)), mul((double)m[ld + 2], v.z));
  r.z = add(add(mul((double)m[2 * ld], v.x), mul((double)m[2 * ld + 1], v.y)), mul((double)m[2 * ld + 2], v.z));
  return r;
}
__device__ __forceinline__ double norm3(D3 a) { return sqrt(add(add(mul(a.x, a.x), mul(a.y, a.y)), mul(a.z, a.z))); }
__device__ __forceinline__ D3 sub3(D3 a, D3 b) { return {add(a.x, -b.x), add(a.y, -b.y), add(a.z, -b.z)}; }

// convert_to_ndc (near = 1): origins shifted to the near plane, then projected; returns origins_ndc, directions_ndc
__device__ __forceinline__ void to_ndc(D3 o, D3 d, double xmult, double ymult, D3& on, D3& dn) {
  const double t = -add(1.0, o.z) / d.z;
  o = {add(o.x, mul(t, d.x)), add(o.y, mul(t, d.y)), add(o.z, mul(t, d.z))};
  on = {mul(xmult, o.x) / o.z, mul(ymult, o.y) / o.z, -1.0};
  const D3 inf = {mul(xmult, d.x) / d.z, mul(ymult, d.y) / d.z, 1.0};
  dn = sub3(inf, on);
}

__global__ void __launch_bounds__(256)
pixels_to_rays_kernel(const int32_t* __restrict__ px, const int32_t* __restrict__ py, const int32_t* __restrict__ cam,
                      const float* __restrict__ pixtocams, const float* __restrict__ camtoworlds,
                      const float* __restrict__ pixtocam_ndc, int64_t n, float* __restrict__ origins,
                      float* __restrict__ directions, float* __restrict__ viewdirs, float* __restrict__ radii,
                      float* __restrict__ imageplane) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int c = cam ? cam[i] : 0;
  const float* p2c = pixtocams + (size_t)c * 9;
  const float* c2w = camtoworlds + (size_t)c * 12;
  const double x = (double)px[i], y = (double)py[i];
  D3 world[3];
  double ipx = 0.0, ipy = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {   // the pixel, its +x neighbour, its +y neighbour
    const D3 pix = {add(x + (k == 1 ? 1.0 : 0.0), 0.5), add(y + (k == 2 ? 1.0 : 0.0), 0.5), 1.0};
    D3 cd = mat3_vec(p2c, 3, pix);
    cd.y = -cd.y;   // OpenCV -> OpenGL
    cd.z = -cd.z;
    if (k == 0) { ipx = cd.x; ipy = cd.y; }
    world[k] = mat3_vec(c2w, 4, cd);
  }
  D3 o = {(double)c2w[3], (double)c2w[7], (double)c2w[11]};
  D3 d = world[0];
  const double dn = norm3(d);
  const D3 v = {d.x / dn, d.y / dn, d.z / dn};
  double dxn, dyn;
  if (!pixtocam_ndc) {
    dxn = norm3(sub3(world[1], d));
    dyn = norm3(sub3(world[2], d));
  } else {
    const double xmult = 1.0 / (double)pixtocam_ndc[2], ymult = 1.0 / (double)pixtocam_ndc[5];
    D3 on, dnn, ox, oy, tmp;
    to_ndc(o, world[1], xmult, ymult, ox, tmp);
    to_ndc(o, world[2], xmult, ymult, oy, tmp);
    to_ndc(o, d, xmult, ymult, on, dnn);
    dxn = norm3(sub3(ox, on));
    dyn = norm3(sub3(oy, on));
    o = on;
    d = dnn;
  }
  const double rad = mul(mul(0.5, add(dxn, dyn)), 2.0) / sqrt(12.0);
  origins[i * 3 + 0] = (float)o.x; origins[i * 3 + 1] = (float)o.y; origins[i * 3 + 2] = (float)o.z;
  directions[i * 3 + 0] = (float)d.x; directions[i * 3 + 1] = (float)d.y; directions[i * 3 + 2] = (float)d.z;
  viewdirs[i * 3 + 0] = (float)v.x; viewdirs[i * 3 + 1] = (float)v.y; viewdirs[i * 3 + 2] = (float)v.z;
  radii[i] = (float)rad;
  imageplane[i * 2 + 0] = (float)ipx;
  imageplane[i * 2 + 1] = (float)ipy;
}

}  // namespace

extern "C" int rn_pixels_to_rays(const int32_t* pix_x, const int32_t* pix_y, const int32_t* cam_idx, const float* pixtocams,
                                 const float* camtoworlds, const float* pixtocam_ndc, int64_t n, float* origins,
                                 float* directions, float* viewdirs, float* radii, float* imageplane, void* stream) {
  if (n < 0 || !pix_x || !pix_y || !pixtocams || !camtoworlds || !origins || !directions || !viewdirs || !radii || !imageplane)
    return rn_set_error(RN_ERR_ARG, "rn_pixels_to_rays: bad arguments");
  if (n == 0) return RN_OK;
  pixels_to_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      pix_x, pix_y, cam_idx, pixtocams, camtoworlds, pixtocam_ndc, n, origins, directions, viewdirs, radii, imageplane);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}
