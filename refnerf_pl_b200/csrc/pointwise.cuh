// Launch wrappers of the non-GEMM kernels of the NerfMLP pipeline (definitions in pointwise.cu).
#pragma once
#include "common.cuh"

namespace rn {

struct MlpScalars {
  int srgb_mapping, srgb_normalization;
  float density_bias, roughness_bias, rgb_premultiplier, rgb_bias, rgb_padding;
};

// K1: cast_rays + lift + IPE for rows [row0, row0+rows) of the level -> out[:, 0:ncols] (cols >= 96 zero)
int launch_encode(int prec, const float* tdist, const float* origins, const float* dirs, const float* radii, int s,
                  int64_t row0, int64_t rows, ActBuf out, int ncols, cudaStream_t st);
// normals = -l2_normalize(d raw_density / d means) from gx0 (+ gx0b if given) = d raw_density / d ipe features
int launch_ipe_grad_normals(const float* gx0, const float* gx0b, int ld, const float* tdist, const float* origins, const float* dirs,
                            const float* radii, int s, int64_t row0, int64_t rows, float* normals_out, float gscale,
                            cudaStream_t st);
// seed of the normals pass: out[r, j] = wd[j] * (a8[r, j] > 0)
int launch_density_grad_seed(int prec, ActBuf a8, const float* wd, ActBuf out, int64_t rows, cudaStream_t st);
// K2: heads activations + reflect + IDE + n.v -> v0[:, 128:256] and the per-sample outputs
int launch_heads_prologue_fwd(int prec, const float* heads_raw, const float* viewdirs, int s, int64_t row0, int64_t rows,
                              MlpScalars sc, ActBuf v0, float* density, float* normals_pred, float* grad_pred,
                              float* roughness, float* tint, cudaStream_t st);
// backward of K2: dv0f [rows,256] f32 (cols 128..200 used) + upstream grads -> d_scal act [rows,16]
// (0 density,1..3 grad_pred,4 roughness,5..7 diffuse,8..10 tint)
int launch_heads_prologue_bwd(int prec, const float* heads_raw, const float* viewdirs, int s, int64_t row0, int64_t rows,
                              MlpScalars sc, const float* dv0f, const float* dcolor, const float* g_density,
                              const float* g_normals_pred, const float* g_grad_pred, const float* g_roughness,
                              const float* g_tint, ActBuf d_scal, const float* dv0_unscale, cudaStream_t st);
// colour combine (models.py:699-729): rgb_raw [rows,4], heads_raw [rows,16] -> rgb, diffuse, specular [.,3]
int launch_color_fwd(const float* rgb_raw, const float* heads_raw, int64_t rows, MlpScalars sc, float* rgb,
                     float* diffuse, float* specular, cudaStream_t st);
// backward: -> d_rgb_raw act [rows,16] (3 real), dcolor f32 [rows,8]: d raw_diffuse(0:3), d tint(3:6)
int launch_color_bwd(int prec, const float* rgb_raw, const float* heads_raw, int64_t rows, MlpScalars sc,
                     const float* g_rgb, const float* g_diffuse, const float* g_specular, ActBuf d_rgb_raw,
                     float* dcolor, cudaStream_t st);
// column sums of an activation-format matrix, atomically added to out[ncols]
int launch_colsum(int prec, ActBuf a, int64_t rows, int ncols, float* out, cudaStream_t st);
// f32 [rows, ld] columns [c0, c0+ncols) -> activation format (ncols % 8 == 0)
int launch_f32_to_act(int prec, const float* src, int ld, int c0, int64_t rows, int ncols, ActBuf dst, cudaStream_t st);
// weight packing: dst(r0+i, c0+j) = src(i,j) (or transposed) for an [nr,nc] fp32 source block
int launch_pack_segment(int prec, const float* src, int src_ld, int nr, int nc, int transpose, void* dst_hi, void* dst_lo,
                        int dst_ld, int r0, int c0, cudaStream_t st);
// grad unpacking: dst[i*dst_ld + j] += src[(r0+i)*src_ld + c0 + j]
int launch_unpack_add(const float* src, int src_ld, int r0, int c0, int nr, int nc, float* dst, int dst_ld,
                      const float* scale, cudaStream_t st);
// table-driven pack / unpack: one launch for every weight block of a NerfMLP
constexpr int kMaxBatchSegs = 64;
struct PackSeg {      // dst[(r0 + i) * dst_ld + c0 + j] (or transposed) = convert(src[i * src_ld + j]),  i < nr, j < nc
  const float* src;
  void* dst_hi;
  void* dst_lo;       // bf16x3 only
  int src_ld, nr, nc, transpose, dst_ld, r0, c0;
  int f32;            // destination is f32 whatever the precision (biases)
};
struct PackTable { int n; PackSeg seg[kMaxBatchSegs]; };
struct UnpackSeg {    // dst[i * dst_ld + j] += src[(r0 + i) * src_ld + c0 + j] * (scale ? *scale : 1)
  const float* src;
  float* dst;
  const float* scale;
  int src_ld, r0, c0, nr, nc, dst_ld;
};
struct UnpackTable { int n; UnpackSeg seg[kMaxBatchSegs]; };
int launch_pack_batch(int prec, const PackTable& t, cudaStream_t st);
int launch_unpack_add_batch(const UnpackTable& t, cudaStream_t st);
// fp16 mode: dynamic power-of-two scaling of the gradient tiles entering the dgrad chains (layout of `scal` in
// pointwise.cu).  amax kernels accumulate max|x| as a uint bit pattern (zero the slot first).
int launch_amax_f32(const float* src, int64_t n, uint32_t* dst, cudaStream_t st);
int launch_amax_f16(const void* src, int ld, int ncols, int64_t rows, uint32_t* dst, cudaStream_t st);
int launch_grad_scale(float* scal, int stage, cudaStream_t st);
int launch_scale_to_f16(const float* src, int ld_src, int64_t rows, int ncols, void* dst, int ld_dst, const float* scale,
                        cudaStream_t st);
int launch_rescale_f16(void* buf, int ld, int ncols, int64_t rows, const float* ratio, cudaStream_t st);
// standalone IDE (unit test surface)
int launch_ide(const float* dirs, const float* kappa_inv, int64_t n, float* out, cudaStream_t st);

}  // namespace rn
