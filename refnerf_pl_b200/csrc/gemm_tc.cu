// tcgen05 / TMEM / TMA GEMMs for sm_100a.
//
//   gemm_tc_kernel : C[128-row tile, n<=256] = A[., k] * B[n, k]^T, both operands K-major bf16 staged by
//                    TMA (128B swizzle) into a multi-stage shared-memory ring; one elected thread issues
//                    tcgen05.mma (kind::f16, M=128, N=n, K=16) into a double-buffered fp32 TMEM
//                    accumulator; four epilogue warps drain TMEM with tcgen05.ld and apply
//                    bias / ReLU / ReLU-mask / format conversion while the next tile's MMAs run.
//                    SPLIT=3 issues hi*hi + lo*hi + hi*lo per K step (split-bf16, ~16-bit mantissa).
//   wgrad_tc_kernel: dW[128 x kx] += dY^T X over a row range: both operands are MN-major views of the
//                    row-major [rows, features] buffers (same TMA boxes, transposed descriptors), split
//                    over rows across CTAs, reduced into fp32 with red.global.add.
//
// Persistent warp-specialised layout (256 threads): warp 0 TMA producer, warp 1 MMA issuer,
// warp 2 TMEM allocator, warps 4-7 epilogue (warp%4 = TMEM lane quadrant).
#include "chain_common.cuh"

namespace rn {
// fp16 mode: number of (wgrad launch, thread) observations of a SATURATED gradient element (|dY| = 65504 after
// cvt.satfinite) since the last reset.  The dgrad chains scale every chain by one power of two chosen from its seed
// tile (pointwise.cu); a gradient that grows more than 16x through a chain would be clamped silently.  The wgrad
// kernels already visit every dY element in shared memory (bias column sums), so the check rides there for free.
__device__ unsigned long long g_fp16_saturated = 0ull;
namespace {
using namespace tc;
using chain::elect_one_sync;
using chain::uniform_warp_idx;

// mask bits for 16 columns: bit i set iff mask[row, col+i] > 0 (bf16 hi plane: sign/zero test on the raw bits)
__device__ __forceinline__ uint32_t mask_bits16(const uint32_t* m) {
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t lo = m[i] & 0xffffu, hi = m[i] >> 16;
    bits |= ((lo != 0u && lo < 0x8000u) ? 1u : 0u) << (2 * i);
    bits |= ((hi != 0u && hi < 0x8000u) ? 1u : 0u) << (2 * i + 1);
  }
  return bits;
}

// epilogue for 16 consecutive columns of one row (bf16 modes); v holds the fp32 accumulators.
template <int PREC>
__device__ __forceinline__ void epi_store16(const GemmEpilogue& e, size_t row, int col, float* v, const uint32_t* mask_raw) {
  if (e.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(e.bias + col);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b = __ldg(b4 + i);
      v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
  }
  if (e.relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (e.mask.hi) {
    const uint32_t bits = mask_bits16(mask_raw);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = ((bits >> i) & 1u) ? v[i] : 0.f;
  }
  if (e.out.hi && col < e.out_cols) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    stg256(reinterpret_cast<uint16_t*>(e.out.hi) + row * e.out.ld + col, h);
    if (PREC == RN_PREC_BF16X3) stg256(reinterpret_cast<uint16_t*>(e.out.lo) + row * e.out.ld + col, l);
  }
  if (e.f32 && col + 16 > e.f32_col0 && col < e.f32_col0 + e.f32_cols) {
    const int c0 = col - e.f32_col0;
    float* p = e.f32 + row * e.f32_ld + c0;
    if (c0 >= 0 && c0 + 16 <= e.f32_cols && (e.f32_ld & 7) == 0) {
      uint32_t a[8];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (e.f32_accum) {
          ldg256(p + 8 * half, a);
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = __float_as_uint(__uint_as_float(a[i]) + v[8 * half + i]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = __float_as_uint(v[8 * half + i]);
        }
        stg256(p + 8 * half, a);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = c0 + i;
        if (c >= 0 && c < e.f32_cols) p[i] = e.f32_accum ? (p[i] + v[i]) : v[i];
      }
    }
  }
}

struct TcMaps {
  CUtensorMap a1_hi, a1_lo, a2_hi, a2_lo, b_hi, b_lo;
};

template <int SPLIT>
struct Cfg {
  static constexpr int kPlanes = SPLIT == 3 ? 2 : 1;
  static constexpr int kStages = SPLIT == 3 ? 2 : 4;
  static constexpr int kABytes = kBM * kBK * 2;        // 16 KB
  static constexpr int kBBytes = 256 * kBK * 2;        // 32 KB (n <= 256)
  static constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// ---------------------------------------------------------------------------------------------
// forward / dgrad GEMM
// ---------------------------------------------------------------------------------------------
template <int SPLIT>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ TcMaps maps, int64_t m, int n, int kb1, int kb2, GemmEpilogue epi) {
  using C = Cfg<SPLIT>;
  constexpr int PREC = SPLIT == 3 ? RN_PREC_BF16X3 : RN_PREC_BF16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* tfull = bars + 2 * C::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int64_t num_tiles = (m + kBM - 1) / kBM;
  const int kbt = kb1 + kb2;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.a1_hi);
    tma_prefetch_desc(&maps.b_hi);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (whole warp runs the loop, one elected lane issues: no divergence waterfall around the
    // uniform-datapath TMA / MMA instructions, see chain_common.cuh) =====
    const uint32_t stage_tx = (uint32_t)C::kPlanes * (uint32_t)(C::kABytes + n * kBK * 2);
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (int)(tile * kBM);
      for (int kb = 0; kb < kbt; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sbase = smem + stage * C::kStageBytes;
        const bool first = kb < kb1;
        const int ka = (first ? kb : kb - kb1) * kBK;
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full[stage], stage_tx);
          tma_load_2d(sbase, first ? &maps.a1_hi : &maps.a2_hi, &full[stage], ka, m0);
          tma_load_2d(sbase + C::kPlanes * C::kABytes, &maps.b_hi, &full[stage], kb * kBK, 0);
          if (SPLIT == 3) {
            tma_load_2d(sbase + C::kABytes, first ? &maps.a1_lo : &maps.a2_lo, &full[stage], ka, m0);
            tma_load_2d(sbase + 2 * C::kABytes + C::kBBytes, &maps.b_lo, &full[stage], kb * kBK, 0);
          }
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp, one elected lane issues) =====
    const uint32_t idesc = make_idesc(n, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int buf = 0;
    uint32_t tphase = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[buf], tphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)buf * 256u;
      for (int kb = 0; kb < kbt; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa_hi = smem_u32(smem + stage * C::kStageBytes);
        const uint32_t sb_hi = sa_hi + C::kPlanes * C::kABytes;
        const uint32_t sa_lo = sa_hi + C::kABytes;
        const uint32_t sb_lo = sb_hi + C::kBBytes;
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kBK / kUmmaK; ++kk) {
            const uint32_t koff = kk * kUmmaK * 2;  // bytes inside the 128B swizzle atom
            const uint64_t da = make_desc(sa_hi + koff, 16, 1024);
            const uint64_t db = make_desc(sb_hi + koff, 16, 1024);
            umma_bf16(tmem_d, da, db, idesc, (kb | kk) ? 1u : 0u);
            if (SPLIT == 3) {
              const uint64_t dal = make_desc(sa_lo + koff, 16, 1024);
              const uint64_t dbl = make_desc(sb_lo + koff, 16, 1024);
              umma_bf16(tmem_d, dal, db, idesc, 1u);
              umma_bf16(tmem_d, da, dbl, idesc, 1u);
            }
          }
          umma_commit(&empty[stage]);  // frees the smem slot once the MMAs above have read it
          if (kb == kbt - 1) umma_commit(&tfull[buf]);   // accumulator complete
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (++buf == 2) { buf = 0; tphase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp - 4;
    int buf = 0;
    uint32_t tphase = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tfull[buf], tphase);
      tc_fence_after();
      const int64_t row = tile * kBM + q * 32 + lane;
      const uint32_t taddr = tmem_base + (uint32_t)buf * 256u + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < n; c0 += 32) {
        uint32_t r[32];
        uint32_t mk[2][8];
        tmem_ld32(taddr + (uint32_t)c0, r);
        if (epi.mask.hi && row < m) {   // issue the ReLU-mask loads while the TMEM load is in flight
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if (c0 + 16 * h < n) ldg256(reinterpret_cast<const uint16_t*>(epi.mask.hi) + (size_t)row * epi.mask.ld + c0 + 16 * h, mk[h]);
        }
        tmem_ld_wait();
        if (row < m) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (c0 + 16 * h < n) {
              float v[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[16 * h + e]);
              epi_store16<PREC>(epi, (size_t)row, c0 + 16 * h, v, mk[h]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
      if (++buf == 2) { buf = 0; tphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// wgrad: dW[n0+i, j] += sum_r dY[r, n0+i] * X[r, j]
// ---------------------------------------------------------------------------------------------
struct WgMaps {
  CUtensorMap dy_hi, dy_lo, x_hi, x_lo;
};

template <int SPLIT>
__global__ void __launch_bounds__(256, 1)
wgrad_tc_kernel(const __grid_constant__ WgMaps maps, int64_t m, int64_t rows_per_cta, int n0, int n_real, int kx,
                int k_real, float* __restrict__ out, int out_ld, float* __restrict__ bias_out) {
  using C = Cfg<SPLIT>;
  const bool do_bias = bias_out != nullptr;   // also accumulate bias_out[n0 + f] += sum_r dY[r, n0 + f] (hi + lo planes)
  constexpr int kRowBlk = 64;                       // rows (reduction) per stage
  constexpr int kBoxBytes = kRowBlk * 128;          // one [64 rows x 64 features] TMA box = 8 KB
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* tfull = bars + 2 * C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = min(m, r_begin + rows_per_cta);
  const int nblk = r_end > r_begin ? (int)((r_end - r_begin + kRowBlk - 1) / kRowBlk) : 0;
  const int xboxes = kx / 64;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.dy_hi);
    tma_prefetch_desc(&maps.x_hi);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], do_bias ? 5 : 1);   // MMA commit (+ the four epilogue warps that sum the staged dY tile)
    }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (nblk > 0) {
    if (warp == 0 && lane == 0) {
      const uint32_t stage_tx = (uint32_t)C::kPlanes * (uint32_t)(2 + xboxes) * kBoxBytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int b = 0; b < nblk; ++b) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sbase = smem + stage * C::kStageBytes;
        mbar_arrive_expect_tx(&full[stage], stage_tx);
        const int r0 = (int)(r_begin + (int64_t)b * kRowBlk);
        // layout per stage: dY_hi (2 boxes) | dY_lo | X_hi (4 boxes) | X_lo
        for (int i = 0; i < 2; ++i) tma_load_2d(sbase + i * kBoxBytes, &maps.dy_hi, &full[stage], n0 + i * 64, r0);
        for (int i = 0; i < xboxes; ++i)
          tma_load_2d(sbase + C::kPlanes * C::kABytes + i * kBoxBytes, &maps.x_hi, &full[stage], i * 64, r0);
        if (SPLIT == 3) {
          for (int i = 0; i < 2; ++i)
            tma_load_2d(sbase + C::kABytes + i * kBoxBytes, &maps.dy_lo, &full[stage], n0 + i * 64, r0);
          for (int i = 0; i < xboxes; ++i)
            tma_load_2d(sbase + 2 * C::kABytes + C::kBBytes + i * kBoxBytes, &maps.x_lo, &full[stage], i * 64, r0);
        }
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1 && lane == 0) {
      const uint32_t idesc = make_idesc(kx, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int b = 0; b < nblk; ++b) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa_hi = smem_u32(smem + stage * C::kStageBytes);
        const uint32_t sb_hi = sa_hi + C::kPlanes * C::kABytes;
        const uint32_t sa_lo = sa_hi + C::kABytes;
        const uint32_t sb_lo = sb_hi + C::kBBytes;
#pragma unroll
        for (int kk = 0; kk < kRowBlk / kUmmaK; ++kk) {
          const uint32_t koff = kk * kUmmaK * 128;  // 16 rows of 128 B
          // MN-major, 128B swizzle: LBO = stride between 64-feature blocks, SBO = stride between 8-row groups
          const uint64_t da = make_desc(sa_hi + koff, kBoxBytes, 1024);
          const uint64_t db = make_desc(sb_hi + koff, kBoxBytes, 1024);
          umma_bf16(tmem_base, da, db, idesc, (b | kk) ? 1u : 0u);
          if (SPLIT == 3) {
            const uint64_t dal = make_desc(sa_lo + koff, kBoxBytes, 1024);
            const uint64_t dbl = make_desc(sb_lo + koff, kBoxBytes, 1024);
            umma_bf16(tmem_base, dal, db, idesc, 1u);
            umma_bf16(tmem_base, da, dbl, idesc, 1u);
          }
        }
        umma_commit(&empty[stage]);
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull[0]);
    } else if (warp >= 4) {
      const int q = warp - 4;
      if (do_bias) {
        // column sums of the dY slab from the staged (128B-swizzled, MN-major) tiles: thread t owns features 2t, 2t+1
        const int t = threadIdx.x - 128;
        const bool active = t < 64;
        const int f = 2 * t;
        const int box = f >> 6, fi = f & 63;
        float s0 = 0.f, s1 = 0.f;
        int stage = 0;
        uint32_t phase = 0;
        for (int b = 0; b < nblk; ++b) {
          mbar_wait(&full[stage], phase);
          if (active) {
#pragma unroll
            for (int pl = 0; pl < C::kPlanes; ++pl) {
              const uint8_t* base = smem + stage * C::kStageBytes + pl * C::kABytes + box * kBoxBytes + (fi & 7) * 2;
#pragma unroll 8
              for (int r = 0; r < kRowBlk; ++r) {
                const uint32_t v = *reinterpret_cast<const uint32_t*>(base + r * 128 + ((((fi >> 3) ^ (r & 7))) << 4));
                s0 += __uint_as_float(v << 16);
                s1 += __uint_as_float(v & 0xffff0000u);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (active) {
          if (n0 + f < n_real) atomicAdd(bias_out + n0 + f, s0);
          if (n0 + f + 1 < n_real) atomicAdd(bias_out + n0 + f + 1, s1);
        }
      }
      mbar_wait(&tfull[0], 0);
      tc_fence_after();
      const int nrow = n0 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < kx; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        if (nrow < n_real) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int j = c0 + e;
            if (j < k_real) atomicAdd(out + (size_t)nrow * out_ld + j, __uint_as_float(r[e]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}



// ---------------------------------------------------------------------------------------------
// wgrad, bf16, full dY width in one pass: dW[0:n_real, 0:kx] += dY^T X with up to two 128-feature slabs
// of dY accumulating into separate TMEM buffers (X is read once), and the bias gradient (column sums
// of dY) computed by the otherwise idle epilogue warps straight from the staged dY tiles.
// ---------------------------------------------------------------------------------------------
// SPLIT = 3 (split-bf16): hi and lo planes of dY and X are staged (32-row blocks so that three stages still fit) and
// every K step issues hi*hi + lo*hi + hi*lo; the bias gradient sums both planes.
template <int NSLAB, int SPLIT>
__global__ void __launch_bounds__(256, 1)
wgrad2_tc_kernel(const __grid_constant__ WgMaps maps, int64_t m, int64_t rows_per_cta, int n_real, int kx, int k_real,
                 int stages, float* __restrict__ out, int out_ld, float* __restrict__ bias_out, int f16,
                 float* __restrict__ partial, int coalesced) {
  constexpr int kPlanes = SPLIT == 3 ? 2 : 1;
  constexpr int kRowBlk = SPLIT == 3 ? 32 : 64;
  constexpr int kBoxBytes = kRowBlk * 128;
  constexpr int kDyPlane = 2 * NSLAB * kBoxBytes;     // one plane of the dY tile
  constexpr int kDyBytes = kPlanes * kDyPlane;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int xboxes = kx / 64;
  const int x_plane = xboxes * kBoxBytes;
  const int stage_bytes = kDyBytes + kPlanes * x_plane;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + 8;
  uint64_t* tfull = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = min(m, r_begin + rows_per_cta);
  const int nblk = r_end > r_begin ? (int)((r_end - r_begin + kRowBlk - 1) / kRowBlk) : 0;
  const bool do_bias = bias_out != nullptr;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.dy_hi);
    tma_prefetch_desc(&maps.x_hi);
    if (SPLIT == 3) {
      tma_prefetch_desc(&maps.dy_lo);
      tma_prefetch_desc(&maps.x_lo);
    }
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], do_bias ? 5 : 1);
    }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, NSLAB * 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (nblk > 0) {
    if (warp == 0) {
      // whole warp runs the loop, one elected lane issues (uniform registers for the TMA operands)
      const uint32_t stage_tx = (uint32_t)stage_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int b = 0; b < nblk; ++b) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sbase = smem + stage * stage_bytes;
        const int r0 = (int)(r_begin + (int64_t)b * kRowBlk);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full[stage], stage_tx);
          // layout per stage: dY_hi | dY_lo | X_hi | X_lo
          for (int i = 0; i < 2 * NSLAB; ++i) tma_load_2d(sbase + i * kBoxBytes, &maps.dy_hi, &full[stage], i * 64, r0);
          for (int i = 0; i < xboxes; ++i) tma_load_2d(sbase + kDyBytes + i * kBoxBytes, &maps.x_hi, &full[stage], i * 64, r0);
          if (SPLIT == 3) {
            for (int i = 0; i < 2 * NSLAB; ++i)
              tma_load_2d(sbase + kDyPlane + i * kBoxBytes, &maps.dy_lo, &full[stage], i * 64, r0);
            for (int i = 0; i < xboxes; ++i)
              tma_load_2d(sbase + kDyBytes + x_plane + i * kBoxBytes, &maps.x_lo, &full[stage], i * 64, r0);
          }
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc(kx, 1, 1, f16, f16);   // A = dY^T, B = X: both bf16 or both fp16
      int stage = 0;
      uint32_t phase = 0;
      for (int b = 0; b < nblk; ++b) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint32_t sb = sa + kDyBytes;
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kRowBlk / kUmmaK; ++kk) {
            const uint32_t koff = kk * kUmmaK * 128;
            const uint64_t db = make_desc(sb + koff, kBoxBytes, 1024);
#pragma unroll
            for (int sl = 0; sl < NSLAB; ++sl) {
              const uint64_t da = make_desc(sa + sl * 2 * kBoxBytes + koff, kBoxBytes, 1024);
              umma_bf16(tmem_base + sl * 256, da, db, idesc, (b | kk) ? 1u : 0u);
              if (SPLIT == 3) {
                umma_bf16(tmem_base + sl * 256, make_desc(sa + kDyPlane + sl * 2 * kBoxBytes + koff, kBoxBytes, 1024), db, idesc, 1u);
                umma_bf16(tmem_base + sl * 256, da, make_desc(sb + x_plane + koff, kBoxBytes, 1024), idesc, 1u);
              }
            }
          }
          umma_commit(&empty[stage]);
          if (b == nblk - 1) umma_commit(&tfull[0]);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    } else if (warp >= 4) {
      const int q = warp - 4;
      if (do_bias) {
        // column sums of dY from the staged (128B-swizzled) tiles: thread t owns features 2t, 2t+1
        const int t = threadIdx.x - 128;
        const bool active = t < NSLAB * 64;
        const int f = 2 * t;
        const int box = f >> 6, fi = f & 63;
        float s0 = 0.f, s1 = 0.f;
        uint32_t sat = 0u;   // fp16: some |dY| reached the largest finite value (0x7bff), i.e. the conversion saturated
        int stage = 0;
        uint32_t phase = 0;
        for (int b = 0; b < nblk; ++b) {
          mbar_wait(&full[stage], phase);
          if (active) {
#pragma unroll
            for (int pl = 0; pl < kPlanes; ++pl) {
              const uint8_t* base = smem + stage * stage_bytes + pl * kDyPlane + box * kBoxBytes + (fi & 7) * 2;
#pragma unroll 8
              for (int r = 0; r < kRowBlk; ++r) {
                const uint32_t v = *reinterpret_cast<const uint32_t*>(base + r * 128 + ((((fi >> 3) ^ (r & 7))) << 4));
                if (f16) {
                  const float2 f2 = unpack_f16x2(v);
                  s0 += f2.x;
                  s1 += f2.y;
                  const uint32_t a = v & 0x7fff7fffu;
                  sat |= (uint32_t)((a & 0xffffu) == 0x7bffu) | (uint32_t)((a >> 16) == 0x7bffu);
                } else {
                  s0 += __uint_as_float(v << 16);
                  s1 += __uint_as_float(v & 0xffff0000u);
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        if (f16 && sat) atomicAdd(&g_fp16_saturated, 1ull);
        if (active) {
          if (partial) {   // deterministic mode: this CTA's bias partial sums, reduced in CTA order afterwards
            float* pb = partial + (size_t)blockIdx.x * (256 * 257) + 256 * 256;
            pb[f] = s0;
            pb[f + 1] = s1;
          } else {
            if (f < n_real) atomicAdd(bias_out + f, s0);
            if (f + 1 < n_real) atomicAdd(bias_out + f + 1, s1);
          }
        }
      }
      mbar_wait(&tfull[0], 0);
      tc_fence_after();
      if (!partial && coalesced) {
        // The accumulator leaves TMEM with one dW row per lane; adding it to global memory from there touches 32 different
        // lines per warp instruction (65 536 scalar reductions per CTA, all CTAs at once at the end of the launch: ~12 % of
        // the kernel).  Transpose through the (now idle) ring instead -- 32 rows x kx floats per warp, 16-byte groups
        // XOR-swizzled by the row so that both the row-per-lane writes and the row-contiguous reads are conflict-free --
        // and add whole rows with red.v4: 512 contiguous bytes per warp instruction.
        asm volatile("bar.sync 1, 128;" ::: "memory");   // every epilogue warp is done with the staged dY tiles (bias sums)
        const uint32_t stg = smem_u32(smem) + (uint32_t)q * (uint32_t)(32 * kx * 4);
        const uint32_t row_pitch = (uint32_t)kx * 4u;
#pragma unroll 1
        for (int sl = 0; sl < NSLAB; ++sl) {
          const uint32_t taddr = tmem_base + sl * 256 + ((uint32_t)(q * 32) << 16);
          for (int c0 = 0; c0 < kx; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + (uint32_t)c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int e4 = 0; e4 < 8; ++e4) {
              const uint32_t g = (uint32_t)(c0 >> 2) + e4;
              sts128(stg + lane * row_pitch + (((g & ~7u) | ((g ^ (uint32_t)lane) & 7u)) << 4), r[4 * e4], r[4 * e4 + 1], r[4 * e4 + 2],
                     r[4 * e4 + 3]);
            }
          }
          __syncwarp();
          for (int rr = 0; rr < 32; ++rr) {
            const int nrow = sl * 128 + q * 32 + rr;
            if (nrow >= n_real) break;
            float* orow = out + (size_t)nrow * out_ld;
            for (int g = lane; 4 * g < k_real; g += 32) {
              const float4 v = lds128f(stg + rr * row_pitch + ((((uint32_t)g & ~7u) | (((uint32_t)g ^ (uint32_t)rr) & 7u)) << 4));
              if (4 * g + 3 < k_real) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + 4 * g), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                             : "memory");
              } else {
                atomicAdd(orow + 4 * g, v.x);
                if (4 * g + 1 < k_real) atomicAdd(orow + 4 * g + 1, v.y);
                if (4 * g + 2 < k_real) atomicAdd(orow + 4 * g + 2, v.z);
              }
            }
          }
          __syncwarp();
        }
      } else
#pragma unroll
      for (int sl = 0; sl < NSLAB; ++sl) {
        const int nrow = sl * 128 + q * 32 + lane;
        const uint32_t taddr = tmem_base + sl * 256 + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < kx; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + (uint32_t)c0, r);
          tmem_ld_wait();
          if (partial) {
            float* ps = partial + (size_t)blockIdx.x * (256 * 257) + (size_t)nrow * 256 + c0;
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              *reinterpret_cast<float4*>(ps + e) = make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]),
                                                               __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
          } else if (nrow < n_real) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int j = c0 + e;
              if (j < k_real) atomicAdd(out + (size_t)nrow * out_ld + j, __uint_as_float(r[e]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, NSLAB * 256);
}

// deterministic mode, second pass: out[n, j] += sum over CTAs (in CTA order) of their partial tiles; same for the bias
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int nctas, int n_real, int n_slab_rows, int k_real, float* __restrict__ out,
                    int out_ld, float* __restrict__ bias_out) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int total = n_real * k_real;
  if (idx < total) {
    const int n = idx / k_real, j = idx - n * k_real;
    float acc = 0.f;
    for (int b = 0; b < nctas; ++b) acc += partial[(size_t)b * (256 * 257) + (size_t)n * 256 + j];
    out[(size_t)n * out_ld + j] += acc;
  } else if (bias_out && idx - total < n_real && idx - total < n_slab_rows) {
    const int f = idx - total;
    float acc = 0.f;
    for (int b = 0; b < nctas; ++b) acc += partial[(size_t)b * (256 * 257) + 256 * 256 + f];
    bias_out[f] += acc;
  }
}

}  // namespace

}  // namespace rn

extern "C" int64_t rn_fp16_saturation_count(int reset) {
  unsigned long long v = 0ull;
  if (cudaMemcpyFromSymbol(&v, rn::g_fp16_saturated, sizeof(v)) != cudaSuccess) return -1;   // (synchronises the device)
  if (reset && v) {
    const unsigned long long z = 0ull;
    cudaMemcpyToSymbol(rn::g_fp16_saturated, &z, sizeof(z));
  }
  return (int64_t)v;
}

namespace rn {

int launch_gemm_tc(const GemmArgs& g, cudaStream_t st) {
  if (g.m <= 0) return RN_OK;
  if (g.prec != RN_PREC_BF16 && g.prec != RN_PREC_BF16X3) return rn_set_error(RN_ERR_ARG, "gemm_tc: bf16 modes only");
  const bool x3 = g.prec == RN_PREC_BF16X3;
  TcMaps maps;
  int rc;
  if ((rc = make_map(&maps.a1_hi, g.a1.hi, g.m, g.a1_valid, g.a1.ld, kBM))) return rc;
  if ((rc = make_map(&maps.a1_lo, x3 ? g.a1.lo : nullptr, g.m, g.a1_valid, g.a1.ld, kBM))) return rc;
  if ((rc = make_map(&maps.a2_hi, g.k2 ? g.a2.hi : nullptr, g.m, g.a2_valid, g.a2.ld, kBM))) return rc;
  if ((rc = make_map(&maps.a2_lo, (x3 && g.k2) ? g.a2.lo : nullptr, g.m, g.a2_valid, g.a2.ld, kBM))) return rc;
  if ((rc = make_map(&maps.b_hi, g.b_hi, g.n, g.k1 + g.k2, g.b_ld, g.n))) return rc;
  if ((rc = make_map(&maps.b_lo, x3 ? g.b_lo : nullptr, g.n, g.k1 + g.k2, g.b_ld, g.n))) return rc;
  const int64_t tiles = (g.m + kBM - 1) / kBM;
  const unsigned grid = (unsigned)(tiles < num_sms() ? tiles : num_sms());
  rn_prof_begin(RN_PROF_GEMM_TC, st, g.algo_flops, (x3 ? 3.0 : 1.0) * 2.0 * (double)g.m * g.n * (g.k1 + g.k2));
  if (x3) {
    static bool once = false;
    if (!once) { if ((rc = set_smem(gemm_tc_kernel<3>, Cfg<3>::kSmemBytes))) return rc; once = true; }
    gemm_tc_kernel<3><<<grid, 256, Cfg<3>::kSmemBytes, st>>>(maps, g.m, g.n, g.k1 / kBK, g.k2 / kBK, g.epi);
  } else {
    static bool once = false;
    if (!once) { if ((rc = set_smem(gemm_tc_kernel<1>, Cfg<1>::kSmemBytes))) return rc; once = true; }
    gemm_tc_kernel<1><<<grid, 256, Cfg<1>::kSmemBytes, st>>>(maps, g.m, g.n, g.k1 / kBK, g.k2 / kBK, g.epi);
  }
  rn_prof_end(RN_PROF_GEMM_TC, st);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_wgrad2_tc(const WgradArgs& g, cudaStream_t st) {
  if (g.x_f16 != g.dy_f16) return rn_set_error(RN_ERR_UNSUPPORTED, "wgrad2: dY and X must share one 16-bit format");
  const bool x3 = g.prec == RN_PREC_BF16X3;
  if (x3 && g.x_f16) return rn_set_error(RN_ERR_UNSUPPORTED, "wgrad2: the split mode is bf16");
  const int row_blk = x3 ? 32 : 64;
  WgMaps maps;
  int rc;
  if ((rc = make_map(&maps.dy_hi, g.dy.hi, g.m, g.dy_valid, g.dy.ld, row_blk))) return rc;
  if ((rc = make_map(&maps.x_hi, g.x.hi, g.m, g.x_valid, g.x.ld, row_blk))) return rc;
  if ((rc = make_map(&maps.dy_lo, x3 ? g.dy.lo : nullptr, g.m, g.dy_valid, g.dy.ld, row_blk))) return rc;
  if ((rc = make_map(&maps.x_lo, x3 ? g.x.lo : nullptr, g.m, g.x_valid, g.x.ld, row_blk))) return rc;
  const int nslab = g.n_real > 128 ? 2 : 1;
  const int box_bytes = row_blk * 128;
  const int stage_bytes = (x3 ? 2 : 1) * (2 * nslab + g.kx / 64) * box_bytes;
  int stages = (232448 - 1024 - 512) / stage_bytes;
  if (stages > 6) stages = 6;
  const int smem = stages * stage_bytes + 1024 + 512;
  int ctas = num_sms();
  int64_t blocks = (g.m + row_blk - 1) / row_blk;
  if (ctas > blocks) ctas = (int)blocks;
  int64_t rows_per = ((blocks + ctas - 1) / ctas) * row_blk;
  const unsigned grid = (unsigned)((g.m + rows_per - 1) / rows_per);
  static bool once = false;
  if (!once) {
    if ((rc = set_smem(wgrad2_tc_kernel<1, 1>, 232448))) return rc;
    if ((rc = set_smem(wgrad2_tc_kernel<2, 1>, 232448))) return rc;
    if ((rc = set_smem(wgrad2_tc_kernel<1, 3>, 232448))) return rc;
    if ((rc = set_smem(wgrad2_tc_kernel<2, 3>, 232448))) return rc;
    once = true;
  }
  const int f16 = (g.x_f16 && g.dy_f16) ? 1 : 0;
  rn_prof_begin(RN_PROF_WGRAD_TC, st, g.algo_flops, (x3 ? 3.0 : 1.0) * 2.0 * (double)g.m * (128.0 * nslab) * g.kx);
  // row-coalesced red.v4 epilogue: needs 16-byte aligned dW rows and 128 x kx floats of staging in the ring
  static const bool scalar_red = getenv("RN_WGRAD_SCALAR_RED") != nullptr;   // (A/B switch for the profiles)
  const int coalesced = (!scalar_red && (reinterpret_cast<uintptr_t>(g.out) & 15u) == 0 && (g.out_ld & 3) == 0 &&
                         (size_t)128 * g.kx * 4 <= (size_t)stages * stage_bytes) ? 1 : 0;
#define RN_WG2(NS, SP) wgrad2_tc_kernel<NS, SP><<<grid, 256, smem, st>>>(maps, g.m, rows_per, g.n_real, g.kx, g.k_real, stages, g.out, g.out_ld, g.bias_out, f16, g.partial, coalesced)
  if (x3) {
    if (nslab == 2) RN_WG2(2, 3); else RN_WG2(1, 3);
  } else {
    if (nslab == 2) RN_WG2(2, 1); else RN_WG2(1, 1);
  }
#undef RN_WG2
  rn_prof_end(RN_PROF_WGRAD_TC, st);
  RN_CUDA_CHECK_LAUNCH();
  if (g.partial) {
    const int total = g.n_real * g.k_real + (g.bias_out ? g.n_real : 0);
    wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(g.partial, (int)grid, g.n_real, 128 * nslab, g.k_real, g.out, g.out_ld, g.bias_out);
    RN_CUDA_CHECK_LAUNCH();
  }
  return RN_OK;
}

int launch_wgrad_tc(const WgradArgs& g, cudaStream_t st) {
  if (g.m <= 0) return RN_OK;
  if (g.prec == RN_PREC_FP16 && !g.all_slabs) return rn_set_error(RN_ERR_UNSUPPORTED, "wgrad_tc: the fp16 mode needs all_slabs");
  if (g.prec != RN_PREC_BF16 && g.prec != RN_PREC_BF16X3 && g.prec != RN_PREC_FP16) return rn_set_error(RN_ERR_ARG, "wgrad_tc: bf16 modes only");
  if (g.all_slabs) {
    if (g.n0 != 0 || g.n_real > 256) return rn_set_error(RN_ERR_ARG, "wgrad_tc: all_slabs needs n0 = 0, n_real <= 256");
    return launch_wgrad2_tc(g, st);
  }
  const bool x3 = g.prec == RN_PREC_BF16X3;
  WgMaps maps;
  int rc;
  if ((rc = make_map(&maps.dy_hi, g.dy.hi, g.m, g.dy_valid, g.dy.ld, 64))) return rc;
  if ((rc = make_map(&maps.dy_lo, x3 ? g.dy.lo : nullptr, g.m, g.dy_valid, g.dy.ld, 64))) return rc;
  if ((rc = make_map(&maps.x_hi, g.x.hi, g.m, g.x_valid, g.x.ld, 64))) return rc;
  if ((rc = make_map(&maps.x_lo, x3 ? g.x.lo : nullptr, g.m, g.x_valid, g.x.ld, 64))) return rc;
  int ctas = num_sms();
  int64_t blocks64 = (g.m + 63) / 64;
  if (ctas > blocks64) ctas = (int)blocks64;
  int64_t rows_per = ((blocks64 + ctas - 1) / ctas) * 64;
  const unsigned grid = (unsigned)((g.m + rows_per - 1) / rows_per);
  rn_prof_begin(RN_PROF_WGRAD_TC, st, g.algo_flops, (x3 ? 3.0 : 1.0) * 2.0 * (double)g.m * 128.0 * g.kx);
  if (x3) {
    static bool once = false;
    if (!once) { if ((rc = set_smem(wgrad_tc_kernel<3>, Cfg<3>::kSmemBytes))) return rc; once = true; }
    wgrad_tc_kernel<3><<<grid, 256, Cfg<3>::kSmemBytes, st>>>(maps, g.m, rows_per, g.n0, g.n_real, g.kx, g.k_real, g.out, g.out_ld, g.bias_out);
  } else {
    static bool once = false;
    if (!once) { if ((rc = set_smem(wgrad_tc_kernel<1>, Cfg<1>::kSmemBytes))) return rc; once = true; }
    wgrad_tc_kernel<1><<<grid, 256, Cfg<1>::kSmemBytes, st>>>(maps, g.m, rows_per, g.n0, g.n_real, g.kx, g.k_real, g.out, g.out_ld, g.bias_out);
  }
  rn_prof_end(RN_PROF_WGRAD_TC, st);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

}  // namespace rn
