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
#include "tc_common.cuh"

namespace rn {
namespace {
using namespace tc;

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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    const uint32_t stage_tx = (uint32_t)C::kPlanes * (uint32_t)(C::kABytes + n * kBK * 2);
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (int)(tile * kBM);
      for (int kb = 0; kb < kbt; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sbase = smem + stage * C::kStageBytes;
        mbar_arrive_expect_tx(&full[stage], stage_tx);
        const bool first = kb < kb1;
        const int ka = (first ? kb : kb - kb1) * kBK;
        tma_load_2d(sbase, first ? &maps.a1_hi : &maps.a2_hi, &full[stage], ka, m0);
        tma_load_2d(sbase + C::kPlanes * C::kABytes, &maps.b_hi, &full[stage], kb * kBK, 0);
        if (SPLIT == 3) {
          tma_load_2d(sbase + C::kABytes, first ? &maps.a1_lo : &maps.a2_lo, &full[stage], ka, m0);
          tma_load_2d(sbase + 2 * C::kABytes + C::kBBytes, &maps.b_lo, &full[stage], kb * kBK, 0);
        }
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
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
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull[buf]);      // accumulator complete
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
                int k_real, float* __restrict__ out, int out_ld) {
  using C = Cfg<SPLIT>;
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
      mbar_init(&empty[i], 1);
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
template <int NSLAB>
__global__ void __launch_bounds__(256, 1)
wgrad2_tc_kernel(const __grid_constant__ WgMaps maps, int64_t m, int64_t rows_per_cta, int n_real, int kx, int k_real,
                 int stages, float* __restrict__ out, int out_ld, float* __restrict__ bias_out) {
  constexpr int kRowBlk = 64;
  constexpr int kBoxBytes = kRowBlk * 128;
  constexpr int kDyBytes = 2 * NSLAB * kBoxBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int xboxes = kx / 64;
  const int stage_bytes = kDyBytes + xboxes * kBoxBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + 8;
  uint64_t* tfull = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = min(m, r_begin + rows_per_cta);
  const int nblk = r_end > r_begin ? (int)((r_end - r_begin + kRowBlk - 1) / kRowBlk) : 0;
  const bool do_bias = bias_out != nullptr;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.dy_hi);
    tma_prefetch_desc(&maps.x_hi);
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
    if (warp == 0 && lane == 0) {
      const uint32_t stage_tx = (uint32_t)stage_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int b = 0; b < nblk; ++b) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sbase = smem + stage * stage_bytes;
        mbar_arrive_expect_tx(&full[stage], stage_tx);
        const int r0 = (int)(r_begin + (int64_t)b * kRowBlk);
        for (int i = 0; i < 2 * NSLAB; ++i) tma_load_2d(sbase + i * kBoxBytes, &maps.dy_hi, &full[stage], i * 64, r0);
        for (int i = 0; i < xboxes; ++i) tma_load_2d(sbase + kDyBytes + i * kBoxBytes, &maps.x_hi, &full[stage], i * 64, r0);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1 && lane == 0) {
      const uint32_t idesc = make_idesc(kx, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int b = 0; b < nblk; ++b) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint32_t sb = sa + kDyBytes;
#pragma unroll
        for (int kk = 0; kk < kRowBlk / kUmmaK; ++kk) {
          const uint32_t koff = kk * kUmmaK * 128;
          const uint64_t db = make_desc(sb + koff, kBoxBytes, 1024);
#pragma unroll
          for (int sl = 0; sl < NSLAB; ++sl)
            umma_bf16(tmem_base + sl * 256, make_desc(sa + sl * 2 * kBoxBytes + koff, kBoxBytes, 1024), db, idesc,
                      (b | kk) ? 1u : 0u);
        }
        umma_commit(&empty[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull[0]);
    } else if (warp >= 4) {
      const int q = warp - 4;
      if (do_bias) {
        // column sums of dY from the staged (128B-swizzled) tiles: thread t owns features 2t, 2t+1
        const int t = threadIdx.x - 128;
        const bool active = t < NSLAB * 64;
        const int f = 2 * t;
        const int box = f >> 6, fi = f & 63;
        float s0 = 0.f, s1 = 0.f;
        int stage = 0;
        uint32_t phase = 0;
        for (int b = 0; b < nblk; ++b) {
          mbar_wait(&full[stage], phase);
          if (active) {
            const uint8_t* base = smem + stage * stage_bytes + box * kBoxBytes + (fi & 7) * 2;
#pragma unroll 8
            for (int r = 0; r < kRowBlk; ++r) {
              const uint32_t v = *reinterpret_cast<const uint32_t*>(base + r * 128 + ((((fi >> 3) ^ (r & 7))) << 4));
              s0 += __uint_as_float(v << 16);
              s1 += __uint_as_float(v & 0xffff0000u);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        if (active) {
          if (f < n_real) atomicAdd(bias_out + f, s0);
          if (f + 1 < n_real) atomicAdd(bias_out + f + 1, s1);
        }
      }
      mbar_wait(&tfull[0], 0);
      tc_fence_after();
#pragma unroll
      for (int sl = 0; sl < NSLAB; ++sl) {
        const int nrow = sl * 128 + q * 32 + lane;
        const uint32_t taddr = tmem_base + sl * 256 + ((uint32_t)(q * 32) << 16);
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, NSLAB * 256);
}

// ---------------------------------------------------------------------------------------------
// Fused forward chain (bf16): up to 9 GEMM layers over one 128-row tile with the activations resident
// in shared memory.  Hidden layers: TMEM accumulator -> bias + ReLU -> bf16 -> written back (128B
// swizzle, K-major) as the next layer's A operand, 64 columns (= one K block) at a time so the next
// layer's MMAs start as soon as its first K block exists.  Weights stream through a TMA ring; the
// input tile (x0 / v0) stays resident for layer 0 and the skip layer.  Only the last layer (heads /
// rgb head) and, in training, the optional activation saves touch global memory.
// ---------------------------------------------------------------------------------------------
constexpr int kChainMaxOps = 12;
struct ChainOp {
  int n;           // MMA N = output columns (multiple of 16)
  int kb_act;      // K blocks read from the resident activation tile (0 or 4)
  int kb_in;       // K blocks read from the resident input tile
  int last_in_use; // this op is the last reader of the input tile
  int kind;        // 0: hidden (result -> next activation tile in smem [+ global save]); 1: global epilogue
  int mode;        // hidden transform: 0 = bias + ReLU (forward), 1 = ReLU mask from a saved activation (backward)
  int gepi;        // kind 1: which global epilogue
  const float* bias;       // mode 0
  const uint16_t* mask;    // mode 1: bf16 [m,256]; positive entries pass the gradient
  void* save_hi;           // optional bf16 [m,256] copy of the hidden result
};
struct ChainParams {
  int num_ops;
  int in_kb;       // K blocks of the input tile (1..4)
  int stages;
  int64_t m;
  ChainOp op[kChainMaxOps];
  GemmEpilogue gepi[2];
};
struct ChainMaps {
  CUtensorMap in;
  CUtensorMap w[kChainMaxOps];
};

__global__ void __launch_bounds__(384, 1)
chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
  constexpr int kStageBytes = 256 * kBK * 2;  // 32 KB weight K block
  constexpr int kBlkBytes = kBM * kBK * 2;    // 16 KB activation K block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_act = smem;                               // 4 K blocks
  uint8_t* s_in = s_act + 4 * kBlkBytes;               // in_kb K blocks
  uint8_t* s_w = s_in + p.in_kb * kBlkBytes;           // weight ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + p.stages * kStageBytes);
  uint64_t* w_full = bars;            // [stages]
  uint64_t* w_empty = bars + 8;       // [stages]
  uint64_t* a_ready = bars + 16;      // [4]
  uint64_t* tfull = bars + 20;        // [2]
  uint64_t* tempty = bars + 22;       // [2]
  uint64_t* in_full = bars + 24;
  uint64_t* in_empty = bars + 25;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  float* bias_s = reinterpret_cast<float*>(bars + 32);  // [256] bias of the op being drained

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t num_tiles = (p.m + kBM - 1) / kBM;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.in);
    for (int i = 0; i < p.num_ops; ++i) tma_prefetch_desc(&maps.w[i]);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&a_ready[i], 8);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    mbar_init(in_full, 1);
    mbar_init(in_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: input tile, then the weight K blocks of every op in order =====
    int stage = 0;
    uint32_t phase = 0, in_phase = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (int)(tile * kBM);
      mbar_wait(in_empty, in_phase ^ 1);
      mbar_arrive_expect_tx(in_full, (uint32_t)(p.in_kb * kBlkBytes));
      for (int kb = 0; kb < p.in_kb; ++kb) tma_load_2d(s_in + kb * kBlkBytes, &maps.in, in_full, kb * kBK, m0);
      in_phase ^= 1;
      for (int l = 0; l < p.num_ops; ++l) {
        const int nkb = p.op[l].kb_act + p.op[l].kb_in;
        const uint32_t tx = (uint32_t)p.op[l].n * kBK * 2;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&w_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&w_full[stage], tx);
          tma_load_2d(s_w + stage * kStageBytes, &maps.w[l], &w_full[stage], kb * kBK, 0);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    int stage = 0;
    uint32_t phase = 0, in_phase = 0;
    uint32_t hidden_done = 0;  // hidden ops issued so far == activation generations requested
    int buf = 0;
    uint32_t tphase = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      bool in_ready = false;
      for (int l = 0; l < p.num_ops; ++l) {
        const ChainOp& L = p.op[l];
        const uint32_t idesc = make_idesc(L.n, 0, 0);
        mbar_wait(&tempty[buf], tphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)buf * 256u;
        const int nkb = L.kb_act + L.kb_in;
        for (int kb = 0; kb < nkb; ++kb) {
          uint32_t sa;
          if (kb < L.kb_act) {
            mbar_wait(&a_ready[kb], (hidden_done - 1u) & 1u);   // generation produced by the latest hidden op
            sa = smem_u32(s_act + kb * kBlkBytes);
          } else {
            if (!in_ready) { mbar_wait(in_full, in_phase); in_ready = true; }
            sa = smem_u32(s_in + (kb - L.kb_act) * kBlkBytes);
          }
          mbar_wait(&w_full[stage], phase);
          tc_fence_after();
          const uint32_t sb = smem_u32(s_w + stage * kStageBytes);
#pragma unroll
          for (int kk = 0; kk < kBK / kUmmaK; ++kk) {
            const uint32_t koff = kk * kUmmaK * 2;
            umma_bf16(tmem_d, make_desc(sa + koff, 16, 1024), make_desc(sb + koff, 16, 1024), idesc, (kb | kk) ? 1u : 0u);
          }
          umma_commit(&w_empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[buf]);
        if (L.last_in_use) umma_commit(in_empty);
        if (L.kind == 0) ++hidden_done;
        if (++buf == 2) { buf = 0; tphase ^= 1; }
      }
      in_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ===== epilogue warps: 8 warps = 4 TMEM lane quadrants x 2 column halves of every 64-column chunk =====
    const int q = (warp - 4) & 3;
    const int half = (warp - 4) >> 2;
    int buf = 0;
    uint32_t tphase = 0;
    const uint32_t s_act_u32 = smem_u32(s_act);
    const uint32_t bias_u32 = smem_u32(bias_s);
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int r_in_tile = q * 32 + lane;
      const int64_t row = tile * kBM + r_in_tile;
      const bool row_ok = row < p.m;
      for (int l = 0; l < p.num_ops; ++l) {
        const ChainOp& L = p.op[l];
        const bool hidden = L.kind == 0;
        if (hidden && L.mode == 0) {
          // stage this op's bias in shared memory (the 8 epilogue warps only: named barrier 1)
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const int t = threadIdx.x - 128;
          if (t < 128) {
            const float2 bv = __ldg(reinterpret_cast<const float2*>(L.bias) + t);
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(bias_u32 + 8u * t), "f"(bv.x), "f"(bv.y) : "memory");
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        mbar_wait(&tfull[buf], tphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)buf * 256u + ((uint32_t)(q * 32) << 16);
        if (hidden) {
          // 4 chunks of 64 columns (this warp: 32 of them) -> swizzled K blocks of the activation tile
          for (int j = 0; j < 4; ++j) {
            const int col0 = j * 64 + half * 32;
            uint32_t r[32];
            uint32_t mk[16];
            tmem_ld32(taddr + (uint32_t)col0, r);
            if (L.mode == 1 && row_ok) {
              const uint16_t* mp = L.mask + (size_t)row * 256 + col0;
              ldg256(mp, mk);
              ldg256(mp + 16, mk + 8);
            }
            tmem_ld_wait();
            uint32_t packed[16];
            if (L.mode == 0) {
              const uint32_t baddr = bias_u32 + (uint32_t)col0 * 4u;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 b = lds128f(baddr + 16u * i);
                const float v0 = fmaxf(__uint_as_float(r[4 * i]) + b.x, 0.f), v1 = fmaxf(__uint_as_float(r[4 * i + 1]) + b.y, 0.f);
                const float v2 = fmaxf(__uint_as_float(r[4 * i + 2]) + b.z, 0.f), v3 = fmaxf(__uint_as_float(r[4 * i + 3]) + b.w, 0.f);
                packed[2 * i] = pack_bf16x2(v0, v1);
                packed[2 * i + 1] = pack_bf16x2(v2, v3);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const uint32_t mlo = mk[i] & 0xffffu, mhi = mk[i] >> 16;
                const bool plo = row_ok && mlo != 0u && mlo < 0x8000u, phi = row_ok && mhi != 0u && mhi < 0x8000u;
                const float v0 = plo ? __uint_as_float(r[2 * i]) : 0.f, v1 = phi ? __uint_as_float(r[2 * i + 1]) : 0.f;
                packed[i] = pack_bf16x2(v0, v1);
              }
            }
            const uint32_t blk = s_act_u32 + (uint32_t)(j * kBlkBytes + r_in_tile * 128);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t pos = (uint32_t)((half * 4 + c) ^ (r_in_tile & 7));
              sts128(blk + pos * 16u, packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
            }
            fence_proxy_async();   // make the generic-proxy smem writes visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[j]);
            if (L.save_hi && row_ok) {
              uint16_t* g = reinterpret_cast<uint16_t*>(L.save_hi) + (size_t)row * 256 + col0;
              stg256(g, packed);
              stg256(g + 16, packed + 8);
            }
          }
        } else {
          const GemmEpilogue& ge = p.gepi[L.gepi];
          for (int c0 = half * 32; c0 < L.n; c0 += 64) {
            uint32_t r[32];
            tmem_ld32(taddr + (uint32_t)c0, r);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (c0 + 16 * h < L.n) {
                  float v[16];
#pragma unroll
                  for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[16 * h + e]);
                  epi_store16<RN_PREC_BF16>(ge, (size_t)row, c0 + 16 * h, v, nullptr);
                }
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launches
// ---------------------------------------------------------------------------------------------
}  // namespace

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
  rn_prof_begin(RN_PROF_GEMM_TC, st, g.algo_flops);
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

int launch_chain_single(const ChainArgs& a, cudaStream_t st) {
  if (a.m <= 0) return RN_OK;
  if (a.num_ops < 1 || a.num_ops > kChainMaxOps) return rn_set_error(RN_ERR_ARG, "chain: 1..12 ops");
  if (a.in_cols % 64 || a.in_cols < 64 || a.in_cols > 256) return rn_set_error(RN_ERR_ARG, "chain: input tile must be 64..256 columns");
  ChainMaps maps;
  ChainParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = make_map(&maps.in, a.in.hi, a.m, a.in_valid, a.in.ld, kBM))) return rc;
  p.num_ops = a.num_ops;
  p.in_kb = a.in_cols / kBK;
  p.m = a.m;
  int last_in = 0;
  for (int l = 0; l < a.num_ops; ++l)
    if (a.op[l].kb_in) last_in = l;
  for (int l = 0; l < kChainMaxOps; ++l) {
    if (l < a.num_ops) {
      const ChainOpArgs& L = a.op[l];
      if (L.n % 16 || L.n > 256 || (L.kind == 0 && L.n != 256)) return rn_set_error(RN_ERR_ARG, "chain: bad op width");
      if ((L.kb_act != 0 && L.kb_act != 4) || (L.kb_in != 0 && L.kb_in != p.in_kb) || L.kb_act + L.kb_in == 0 ||
          (l == 0 && L.kb_act != 0))
        return rn_set_error(RN_ERR_ARG, "chain: bad K structure");
      const int ktot = (L.kb_act + L.kb_in) * kBK;
      if ((rc = make_map(&maps.w[l], L.w, L.n, ktot, L.w_ld, L.n))) return rc;
      ChainOp& o = p.op[l];
      o.n = L.n; o.kb_act = L.kb_act; o.kb_in = L.kb_in; o.last_in_use = (l == last_in) ? 1 : 0;
      o.kind = L.kind; o.mode = L.mode; o.gepi = L.gepi; o.bias = L.bias;
      o.mask = reinterpret_cast<const uint16_t*>(L.mask); o.save_hi = L.save_hi;
      if (L.kind == 0 && L.mode == 0 && !L.bias) return rn_set_error(RN_ERR_ARG, "chain: forward op without bias");
      if (L.kind == 0 && L.mode == 1 && !L.mask) return rn_set_error(RN_ERR_ARG, "chain: backward op without mask");
    } else {
      memset(&maps.w[l], 0, sizeof(CUtensorMap));
    }
  }
  p.gepi[0] = a.gepi[0];
  p.gepi[1] = a.gepi[1];
  const int fixed = 4 * kBM * kBK * 2 + p.in_kb * kBM * kBK * 2;
  p.stages = (232448 - 1024 - 1280 - fixed) / (256 * kBK * 2);
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) return rn_set_error(RN_ERR_ARG, "chain: not enough shared memory for the weight ring");
  const int smem = fixed + p.stages * 256 * kBK * 2 + 1024 + 1280;
  static bool smem_set = false;
  if (!smem_set) {
    if ((rc = set_smem(chain_kernel, 232448))) return rc;
    smem_set = true;
  }
  const int64_t tiles = (a.m + kBM - 1) / kBM;
  const unsigned grid = (unsigned)(tiles < num_sms() ? tiles : num_sms());
  rn_prof_begin(RN_PROF_CHAIN_TC, st, a.algo_flops);
  chain_kernel<<<grid, 384, smem, st>>>(maps, p);
  rn_prof_end(RN_PROF_CHAIN_TC, st);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_wgrad2_tc(const WgradArgs& g, cudaStream_t st) {
  WgMaps maps;
  int rc;
  if ((rc = make_map(&maps.dy_hi, g.dy.hi, g.m, g.dy_valid, g.dy.ld, 64))) return rc;
  if ((rc = make_map(&maps.x_hi, g.x.hi, g.m, g.x_valid, g.x.ld, 64))) return rc;
  memset(&maps.dy_lo, 0, sizeof(CUtensorMap));
  memset(&maps.x_lo, 0, sizeof(CUtensorMap));
  const int nslab = g.n_real > 128 ? 2 : 1;
  const int stage_bytes = (2 * nslab + g.kx / 64) * 8192;
  int stages = (232448 - 1024 - 512) / stage_bytes;
  if (stages > 6) stages = 6;
  const int smem = stages * stage_bytes + 1024 + 512;
  int ctas = num_sms();
  int64_t blocks64 = (g.m + 63) / 64;
  if (ctas > blocks64) ctas = (int)blocks64;
  int64_t rows_per = ((blocks64 + ctas - 1) / ctas) * 64;
  const unsigned grid = (unsigned)((g.m + rows_per - 1) / rows_per);
  static bool once = false;
  if (!once) {
    if ((rc = set_smem(wgrad2_tc_kernel<1>, 232448))) return rc;
    if ((rc = set_smem(wgrad2_tc_kernel<2>, 232448))) return rc;
    once = true;
  }
  rn_prof_begin(RN_PROF_WGRAD_TC, st, g.algo_flops);
  if (nslab == 2)
    wgrad2_tc_kernel<2><<<grid, 256, smem, st>>>(maps, g.m, rows_per, g.n_real, g.kx, g.k_real, stages, g.out, g.out_ld, g.bias_out);
  else
    wgrad2_tc_kernel<1><<<grid, 256, smem, st>>>(maps, g.m, rows_per, g.n_real, g.kx, g.k_real, stages, g.out, g.out_ld, g.bias_out);
  rn_prof_end(RN_PROF_WGRAD_TC, st);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

int launch_wgrad_tc(const WgradArgs& g, cudaStream_t st) {
  if (g.m <= 0) return RN_OK;
  if (g.prec != RN_PREC_BF16 && g.prec != RN_PREC_BF16X3) return rn_set_error(RN_ERR_ARG, "wgrad_tc: bf16 modes only");
  if (g.all_slabs) {
    if (g.prec != RN_PREC_BF16 || g.n0 != 0 || g.n_real > 256) return rn_set_error(RN_ERR_ARG, "wgrad_tc: all_slabs needs bf16, n0 = 0, n_real <= 256");
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
  rn_prof_begin(RN_PROF_WGRAD_TC, st, g.algo_flops);
  if (x3) {
    static bool once = false;
    if (!once) { if ((rc = set_smem(wgrad_tc_kernel<3>, Cfg<3>::kSmemBytes))) return rc; once = true; }
    wgrad_tc_kernel<3><<<grid, 256, Cfg<3>::kSmemBytes, st>>>(maps, g.m, rows_per, g.n0, g.n_real, g.kx, g.k_real, g.out, g.out_ld);
  } else {
    static bool once = false;
    if (!once) { if ((rc = set_smem(wgrad_tc_kernel<1>, Cfg<1>::kSmemBytes))) return rc; once = true; }
    wgrad_tc_kernel<1><<<grid, 256, Cfg<1>::kSmemBytes, st>>>(maps, g.m, rows_per, g.n0, g.n_real, g.kx, g.k_real, g.out, g.out_ld);
  }
  rn_prof_end(RN_PROF_WGRAD_TC, st);
  RN_CUDA_CHECK_LAUNCH();
  return RN_OK;
}

}  // namespace rn
