// Fused split-bf16 GEMM chain on CTA pairs (sm_100a, tcgen05.mma cta_group::2): the arithmetic that matches the
// reference's fp32 nn.Linear within BASELINE.json's tolerances (models.py:576-582, 689-700) at tensor-core speed.
//
// Every operand is a bf16 hi plane + a bf16 lo plane (x ~= hi + lo, 16-bit significand) and every K step issues
// three MMAs, hi*hi + lo*hi + hi*lo, into an fp32 TMEM accumulator.  Both planes of the running activation stay
// resident in shared memory (2 x 64 KB per CTA), so a cluster works on ONE 256-row tile (128 rows per CTA) and the
// overlap that chain_pair.cu gets from a second row tile comes from splitting every 256-wide op into two COLUMN
// HALVES instead (M = 256, N = 128 MMAs: 64 cycles each, measured with tools/micro/umma_contention.cu):
//
//   * TMEM holds two 256-column accumulators used by alternate ops.  The MMA warp issues all K blocks of output
//     columns 0..127 of op l, commits, then all K blocks of columns 128..255.  While the second half runs on the
//     tensor pipe the epilogue warps drain the first half (bias + ReLU / ReLU bit mask, hi/lo split) and write it back
//     as K blocks 0 and 1 of the next A operand; each finished 64-column K block is handed to the MMA warp through its
//     own mbarrier (act_ready[c]).  Op l+1 therefore starts on K blocks 0, 1 the moment op l's last MMA is issued and
//     reaches K blocks 2, 3 (1 536 cycles later) after the epilogue of the second half has written them.
//     (RN_X3_ORDER=1 selects the alternative issue order of res_order(): K blocks 0, 1 of both halves first.)
//   * the chain input of the first op is loaded by TMA straight into the (then dead) activation tile, K block by K
//     block, as soon as the previous tile's last MMA has retired, after an L2 prefetch issued one tile ahead; the skip
//     layer's input K blocks stream through the ring and feed both column halves while they are resident.
//   * weight items [64 rows per CTA x 64] x {hi, lo} (one K block of one column half) stream through a 6-stage TMA
//     ring (16 KB stages).
//   * saves for the wgrad kernels (hi plane only: wgrad runs on bf16 operands) and activation-format outputs of
//     global ops (the bottleneck -> v0 hi/lo) leave as TMA stores issued by a dedicated store warp straight from
//     the swizzled activation tile; ReLU masks travel as 1 bit per activation in a warp-coalesced layout.
//   * dgrad chains: the seed tile of the NEXT row tile is generated before the last op's accumulator is drained, so the
//     f32 stores of that epilogue overlap the first MMAs of the next tile.
//
// Roles per CTA (384 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator,
// warp 3 store warp, warps 4-11 epilogue (warp % 4 = TMEM lane quadrant, (warp - 4) / 4 = 32-column half of every
// 64-column K block).
#include <stdio.h>
#include <stdlib.h>

#include "chain_common.cuh"

namespace rn {
namespace {
using namespace tc;
using namespace chain;

constexpr int kStages = 6;
constexpr int kTraceSlots = 32;   // RN_CHAIN_TRACE: 64-bit time stamps per (op, tile) of CTA 0
constexpr int kItemPlane = 8192;               // [64 x 64] bf16: one plane of a weight item
constexpr int kStageBytes = 2 * kItemPlane;    // hi | lo  (an input item is one [128 x 64] plane = the whole stage)
constexpr int kPlaneBytes = 16384;             // [128 x 64] bf16, 128B swizzle: one K block of one activation plane
constexpr int kActPlane = 4 * kPlaneBytes;     // 64 KB: one plane of the activation tile (4 K blocks)
constexpr int kSmemRing = 2 * kActPlane;       // act_hi | act_lo | ring
constexpr int kSmemBars = kSmemRing + kStages * kStageBytes;
constexpr int kSmemBias = kSmemBars + 512;     // [op parity][column half h][4 K blocks x 32 floats]
constexpr int kSmemTotal = kSmemBias + 2048;
static_assert(kSmemTotal <= 232448, "shared memory budget");

// Issue order of the (column half, resident K block) pairs of an op.  split = 0: half 0 (K blocks 0..3), then half 1.
// split = 1: a 256-wide op over a full activation tile runs K blocks 0, 1 of BOTH column halves before K blocks 2, 3 --
// the previous op's second-half epilogue (which writes blocks 2 and 3 once that op's last MMA has retired) then has
// 3 072 / 3 840 cycles of tensor work to hide behind instead of 1 536 / 2 304, at the price of the first half completing
// only 1 536 cycles before the op does.  Per accumulator element the K order is the same (bit-identical results).
__device__ __forceinline__ void res_order(int split, int i, int kb_res, int nhalf, int& half, int& kb) {
  if (split && nhalf == 2 && kb_res == 4) {
    half = (i >> 1) & 1;
    kb = (i & 1) | ((i >> 2) << 1);
  } else {   // (no integer division on the MMA warp's critical path)
    half = i >= kb_res ? 1 : 0;
    kb = half ? i - kb_res : i;
  }
}

struct X3Maps {
  CUtensorMap in_hi, in_lo;       // chain input [m, in_cols], box [128 x 64]
  CUtensorMap gout_hi, gout_lo;   // activation-format output of a global op, box [128 x 64]
  CUtensorMap w_hi[kMaxOps], w_lo[kMaxOps];   // box [64 rows (hidden ops) or n/2 rows (global ops) x 64]
  CUtensorMap save[kMaxOps];      // hi-plane saves of hidden ops, box [128 x 64]
};

// MODE 0: forward chain (hidden ops: bias + ReLU); MODE 1: dgrad chain (hidden ops: ReLU bit mask; seed ops)
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
chain_x3_kernel(const __grid_constant__ X3Maps maps, const __grid_constant__ PairParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBars);
  uint64_t* ring_full = bars;          // [6] leader's are used (TMA bytes of both CTAs land there)
  uint64_t* ring_empty = bars + 6;     // [6] per CTA, multicast MMA commit
  uint64_t* acc_full = bars + 12;      // [2 accumulators][2 column halves] per CTA, multicast MMA commit
  uint64_t* acc_free = bars + 16;      // [2] leader's: 16 arrivals (8 epilogue warps x 2 CTAs)
  uint64_t* act_ready = bars + 18;     // [4] leader's: 16 arrivals: K block c of the activation tile is written
  // store events, one channel per pair of activation K blocks (0: blocks 0-1 = first column half, staged outputs;
  // 1: blocks 2-3): a half is saved as soon as it is written, so that its TMA store has the whole next op to drain
  uint64_t* written = bars + 32;       // [2] per CTA: 8 arrivals: the blocks a store needs are in shared memory
  uint64_t* drained = bars + 34;       // [2] per CTA: the store warp's TMA stores have read those blocks
  uint64_t* in_full = bars + 24;       // [4] leader's: K block kb of the first op's chain input has landed (tx bytes)
  uint64_t* saves_drained = bars + 28; // per CTA: the last save of the tile has been read (the next tile's input may land)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);
  uint64_t* blk_read = bars + 30;      // [2] per CTA, MMA commit: the second column half's MMAs have read activation K block
                                       //     0 / 1 of a hidden op, i.e. the first half's epilogue may overwrite it in place
  // the first op reads the chain input only: it is loaded straight into activation K blocks [in_blk0, 4)
  const bool direct_in = p.op[0].kind != 2 && p.op[0].kb_in > 0;
  const int in_blk0 = 4 - p.op[0].kb_in;
  int last_save = -1;
  for (int l = 0; l < p.num_ops; ++l)
    if (p.op[l].kind == 0 && p.op[l].save) last_save = l;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t num_tiles = (p.m + 255) / 256;
  const int64_t cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const uint32_t smem_base = smem_u32(smem);

  if (threadIdx.x == 0) {
    if (smem_base & 1023u) __trap();
    tma_prefetch_desc(&maps.in_hi);
    tma_prefetch_desc(&maps.in_lo);
    for (int i = 0; i < p.num_ops; ++i) {
      tma_prefetch_desc(&maps.w_hi[i]);
      tma_prefetch_desc(&maps.w_lo[i]);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&act_ready[i], 16);
      mbar_init(&in_full[i], 1);
    }
    mbar_init(&acc_free[0], 16);
    mbar_init(&acc_free[1], 16);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&written[i], 8);
      mbar_init(&drained[i], 1);
    }
    mbar_init(saves_drained, 1);
    mbar_init(&blk_read[0], 1);
    mbar_init(&blk_read[1], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: this CTA's rows of every weight item (hi, lo) and of the ring-fed input K blocks =====
    uint32_t pos = 0;
    const uint32_t ring_full_leader = map_to_cta(smem_u32(&ring_full[0]), 0);
    const uint32_t in_full_leader = map_to_cta(smem_u32(&in_full[0]), 0);
    auto acquire = [&]() -> uint32_t {
      const uint32_t s = pos % kStages, ph = (pos / kStages) & 1u;
      mbar_wait(&ring_empty[s], ph ^ 1u);
      ++pos;
      return s;
    };
    // weight item: K block `kblk` of output-column half `half` of op l (global ops: their only "half")
    auto load_w = [&](int l, const PairOp& L, int kblk, int half) {
      const uint32_t s = acquire();
      const int rows = L.n >= 256 ? 64 : (L.n >> 1);
      if (p.whatif & 2) {   // timing experiment: no weight traffic at all
        if (rank == 0 && elect_one_sync()) mbar_arrive(&ring_full[s]);
        __syncwarp();
        return;
      }
      if (elect_one_sync()) {
        if (rank == 0) mbar_arrive_expect_tx(&ring_full[s], (uint32_t)rows * 512u);   // 2 planes x 2 CTAs x rows x 128 B
        const uint32_t dst = smem_base + kSmemRing + s * kStageBytes;
        const int r0 = half * 128 + (int)rank * rows;
        tma_load_2d_pair(dst, &maps.w_hi[l], ring_full_leader + 8u * s, kblk * kBK, r0);
        tma_load_2d_pair(dst + kItemPlane, &maps.w_lo[l], ring_full_leader + 8u * s, kblk * kBK, r0);
      }
      __syncwarp();
    };
    auto load_in = [&](const CUtensorMap* map, int kb, int row0) {   // one plane of an input K block: a whole stage
      const uint32_t s = acquire();
      if (elect_one_sync()) {
        if (rank == 0) mbar_arrive_expect_tx(&ring_full[s], 2u * kPlaneBytes);
        tma_load_2d_pair(smem_base + kSmemRing + s * kStageBytes, map, ring_full_leader + 8u * s, kb * kBK, row0);
      }
      __syncwarp();
    };
    uint32_t tile_iter = 0;
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tile_iter) {
      const int row0 = (int)(tile * 256 + (int64_t)rank * 128);
      for (int l = 0; l < p.num_ops; ++l) {
        const PairOp& L = p.op[l];
        if (L.kind == 2) continue;
        const int nhalf = L.n >= 256 ? 2 : 1;
        // L2 prefetches, issued about one op (6 000+ cycles) ahead of the TMA loads they serve -- early enough to cover the
        // HBM latency, late enough that the save traffic streaming through L2 has not evicted the lines again:
        //   * the next op's ring-fed input K blocks (the skip layer re-reads the chain input 5 ops after the first op did);
        //   * the NEXT tile's chain input, three ops before the tile boundary where its load sits on the critical path
        if (p.in_kb && elect_one_sync()) {
          if (l + 1 < p.num_ops && l + 1 > 0 && p.op[l + 1].kind != 2 && p.op[l + 1].kb_in) {
            for (int kb = 0; kb < p.in_kb; ++kb) {
              tma_prefetch_l2_2d(&maps.in_hi, kb * kBK, row0);
              tma_prefetch_l2_2d(&maps.in_lo, kb * kBK, row0);
            }
          }
          if (l == (p.num_ops >= 3 ? p.num_ops - 3 : 0) && tile + num_clusters < num_tiles) {
            const int rown = (int)((tile + num_clusters) * 256 + (int64_t)rank * 128);
            for (int kb = 0; kb < p.in_kb; ++kb) {
              tma_prefetch_l2_2d(&maps.in_hi, kb * kBK, rown);
              tma_prefetch_l2_2d(&maps.in_lo, kb * kBK, rown);
            }
          }
        }
        __syncwarp();
        if (l == 0 && direct_in) {
          // the first weight items first, then the input tile (it has to wait for the previous tile), then the rest
          const int nres = nhalf * L.kb_in;
          const int early = nres < 3 ? nres : 3;
          int oh, ok;
          for (int i = 0; i < early; ++i) {
            res_order(p.split_order, i, L.kb_in, nhalf, oh, ok);
            load_w(l, L, ok, oh);
          }
          if (tile_iter) {
            // every MMA of the previous tile has retired (release of its last ring item) and its saves have been read
            mbar_wait(&ring_empty[(pos - early - 1) % kStages], ((pos - early - 1) / kStages) & 1u);
            if (last_save >= 0) mbar_wait(saves_drained, (tile_iter - 1) & 1u);
          }
          if (elect_one_sync()) {
            for (int kb = 0; kb < L.kb_in; ++kb) {
              if (rank == 0) mbar_arrive_expect_tx(&in_full[kb], 4u * kPlaneBytes);
              tma_load_2d_pair(smem_base + (in_blk0 + kb) * kPlaneBytes, &maps.in_hi, in_full_leader + 8u * kb, kb * kBK, row0);
              tma_load_2d_pair(smem_base + kActPlane + (in_blk0 + kb) * kPlaneBytes, &maps.in_lo, in_full_leader + 8u * kb, kb * kBK, row0);
            }
          }
          __syncwarp();
          for (int i = early; i < nres; ++i) {
            res_order(p.split_order, i, L.kb_in, nhalf, oh, ok);
            load_w(l, L, ok, oh);
          }
          continue;
        }
        // ring-fed input K blocks (skip layers) feed both column halves while they are resident
        for (int kb = 0; kb < L.kb_in; ++kb) {
          load_in(&maps.in_hi, kb, row0);
          load_in(&maps.in_lo, kb, row0);
          for (int half = 0; half < nhalf; ++half) load_w(l, L, L.kb_act + kb, half);
        }
        for (int i = 0; i < nhalf * L.kb_act; ++i) {
          int oh, ok;
          res_order(p.split_order, i, L.kb_act, nhalf, oh, ok);
          const bool trp = p.trace && blockIdx.x == 0 && lane == 0 && (i == 0 || i == 4) && tile_iter * p.num_ops + l < 64;
          if (trp) {   // (debug timeline) when this item's ring stage became free / when its TMA loads were issued
            mbar_wait(&ring_empty[pos % kStages], ((pos / kStages) & 1u) ^ 1u);
            p.trace[(tile_iter * p.num_ops + l) * kTraceSlots + 26 + (i ? 2 : 0)] = clock64();
          }
          load_w(l, L, ok, oh);
          if (trp) p.trace[(tile_iter * p.num_ops + l) * kTraceSlots + 27 + (i ? 2 : 0)] = clock64();
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader CTA; whole warp runs the loop, one elected lane issues) =====
    uint32_t pos = 0;        // ring position, mirrors the producer's item order
    uint32_t gemm_idx = 0;   // GEMM ops issued so far: accumulator = gemm_idx & 1
    uint32_t aver = 0;       // activation-tile versions acquired so far (act_ready phases)
    bool need_acq = false;   // the activation tile has been (is being) rewritten since it was last acquired
    uint32_t opcount = 0;
    auto wait_full = [&](uint32_t q) { mbar_wait(&ring_full[q % kStages], (q / kStages) & 1u); };
    auto stage_addr = [&](uint32_t q) -> uint32_t { return smem_base + kSmemRing + (q % kStages) * kStageBytes; };
    // the 12 MMAs of one K block of one column half (hi*hi + lo*hi + hi*lo per K step), then the commits
    auto mma_kblock = [&](uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t idesc,
                          bool first, uint64_t* commit0, uint64_t* commit1, uint64_t* commit2) {
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < kBK / kUmmaK; ++kk) {
          const uint32_t koff = kk * kUmmaK * 2;
          const uint64_t dah = make_desc(a_hi + koff, 16, 1024), dal = make_desc(a_lo + koff, 16, 1024);
          const uint64_t dbh = make_desc(b_hi + koff, 16, 1024), dbl = make_desc(b_lo + koff, 16, 1024);
          umma2_bf16(tmem_d, dah, dbh, idesc, (!first || kk) ? 1u : 0u);
          umma2_bf16(tmem_d, dal, dbh, idesc, 1u);
          umma2_bf16(tmem_d, dah, dbl, idesc, 1u);
        }
        if (commit0) umma2_commit(commit0);
        if (commit1) umma2_commit(commit1);
        if (commit2) umma2_commit(commit2);
      }
      __syncwarp();
    };
    uint32_t tile_iter = 0;
    // (the op descriptor of op l+1 is fetched while op l is issued, as in chain_x3t.cu)
    struct OpLite { int n, kb_act, kb_in, kind; };
    auto fetch = [&](int l) -> OpLite {
      const PairOp& o = p.op[l];
      return OpLite{o.n, o.kb_act, o.kb_in, o.kind};
    };
    OpLite nxt = fetch(0);
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tile_iter) {
      for (int l = 0; l < p.num_ops; ++l, ++opcount) {
        const OpLite L = nxt;
        nxt = fetch(l + 1 < p.num_ops ? l + 1 : 0);
        if (L.kind == 2) {   // seed op: the epilogue warps generate the activation tile, no MMA
          need_acq = true;
          continue;
        }
        const int nhalf = L.n >= 256 ? 2 : 1;
        const uint32_t idesc = make_idesc2(nhalf == 2 ? 128 : L.n);
        const uint32_t buf = gemm_idx & 1u, use = gemm_idx >> 1;
        const bool tr = p.trace && blockIdx.x == 0 && opcount < 64 && lane == 0;
        if (tr) p.trace[opcount * kTraceSlots + 8] = clock64();
        mbar_wait_cluster(&acc_free[buf], (use & 1u) ^ 1u);   // epilogue of the op two before has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * 256u;
        if (tr) p.trace[opcount * kTraceSlots + 0] = clock64();
        const bool direct = l == 0 && direct_in;
        const int kb_res = direct ? L.kb_in : L.kb_act;   // K blocks read from the activation tile
        const int blk0 = direct ? in_blk0 : 0;
        if (!direct) {
          for (int kb = 0; kb < L.kb_in; ++kb) {
            // items: in_hi, in_lo, W(half 0)[, W(half 1)]
            for (int i = 0; i < 2 + nhalf; ++i) wait_full(pos + i);
            tc_fence_after();
            const uint32_t sah = stage_addr(pos), sal = stage_addr(pos + 1);
            for (int half = 0; half < nhalf; ++half) {
              const uint32_t sb = stage_addr(pos + 2 + half);
              const bool last_half = half == nhalf - 1;
              const bool completes = !L.kb_act && kb == L.kb_in - 1;   // (never for the chains built today)
              mma_kblock(tmem_d + half * 128u, sah, sal, sb, sb + kItemPlane, idesc, kb == 0, &ring_empty[(pos + 2 + half) % kStages],
                         last_half ? &ring_empty[pos % kStages] : nullptr, last_half ? &ring_empty[(pos + 1) % kStages] : nullptr);
              if (completes && elect_one_sync()) umma2_commit(&acc_full[buf * 2 + half]);
              __syncwarp();
            }
            pos += 2 + nhalf;
          }
        }
        for (int i = 0; i < nhalf * kb_res; ++i) {
          int half, kb;
          res_order(p.split_order, i, kb_res, nhalf, half, kb);
          {
            if (half == 0) {
              if (direct) mbar_wait_cluster(&in_full[kb], tile_iter & 1u);
              else if (need_acq) mbar_wait_cluster(&act_ready[kb], aver & 1u);
            }
            if (tr && (i == 0 || i == 4)) p.trace[opcount * kTraceSlots + (i ? 31 : 30)] = clock64();   // before the weight wait
            wait_full(pos);
            tc_fence_after();
            if (tr && kb == 0 && half == 0) p.trace[opcount * kTraceSlots + 1] = clock64();
            if (tr && kb == kb_res - 1 && half == 0) p.trace[opcount * kTraceSlots + 7] = clock64();
            if (tr && i < 8) p.trace[opcount * kTraceSlots + 9 + i] = clock64();   // item i: every wait passed
            const uint32_t sb = stage_addr(pos);
            // (one-half ops complete both half barriers of their accumulator so that the two phase counts stay in step)
            const bool last_kb = kb == kb_res - 1;
            uint64_t* c1 = last_kb ? &acc_full[buf * 2 + half] : nullptr;
            uint64_t* c2 = (last_kb && nhalf == 1) ? &acc_full[buf * 2 + 1] : nullptr;
            if (half == 1) {
              // hidden ops: the first half's epilogue rewrites K blocks 0 and 1 in place while this half still reads the
              // tile; tell it when those two blocks have been read (a chain input that starts at block 2 never uses them)
              if (blk0 >= 2) {
                if (kb == 0) { c1 = &blk_read[0]; c2 = &blk_read[1]; }   // (kb_res == 2: kb 0 is not the last block)
              } else if (blk0 + kb < 2) {
                c2 = &blk_read[blk0 + kb];                                // (kb <= 1 < kb_res - 1: c2 is free)
              }
            }
            mma_kblock(tmem_d + half * 128u, smem_base + (blk0 + kb) * kPlaneBytes, smem_base + kActPlane + (blk0 + kb) * kPlaneBytes,
                       sb, sb + kItemPlane, idesc, kb == 0 && (direct || !L.kb_in), &ring_empty[pos % kStages], c1, c2);
            ++pos;
          }
        }
        if (tr) p.trace[opcount * kTraceSlots + 2] = clock64();
        if (!direct && L.kb_act && need_acq) {
          ++aver;
          need_acq = false;
        }
        if (L.kind == 0) need_acq = true;
        ++gemm_idx;
      }
    }
  } else if (warp == 3) {
    // ===== store warp: TMA stores of the hi-plane saves and of staged activation-format outputs =====
    uint32_t nw[2] = {0u, 0u};
    bool any = false;
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int64_t row_c = tile * 256 + (int64_t)rank * 128;
      for (int l = 0; l < p.num_ops; ++l) {
        const PairOp& L = p.op[l];
        const bool sv = L.kind == 0 && L.save;
        const bool staged = L.kind == 1 && p.gepi[L.gepi].out.hi != nullptr;
        if (!sv && !staged) continue;
        for (int ch = 0; ch < (sv ? 2 : 1); ++ch) {
          mbar_wait(&written[ch], nw[ch] & 1u);
          ++nw[ch];
          if (elect_one_sync()) {
            if (row_c < p.m && !(p.whatif & 4)) {
              if (sv) {
#pragma unroll
                for (int c = 2 * ch; c < 2 * ch + 2; ++c) tma_store_2d(&maps.save[l], smem_base + c * kPlaneBytes, c * 64, (int)row_c);
              } else {
                const int nb = p.gepi[L.gepi].out_cols >> 6;
                for (int c = 0; c < nb; ++c) {
                  tma_store_2d(&maps.gout_hi, smem_base + c * kPlaneBytes, c * 64, (int)row_c);
                  tma_store_2d(&maps.gout_lo, smem_base + kActPlane + c * kPlaneBytes, c * 64, (int)row_c);
                }
              }
              tma_store_commit();
              tma_store_wait_read<0>();
              any = true;
            }
            mbar_arrive(&drained[ch]);
            if (l == last_save && ch == 1) mbar_arrive(saves_drained);
          }
          __syncwarp();
        }
      }
    }
    if (any) tma_store_wait_all();
  } else if (warp >= 4) {
    // ===== epilogue warps: q = TMEM lane quadrant (32 rows), h = 32-column half of every 64-column K block =====
    const int q = (warp - 4) & 3;
    const int h = (warp - 4) >> 2;
    const int r_in_tile = q * 32 + lane;
    const uint32_t free_addr0 = map_to_cta(smem_u32(&acc_free[0]), 0);
    const uint32_t ready_addr0 = map_to_cta(smem_u32(&act_ready[0]), 0);
    const uint32_t act_row_hi = smem_base + (uint32_t)(r_in_tile * 128);
    const uint32_t act_row_lo = act_row_hi + kActPlane;
    const uint32_t swz = (uint32_t)(r_in_tile & 7);
    // bias / seed-vector staging: two buffers (op parity) x two column halves x 128 floats (this warp's 32 columns of
    // each K block).  The four warps of a column half write identical values, so no cross-warp barrier is needed:
    // every warp only relies on its own stores; a warp two ops ahead cannot exist (acc_free has 16 arrivals).
    const uint32_t bias_half = smem_base + kSmemBias + (uint32_t)h * 512u;
    uint32_t gemm_idx = 0, opcount = 0, nhid = 0;
    uint32_t ndrain[2] = {0u, 0u};
    bool save_outstanding[2] = {false, false};
    bool seed_early = false;   // the seed tile of the current row tile was generated at the end of the previous one
    auto wait_drained = [&](int ch) {
      if (save_outstanding[ch]) {
        mbar_wait(&drained[ch], ndrain[ch] & 1u);
        ++ndrain[ch];
        save_outstanding[ch] = false;
      }
    };
    // write this thread's 32 columns (half h of K block c) of both planes
    auto store_block = [&](int c, const uint32_t* hi, const uint32_t* lo) {
      const uint32_t bh = act_row_hi + (uint32_t)(c * kPlaneBytes), bl = act_row_lo + (uint32_t)(c * kPlaneBytes);
      if (p.whatif & 1) {   // timing experiment: keep the arithmetic alive, skip the stores
        uint32_t x = 0;
#pragma unroll
        for (int u = 0; u < 16; ++u) x ^= hi[u] ^ lo[u];
        if (x == 0x12345u) sts128(bh, x, x, x, x);
        fence_proxy_async();
        return;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t pos16 = ((uint32_t)(h * 4 + u) ^ swz) * 16u;
        sts128(bh + pos16, hi[4 * u], hi[4 * u + 1], hi[4 * u + 2], hi[4 * u + 3]);
        sts128(bl + pos16, lo[4 * u], lo[4 * u + 1], lo[4 * u + 2], lo[4 * u + 3]);
      }
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core and to TMA stores
    };
    auto stage_bias = [&](const float* ptr, int n, uint32_t buf_addr) {
      const int col = 64 * (lane >> 3) + 32 * h + 4 * (lane & 7);
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col < n) bv = __ldg(reinterpret_cast<const float4*>(ptr + col));
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(buf_addr + 16u * lane), "f"(bv.x), "f"(bv.y), "f"(bv.z), "f"(bv.w)
                   : "memory");
      __syncwarp();
    };
    // seed tile of a dgrad chain: vec[col] where the ReLU bit is set, generated in place (no accumulator)
    auto do_seed = [&](const PairOp& L, int64_t tile_s, uint32_t vec_buf) {
      const int64_t row_ws = tile_s * 256 + (int64_t)rank * 128 + q * 32;
      uint32_t bws[4] = {0u, 0u, 0u, 0u};
      if (row_ws < p.m) {
#pragma unroll
        for (int c = 0; c < 4; ++c) bws[c] = __ldg(L.mask_bits + (size_t)(row_ws >> 5) * 256 + lane + 32 * (2 * c + h));
      }
      stage_bias(L.bias, 256, vec_buf);
      wait_drained(0);
      wait_drained(1);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = lds128f(vec_buf + (uint32_t)(c * 128 + 16 * i));
          const float v0 = (bws[c] >> (2 * i) & 1u) ? b.x * p.seed_scale : 0.f;
          const float v1 = (bws[c] >> (16 + 2 * i) & 1u) ? b.y * p.seed_scale : 0.f;
          const float v2 = (bws[c] >> (2 * i + 1) & 1u) ? b.z * p.seed_scale : 0.f;
          const float v3 = (bws[c] >> (17 + 2 * i) & 1u) ? b.w * p.seed_scale : 0.f;
          split2(v0, v1, hi[2 * i], lo[2 * i]);
          split2(v2, v3, hi[2 * i + 1], lo[2 * i + 1]);
        }
        store_block(c, hi, lo);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_addr(ready_addr0 + 8u * c);
      }
    };
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int64_t row_w = tile * 256 + (int64_t)rank * 128 + q * 32;   // first row of this warp
      const int64_t row = row_w + lane;
      const bool row_ok = row < p.m;
      const size_t bits_base = (size_t)(row_w >> 5) * 256 + lane;        // word w of the 32-row group at + 32 w
      for (int l = 0; l < p.num_ops; ++l, ++opcount) {
        const PairOp& L = p.op[l];
        const bool seed = MODE == 1 && L.kind == 2;
        const uint32_t bias_buf = bias_half + (opcount & 1u) * 1024u;
        if (seed) {
          if (!seed_early) do_seed(L, tile, bias_buf);
          seed_early = false;
          continue;
        }
        const float* bias_ptr = L.kind == 1 ? p.gepi[L.gepi].bias : (MODE == 0 ? L.bias : nullptr);
        if (bias_ptr) stage_bias(bias_ptr, L.n, bias_buf);
        uint32_t bw[4] = {0u, 0u, 0u, 0u};
        if (MODE == 1 && L.kind == 0 && row_w < p.m) {
#pragma unroll
          for (int c = 0; c < 4; ++c) bw[c] = __ldg(L.mask_bits + bits_base + 32 * (2 * c + h));
        }
        const uint32_t buf = gemm_idx & 1u, use = gemm_idx >> 1;
        ++gemm_idx;
        const bool tr = p.trace && blockIdx.x == 0 && warp == 4 && lane == 0 && opcount < 64;
        if (tr) p.trace[opcount * kTraceSlots + 3] = clock64();
        const uint32_t taddr = tmem_base + buf * 256u + ((uint32_t)(q * 32) << 16);
        if (L.kind == 0) {
          uint32_t bits_out[4];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            mbar_wait(&acc_full[buf * 2 + half], use & 1u);
            tc_fence_after();
            if (half == 0 && tr) p.trace[opcount * kTraceSlots + 4] = clock64();
            wait_drained(half);
            uint32_t ra[32], rb[32];
            tmem_ld32(taddr + (uint32_t)(128 * half + 32 * h), ra);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const int c = 2 * half + cc;
              uint32_t* cur = cc ? rb : ra;
              tmem_ld_wait();
              if (cc == 0) tmem_ld32(taddr + (uint32_t)(64 * (c + 1) + 32 * h), rb);
              uint32_t hi[16], lo[16];
              if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 b = lds128f(bias_buf + (uint32_t)(c * 128 + 16 * i));
                  const float v0 = fmaxf(__uint_as_float(cur[4 * i]) + b.x, 0.f);
                  const float v1 = fmaxf(__uint_as_float(cur[4 * i + 1]) + b.y, 0.f);
                  const float v2 = fmaxf(__uint_as_float(cur[4 * i + 2]) + b.z, 0.f);
                  const float v3 = fmaxf(__uint_as_float(cur[4 * i + 3]) + b.w, 0.f);
                  split2(v0, v1, hi[2 * i], lo[2 * i]);
                  split2(v2, v3, hi[2 * i + 1], lo[2 * i + 1]);
                }
                if (L.save_bits) bits_out[c] = relu_bits_of(hi);
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float v0 = (bw[c] >> i & 1u) ? __uint_as_float(cur[2 * i]) : 0.f;
                  const float v1 = (bw[c] >> (16 + i) & 1u) ? __uint_as_float(cur[2 * i + 1]) : 0.f;
                  split2(v0, v1, hi[i], lo[i]);
                }
              }
              if (half == 0) mbar_wait(&blk_read[c], nhid & 1u);   // the second half's MMAs are done with the old block c
              store_block(c, hi, lo);
              __syncwarp();
              if (lane == 0) {
                // (the store warp's signal goes first: no warp can reach the next store event before every warp has
                // signalled this one, because the next op's accumulator needs this op's last block from all of them)
                if (cc == 1 && L.save) mbar_arrive(&written[half]);
                mbar_arrive_cluster_addr(ready_addr0 + 8u * c);
              }
              if (tr && c == 0) p.trace[opcount * kTraceSlots + 5] = clock64();
              if (p.trace && blockIdx.x == 0 && lane == 0 && opcount < 64 && c == 3) p.trace[opcount * kTraceSlots + 18 + (warp - 4)] = clock64();
            }
          }
          ++nhid;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_addr(free_addr0 + 8u * buf);
          if (L.save) save_outstanding[0] = save_outstanding[1] = true;
          if (MODE == 0 && L.save_bits && row_w < p.m) {
#pragma unroll
            for (int c = 0; c < 4; ++c) L.save_bits[bits_base + 32 * (2 * c + h)] = bits_out[c];
          }
        } else {
          // global op (n <= 128, one column half): columns [0, out_cols) leave in activation format (hi/lo) through the
          // activation tile + TMA stores, the f32 range through ordinary stores
          GemmEpilogue ge = p.gepi[L.gepi];
          const bool staged = ge.out.hi != nullptr;
          const int out_cols = staged ? ge.out_cols : 0;
          ge.out.hi = nullptr;
          mbar_wait(&acc_full[buf * 2], use & 1u);
          tc_fence_after();
          if (tr) p.trace[opcount * kTraceSlots + 4] = clock64();
          // dgrad chains: the next row tile's seed only needs the activation tile, which this op's MMAs have released
          if (MODE == 1 && l == p.num_ops - 1 && p.op[0].kind == 2 && tile + num_clusters < num_tiles) {
            do_seed(p.op[0], tile + num_clusters, bias_half + ((opcount + 1) & 1u) * 1024u);
            seed_early = true;
          }
          if (staged) wait_drained(0);
          for (int gi = h; gi * 32 < L.n; gi += 2) {
            const int col0 = gi * 32;
            uint32_t r[32];
            tmem_ld32(taddr + (uint32_t)col0, r);
            tmem_ld_wait();
            const uint32_t bsm = bias_buf + (uint32_t)((gi >> 1) * 128);   // this group's 32 bias values
            if (col0 < out_cols) {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias_ptr) b = lds128f(bsm + 16u * i);
                split2(__uint_as_float(r[4 * i]) + b.x, __uint_as_float(r[4 * i + 1]) + b.y, hi[2 * i], lo[2 * i]);
                split2(__uint_as_float(r[4 * i + 2]) + b.z, __uint_as_float(r[4 * i + 3]) + b.w, hi[2 * i + 1], lo[2 * i + 1]);
              }
              store_block(col0 >> 6, hi, lo);   // (col0 >> 5) & 1 == h by construction
            } else if (row_ok) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                if (col0 + 16 * hh < L.n) {
                  float v[16];
#pragma unroll
                  for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[16 * hh + e]);
                  if (bias_ptr) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                      const float4 b = lds128f(bsm + (uint32_t)(64 * hh + 16 * i));
                      v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
                    }
                  }
                  epi_global16<0>(ge, (size_t)row, col0 + 16 * hh, v);
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (staged) mbar_arrive(&written[0]);
            mbar_arrive_cluster_addr(free_addr0 + 8u * buf);
          }
          if (staged) save_outstanding[0] = true;
        }
        if (tr) p.trace[opcount * kTraceSlots + 6] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc2(tmem_base, kTmemCols);
}

}  // namespace

int launch_chain_x3(const ChainArgs& a, cudaStream_t st) {
  if (a.m <= 0) return RN_OK;
  // default: the tensor-memory-resident variant (chain_x3t.cu), same op list and results; RN_X3_TS=0 runs this file's kernel
  static const int use_ts = getenv("RN_X3_TS") ? atoi(getenv("RN_X3_TS")) : 1;
  if (use_ts) return launch_chain_x3t(a, st);
  if (a.act_f16 || a.w_f16) return rn_set_error(RN_ERR_UNSUPPORTED, "chain_x3: bf16 planes only");
  if (a.num_ops < 1 || a.num_ops > kMaxOps) return rn_set_error(RN_ERR_ARG, "chain_x3: 1..12 ops");
  if (a.in.hi && (!a.in.lo || a.in_cols % 64 || a.in_cols < 64 || a.in_cols > 256))
    return rn_set_error(RN_ERR_ARG, "chain_x3: input tile must have both planes and 64..256 columns");
  if (a.in2.hi) return rn_set_error(RN_ERR_UNSUPPORTED, "chain_x3: no second input");
  if (a.m + 256 > 0x7fffffffLL) return rn_set_error(RN_ERR_ARG, "chain_x3: too many rows for one launch");
  X3Maps maps;
  PairParams p;
  memset(&p, 0, sizeof(p));
  memset(&maps, 0, sizeof(maps));
  int rc;
  int mode = -1;
  if ((rc = tc::make_map(&maps.in_hi, a.in.hi, a.m, a.in_valid, a.in.ld, kBM))) return rc;
  if ((rc = tc::make_map(&maps.in_lo, a.in.lo, a.m, a.in_valid, a.in.ld, kBM))) return rc;
  p.in2_sync_op = -1;
  // (measured on the B200: both orders run the 16 384-ray step within 0.5 % of each other -- the chains are paced by the
  // weight ring's L2 latency and the power cap, not by the epilogue hand-over -- so the plain order stays the default)
  static const int split_order = getenv("RN_X3_ORDER") ? atoi(getenv("RN_X3_ORDER")) : 0;
  p.split_order = split_order;
  static const int whatif = getenv("RN_X3_WHATIF") ? atoi(getenv("RN_X3_WHATIF")) : 0;
  p.whatif = whatif;
  if (whatif) {
    static bool warned = false;
    if (!warned) fprintf(stderr, "refnerf_b200: RN_X3_WHATIF=%d -- TIMING EXPERIMENT, the chain kernels skip work and their results are GARBAGE\n", whatif);
    warned = true;
  }
  p.seed_scale = a.seed_scale;
  p.num_ops = a.num_ops;
  p.in_kb = a.in.hi ? a.in_cols / kBK : 0;
  p.m = a.m;
  bool awrite_pending = false;   // a hidden / seed op whose activation tile no later op has read yet
  int staged_gepi = -1;
  for (int l = 0; l < a.num_ops; ++l) {
    const ChainOpArgs& L = a.op[l];
    PairOp& o = p.op[l];
    if (L.kind == 2) {
      if (!L.bias || !L.mask_bits || L.kb_act || L.kb_in || L.save_hi || awrite_pending || l != 0)
        return rn_set_error(RN_ERR_ARG, "chain_x3: bad seed op (first op of a backward chain only)");
      if (mode < 0) mode = 1;
      if (mode != 1) return rn_set_error(RN_ERR_ARG, "chain_x3: seed ops belong to backward chains");
      o.n = 256; o.kind = 2; o.bias = L.bias; o.mask_bits = L.mask_bits;
      awrite_pending = true;
      continue;
    }
    if (L.n % 16 || L.n < 16 || (L.kind == 0 && L.n != 256) || (L.kind != 0 && L.n > 128))
      return rn_set_error(RN_ERR_ARG, "chain_x3: hidden ops are 256 wide, global ops at most 128");
    if ((L.kb_act != 0 && L.kb_act != 4) || (L.kb_in != 0 && L.kb_in != p.in_kb) || L.kb_act + L.kb_in == 0 || L.in2 ||
        (L.kb_in && !a.in.hi) || !L.w || !L.w_lo)
      return rn_set_error(RN_ERR_ARG, "chain_x3: bad K structure / missing weight plane");
    if (l > 0 && !L.kb_act) return rn_set_error(RN_ERR_ARG, "chain_x3: only the first op may read the chain input alone");
    if (awrite_pending && !L.kb_act)
      return rn_set_error(RN_ERR_ARG, "chain_x3: the op after a hidden / seed op must read the activation tile");
    if (L.kb_act) awrite_pending = false;
    const int ktot = (L.kb_act + L.kb_in) * kBK;
    const int box_rows = L.n >= 256 ? 64 : L.n / 2;
    if ((rc = tc::make_map(&maps.w_hi[l], L.w, L.n, ktot, L.w_ld, box_rows))) return rc;
    if ((rc = tc::make_map(&maps.w_lo[l], L.w_lo, L.n, ktot, L.w_ld, box_rows))) return rc;
    o.n = L.n; o.kb_act = L.kb_act; o.kb_in = L.kb_in; o.in2 = 0;
    o.kind = L.kind; o.gepi = L.gepi; o.bias = L.bias;
    o.mask_bits = L.mask_bits; o.save_bits = L.save_bits;
    o.save = (L.kind == 0 && L.save_hi) ? 1 : 0;
    if (o.save && (rc = tc::make_map(&maps.save[l], L.save_hi, a.m, 256, 256, kBM))) return rc;
    if (L.kind == 0) {
      if (mode < 0) mode = L.mode;
      if (L.mode != mode) return rn_set_error(RN_ERR_ARG, "chain_x3: forward and backward hidden ops cannot be mixed");
      if (L.mode == 0 && !L.bias) return rn_set_error(RN_ERR_ARG, "chain_x3: forward op without bias");
      if (L.mode == 1 && !L.mask_bits) return rn_set_error(RN_ERR_ARG, "chain_x3: backward op without ReLU bits");
      awrite_pending = true;
    } else {
      if (L.gepi < 0 || L.gepi > 1) return rn_set_error(RN_ERR_ARG, "chain_x3: bad epilogue index");
      const GemmEpilogue& e = a.gepi[L.gepi];
      if (e.relu || e.mask.hi) return rn_set_error(RN_ERR_UNSUPPORTED, "chain_x3: global ops are linear");
      if (e.out.hi) {
        if (!e.out.lo || e.out_cols % 64 || e.out_cols < 64 || e.out_cols > L.n || (staged_gepi >= 0 && staged_gepi != L.gepi))
          return rn_set_error(RN_ERR_ARG, "chain_x3: a staged output needs both planes, 64..128 columns, one epilogue");
        if (staged_gepi < 0) {
          if ((rc = tc::make_map(&maps.gout_hi, e.out.hi, a.m, e.out_cols, e.out.ld, kBM))) return rc;
          if ((rc = tc::make_map(&maps.gout_lo, e.out.lo, a.m, e.out_cols, e.out.ld, kBM))) return rc;
          staged_gepi = L.gepi;
        }
      }
    }
  }
  if (a.op[0].kind != 2 && a.op[0].kb_act) return rn_set_error(RN_ERR_ARG, "chain_x3: the first op reads the chain input only");
  if (a.op[0].kind == 0 && a.op[0].kb_in != 2 && a.op[0].kb_in != 4)
    return rn_set_error(RN_ERR_ARG, "chain_x3: a hidden first op takes a 128- or 256-column chain input");
  if (awrite_pending) return rn_set_error(RN_ERR_ARG, "chain_x3: the last hidden op's result is never read");
  if (a.op[a.num_ops - 1].kind != 1) return rn_set_error(RN_ERR_ARG, "chain_x3: the last op must be a global op");
  if (mode < 0) mode = 0;
  p.gepi[0] = a.gepi[0];
  p.gepi[1] = a.gepi[1];
  static bool smem_set = false;
  if (!smem_set) {
    if ((rc = tc::set_smem(chain_x3_kernel<0>, kSmemTotal))) return rc;
    if ((rc = tc::set_smem(chain_x3_kernel<1>, kSmemTotal))) return rc;
    smem_set = true;
  }
  const int64_t tiles = (a.m + 255) / 256;
  const int max_clusters = tc::num_sms() / 2;
  const unsigned grid = 2u * (unsigned)(tiles < max_clusters ? tiles : max_clusters);
  // debug: RN_CHAIN_TRACE=<n> prints the per-op timeline of CTA 0 for the first n launches
  static long long* trace_buf = nullptr;
  static int trace_left = getenv("RN_CHAIN_TRACE") ? atoi(getenv("RN_CHAIN_TRACE")) : 0;
  p.trace = nullptr;
  if (trace_left > 0) {
    if (!trace_buf) cudaMalloc(&trace_buf, 64 * kTraceSlots * sizeof(long long));
    cudaMemsetAsync(trace_buf, 0, 64 * kTraceSlots * sizeof(long long), st);
    p.trace = trace_buf;
  }
  double exec_flops = 0.0;   // three MMAs per K step
  for (int l = 0; l < a.num_ops; ++l)
    if (a.op[l].kind != 2) exec_flops += 3.0 * 2.0 * (double)a.m * a.op[l].n * (a.op[l].kb_act + a.op[l].kb_in) * kBK;
  rn_prof_begin(RN_PROF_CHAIN_TC, st, a.algo_flops, exec_flops);
  if (mode == 0)
    chain_x3_kernel<0><<<grid, 384, kSmemTotal, st>>>(maps, p);
  else
    chain_x3_kernel<1><<<grid, 384, kSmemTotal, st>>>(maps, p);
  rn_prof_end(RN_PROF_CHAIN_TC, st);
  RN_CUDA_CHECK_LAUNCH();
  if (p.trace) {
    --trace_left;
    cudaStreamSynchronize(st);
    static long long hbuf[64 * kTraceSlots];
    cudaMemcpy(hbuf, trace_buf, sizeof(hbuf), cudaMemcpyDeviceToHost);
    long long t0 = 0;
    for (int i = 0; i < 64 * kTraceSlots && !t0; ++i) t0 = hbuf[i];
    for (int i = 0; i < 64 * kTraceSlots; ++i)
      if (hbuf[i] && hbuf[i] < t0) t0 = hbuf[i];
    printf("chain_x3 trace (mode %d, %d ops, m=%lld, order %d): per op: mma_acc_free first_blk_ready mma_issued | epi_begin acc_full(half 0) blk0_handed epi_end | last_blk_ready(half 0)"
           " || MMA warp: before acc_free wait, waits of items 0..7 passed || block 3 handed by epilogue warps 4..11"
           " || producer: stage free / TMA issued for items 0 and 4 || MMA warp: before the weight wait of items 0 and 4  [cycles since first event]\n",
           mode, a.num_ops, (long long)a.m, p.split_order);
    for (int i = 0; i < 44; ++i) {
      printf("  op %2d:", i);
      for (int j = 0; j < 32; ++j) {
        if (j == 3 || j == 7) printf(" |");
        if (j == 8 || j == 18 || j == 26 || j == 30) printf(" ||");
        if (hbuf[i * kTraceSlots + j]) printf(" %7lld", hbuf[i * kTraceSlots + j] - t0); else printf("       -");
      }
      printf("\n");
    }
  }
  return RN_OK;
}

}  // namespace rn
