// Fused split-bf16 GEMM chain with the running activation in TENSOR MEMORY (sm_100a, tcgen05.mma cta_group::2, TS form):
// the same arithmetic, op list and results as chain_x3.cu (hi*hi + lo*hi + hi*lo per K step, fp32 accumulate, same K order
// per accumulator element), re-laid-out because chain_x3.cu is bound by the 128 B/clk shared-memory port, not by the tensor
// pipe (ncu, profiles/r02b_step_ncu_full.txt: tensor-core operand reads 4.5 M + LSU 1.9 M + TMA 1.6 M wavefronts per SM
// against 8.7 M cycles of a spatial forward launch = 92 % of the port; profiles/r02b_chain_whatif.txt).  Two thirds of the
// operand reads are the A operand (activation rows: 4 KB of the 6 KB an M256 x N128 x K16 MMA fetches per CTA).  Here:
//
//   * TMEM (512 columns) = ONE 256-column fp32 accumulator (two column halves) | A_hi (128 columns) | A_lo (128 columns).
//     The A planes hold the 128 rows x 256 K elements of the running activation, two bf16 K elements per 32-bit column,
//     row = TMEM lane; the MMAs read them in place (tcgen05.mma [d], [a_tmem], b_desc), so only the weight operand (2 KB
//     per MMA and CTA) crosses the shared-memory port.
//   * the accumulator is single-buffered: op l+1 may start on column half h as soon as the epilogue warps have READ half h
//     of op l (acc_read[h]); with the column-half pipeline of chain_x3.cu that is already the order of events -- half 0 of
//     op l is drained while half 1 runs, half 1 while op l+1 works on its own half 0.
//   * the epilogue writes bias + ReLU / ReLU-masked values, split into hi / lo, back with tcgen05.st as K blocks of the next
//     A operand (same hand-over barriers as chain_x3.cu: act_ready[c], blk_read[c] for the in-place rewrite of K blocks 0, 1).
//   * shared memory carries the weight ring (8 stages instead of 6), the ring-fed chain-input K blocks (first op and skip
//     layer, SS-form MMAs), and a staging tile for what leaves by TMA store: the hi plane of a training forward's saves and
//     the hi / lo planes of an activation-format global output (the bottleneck -> v0).
//   * a one-half (global) op runs in accumulator half 1 unless it follows another global op (then half 0) and hands the
//     half it does not use back at once, so neither the second heads op nor the hidden op after a global op waits for a
//     global epilogue; an activation-format global output is staged in hi-plane blocks 2, 3 + the lo-plane blocks (store
//     channel 1), which the next row tile overwrites 3 000 cycles later than blocks 0, 1.
//   * K steps of the chain input that are zero padding are not issued (96 of 128 IPE columns, 201 of 256 view-net
//     inputs hold data), and the MMA warp fetches the descriptor of op l+1 while it issues op l (constant-bank loads with
//     a run-time op index on the issue path of every item cost 1-2 % of the step).
//   * an op that reads both the activation and the chain input (the skip layer) is issued as: half 0 over the activation K
//     blocks, then the input K blocks for both halves, then half 1 over the activation K blocks -- half 1 of the
//     accumulator is then first touched 3 072 cycles into the op, when the previous op's half-1 epilogue is long done.
//
// Roles per CTA (384 threads) as in chain_x3.cu: warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warp 2 TMEM
// allocator, warp 3 store warp, warps 4-11 epilogue (warp % 4 = TMEM lane quadrant, (warp - 4) / 4 = 32-column half of every
// 64-column K block).
#include <stdio.h>
#include <stdlib.h>

#include "chain_common.cuh"

namespace rn {
namespace {
using namespace tc;
using namespace chain;

constexpr int kStages = 8;
constexpr int kTraceSlots = 8;
constexpr int kItemPlane = 8192;               // [64 x 64] bf16: one plane of a weight item
constexpr int kStageBytes = 2 * kItemPlane;    // hi | lo  (an input item is one [128 x 64] plane = the whole stage)
constexpr int kPlaneBytes = 16384;             // [128 x 64] bf16, 128B swizzle: one K block of one staging plane
constexpr int kStgHi = 0;                      // staging tile, hi plane: 4 K blocks (saves, staged outputs)
constexpr int kStgLo = 4 * kPlaneBytes;        // staging tile, lo plane: K blocks 0, 1 (staged outputs only)
constexpr int kSmemRing = 6 * kPlaneBytes;
constexpr int kSmemBars = kSmemRing + kStages * kStageBytes;
constexpr int kSmemBias = kSmemBars + 512;     // [op parity][column half h][4 K blocks x 32 floats]
constexpr int kSmemTotal = kSmemBias + 2048;
static_assert(kSmemTotal <= 232448, "shared memory budget");
constexpr uint32_t kAccCol = 0, kAHiCol = 256, kALoCol = 384;   // TMEM columns

struct X3Maps {
  CUtensorMap in_hi, in_lo;       // chain input [m, in_cols], box [128 x 64]
  CUtensorMap gout_hi, gout_lo;   // activation-format output of a global op, box [128 x 64]
  CUtensorMap w_hi[kMaxOps], w_lo[kMaxOps];   // box [64 rows (hidden ops) or n/2 rows (global ops) x 64]
  CUtensorMap save[kMaxOps];      // hi-plane saves of hidden ops, box [128 x 64]
};

// A operand from tensor memory (row = lane, two K elements per 32-bit column)
__device__ __forceinline__ void umma2_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE 0: forward chain (hidden ops: bias + ReLU); MODE 1: dgrad chain (hidden ops: ReLU bit mask; seed ops)
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
chain_x3t_kernel(const __grid_constant__ X3Maps maps, const __grid_constant__ PairParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBars);
  uint64_t* ring_full = bars;          // [8] leader's are used (TMA bytes of both CTAs land there)
  uint64_t* ring_empty = bars + 8;     // [8] per CTA, multicast MMA commit
  uint64_t* acc_full = bars + 16;      // [2 column halves] per CTA, multicast MMA commit
  uint64_t* acc_read = bars + 18;      // [2] leader's: 16 arrivals (8 epilogue warps x 2 CTAs): half h has been read out
  uint64_t* act_ready = bars + 20;     // [4] leader's: 16 arrivals: K block c of the A planes is written
  uint64_t* written = bars + 24;       // [2] per CTA: 8 arrivals: the staging blocks a store needs are in shared memory
  uint64_t* drained = bars + 26;       // [2] per CTA: the store warp's TMA stores have read those blocks
  uint64_t* blk_read = bars + 28;      // [2] per CTA, MMA commit: the second column half's MMAs have read A K block 0 / 1 of
                                       //     a hidden op, i.e. the first half's epilogue may overwrite it in place
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t num_tiles = (p.m + 255) / 256;
  const int64_t cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const uint32_t smem_base = smem_u32(smem);
  const int sch = p.split_order ? 1 : 0;   // store channel (and hi-plane block pair) of a staged activation-format output

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
    for (int i = 0; i < 4; ++i) mbar_init(&act_ready[i], 16);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_read[i], 16);
      mbar_init(&written[i], 8);
      mbar_init(&drained[i], 1);
      mbar_init(&blk_read[i], 1);
    }
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
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int row0 = (int)(tile * 256 + (int64_t)rank * 128);
      for (int l = 0; l < p.num_ops; ++l) {
        const PairOp& L = p.op[l];
        if (L.kind == 2) continue;
        const int nhalf = L.n >= 256 ? 2 : 1;
        // L2 prefetches, about one op ahead of the TMA loads they serve: the next op's ring-fed input K blocks, and the NEXT
        // tile's chain input three ops before the tile boundary
        if (p.in_kb && elect_one_sync()) {
          if (l + 1 < p.num_ops && p.op[l + 1].kind != 2 && p.op[l + 1].kb_in) {
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
        // item order = the MMA warp's: half 0 over the activation K blocks, the input K blocks (both halves while the
        // block is resident), the remaining halves over the activation K blocks
        for (int kb = 0; kb < L.kb_act; ++kb) load_w(l, L, kb, 0);
        for (int kb = 0; kb < L.kb_in; ++kb) {
          load_in(&maps.in_hi, kb, row0);
          load_in(&maps.in_lo, kb, row0);
          for (int half = 0; half < nhalf; ++half) load_w(l, L, L.kb_act + kb, half);
        }
        for (int half = 1; half < nhalf; ++half)
          for (int kb = 0; kb < L.kb_act; ++kb) load_w(l, L, kb, half);
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader CTA; whole warp runs the loop, one elected lane issues) =====
    uint32_t pos = 0;        // ring position, mirrors the producer's item order
    uint32_t gemm_idx = 0;   // GEMM ops issued so far (phase of acc_full / acc_read)
    uint32_t aver = 0;       // A-tile versions acquired so far (act_ready phases)
    bool need_acq = false;   // the A tile has been (is being) rewritten since it was last acquired
    uint32_t opcount = 0;
    // (Moving an item's barrier waits one item ahead -- before the MMAs of the item in front of it -- was measured and is
    // slower: a passing mbarrier.try_wait costs the MMA warp 150-250 cycles under load wherever it stands, and the tensor
    // pipe's queue is too short to cover it; profiles/r02b_chain_x3t_trace_train.txt.)
    auto wait_full = [&](uint32_t q) { mbar_wait(&ring_full[q % kStages], (q / kStages) & 1u); };
    auto stage_addr = [&](uint32_t q) -> uint32_t { return smem_base + kSmemRing + (q % kStages) * kStageBytes; };
    // the 12 MMAs of one K block of one column half (hi*hi + lo*hi + hi*lo per K step), then up to four commits.
    // ts: the A planes are the TMEM columns a_hi / a_lo (8 columns per K step); else shared-memory addresses (SS form)
    auto mma_kblock = [&](bool ts, uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t idesc,
                          bool first, uint64_t* c0, uint64_t* c1, uint64_t* c2, uint64_t* c3, uint64_t* c4 = nullptr,
                          uint64_t* c5 = nullptr, int nsteps = kBK / kUmmaK) {
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < kBK / kUmmaK; ++kk) {
          if (kk >= nsteps) break;   // (the rest of this input K block is zero padding: x + 0 * w = x, exactly)
          const uint32_t koff = kk * kUmmaK * 2;
          const uint64_t dbh = make_desc(b_hi + koff, 16, 1024), dbl = make_desc(b_lo + koff, 16, 1024);
          const uint32_t acc = (!first || kk) ? 1u : 0u;
          if (ts) {
            umma2_ts_bf16(tmem_d, a_hi + 8u * kk, dbh, idesc, acc);
            umma2_ts_bf16(tmem_d, a_lo + 8u * kk, dbh, idesc, 1u);
            umma2_ts_bf16(tmem_d, a_hi + 8u * kk, dbl, idesc, 1u);
          } else {
            const uint64_t dah = make_desc(a_hi + koff, 16, 1024), dal = make_desc(a_lo + koff, 16, 1024);
            umma2_bf16(tmem_d, dah, dbh, idesc, acc);
            umma2_bf16(tmem_d, dal, dbh, idesc, 1u);
            umma2_bf16(tmem_d, dah, dbl, idesc, 1u);
          }
        }
        if (c0) umma2_commit(c0);
        if (c1) umma2_commit(c1);
        if (c2) umma2_commit(c2);
        if (c3) umma2_commit(c3);
        if (c4) umma2_commit(c4);
        if (c5) umma2_commit(c5);
      }
      __syncwarp();
    };
    // The op descriptor of op l+1 is fetched while op l is issued: read at the op boundary, the dependent constant-bank
    // loads (run-time op index) sit on the MMA warp's critical path while the tensor pipe's queue runs dry.
    struct OpLite { int n, kb_act, kb_in, kind, acc_half; };
    auto fetch = [&](int l) -> OpLite {
      const PairOp& o = p.op[l];
      return OpLite{o.n, o.kb_act, o.kb_in, o.kind, o.acc_half};
    };
    OpLite nxt = fetch(0);
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      for (int l = 0; l < p.num_ops; ++l, ++opcount) {
        const OpLite L = nxt;
        nxt = fetch(l + 1 < p.num_ops ? l + 1 : 0);
        if (L.kind == 2) {   // seed op: the epilogue warps generate the A tile, no MMA
          need_acq = true;
          continue;
        }
        const int nhalf = L.n >= 256 ? 2 : 1;
        const uint32_t idesc = make_idesc2(nhalf == 2 ? 128 : L.n);
        const uint32_t rd_par = (gemm_idx & 1u) ^ 1u;   // acc_read phase of the previous GEMM op (passes at once for the first)
        const uint32_t top = opcount - (uint32_t)p.whatif;   // (RN_CHAIN_TRACE_SKIP ops of this CTA are not recorded)
        const bool tr = p.trace && blockIdx.x == 0 && top < 64u && lane == 0;
        if (tr) p.trace[top * kTraceSlots + 0] = clock64();
        bool acc_ok[2] = {false, false}, started[2] = {false, false};
        const int gh = nhalf == 1 ? L.acc_half : 0;   // one-half ops: the accumulator half they run in
        // column half `half` of the accumulator: the previous op's epilogue must have read it out
        auto acc_acquire = [&](int half) {
          if (!acc_ok[half]) {
            mbar_wait(&acc_read[half], rd_par);
            tc_fence_after();
            acc_ok[half] = true;
          }
        };
        // one (half, activation K block) item: TS-form MMAs on A K block kb
        auto act_item = [&](int half, int kb) {
          if (half == 0 && need_acq) mbar_wait(&act_ready[kb], aver & 1u);
          acc_acquire(half + gh);
          if (nhalf == 1 && kb == L.kb_act - 1) acc_acquire(gh ^ 1);   // (phase protection of the half this op does not use)
          if (tr && kb == 0 && half == 0) p.trace[top * kTraceSlots + 7] = clock64();   // before the weight wait
          wait_full(pos);
          tc_fence_after();
          if (tr && kb == 0 && half == 0) p.trace[top * kTraceSlots + 1] = clock64();
          const uint32_t sb = stage_addr(pos);
          const bool last = kb == L.kb_act - 1 && (half > 0 || L.kb_in == 0);   // last item of this column half
          uint64_t* c1 = last ? &acc_full[half] : nullptr;
          uint64_t* c2 = (last && nhalf == 1) ? &acc_full[1] : nullptr;
          // hidden ops: the first half's epilogue rewrites A K blocks 0 and 1 in place while the second half still reads them
          uint64_t* c3 = (half == 1 && kb < 2) ? &blk_read[kb] : nullptr;
          mma_kblock(true, tmem_base + kAccCol + (half + gh) * 128u, tmem_base + kAHiCol + 32u * kb, tmem_base + kALoCol + 32u * kb,
                     sb, sb + kItemPlane, idesc, !started[half], &ring_empty[pos % kStages], c1, c2, c3);
          started[half] = true;
          ++pos;
        };
        for (int kb = 0; kb < L.kb_act; ++kb) act_item(0, kb);
        for (int kb = 0; kb < L.kb_in; ++kb) {
          // items: in_hi, in_lo, W(half 0)[, W(half 1)]: SS-form MMAs on the ring-fed input K block
          for (int i = 0; i < 2 + nhalf; ++i) wait_full(pos + i);
          tc_fence_after();
          const uint32_t sah = stage_addr(pos), sal = stage_addr(pos + 1);
          for (int half = 0; half < nhalf; ++half) {
            acc_acquire(half + gh);
            const uint32_t sb = stage_addr(pos + 2 + half);
            const bool last_half = half == nhalf - 1;
            const bool last_in = kb == L.kb_in - 1;
            // half 0 is complete after its last input block; any half of an op without activation blocks likewise
            const bool completes = last_in && (half == 0 || L.kb_act == 0);
            uint64_t* c1 = last_half ? &ring_empty[pos % kStages] : nullptr;
            uint64_t* c2 = last_half ? &ring_empty[(pos + 1) % kStages] : nullptr;
            uint64_t* c3 = completes ? &acc_full[half] : nullptr;
            // keep the phase counts of the barriers this op does not otherwise complete in step: a one-half op completes
            // acc_full[1] as well; a hidden op that never reads the A tile (the first op) releases A K blocks 0, 1 at once
            uint64_t* c4 = (completes && nhalf == 1) ? &acc_full[1] : nullptr;
            uint64_t* c5 = nullptr;
            if (L.kb_act == 0 && L.kind == 0 && kb == 0 && half == 0) {
              c4 = &blk_read[0];   // (nhalf == 2 here, so c4 is free)
              c5 = &blk_read[1];
            }
            // K steps of 16 columns that hold chain-input data (96 of 128 IPE columns, 201 of 256 view-net inputs)
            int nsteps = (p.in_valid - kb * kBK + kUmmaK - 1) / kUmmaK;
            nsteps = nsteps < 1 ? 1 : (nsteps > kBK / kUmmaK ? kBK / kUmmaK : nsteps);
            mma_kblock(false, tmem_base + kAccCol + (half + gh) * 128u, sah, sal, sb, sb + kItemPlane, idesc, !started[half],
                       &ring_empty[(pos + 2 + half) % kStages], c1, c2, c3, c4, c5, nsteps);
            started[half] = true;
          }
          pos += 2 + nhalf;
        }
        for (int half = 1; half < nhalf; ++half)
          for (int kb = 0; kb < L.kb_act; ++kb) act_item(half, kb);
        if (tr) p.trace[top * kTraceSlots + 2] = clock64();
        if (L.kb_act && need_acq) {
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
        for (int ch = sv ? 0 : sch; ch < (sv ? 2 : sch + 1); ++ch) {   // (a staged output is one event on channel sch)
          mbar_wait(&written[ch], nw[ch] & 1u);
          ++nw[ch];
          if (elect_one_sync()) {
            if (row_c < p.m) {
              if (sv) {
#pragma unroll
                for (int c = 2 * ch; c < 2 * ch + 2; ++c)
                  tma_store_2d(&maps.save[l], smem_base + kStgHi + c * kPlaneBytes, c * 64, (int)row_c);
              } else {
                const int nb = p.gepi[L.gepi].out_cols >> 6;
                for (int c = 0; c < nb; ++c) {
                  tma_store_2d(&maps.gout_hi, smem_base + kStgHi + (2 * sch + c) * kPlaneBytes, c * 64, (int)row_c);
                  tma_store_2d(&maps.gout_lo, smem_base + kStgLo + c * kPlaneBytes, c * 64, (int)row_c);
                }
              }
              tma_store_commit();
              tma_store_wait_read<0>();
              any = true;
            }
            mbar_arrive(&drained[ch]);
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
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t read_addr0 = map_to_cta(smem_u32(&acc_read[0]), 0);
    const uint32_t ready_addr0 = map_to_cta(smem_u32(&act_ready[0]), 0);
    const uint32_t stg_row_hi = smem_base + kStgHi + (uint32_t)(r_in_tile * 128);
    const uint32_t stg_row_lo = smem_base + kStgLo + (uint32_t)(r_in_tile * 128);
    const uint32_t swz = (uint32_t)(r_in_tile & 7);
    // bias / seed-vector staging: two buffers (op parity) x two column halves x 128 floats (this warp's 32 columns of
    // each K block).  The four warps of a column half write identical values, so no cross-warp barrier is needed:
    // every warp only relies on its own stores; a warp two ops ahead cannot exist (acc_read needs all 16 warps).
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
    // this thread's 32 columns (half h of K block c) of both planes -> the A planes in tensor memory
    auto store_tmem = [&](int c, const uint32_t* hi, const uint32_t* lo) {
      tmem_st16(tmem_base + kAHiCol + lane_sel + (uint32_t)(32 * c + 16 * h), hi);
      tmem_st16(tmem_base + kALoCol + lane_sel + (uint32_t)(32 * c + 16 * h), lo);
    };
    // ... -> the staging tile in shared memory (128B swizzle, read by the store warp's TMA stores); lo == nullptr: hi plane only
    auto store_stage = [&](int c_hi, int c_lo, const uint32_t* hi, const uint32_t* lo) {
      const uint32_t bh = stg_row_hi + (uint32_t)(c_hi * kPlaneBytes), bl = stg_row_lo + (uint32_t)(c_lo * kPlaneBytes);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t pos16 = ((uint32_t)(h * 4 + u) ^ swz) * 16u;
        sts128(bh + pos16, hi[4 * u], hi[4 * u + 1], hi[4 * u + 2], hi[4 * u + 3]);
        if (lo) sts128(bl + pos16, lo[4 * u], lo[4 * u + 1], lo[4 * u + 2], lo[4 * u + 3]);
      }
      fence_proxy_async();   // generic-proxy smem writes -> visible to the TMA stores
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
        store_tmem(c, hi, lo);
        tmem_st_wait();
        tc_fence_before();
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
        const uint32_t full_par = gemm_idx & 1u;
        ++gemm_idx;
        const uint32_t top = opcount - (uint32_t)p.whatif;
        const bool tr = p.trace && blockIdx.x == 0 && warp == 4 && lane == 0 && top < 64u;
        if (tr) p.trace[top * kTraceSlots + 3] = clock64();
        const uint32_t taddr = tmem_base + kAccCol + lane_sel;
        if (L.kind == 0) {
          uint32_t bits_out[4];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            // (the drain of the previous save does not depend on this op: waited for BEFORE the accumulator, so that the
            // passing wait's latency is not on the path from acc_full to the first hand-over)
            if (L.save) wait_drained(half);
            mbar_wait(&acc_full[half], full_par);
            tc_fence_after();
            if (half == 0 && tr) p.trace[top * kTraceSlots + 4] = clock64();
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
              if (cc == 1) {
                // both 64-column groups of this accumulator half are in registers: the next op may overwrite the half
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_addr(read_addr0 + 8u * half);
              }
              if (half == 0) mbar_wait(&blk_read[c], nhid & 1u);   // the second half's MMAs are done with the old A block c
              store_tmem(c, hi, lo);
              if (L.save) store_stage(c, 0, hi, nullptr);
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                // (the store warp's signal goes first: no warp can reach the next store event before every warp has
                // signalled this one, because the next op's accumulator needs this op's last block from all of them)
                if (cc == 1 && L.save) mbar_arrive(&written[half]);
                mbar_arrive_cluster_addr(ready_addr0 + 8u * c);
              }
              if (tr && c == 0) p.trace[top * kTraceSlots + 5] = clock64();
            }
          }
          ++nhid;
          if (L.save) save_outstanding[0] = save_outstanding[1] = true;
          if (MODE == 0 && L.save_bits && row_w < p.m) {
#pragma unroll
            for (int c = 0; c < 4; ++c) L.save_bits[bits_base + 32 * (2 * c + h)] = bits_out[c];
          }
        } else {
          // global op (n <= 128, one column half): columns [0, out_cols) leave in activation format (hi/lo) through the
          // staging tile + TMA stores, the f32 range through ordinary stores
          GemmEpilogue ge = p.gepi[L.gepi];
          const bool staged = ge.out.hi != nullptr;
          const int out_cols = staged ? ge.out_cols : 0;
          ge.out.hi = nullptr;
          const int gh = L.acc_half;
          mbar_wait(&acc_full[0], full_par);
          tc_fence_after();
          // a one-half op completes both accumulator-half barriers: the half it does not use is handed back as soon as the
          // op's MMAs are done (not earlier: the MMA warp acquires both barriers of the PREVIOUS op before this op's last K
          // block, so acc_full also tells that every warp's arrivals of the previous phase are in)
          if (lane == 0) mbar_arrive_cluster_addr(read_addr0 + 8u * (gh ^ 1));
          if (tr) p.trace[top * kTraceSlots + 4] = clock64();
          // this warp's (up to two) 32-column groups: read them out, hand the accumulator back, then process
          uint32_t rg[2][32];
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if ((h + 2 * j) * 32 < L.n) tmem_ld32(taddr + (uint32_t)(128 * gh + (h + 2 * j) * 32), rg[j]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_addr(read_addr0 + 8u * gh);
          // dgrad chains: the next row tile's seed only needs the A tile, which this op's MMAs have released
          if (MODE == 1 && l == p.num_ops - 1 && p.op[0].kind == 2 && tile + num_clusters < num_tiles) {
            do_seed(p.op[0], tile + num_clusters, bias_half + ((opcount + 1) & 1u) * 1024u);
            seed_early = true;
          }
          if (staged) wait_drained(sch);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int gi = h + 2 * j;
            const int col0 = gi * 32;
            if (col0 >= L.n) continue;
            const uint32_t* r = rg[j];
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
              store_stage(2 * sch + (col0 >> 6), col0 >> 6, hi, lo);   // (col0 >> 5) & 1 == h by construction
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
          __syncwarp();
          if (staged && lane == 0) mbar_arrive(&written[sch]);
          if (staged) save_outstanding[sch] = true;
        }
        if (tr) p.trace[top * kTraceSlots + 6] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc2(tmem_base, kTmemCols);
}

}  // namespace

int launch_chain_x3t(const ChainArgs& a, cudaStream_t st) {
  if (a.m <= 0) return RN_OK;
  if (a.act_f16 || a.w_f16) return rn_set_error(RN_ERR_UNSUPPORTED, "chain_x3t: bf16 planes only");
  if (a.num_ops < 1 || a.num_ops > kMaxOps) return rn_set_error(RN_ERR_ARG, "chain_x3t: 1..12 ops");
  if (a.in.hi && (!a.in.lo || a.in_cols % 64 || a.in_cols < 64 || a.in_cols > 256))
    return rn_set_error(RN_ERR_ARG, "chain_x3t: input tile must have both planes and 64..256 columns");
  if (a.in2.hi) return rn_set_error(RN_ERR_UNSUPPORTED, "chain_x3t: no second input");
  if (a.m + 256 > 0x7fffffffLL) return rn_set_error(RN_ERR_ARG, "chain_x3t: too many rows for one launch");
  X3Maps maps;
  PairParams p;
  memset(&p, 0, sizeof(p));
  memset(&maps, 0, sizeof(maps));
  int rc;
  int mode = -1;
  if ((rc = tc::make_map(&maps.in_hi, a.in.hi, a.m, a.in_valid, a.in.ld, kBM))) return rc;
  if ((rc = tc::make_map(&maps.in_lo, a.in.lo, a.m, a.in_valid, a.in.ld, kBM))) return rc;
  p.in2_sync_op = -1;
  static const int x3t_opt = getenv("RN_X3T_OPT") ? atoi(getenv("RN_X3T_OPT")) : 3;   // bit 0: global ops in accumulator half 1,
  p.split_order = (x3t_opt >> 1) & 1;                                                 // bit 1: staged outputs on store channel 1
  p.seed_scale = a.seed_scale;
  static const int trace_skip = getenv("RN_CHAIN_TRACE_SKIP") ? atoi(getenv("RN_CHAIN_TRACE_SKIP")) : 0;
  p.whatif = trace_skip;   // (chain_x3t.cu uses this field as the number of ops the debug timeline skips)
  p.num_ops = a.num_ops;
  p.in_kb = a.in.hi ? a.in_cols / kBK : 0;
  p.in_valid = a.in_valid > 0 ? a.in_valid : a.in_cols;
  p.m = a.m;
  bool awrite_pending = false;   // a hidden / seed op whose A tile no later op has read yet
  int staged_gepi = -1;
  for (int l = 0; l < a.num_ops; ++l) {
    const ChainOpArgs& L = a.op[l];
    PairOp& o = p.op[l];
    if (L.kind == 2) {
      if (!L.bias || !L.mask_bits || L.kb_act || L.kb_in || L.save_hi || awrite_pending || l != 0)
        return rn_set_error(RN_ERR_ARG, "chain_x3t: bad seed op (first op of a backward chain only)");
      if (mode < 0) mode = 1;
      if (mode != 1) return rn_set_error(RN_ERR_ARG, "chain_x3t: seed ops belong to backward chains");
      o.n = 256; o.kind = 2; o.bias = L.bias; o.mask_bits = L.mask_bits;
      awrite_pending = true;
      continue;
    }
    if (L.n % 16 || L.n < 16 || (L.kind == 0 && L.n != 256) || (L.kind != 0 && L.n > 128))
      return rn_set_error(RN_ERR_ARG, "chain_x3t: hidden ops are 256 wide, global ops at most 128");
    if ((L.kb_act != 0 && L.kb_act != 4) || (L.kb_in != 0 && L.kb_in != p.in_kb) || L.kb_act + L.kb_in == 0 || L.in2 ||
        (L.kb_in && !a.in.hi) || !L.w || !L.w_lo)
      return rn_set_error(RN_ERR_ARG, "chain_x3t: bad K structure / missing weight plane");
    if (l > 0 && !L.kb_act) return rn_set_error(RN_ERR_ARG, "chain_x3t: only the first op may read the chain input alone");
    if (awrite_pending && !L.kb_act)
      return rn_set_error(RN_ERR_ARG, "chain_x3t: the op after a hidden / seed op must read the activation tile");
    if (L.kb_act) awrite_pending = false;
    const int ktot = (L.kb_act + L.kb_in) * kBK;
    const int box_rows = L.n >= 256 ? 64 : L.n / 2;
    if ((rc = tc::make_map(&maps.w_hi[l], L.w, L.n, ktot, L.w_ld, box_rows))) return rc;
    if ((rc = tc::make_map(&maps.w_lo[l], L.w_lo, L.n, ktot, L.w_ld, box_rows))) return rc;
    o.n = L.n; o.kb_act = L.kb_act; o.kb_in = L.kb_in; o.in2 = 0;
    o.kind = L.kind; o.gepi = L.gepi; o.bias = L.bias;
    o.mask_bits = L.mask_bits; o.save_bits = L.save_bits;
    o.save = (L.kind == 0 && L.save_hi) ? 1 : 0;
    // a global op runs in accumulator half 1 unless the op before it is a global op that does
    o.acc_half = ((x3t_opt & 1) && L.kind == 1 && !(l > 0 && p.op[l - 1].kind == 1 && p.op[l - 1].acc_half == 1)) ? 1 : 0;
    if (o.save && (rc = tc::make_map(&maps.save[l], L.save_hi, a.m, 256, 256, kBM))) return rc;
    if (L.kind == 0) {
      if (mode < 0) mode = L.mode;
      if (L.mode != mode) return rn_set_error(RN_ERR_ARG, "chain_x3t: forward and backward hidden ops cannot be mixed");
      if (L.mode == 0 && !L.bias) return rn_set_error(RN_ERR_ARG, "chain_x3t: forward op without bias");
      if (L.mode == 1 && !L.mask_bits) return rn_set_error(RN_ERR_ARG, "chain_x3t: backward op without ReLU bits");
      awrite_pending = true;
    } else {
      if (L.gepi < 0 || L.gepi > 1) return rn_set_error(RN_ERR_ARG, "chain_x3t: bad epilogue index");
      const GemmEpilogue& e = a.gepi[L.gepi];
      if (e.relu || e.mask.hi) return rn_set_error(RN_ERR_UNSUPPORTED, "chain_x3t: global ops are linear");
      if (e.out.hi) {
        if (!e.out.lo || e.out_cols % 64 || e.out_cols < 64 || e.out_cols > L.n || (staged_gepi >= 0 && staged_gepi != L.gepi))
          return rn_set_error(RN_ERR_ARG, "chain_x3t: a staged output needs both planes, 64..128 columns, one epilogue");
        if (staged_gepi < 0) {
          if ((rc = tc::make_map(&maps.gout_hi, e.out.hi, a.m, e.out_cols, e.out.ld, kBM))) return rc;
          if ((rc = tc::make_map(&maps.gout_lo, e.out.lo, a.m, e.out_cols, e.out.ld, kBM))) return rc;
          staged_gepi = L.gepi;
        }
      }
    }
  }
  if (a.op[0].kind != 2 && a.op[0].kb_act) return rn_set_error(RN_ERR_ARG, "chain_x3t: the first op reads the chain input only");
  if (a.op[0].kind == 0 && a.op[0].kb_in != 2 && a.op[0].kb_in != 4)
    return rn_set_error(RN_ERR_ARG, "chain_x3t: a hidden first op takes a 128- or 256-column chain input");
  if (awrite_pending) return rn_set_error(RN_ERR_ARG, "chain_x3t: the last hidden op's result is never read");
  if (a.op[a.num_ops - 1].kind != 1) return rn_set_error(RN_ERR_ARG, "chain_x3t: the last op must be a global op");
  if (mode < 0) mode = 0;
  p.gepi[0] = a.gepi[0];
  p.gepi[1] = a.gepi[1];
  static bool smem_set = false;
  if (!smem_set) {
    if ((rc = tc::set_smem(chain_x3t_kernel<0>, kSmemTotal))) return rc;
    if ((rc = tc::set_smem(chain_x3t_kernel<1>, kSmemTotal))) return rc;
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
  double exec_flops = 0.0;   // three MMAs per executed K step (input K steps of pure padding are skipped)
  int in_steps = 0;
  for (int kb = 0; kb < p.in_kb; ++kb) {
    int ns = (p.in_valid - kb * kBK + kUmmaK - 1) / kUmmaK;
    in_steps += ns < 1 ? 1 : (ns > kBK / kUmmaK ? kBK / kUmmaK : ns);
  }
  for (int l = 0; l < a.num_ops; ++l)
    if (a.op[l].kind != 2)
      exec_flops += 3.0 * 2.0 * (double)a.m * a.op[l].n * (a.op[l].kb_act * kBK + (a.op[l].kb_in ? in_steps * kUmmaK : 0));
  rn_prof_begin(RN_PROF_CHAIN_TC, st, a.algo_flops, exec_flops);
  if (mode == 0)
    chain_x3t_kernel<0><<<grid, 384, kSmemTotal, st>>>(maps, p);
  else
    chain_x3t_kernel<1><<<grid, 384, kSmemTotal, st>>>(maps, p);
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
    printf("chain_x3t trace (mode %d, %d ops, m=%lld): per op: mma_begin first_item_ready mma_issued | epi_begin acc_full(half 0) blk0_handed epi_end | first item: A block + accumulator acquired (then the weight wait)  [cycles since first event]\n",
           mode, a.num_ops, (long long)a.m);
    for (int i = 0; i < 44; ++i) {
      printf("  op %2d:", i);
      for (int j = 0; j < 8; ++j) {
        if (j == 3 || j == 7) printf(" |");
        if (hbuf[i * kTraceSlots + j]) printf(" %8lld", hbuf[i * kTraceSlots + j] - t0); else printf("        -");
      }
      printf("\n");
    }
  }
  return RN_OK;
}

}  // namespace rn
