// Fused split-bf16 GEMM chain on CTA pairs (sm_100a, tcgen05.mma cta_group::2): the arithmetic that matches the
// reference's fp32 nn.Linear within BASELINE.json's tolerances (models.py:576-582, 689-700) at tensor-core speed.
//
// Every operand is a bf16 hi plane + a bf16 lo plane (x ~= hi + lo, 16-bit significand) and every K step issues
// three MMAs, hi*hi + lo*hi + hi*lo, into an fp32 TMEM accumulator.  Both planes of the running activation stay
// resident in shared memory (2 x 64 KB per CTA), so a cluster works on ONE 256-row tile (128 rows per CTA) and the
// overlap that chain_pair.cu gets from a second row tile comes from the accumulators instead:
//
//   * TMEM holds two 256-column accumulators used by alternate ops.  The epilogue of op l drains accumulator
//     l & 1 one 64-column K block at a time (bias + ReLU / ReLU bit mask, hi/lo split) and hands each finished K
//     block of the next A operand to the MMA warp through its own mbarrier (act_ready[c]); the MMAs of op l+1 on K
//     block c therefore start as soon as that block is written and run (3 x 512 cycles per K block) while the
//     epilogue warps finish the remaining blocks.  Only the first block's epilogue is exposed.
//   * the chain input of the first op is loaded by TMA straight into the (then dead) activation tile as soon as the
//     previous tile's last MMA has retired, after an L2 prefetch issued one tile ahead; the skip layer's input K
//     blocks stream through the ring, interleaved with its activation K blocks so that their load latency hides
//     behind the MMAs of a resident block.
//   * weight K blocks [N/2 rows per CTA x 64] x {hi, lo} stream through a 3-stage TMA ring (32 KB stages).
//   * saves for the wgrad kernels (hi plane only: wgrad runs on bf16 operands) and activation-format outputs of
//     global ops (the bottleneck -> v0 hi/lo) leave as TMA stores issued by a dedicated store warp straight from
//     the swizzled activation tile; ReLU masks travel as 1 bit per activation in a warp-coalesced layout.
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

constexpr int kStages = 3;
constexpr int kPlaneBytes = 16384;             // [128 x 64] bf16, 128B swizzle
constexpr int kStageBytes = 2 * kPlaneBytes;   // hi | lo
constexpr int kActPlane = 4 * kPlaneBytes;     // 64 KB: one plane of the activation tile (4 K blocks)
constexpr int kSmemRing = 2 * kActPlane;       // act_hi | act_lo | ring
constexpr int kSmemBars = kSmemRing + kStages * kStageBytes;
constexpr int kSmemBias = kSmemBars + 256;     // [op parity][column half h][4 K blocks x 32 floats]
constexpr int kSmemTotal = kSmemBias + 2048;
static_assert(kSmemTotal <= 232448, "shared memory budget");

struct X3Maps {
  CUtensorMap in_hi, in_lo;       // chain input [m, in_cols], box [128 x 64]
  CUtensorMap gout_hi, gout_lo;   // activation-format output of a global op, box [128 x 64]
  CUtensorMap w_hi[kMaxOps], w_lo[kMaxOps];
  CUtensorMap save[kMaxOps];      // hi-plane saves of hidden ops, box [128 x 64]
};

// MODE 0: forward chain (hidden ops: bias + ReLU); MODE 1: dgrad chain (hidden ops: ReLU bit mask; seed ops)
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
chain_x3_kernel(const __grid_constant__ X3Maps maps, const __grid_constant__ PairParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBars);
  uint64_t* ring_full = bars;          // [3] leader's are used (TMA bytes of both CTAs land there)
  uint64_t* ring_empty = bars + 4;     // [3] per CTA, multicast MMA commit
  uint64_t* acc_full = bars + 8;       // [2] per CTA, multicast MMA commit
  uint64_t* acc_free = bars + 10;      // [2] leader's: 16 arrivals (8 epilogue warps x 2 CTAs)
  uint64_t* act_ready = bars + 12;     // [4] leader's: 16 arrivals: K block c of the activation tile is written
  uint64_t* written = bars + 16;       // per CTA: 8 arrivals: the tile part a store needs is in shared memory
  uint64_t* drained = bars + 17;       // per CTA: the store warp's TMA stores have read the tile
  uint64_t* in_full = bars + 18;       // leader's: the first op's chain input has landed in the activation tile (tx bytes)
  uint64_t* saves_drained = bars + 19; // per CTA: the last save of the tile has been read (the next tile's input may land)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_free[i], 16);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&act_ready[i], 16);
    mbar_init(written, 8);
    mbar_init(drained, 1);
    mbar_init(in_full, 1);
    mbar_init(saves_drained, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: this CTA's half of every weight K block (hi, lo) and its rows of the input K blocks =====
    uint32_t pos = 0;
    const uint32_t ring_full_leader = map_to_cta(smem_u32(&ring_full[0]), 0);
    auto load2 = [&](uint32_t total_bytes, const CUtensorMap* mh, const CUtensorMap* ml, int c0, int c1) {
      const uint32_t s = pos % kStages, ph = (pos / kStages) & 1u;
      mbar_wait(&ring_empty[s], ph ^ 1u);
      ++pos;
      if (elect_one_sync()) {
        if (rank == 0) mbar_arrive_expect_tx(&ring_full[s], total_bytes);
        const uint32_t dst = smem_base + kSmemRing + s * kStageBytes;
        tma_load_2d_pair(dst, mh, ring_full_leader + 8u * s, c0, c1);
        tma_load_2d_pair(dst + kPlaneBytes, ml, ring_full_leader + 8u * s, c0, c1);
      }
      __syncwarp();
    };
    const uint32_t in_full_leader = map_to_cta(smem_u32(in_full), 0);
    uint32_t tile_iter = 0;
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tile_iter) {
      const int row0 = (int)(tile * 256 + (int64_t)rank * 128);
      if (p.in_kb && tile + num_clusters < num_tiles && elect_one_sync()) {
        // next tile's chain input -> L2 (its TMA load is issued at the tile boundary and sits on the critical path)
        const int rown = (int)((tile + num_clusters) * 256 + (int64_t)rank * 128);
        for (int kb = 0; kb < p.in_kb; ++kb) {
          tma_prefetch_l2_2d(&maps.in_hi, kb * kBK, rown);
          tma_prefetch_l2_2d(&maps.in_lo, kb * kBK, rown);
        }
      }
      __syncwarp();
      for (int l = 0; l < p.num_ops; ++l) {
        const PairOp& L = p.op[l];
        if (L.kind == 2) continue;
        const int nh = L.n >> 1;
        const uint32_t wbytes = (uint32_t)L.n * 256u;   // 2 planes x 2 CTAs x (n/2 rows x 128 B)
        if (l == 0 && direct_in) {
          // first weight block(s) first, then the input tile (it has to wait for the previous tile), then the rest
          const int early = L.kb_in < 2 ? L.kb_in : 2;
          for (int kb = 0; kb < early; ++kb) load2(wbytes, &maps.w_hi[l], &maps.w_lo[l], kb * kBK, (int)rank * nh);
          if (tile_iter) {
            // every MMA of the previous tile has retired (release of its last ring item) and its saves have been read
            mbar_wait(&ring_empty[(pos - early - 1) % kStages], ((pos - early - 1) / kStages) & 1u);
            if (last_save >= 0) mbar_wait(saves_drained, (tile_iter - 1) & 1u);
          }
          if (elect_one_sync()) {
            if (rank == 0) mbar_arrive_expect_tx(in_full, (uint32_t)L.kb_in * 4u * kPlaneBytes);
            for (int kb = 0; kb < L.kb_in; ++kb) {
              tma_load_2d_pair(smem_base + (in_blk0 + kb) * kPlaneBytes, &maps.in_hi, in_full_leader, kb * kBK, row0);
              tma_load_2d_pair(smem_base + kActPlane + (in_blk0 + kb) * kPlaneBytes, &maps.in_lo, in_full_leader, kb * kBK, row0);
            }
          }
          __syncwarp();
          for (int kb = early; kb < L.kb_in; ++kb) load2(wbytes, &maps.w_hi[l], &maps.w_lo[l], kb * kBK, (int)rank * nh);
          continue;
        }
        // skip layers: input K block i, then activation K block i (the MMAs of the resident block cover the loads)
        const int nmax = L.kb_in > L.kb_act ? L.kb_in : L.kb_act;
        for (int kb = 0; kb < nmax; ++kb) {
          if (kb < L.kb_in) {
            load2(4u * kPlaneBytes, &maps.in_hi, &maps.in_lo, kb * kBK, row0);
            load2(wbytes, &maps.w_hi[l], &maps.w_lo[l], (L.kb_act + kb) * kBK, (int)rank * nh);
          }
          if (kb < L.kb_act) load2(wbytes, &maps.w_hi[l], &maps.w_lo[l], kb * kBK, (int)rank * nh);
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
    // the 12 MMAs of one K block (hi*hi + lo*hi + hi*lo per K step), then the commits that depend on them
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
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tile_iter) {
      for (int l = 0; l < p.num_ops; ++l, ++opcount) {
        const PairOp& L = p.op[l];
        if (L.kind == 2) {   // seed op: the epilogue warps generate the activation tile, no MMA
          need_acq = true;
          continue;
        }
        const uint32_t idesc = make_idesc2(L.n);
        const uint32_t buf = gemm_idx & 1u, use = gemm_idx >> 1;
        mbar_wait_cluster(&acc_free[buf], (use & 1u) ^ 1u);   // epilogue of the op two before has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * 256u;
        const bool tr = p.trace && blockIdx.x == 0 && opcount < 64 && lane == 0;
        if (tr) p.trace[opcount * 8 + 0] = clock64();
        bool first = true;
        const int nkb = L.kb_in + L.kb_act;
        int done = 0;
        if (l == 0 && direct_in) {
          // the chain input sits in activation K blocks [in_blk0, 4): only the weights come through the ring
          mbar_wait_cluster(in_full, tile_iter & 1u);
          for (int kb = 0; kb < L.kb_in; ++kb) {
            wait_full(pos);
            tc_fence_after();
            if (tr && kb == 0) p.trace[opcount * 8 + 1] = clock64();
            ++done;
            const uint32_t sb = stage_addr(pos);
            mma_kblock(tmem_d, smem_base + (in_blk0 + kb) * kPlaneBytes, smem_base + kActPlane + (in_blk0 + kb) * kPlaneBytes, sb,
                       sb + kPlaneBytes, idesc, first, &ring_empty[pos % kStages], done == nkb ? &acc_full[buf] : nullptr, nullptr);
            ++pos;
            first = false;
          }
        } else {
          const int nmax = L.kb_in > L.kb_act ? L.kb_in : L.kb_act;
          for (int kb = 0; kb < nmax; ++kb) {
            if (kb < L.kb_in) {
              wait_full(pos);
              wait_full(pos + 1);
              tc_fence_after();
              ++done;
              const uint32_t sa = stage_addr(pos), sb = stage_addr(pos + 1);
              mma_kblock(tmem_d, sa, sa + kPlaneBytes, sb, sb + kPlaneBytes, idesc, first, &ring_empty[pos % kStages],
                         &ring_empty[(pos + 1) % kStages], done == nkb ? &acc_full[buf] : nullptr);
              pos += 2;
              first = false;
            }
            if (kb < L.kb_act) {
              if (need_acq) mbar_wait_cluster(&act_ready[kb], aver & 1u);
              wait_full(pos);
              tc_fence_after();
              if (tr && kb == 0) p.trace[opcount * 8 + 1] = clock64();
              ++done;
              const uint32_t sb = stage_addr(pos);
              mma_kblock(tmem_d, smem_base + kb * kPlaneBytes, smem_base + kActPlane + kb * kPlaneBytes, sb, sb + kPlaneBytes,
                         idesc, first, &ring_empty[pos % kStages], done == nkb ? &acc_full[buf] : nullptr, nullptr);
              ++pos;
              first = false;
            }
          }
        }
        if (tr) p.trace[opcount * 8 + 2] = clock64();
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
    uint32_t nw = 0;
    bool any = false;
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int64_t row_c = tile * 256 + (int64_t)rank * 128;
      for (int l = 0; l < p.num_ops; ++l) {
        const PairOp& L = p.op[l];
        const bool sv = L.kind == 0 && L.save;
        const bool staged = L.kind == 1 && p.gepi[L.gepi].out.hi != nullptr;
        if (!sv && !staged) continue;
        mbar_wait(written, nw & 1u);
        ++nw;
        if (elect_one_sync()) {
          if (row_c < p.m) {
            if (sv) {
#pragma unroll
              for (int c = 0; c < 4; ++c) tma_store_2d(&maps.save[l], smem_base + c * kPlaneBytes, c * 64, (int)row_c);
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
          mbar_arrive(drained);
          if (l == last_save) mbar_arrive(saves_drained);
        }
        __syncwarp();
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
    uint32_t gemm_idx = 0, ndrain = 0, opcount = 0;
    bool save_outstanding = false;
    auto wait_drained = [&]() {
      if (save_outstanding) {
        mbar_wait(drained, ndrain & 1u);
        ++ndrain;
        save_outstanding = false;
      }
    };
    // write this thread's 32 columns (half h of K block c) of both planes, then hand the block over
    auto store_block = [&](int c, const uint32_t* hi, const uint32_t* lo) {
      const uint32_t bh = act_row_hi + (uint32_t)(c * kPlaneBytes), bl = act_row_lo + (uint32_t)(c * kPlaneBytes);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t pos16 = ((uint32_t)(h * 4 + u) ^ swz) * 16u;
        sts128(bh + pos16, hi[4 * u], hi[4 * u + 1], hi[4 * u + 2], hi[4 * u + 3]);
        sts128(bl + pos16, lo[4 * u], lo[4 * u + 1], lo[4 * u + 2], lo[4 * u + 3]);
      }
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core and to TMA stores
    };
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int64_t row_w = tile * 256 + (int64_t)rank * 128 + q * 32;   // first row of this warp
      const int64_t row = row_w + lane;
      const bool row_ok = row < p.m;
      const size_t bits_base = (size_t)(row_w >> 5) * 256 + lane;        // word w of the 32-row group at + 32 w
      for (int l = 0; l < p.num_ops; ++l, ++opcount) {
        const PairOp& L = p.op[l];
        const bool seed = MODE == 1 && L.kind == 2;
        const float* bias_ptr = L.kind == 1 ? p.gepi[L.gepi].bias : ((MODE == 0 || seed) ? L.bias : nullptr);
        const uint32_t bias_buf = bias_half + (opcount & 1u) * 1024u;
        if (bias_ptr) {
          const int col = 64 * (lane >> 3) + 32 * h + 4 * (lane & 7);
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col < L.n) bv = __ldg(reinterpret_cast<const float4*>(bias_ptr + col));
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(bias_buf + 16u * lane), "f"(bv.x), "f"(bv.y), "f"(bv.z),
                       "f"(bv.w)
                       : "memory");
          __syncwarp();
        }
        uint32_t bw[4] = {0u, 0u, 0u, 0u};
        if (MODE == 1 && (L.kind == 0 || seed) && row_w < p.m) {
#pragma unroll
          for (int c = 0; c < 4; ++c) bw[c] = __ldg(L.mask_bits + bits_base + 32 * (2 * c + h));
        }
        if (seed) {
          // seed tile of a dgrad chain: vec[col] where the ReLU bit is set, generated in place (no accumulator)
          wait_drained();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = lds128f(bias_buf + (uint32_t)(c * 128 + 16 * i));
              const float v0 = (bw[c] >> (2 * i) & 1u) ? b.x * p.seed_scale : 0.f;
              const float v1 = (bw[c] >> (16 + 2 * i) & 1u) ? b.y * p.seed_scale : 0.f;
              const float v2 = (bw[c] >> (2 * i + 1) & 1u) ? b.z * p.seed_scale : 0.f;
              const float v3 = (bw[c] >> (17 + 2 * i) & 1u) ? b.w * p.seed_scale : 0.f;
              split2(v0, v1, hi[2 * i], lo[2 * i]);
              split2(v2, v3, hi[2 * i + 1], lo[2 * i + 1]);
            }
            store_block(c, hi, lo);
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_addr(ready_addr0 + 8u * c);
          }
          continue;
        }
        const uint32_t buf = gemm_idx & 1u, use = gemm_idx >> 1;
        ++gemm_idx;
        const bool tr = p.trace && blockIdx.x == 0 && warp == 4 && lane == 0 && opcount < 64;
        if (tr) p.trace[opcount * 8 + 3] = clock64();
        mbar_wait(&acc_full[buf], use & 1u);
        tc_fence_after();
        if (tr) p.trace[opcount * 8 + 4] = clock64();
        const uint32_t taddr = tmem_base + buf * 256u + ((uint32_t)(q * 32) << 16);
        if (L.kind == 0) {
          wait_drained();
          uint32_t bits_out[4];
          uint32_t ra[32], rb[32];
          tmem_ld32(taddr + (uint32_t)(32 * h), ra);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t* cur = (c & 1) ? rb : ra;
            uint32_t* nxt = (c & 1) ? ra : rb;
            tmem_ld_wait();
            if (c < 3) tmem_ld32(taddr + (uint32_t)(64 * (c + 1) + 32 * h), nxt);
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
            store_block(c, hi, lo);
            __syncwarp();
            if (lane == 0) {
              // (the store warp's signal goes first: no warp can reach the next store event before every warp has
              // signalled this one, because the next op's accumulator needs this op's last block from all of them)
              if (c == 3 && L.save) mbar_arrive(written);
              mbar_arrive_cluster_addr(ready_addr0 + 8u * c);
            }
            if (tr && c == 0) p.trace[opcount * 8 + 5] = clock64();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_addr(free_addr0 + 8u * buf);
          if (L.save) save_outstanding = true;
          if (MODE == 0 && L.save_bits && row_w < p.m) {
#pragma unroll
            for (int c = 0; c < 4; ++c) L.save_bits[bits_base + 32 * (2 * c + h)] = bits_out[c];
          }
        } else {
          // global op: columns [0, out_cols) leave in activation format (hi/lo) through the activation tile + TMA
          // stores, the f32 range through ordinary stores
          GemmEpilogue ge = p.gepi[L.gepi];
          const bool staged = ge.out.hi != nullptr;
          const int out_cols = staged ? ge.out_cols : 0;
          ge.out.hi = nullptr;
          if (staged) wait_drained();
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
            mbar_arrive_cluster_addr(free_addr0 + 8u * buf);
            if (staged) mbar_arrive(written);
          }
          if (staged) save_outstanding = true;
        }
        if (tr) p.trace[opcount * 8 + 6] = clock64();
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
      if (!L.bias || !L.mask_bits || L.kb_act || L.kb_in || L.save_hi || awrite_pending)
        return rn_set_error(RN_ERR_ARG, "chain_x3: bad seed op");
      if (mode < 0) mode = 1;
      if (mode != 1) return rn_set_error(RN_ERR_ARG, "chain_x3: seed ops belong to backward chains");
      o.n = 256; o.kind = 2; o.bias = L.bias; o.mask_bits = L.mask_bits;
      awrite_pending = true;
      continue;
    }
    if (L.n % 16 || L.n < 16 || L.n > 256 || (L.kind == 0 && L.n != 256)) return rn_set_error(RN_ERR_ARG, "chain_x3: bad op width");
    if ((L.kb_act != 0 && L.kb_act != 4) || (L.kb_in != 0 && L.kb_in != p.in_kb) || L.kb_act + L.kb_in == 0 || L.in2 ||
        (L.kb_in && !a.in.hi) || !L.w || !L.w_lo)
      return rn_set_error(RN_ERR_ARG, "chain_x3: bad K structure / missing weight plane");
    if (awrite_pending && !L.kb_act)
      return rn_set_error(RN_ERR_ARG, "chain_x3: the op after a hidden / seed op must read the activation tile");
    if (L.kb_act) awrite_pending = false;
    const int ktot = (L.kb_act + L.kb_in) * kBK;
    if ((rc = tc::make_map(&maps.w_hi[l], L.w, L.n, ktot, L.w_ld, L.n / 2))) return rc;
    if ((rc = tc::make_map(&maps.w_lo[l], L.w_lo, L.n, ktot, L.w_ld, L.n / 2))) return rc;
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
        if (!e.out.lo || e.out_cols % 64 || e.out_cols < 64 || e.out_cols > 256 || e.out_cols > L.n || (staged_gepi >= 0 && staged_gepi != L.gepi))
          return rn_set_error(RN_ERR_ARG, "chain_x3: a staged output needs both planes, 64..256 columns, one epilogue");
        if (staged_gepi < 0) {
          if ((rc = tc::make_map(&maps.gout_hi, e.out.hi, a.m, e.out_cols, e.out.ld, kBM))) return rc;
          if ((rc = tc::make_map(&maps.gout_lo, e.out.lo, a.m, e.out_cols, e.out.ld, kBM))) return rc;
          staged_gepi = L.gepi;
        }
      }
    }
  }
  // across tiles the chain repeats: the first op must not read a tile its own launch has not produced
  if (a.op[0].kind != 2 && a.op[0].kb_act) return rn_set_error(RN_ERR_ARG, "chain_x3: the first op reads the chain input only");
  if (awrite_pending) return rn_set_error(RN_ERR_ARG, "chain_x3: the last hidden op's result is never read");
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
    if (!trace_buf) cudaMalloc(&trace_buf, 64 * 8 * sizeof(long long));
    cudaMemsetAsync(trace_buf, 0, 64 * 8 * sizeof(long long), st);
    p.trace = trace_buf;
  }
  rn_prof_begin(RN_PROF_CHAIN_TC, st, a.algo_flops);
  if (mode == 0)
    chain_x3_kernel<0><<<grid, 384, kSmemTotal, st>>>(maps, p);
  else
    chain_x3_kernel<1><<<grid, 384, kSmemTotal, st>>>(maps, p);
  rn_prof_end(RN_PROF_CHAIN_TC, st);
  RN_CUDA_CHECK_LAUNCH();
  if (p.trace) {
    --trace_left;
    cudaStreamSynchronize(st);
    static long long hbuf[64 * 8];
    cudaMemcpy(hbuf, trace_buf, sizeof(hbuf), cudaMemcpyDeviceToHost);
    const long long t0 = hbuf[0];
    printf("chain_x3 trace (mode %d, %d ops, m=%lld): per op: mma_acc_free act_blk0_ready mma_issued | epi_wait_begin acc_full blk0_handed epi_end  [cycles since first]\n",
           mode, a.num_ops, (long long)a.m);
    for (int i = 0; i < 40; ++i)
      printf("  op %2d: %8lld %8lld %8lld | %8lld %8lld %8lld %8lld\n", i, hbuf[i * 8] - t0, hbuf[i * 8 + 1] - t0, hbuf[i * 8 + 2] - t0,
             hbuf[i * 8 + 3] - t0, hbuf[i * 8 + 4] - t0, hbuf[i * 8 + 5] - t0, hbuf[i * 8 + 6] - t0);
  }
  return RN_OK;
}

}  // namespace rn
