// Fused GEMM chain on CTA pairs (sm_100a, tcgen05.mma cta_group::2).
//
// One cluster of two CTAs (one per SM of a TPC) walks "super tiles" of 512 rows = 2 row tiles (A, B) x 256
// rows, each CTA holding 128 rows of each tile.  Per chain op (one Linear layer of the NerfMLP, forward or
// dgrad) the pair issues UMMA M=256 x N<=256 x K=16 instructions: every CTA supplies its 128 rows of the
// A operand (the running activation, resident in its shared memory) and HALF of the weight rows (N/2),
// so each SM stages only half of every weight K block and the pair shares it.
//
//   * ping-pong: the two row tiles alternate on the tensor pipe.  While the MMAs of tile B run, the eight
//     epilogue warps of both CTAs drain tile A's TMEM accumulator (bias + ReLU / ReLU mask, bf16) and
//     write it back as the next op's A operand (128B-swizzled, K-major), and vice versa; TMEM holds
//     exactly the two 256-column fp32 accumulators.
//   * weight ring: 16 KB stages; the K blocks of an op stay resident while tile A and then tile B consume
//     them, so one L2 -> SMEM weight transfer serves 512 rows.  Input-tile K blocks (layer 0 and the skip
//     layer) stream through the same ring just in time.
//   * saves for the wgrad kernels (training) leave as TMA stores straight from the swizzled activation
//     tile; ReLU masks travel as 1 bit per activation (32 B per row and layer), written by the forward chain and
//     prefetched one tile ahead by the dgrad chains; eval touches HBM only for the chain input and the head outputs.
//
// Roles per CTA (384 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warp 2 TMEM
// allocator, warp 3 store warp (TMA saves), warps 4-11 epilogue (warp % 4 = TMEM lane quadrant, (warp - 4) / 4 = column half).
#include <stdio.h>
#include <stdlib.h>

#include "chain_common.cuh"

namespace rn {
namespace {
using namespace tc;
using namespace chain;

constexpr int kRingStages = 6;
constexpr int kStageBytes = 16384;            // one ring item: [128 x 64] bf16
constexpr int kBlkBytes = kBM * kBK * 2;      // 16 KB activation K block
constexpr int kActBytes = 4 * kBlkBytes;      // 64 KB activation tile
constexpr int kSmemRing = 2 * kActBytes;
constexpr int kSmemBars = kSmemRing + kRingStages * kStageBytes;
constexpr int kSmemBias = kSmemBars + 256;
constexpr int kSmemTotal = kSmemBias + 2 * 1024;
static_assert(kSmemTotal <= 232448, "shared memory budget");

struct PairMaps {
  CUtensorMap in;
  CUtensorMap in2;
  CUtensorMap w[kMaxOps];
  CUtensorMap w2[kMaxOps];   // lo plane of split-bf16 weights (w_planes = 2)
  CUtensorMap save[kMaxOps];
};

// MODE 0: forward chain (hidden ops: bias + ReLU); MODE 1: backward / dgrad chain (hidden ops: ReLU bit mask)
// F16: format of the activation tile and of the activation-format outputs (0 bf16, 1 fp16)
template <int MODE, int F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
chain_pair_kernel(const __grid_constant__ PairMaps maps, const __grid_constant__ PairParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBars);
  uint64_t* ring_full = bars;          // [8]  leader's are used (TMA bytes of both CTAs land there)
  uint64_t* ring_empty = bars + 8;     // [8]  per CTA, signalled by the multicast MMA commit
  uint64_t* acc_full = bars + 16;      // [2]  per CTA, multicast MMA commit
  uint64_t* acc_free = bars + 18;      // [2]  leader's: 16 arrivals (8 epilogue warps x 2 CTAs)
  uint64_t* in2_ready = bars + 20;     // per CTA: 8 arrivals (epilogue warps: their saves of this super tile are complete)
  uint64_t* seed_done = bars + 22;     // [2]  leader's: 16 arrivals (seed op: activation tile generated).  A separate
                                       //      barrier: a warp's seed arrival must not be able to land in the acc_free
                                       //      phase of the op before it (no wait separates the two arrivals)
  uint64_t* save_ready = bars + 24;    // [2]  per CTA: 8 arrivals (epilogue warps: their part of act[t] is written and fenced)
  uint64_t* save_drained = bars + 26;  // [2]  per CTA: 1 arrival (store warp: the TMA saves of act[t] have read shared memory)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t num_super = (p.m + 511) / 512;
  const int64_t cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const uint32_t smem_base = smem_u32(smem);

  if (threadIdx.x == 0) {
    if (smem_base & 1023u) __trap();
    tma_prefetch_desc(&maps.in);
    for (int i = 0; i < p.num_ops; ++i) tma_prefetch_desc(&maps.w[i]);
    for (int i = 0; i < kRingStages; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_free[i], 16);
    }
    mbar_init(in2_ready, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&save_ready[i], 8);
      mbar_init(&save_drained[i], 1);
    }
    mbar_init(&seed_done[0], 16);
    mbar_init(&seed_done[1], 16);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (whole warp, one elected lane issues): this CTA's half of every weight K block and its
    // rows of the input K blocks =====
    uint32_t pos = 0;
    const uint32_t ring_full_leader = map_to_cta(smem_u32(&ring_full[0]), 0);
    auto acquire = [&](uint32_t total_bytes) -> uint32_t {
      const uint32_t s = pos % kRingStages, ph = (pos / kRingStages) & 1u;
      mbar_wait(&ring_empty[s], ph ^ 1u);
      ++pos;
      return s;
    };
    auto load = [&](uint32_t s, uint32_t total_bytes, const CUtensorMap* map, int c0, int c1) {
      if (elect_one_sync()) {
        if (rank == 0) mbar_arrive_expect_tx(&ring_full[s], total_bytes);
        tma_load_2d_pair(smem_base + kSmemRing + s * kStageBytes, map, ring_full_leader + 8u * s, c0, c1);
      }
      __syncwarp();
    };
    uint32_t st_iter = 0;
    for (int64_t st = cluster_id; st < num_super; st += num_clusters, ++st_iter) {
      for (int l = 0; l < p.num_ops; ++l) {
        const PairOp& L = p.op[l];
        const int nh = L.n >> 1;
        for (int kb = 0; kb < L.kb_act; ++kb) {
          const uint32_t s = acquire(0);
          load(s, (uint32_t)L.n * 128u, &maps.w[l], kb * kBK, (int)rank * nh);
          if (p.w_planes == 2) {
            const uint32_t s2 = acquire(0);
            load(s2, (uint32_t)L.n * 128u, &maps.w2[l], kb * kBK, (int)rank * nh);
          }
        }
        // a second input that is a save of this launch: its TMA stores must have completed (see the epilogue)
        if (L.kb_in && L.in2 && p.in2_sync_op >= 0) mbar_wait(in2_ready, st_iter & 1u);
        for (int kb = 0; kb < L.kb_in; ++kb) {
          for (int t = 0; t < 2; ++t) {
            const uint32_t s = acquire(0);
            const int64_t row0 = st * 512 + t * 256 + (int64_t)rank * 128;
            load(s, 2u * kStageBytes, L.in2 ? &maps.in2 : &maps.in, kb * kBK, (int)row0);
            if (t == 0) {
              const uint32_t sw = acquire(0);
              load(sw, (uint32_t)L.n * 128u, &maps.w[l], (L.kb_act + kb) * kBK, (int)rank * nh);
              if (p.w_planes == 2) {
                const uint32_t sw2 = acquire(0);
                load(sw2, (uint32_t)L.n * 128u, &maps.w2[l], (L.kb_act + kb) * kBK, (int)rank * nh);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader CTA; whole warp, one elected lane issues) =====
    uint32_t pos = 0;       // ring position, mirrors the producer's item order
    uint32_t opcount = 0;   // ops walked so far (trace index)
    uint32_t nreal = 0;     // GEMM ops issued so far on each tile (acc_free phases)
    uint32_t nseed = 0;     // seed ops passed so far (seed_done phases)
    bool seed_pending = false;
    auto wait_full = [&](uint32_t q) {
      mbar_wait(&ring_full[q % kRingStages], (q / kRingStages) & 1u);
    };
    auto stage_addr = [&](uint32_t q) -> uint32_t { return smem_base + kSmemRing + (q % kRingStages) * kStageBytes; };
    // 4 MMAs of one K block, then the commits that depend on them; issued by one elected lane
    auto mma_kblock = [&](uint32_t tmem_d, uint32_t sa, uint32_t sb, uint32_t idesc, bool first, uint64_t* commit0,
                          uint64_t* commit1, uint64_t* commit2) {
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < kBK / kUmmaK; ++kk) {
          const uint32_t koff = kk * kUmmaK * 2;
          umma2_bf16(tmem_d, make_desc(sa + koff, 16, 1024), make_desc(sb + koff, 16, 1024), idesc, (!first || kk) ? 1u : 0u);
        }
        if (commit0) umma2_commit(commit0);
        if (commit1) umma2_commit(commit1);
        if (commit2) umma2_commit(commit2);
      }
      __syncwarp();
    };
    // The op descriptor of op l+1 is fetched while op l is issued (as in chain_x3t.cu): read at the op boundary and inside
    // the K loops, the dependent constant-bank loads (run-time op index) sit on the MMA warp's critical path.
    struct OpLite { int n, kb_act, kb_in, kind; };
    auto fetch = [&](int l) -> OpLite {
      const PairOp& o = p.op[l];
      return OpLite{o.n, o.kb_act, o.kb_in, o.kind};
    };
    OpLite nxt = fetch(0);
    const int w_planes = p.w_planes, a_f16 = p.a_f16, b_f16 = p.b_f16;
    for (int64_t st = cluster_id; st < num_super; st += num_clusters) {
      for (int l = 0; l < p.num_ops; ++l, ++opcount) {
        const OpLite L = nxt;
        nxt = fetch(l + 1 < p.num_ops ? l + 1 : 0);
        if (L.kind == 2) {           // seed op: the epilogue warps generate the activation tile, no MMA
          ++nseed;
          seed_pending = true;
          continue;
        }
        const uint32_t idesc = make_idesc2(L.n, a_f16, b_f16);
        const uint32_t free_parity = (nreal & 1u) ^ 1u;   // epilogue of the previous GEMM op on this tile done
        ++nreal;
        if (w_planes == 2) {
          // split-bf16 weights against one-plane activations: every K block is a (hi, lo) pair of ring items and
          // both row tiles consume it before the next one (eight resident items would not fit the ring), so the two
          // accumulators fill side by side: dY * W_hi + dY * W_lo
          for (int kb = 0; kb < L.kb_act; ++kb) {
            wait_full(pos);
            wait_full(pos + 1);
            tc_fence_after();
            for (int t = 0; t < 2; ++t) {
              if (kb == 0) {
                mbar_wait_cluster(&acc_free[t], free_parity);
                if (seed_pending) mbar_wait_cluster(&seed_done[t], (nseed - 1u) & 1u);
                tc_fence_after();
              }
              const uint32_t tmem_d = tmem_base + (uint32_t)t * 256u;
              const uint32_t sa = smem_base + t * kActBytes + kb * kBlkBytes;
              mma_kblock(tmem_d, sa, stage_addr(pos), idesc, kb == 0, t == 1 ? &ring_empty[pos % kRingStages] : nullptr, nullptr,
                         nullptr);
              mma_kblock(tmem_d, sa, stage_addr(pos + 1), idesc, false, t == 1 ? &ring_empty[(pos + 1) % kRingStages] : nullptr,
                         (kb == L.kb_act - 1 && !L.kb_in) ? &acc_full[t] : nullptr, nullptr);
            }
            pos += 2;
          }
          if (L.kb_act) seed_pending = false;
          for (int kb = 0; kb < L.kb_in; ++kb) {
            const uint32_t px0 = pos, pwh = pos + 1, pwl = pos + 2, px1 = pos + 3;
            pos += 4;
            for (int t = 0; t < 2; ++t) {
              if (!L.kb_act && kb == 0) {
                mbar_wait_cluster(&acc_free[t], free_parity);
                tc_fence_after();
              }
              const uint32_t px = t ? px1 : px0;
              wait_full(px);
              if (t == 0) {
                wait_full(pwh);
                wait_full(pwl);
              }
              tc_fence_after();
              const uint32_t tmem_d = tmem_base + (uint32_t)t * 256u;
              mma_kblock(tmem_d, stage_addr(px), stage_addr(pwh), idesc, !L.kb_act && kb == 0,
                         t == 1 ? &ring_empty[pwh % kRingStages] : nullptr, nullptr, nullptr);
              mma_kblock(tmem_d, stage_addr(px), stage_addr(pwl), idesc, false, &ring_empty[px % kRingStages],
                         t == 1 ? &ring_empty[pwl % kRingStages] : nullptr, kb == L.kb_in - 1 ? &acc_full[t] : nullptr);
            }
          }
          continue;
        }
        // --- K blocks from the resident activation tiles: tile 0 then tile 1 over the same weight stages
        const uint32_t pos_w = pos;
        for (int t = 0; t < 2; ++t) {
          if (L.kb_act) {
            mbar_wait_cluster(&acc_free[t], free_parity);
            if (seed_pending) mbar_wait_cluster(&seed_done[t], (nseed - 1u) & 1u);
            tc_fence_after();
          }
          if (p.trace && blockIdx.x == 0 && opcount < 64 && lane == 0) p.trace[(opcount * 2 + t) * 8 + 0] = clock64();
          const uint32_t tmem_d = tmem_base + (uint32_t)t * 256u;
          for (int kb = 0; kb < L.kb_act; ++kb) {
            if (t == 0) {
              wait_full(pos_w + kb);
              tc_fence_after();
            }
            mma_kblock(tmem_d, smem_base + t * kActBytes + kb * kBlkBytes, stage_addr(pos_w + kb), idesc, kb == 0,
                       t == 1 ? &ring_empty[(pos_w + kb) % kRingStages] : nullptr,
                       (kb == L.kb_act - 1 && !L.kb_in) ? &acc_full[t] : nullptr, nullptr);
          }
          if (p.trace && blockIdx.x == 0 && opcount < 64 && lane == 0) p.trace[(opcount * 2 + t) * 8 + 1] = clock64();
        }
        pos += L.kb_act;
        if (L.kb_act) seed_pending = false;
        // --- K blocks from the chain input: x(tile 0), W, x(tile 1) per K block
        for (int kb = 0; kb < L.kb_in; ++kb) {
          const uint32_t px0 = pos, pw = pos + 1, px1 = pos + 2;
          pos += 3;
          for (int t = 0; t < 2; ++t) {
            if (!L.kb_act && kb == 0) {
              mbar_wait_cluster(&acc_free[t], free_parity);
              tc_fence_after();
            }
            const uint32_t px = t ? px1 : px0;
            wait_full(px);
            if (t == 0) wait_full(pw);
            tc_fence_after();
            mma_kblock(tmem_base + (uint32_t)t * 256u, stage_addr(px), stage_addr(pw), idesc, !L.kb_act && kb == 0,
                       &ring_empty[px % kRingStages], t == 1 ? &ring_empty[pw % kRingStages] : nullptr,
                       kb == L.kb_in - 1 ? &acc_full[t] : nullptr);
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===== store warp: issues the TMA saves of both tiles (one [128 x 64] box per K block) so that the epilogue warps
    // never block on the TMA queue; signals save_drained[t] once the stores of act[t] have READ shared memory, and
    // in2_ready once a save that this launch re-loads is complete in global memory =====
    uint32_t nsave0 = 0, nsave1 = 0;
    bool any = false;
    for (int64_t st = cluster_id; st < num_super; st += num_clusters) {
      for (int l = 0; l < p.num_ops; ++l) {
        const PairOp& L = p.op[l];
        const bool hidden = L.kind == 0 || (MODE == 1 && L.kind == 2);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (hidden && L.save) {
            const uint32_t n = t ? nsave1 : nsave0;
            mbar_wait(&save_ready[t], n & 1u);
            if (t) ++nsave1; else ++nsave0;
            const int64_t row0 = st * 512 + t * 256 + (int64_t)rank * 128;
            if (elect_one_sync()) {
              if (row0 < p.m) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  tma_store_2d(&maps.save[l], smem_base + (uint32_t)(t * kActBytes + c * kBlkBytes), c * 64, (int)row0);
              }
              tma_store_commit();
              // the next save of this tile is at least one op (> 2 000 cycles) away: waiting here costs nothing and keeps
              // the hand-back unconditional (no dependence on a later save being issued)
              tma_store_wait_read<0>();
              mbar_arrive(&save_drained[t]);
            }
            __syncwarp();
            any = true;
          }
          if (l == p.in2_sync_op && t == 1) {
            // the save that a later op of this super tile re-loads (at least 3 ops older) must be complete in global memory
            if (elect_one_sync()) {
              if (hidden && L.save) asm volatile("cp.async.bulk.wait_group 2;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
              mbar_arrive(in2_ready);
            }
            __syncwarp();
          }
        }
      }
    }
    if (any) {
      if (elect_one_sync()) tma_store_wait_all();
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== epilogue warps: q = TMEM lane quadrant (32 rows), h = column half (128 columns = K blocks 2h, 2h+1) =====
    const int q = (warp - 4) & 3;
    const int h = (warp - 4) >> 2;
    const int r_in_tile = q * 32 + lane;
    // bias staging: two buffers (op parity) x two column halves x 128 floats.  The four warps of a column half
    // write identical values, so no cross-warp barrier is needed: every warp only relies on its own stores.
    const uint32_t bias_half = smem_base + kSmemBias + (uint32_t)h * 512u;
    const uint32_t free_addr0 = map_to_cta(smem_u32(&acc_free[0]), 0);   // leader's acc_free[0]; [1] is 8 bytes on
    const uint32_t seed_addr0 = map_to_cta(smem_u32(&seed_done[0]), 0);
    uint32_t opcount = 0;
    uint32_t nfull = 0;                  // accumulator commits consumed so far (seed ops have none)
    uint32_t outstanding = 0;            // bit t: a save of act[t] has been handed to the store warp and not yet seen drained
    uint32_t drained_phase = 0;          // bit t: parity of the next save_drained[t] phase to wait for
    uint4 bits_next = make_uint4(0, 0, 0, 0);
    auto load_bits = [&](int64_t st, int l, int t) -> uint4 {
      // ReLU bits of this thread's row for op l (this warp's 128 columns = words 4h .. 4h+3; relu_bits_index layout:
      // the 32 rows of the warp are contiguous per word, so each of the four loads is one 128-byte line)
      uint4 b = make_uint4(0, 0, 0, 0);
      if (l < p.num_ops && st < num_super && p.op[l].kind != 1) {
        const int64_t row_w = st * 512 + t * 256 + (int64_t)rank * 128 + q * 32;
        if (row_w < p.m) {
          const uint32_t* bp = p.op[l].mask_bits + (size_t)(row_w >> 5) * 256 + (4 * h) * 32 + lane;
          b = make_uint4(__ldg(bp), __ldg(bp + 32), __ldg(bp + 64), __ldg(bp + 96));
        }
      }
      return b;
    };
    if (MODE == 1) bits_next = load_bits(cluster_id, 0, 0);
    for (int64_t st = cluster_id; st < num_super; st += num_clusters) {
      for (int l = 0; l < p.num_ops; ++l, ++opcount) {
        const PairOp& L = p.op[l];
        const bool seed = MODE == 1 && L.kind == 2;   // act tile = bits ? vec : 0 (vec staged like a bias), no accumulator
        const bool hidden = L.kind == 0 || seed;
        const float* bias_ptr = (L.kind == 0) ? (MODE == 0 ? L.bias : nullptr) : (seed ? L.bias : p.gepi[L.gepi].bias);
        const uint32_t bias_buf = bias_half + (opcount & 1u) * 1024u;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias_ptr && h * 128 + 4 * lane < L.n) bv = __ldg(reinterpret_cast<const float4*>(bias_ptr + h * 128) + lane);
        for (int t = 0; t < 2; ++t) {
          const int64_t row = st * 512 + t * 256 + (int64_t)rank * 128 + r_in_tile;
          const bool row_ok = row < p.m;
          uint4 bits_cur = bits_next;
          if (MODE == 1) {
            // prefetch the bits of the next (op, tile) this warp will process
            if (t == 0) bits_next = load_bits(st, l, 1);
            else if (l + 1 < p.num_ops) bits_next = load_bits(st, l + 1, 0);
            else bits_next = load_bits(st + num_clusters, 0, 0);
          }
          const bool tr = p.trace && blockIdx.x == 0 && warp == 4 && lane == 0 && opcount < 64;
          if (tr) p.trace[(opcount * 2 + t) * 8 + 2] = clock64();
          if (!seed) {
            mbar_wait(&acc_full[t], nfull & 1u);
            tc_fence_after();
          }
          if (tr) p.trace[(opcount * 2 + t) * 8 + 3] = clock64();
          if (t == 0 && bias_ptr) {
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(bias_buf + 16u * lane), "f"(bv.x), "f"(bv.y), "f"(bv.z),
                         "f"(bv.w)
                         : "memory");
            __syncwarp();
          }
          const uint32_t taddr = tmem_base + (uint32_t)t * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 128);
          if (hidden) {
            if (outstanding >> t & 1u) {
              // the TMA stores that read act[t] must have drained it before it is overwritten
              mbar_wait(&save_drained[t], drained_phase >> t & 1u);
              drained_phase ^= 1u << t;
              outstanding &= ~(1u << t);
            }
            const uint32_t act_row = smem_base + (uint32_t)(t * kActBytes + r_in_tile * 128);
            const uint32_t bits_arr[4] = {bits_cur.x, bits_cur.y, bits_cur.z, bits_cur.w};
            uint32_t bits_out[4];
            uint32_t ra[32], rb[32];
            if (!seed) tmem_ld32(taddr, ra);
#pragma unroll
            for (int g = 0; g < 4; ++g) {   // 32-column groups of this warp's 128 columns; TMEM loads run one group ahead
              uint32_t* cur = (g & 1) ? rb : ra;
              uint32_t* nxt = (g & 1) ? ra : rb;
              if (!seed) {
                tmem_ld_wait();
                if (g < 3) tmem_ld32(taddr + (uint32_t)((g + 1) * 32), nxt);
              }
              uint32_t packed[16];
              if (MODE == 0) {
                const uint32_t baddr = bias_buf + (uint32_t)g * 128u;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 b = lds128f(baddr + 16u * i);
                  packed[2 * i] = pack_relu_act<F16>(__uint_as_float(cur[4 * i]) + b.x, __uint_as_float(cur[4 * i + 1]) + b.y);
                  packed[2 * i + 1] = pack_relu_act<F16>(__uint_as_float(cur[4 * i + 2]) + b.z, __uint_as_float(cur[4 * i + 3]) + b.w);
                }
                if (L.save_bits) bits_out[g] = relu_bits_of(packed);
              } else if (seed) {
                // seed tile of a dgrad chain: vec[col] where the ReLU bit is set (vec staged in the bias buffer)
                const uint32_t baddr = bias_buf + (uint32_t)g * 128u;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 b = lds128f(baddr + 16u * i);
                  packed[2 * i] = apply_relu_bits(bits_arr[g], 2 * i, pack_act<F16>(b.x * p.seed_scale, b.y * p.seed_scale));
                  packed[2 * i + 1] = apply_relu_bits(bits_arr[g], 2 * i + 1, pack_act<F16>(b.z * p.seed_scale, b.w * p.seed_scale));
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  packed[i] = apply_relu_bits(bits_arr[g], i, pack_act<F16>(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])));
              }
              const int c = 2 * h + (g >> 1);   // activation K block
              const uint32_t blk = act_row + (uint32_t)(c * kBlkBytes);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint32_t pos16 = (uint32_t)(((g & 1) * 4 + u) ^ (r_in_tile & 7));
                sts128(blk + pos16 * 16u, packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
              }
            }
            fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core and to TMA stores
            tc_fence_before();
            __syncwarp();
            if (tr) p.trace[(opcount * 2 + t) * 8 + 4] = clock64();
            if (lane == 0) {
              // (the store warp's signal goes first: see chain_x3.cu for why no warp can then reach the next save early)
              if (L.save) mbar_arrive(&save_ready[t]);
              mbar_arrive_cluster_addr((seed ? seed_addr0 : free_addr0) + 8u * t);
            }
            if (tr) p.trace[(opcount * 2 + t) * 8 + 5] = clock64();   // arrive + TMA store issue done (lane 0 of warp 4)
            if (L.save) outstanding |= 1u << t;
            if (MODE == 0 && L.save_bits && row - lane < p.m) {
              uint32_t* bp = L.save_bits + (size_t)((row - lane) >> 5) * 256 + (4 * h) * 32 + lane;
#pragma unroll
              for (int g = 0; g < 4; ++g) bp[32 * g] = bits_out[g];
            }
            if (tr) p.trace[(opcount * 2 + t) * 8 + 6] = clock64();   // ReLU-bit store issued
          } else {
            const GemmEpilogue& ge = p.gepi[L.gepi];
            const int c_end = L.n < h * 128 + 128 ? L.n : h * 128 + 128;
            for (int c0 = h * 128; c0 < c_end; c0 += 32) {
              uint32_t r[32];
              tmem_ld32(taddr + (uint32_t)(c0 - h * 128), r);
              tmem_ld_wait();
              if (row_ok) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  if (c0 + 16 * hh < L.n) {
                    float v[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[16 * hh + e]);
                    if (bias_ptr) {
                      const uint32_t baddr = bias_buf + (uint32_t)(c0 - h * 128 + 16 * hh) * 4u;
#pragma unroll
                      for (int i = 0; i < 4; ++i) {
                        const float4 b = lds128f(baddr + 16u * i);
                        v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
                      }
                    }
                    epi_global16<F16>(ge, (size_t)row, c0 + 16 * hh, v);
                  }
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_addr(free_addr0 + 8u * t);
          }
        }
        if (!seed) ++nfull;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc2(tmem_base, kTmemCols);
}

}  // namespace

int launch_chain(const ChainArgs& a, cudaStream_t st) {
  if (a.impl == 2) return launch_chain_x3(a, st);
  if (a.impl != 0) return rn_set_error(RN_ERR_ARG, "chain: unknown implementation");
  if (a.w_planes != 1 && a.w_planes != 2) return rn_set_error(RN_ERR_ARG, "chain: w_planes is 1 or 2");
  if (a.w_planes == 2 && (a.act_f16 || a.w_f16)) return rn_set_error(RN_ERR_UNSUPPORTED, "chain: split weights are bf16");
  if (a.m <= 0) return RN_OK;
  if (a.act_f16 != a.w_f16) return rn_set_error(RN_ERR_UNSUPPORTED, "chain: activations and weights must share one 16-bit format");
  if (a.num_ops < 1 || a.num_ops > kMaxOps) return rn_set_error(RN_ERR_ARG, "chain: 1..12 ops");
  if (a.in.hi && (a.in_cols % 64 || a.in_cols < 64 || a.in_cols > 256)) return rn_set_error(RN_ERR_ARG, "chain: input tile must be 64..256 columns");
  if (a.m + 512 > 0x7fffffffLL) return rn_set_error(RN_ERR_ARG, "chain: too many rows for one launch");
  PairMaps maps;
  PairParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  int mode = -1;
  if ((rc = tc::make_map(&maps.in, a.in.hi, a.m, a.in_valid, a.in.ld, kBM))) return rc;
  if ((rc = tc::make_map(&maps.in2, a.in2.hi, a.m, a.in2_valid, a.in2.ld, kBM))) return rc;
  p.in2_sync_op = -1;
  p.a_f16 = a.act_f16;
  p.b_f16 = a.w_f16;
  p.w_planes = a.w_planes;
  p.seed_scale = a.seed_scale;
  p.num_ops = a.num_ops;
  p.in_kb = a.in.hi ? a.in_cols / kBK : 0;
  p.m = a.m;
  for (int l = 0; l < kMaxOps; ++l) {
    memset(&maps.w[l], 0, sizeof(CUtensorMap));
    memset(&maps.w2[l], 0, sizeof(CUtensorMap));
    memset(&maps.save[l], 0, sizeof(CUtensorMap));
    if (l >= a.num_ops) continue;
    const ChainOpArgs& L = a.op[l];
    if (L.kind == 2) {
      // seed op of a dgrad chain: activation tile = ReLU bit ? vec[col] : 0 (no GEMM)
      if (!L.bias || !L.mask_bits || L.kb_act || L.kb_in || L.save_hi || l + 1 >= a.num_ops || a.op[l + 1].kb_act != 4)
        return rn_set_error(RN_ERR_ARG, "chain: bad seed op (needs vec + bits and a following op that reads the activation tile)");
      if (mode < 0) mode = 1;
      if (mode != 1) return rn_set_error(RN_ERR_ARG, "chain: seed ops belong to backward chains");
      PairOp& o = p.op[l];
      o.n = 256; o.kind = 2; o.bias = L.bias; o.mask_bits = L.mask_bits;
      continue;
    }
    if (L.n % 16 || L.n < 16 || L.n > 256 || (L.kind == 0 && L.n != 256)) return rn_set_error(RN_ERR_ARG, "chain: bad op width");
    const int in_kb_expected = L.in2 ? a.in2_cols / kBK : p.in_kb;
    if ((L.kb_act != 0 && L.kb_act != 4) || (L.kb_in != 0 && L.kb_in != in_kb_expected) || L.kb_act + L.kb_in == 0 ||
        (l == 0 && L.kb_act != 0) || (L.in2 && (!a.in2.hi || a.in2_cols % 64)) || (L.kb_in && !L.in2 && !a.in.hi))
      return rn_set_error(RN_ERR_ARG, "chain: bad K structure");
    if (L.in2 && L.kb_in) {
      // if the second input is a save of an earlier op of this launch, op l-1's epilogue orders store -> load
      for (int j = 0; j < l; ++j)
        if (a.op[j].kind == 0 && a.op[j].save_hi == a.in2.hi) {
          if (j > l - 3) return rn_set_error(RN_ERR_ARG, "chain: a re-loaded save must be produced at least 3 ops earlier");
          if (a.op[l - 1].kind != 0) return rn_set_error(RN_ERR_ARG, "chain: the op before a re-loaded save's reader must be hidden");
          p.in2_sync_op = l - 1;
        }
    }
    const int ktot = (L.kb_act + L.kb_in) * kBK;
    if ((rc = tc::make_map(&maps.w[l], L.w, L.n, ktot, L.w_ld, L.n / 2))) return rc;
    if (a.w_planes == 2) {
      if (!L.w_lo) return rn_set_error(RN_ERR_ARG, "chain: missing lo weight plane");
      if ((rc = tc::make_map(&maps.w2[l], L.w_lo, L.n, ktot, L.w_ld, L.n / 2))) return rc;
    }
    PairOp& o = p.op[l];
    o.n = L.n; o.kb_act = L.kb_act; o.kb_in = L.kb_in; o.in2 = L.in2;
    o.kind = L.kind; o.gepi = L.gepi; o.bias = L.bias;
    o.mask_bits = L.mask_bits; o.save_bits = L.save_bits;
    o.save = (L.kind == 0 && L.save_hi) ? 1 : 0;
    if (o.save && (rc = tc::make_map(&maps.save[l], L.save_hi, a.m, 256, 256, 128))) return rc;
    if (L.kind == 0) {
      if (mode < 0) mode = L.mode;
      if (L.mode != mode) return rn_set_error(RN_ERR_ARG, "chain: forward and backward hidden ops cannot be mixed");
      if (L.mode == 0 && !L.bias) return rn_set_error(RN_ERR_ARG, "chain: forward op without bias");
      if (L.mode == 1 && !L.mask_bits) return rn_set_error(RN_ERR_ARG, "chain: backward op without ReLU bits");
    }
  }
  if (mode < 0) mode = 0;
  p.gepi[0] = a.gepi[0];
  p.gepi[1] = a.gepi[1];
  static bool smem_set = false;
  if (!smem_set) {
    if ((rc = tc::set_smem(chain_pair_kernel<0, 0>, kSmemTotal))) return rc;
    if ((rc = tc::set_smem(chain_pair_kernel<1, 0>, kSmemTotal))) return rc;
    if ((rc = tc::set_smem(chain_pair_kernel<0, 1>, kSmemTotal))) return rc;
    if ((rc = tc::set_smem(chain_pair_kernel<1, 1>, kSmemTotal))) return rc;
    smem_set = true;
  }
  const int64_t supers = (a.m + 511) / 512;
  const int max_clusters = tc::num_sms() / 2;
  const unsigned grid = 2u * (unsigned)(supers < max_clusters ? supers : max_clusters);
  // debug: RN_CHAIN_TRACE=<n> prints the per-(op, tile) timeline of CTA 0 for the first n launches
  static long long* trace_buf = nullptr;
  static int trace_left = getenv("RN_CHAIN_TRACE") ? atoi(getenv("RN_CHAIN_TRACE")) : 0;
  p.trace = nullptr;
  if (trace_left > 0) {
    if (!trace_buf) cudaMalloc(&trace_buf, 64 * 2 * 8 * sizeof(long long));
    cudaMemsetAsync(trace_buf, 0, 64 * 2 * 8 * sizeof(long long), st);
    p.trace = trace_buf;
  }
  double exec_flops = 0.0;
  for (int l = 0; l < a.num_ops; ++l)
    if (a.op[l].kind != 2) exec_flops += 2.0 * (double)a.m * a.op[l].n * (a.op[l].kb_act + a.op[l].kb_in) * kBK * a.w_planes;
  rn_prof_begin(RN_PROF_CHAIN_TC, st, a.algo_flops, exec_flops);
  if (mode == 0 && !a.act_f16)
    chain_pair_kernel<0, 0><<<grid, 384, kSmemTotal, st>>>(maps, p);
  else if (mode == 0)
    chain_pair_kernel<0, 1><<<grid, 384, kSmemTotal, st>>>(maps, p);
  else if (!a.act_f16)
    chain_pair_kernel<1, 0><<<grid, 384, kSmemTotal, st>>>(maps, p);
  else
    chain_pair_kernel<1, 1><<<grid, 384, kSmemTotal, st>>>(maps, p);
  rn_prof_end(RN_PROF_CHAIN_TC, st);
  RN_CUDA_CHECK_LAUNCH();
  if (p.trace) {
    --trace_left;
    cudaStreamSynchronize(st);
    static long long h[64 * 2 * 8];
    cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = h[0];
    printf("chain trace (mode %d, %d ops, m=%lld): per (op,tile): mma_wait_done mma_issued | epi_wait_begin epi_wait_end epi_end | stores_issued bits_stored  [cycles since first]\n", mode, a.num_ops, (long long)a.m);
    for (int i = 0; i < 40; ++i)
      printf("  op %2d tile %d: %8lld %8lld | %8lld %8lld %8lld | %8lld %8lld\n", i / 2, i % 2, h[i * 8] - t0, h[i * 8 + 1] - t0, h[i * 8 + 2] - t0,
             h[i * 8 + 3] - t0, h[i * 8 + 4] - t0, h[i * 8 + 5] - t0, h[i * 8 + 6] - t0);
  }
  return RN_OK;
}

}  // namespace rn
