// Fused GEMM chain with the A operand in tensor memory (sm_100a, tcgen05.mma cta_group::2, TS form).
//
// The SS chain (chain_pair.cu) is bound by the 128 B/clk shared-memory port: per 128-row tile and layer the MMA
// operand reads (A + half of B), the epilogue's activation stores, the bias loads and the weight TMA writes need
// ~2 900 wavefronts against the 2 048 cycles the tensor pipe needs.  Here the running activation never touches
// shared memory:
//
//   * one 128-row tile per CTA (256 rows per CTA pair, UMMA M = 256); every op is issued as two column halves
//     (N <= 128 each) into two 128-column fp32 accumulators, so TMEM holds
//         ACC0 | ACC1 | A-buffer 0 | A-buffer 1      (4 x 128 columns)
//     and the epilogue of half 0 overlaps the MMAs of half 1;
//   * the epilogue turns an accumulator half into bf16 and writes it with tcgen05.st into the A buffer of the NEXT
//     op (row = TMEM lane, two K elements per 32-bit column); the next op reads its A operand from there
//     (tcgen05.mma [d], [a_tmem], b_desc);
//   * its first two K blocks only depend on half 0 of the previous op, so they are issued while the epilogue of
//     half 1 is still running;
//   * shared memory carries only the weights (8 KB stages: one K block of one column half, half of its rows per
//     CTA), the chain-input tile (layer 0 / skip layer / seed gradient, SS form) and, in training, the staging
//     buffers of the TMA-store saves.
//
// Roles per CTA (384 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator,
// warps 4-11 epilogue (warp % 4 = TMEM lane quadrant, (warp - 4) / 4 = 64-column half of the accumulator half).
#include <stdio.h>
#include <stdlib.h>

#include "chain_common.cuh"

namespace rn {
namespace {
using namespace tc;
using namespace chain;

constexpr int kRing = 12;                     // weight ring stages
constexpr int kWStage = 8192;                 // [64 rows x 64 K] bf16
constexpr int kBlkBytes = kBM * kBK * 2;      // 16 KB input-tile K block
constexpr int kSmemIn = 0;                    // input tile: up to 4 K blocks
constexpr int kSmemRing = 4 * kBlkBytes;
constexpr int kSmemStage = kSmemRing + kRing * kWStage;   // save staging: 2 x [128 x 128] bf16 (two swizzled K blocks each)
constexpr int kSmemBars = kSmemStage + 2 * 2 * kBlkBytes;
constexpr int kSmemBias = kSmemBars + 256;
constexpr int kSmemTotal = kSmemBias + 2 * 1024;
static_assert(kSmemTotal <= 232448, "shared memory budget");

constexpr uint32_t kAcc0 = 0, kABuf0 = 256;   // TMEM columns: ACC h at 128 h, A buffer b at 256 + 128 b

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

struct TsMaps {
  CUtensorMap in;
  CUtensorMap w0[kMaxOps];     // column half 0 (box rows = width / 2)
  CUtensorMap w1[kMaxOps];     // column half 1
  CUtensorMap save[kMaxOps];
};
struct TsHalf {
  int n0, n1;   // widths of the two column halves (multiples of 32; n1 = 0: single half)
};
struct TsParams {
  PairParams p;
  TsHalf half[kMaxOps];
  int last_in_op;   // last op that reads the chain-input tile
};

// MODE 0: forward chain (hidden ops: bias + ReLU); MODE 1: backward / dgrad chain (hidden ops: ReLU bit mask)
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
chain_ts_kernel(const __grid_constant__ TsMaps maps, const __grid_constant__ TsParams tp) {
  const PairParams& p = tp.p;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBars);
  uint64_t* ring_full = bars;          // [12] leader's are used (TMA bytes of both CTAs land there)
  uint64_t* ring_empty = bars + 12;    // [12] per CTA, multicast MMA commit
  uint64_t* acc_full = bars + 24;      // [2]  per CTA, multicast MMA commit (accumulator half complete)
  uint64_t* epi_done = bars + 26;      // [2]  leader's: 16 arrivals (accumulator half drained, next A half written)
  uint64_t* in_full = bars + 28;       // leader's
  uint64_t* in_empty = bars + 29;      // per CTA, multicast MMA commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t num_tiles = (p.m + 255) / 256;
  const int64_t cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const uint32_t smem_base = smem_u32(smem);

  if (threadIdx.x == 0) {
    if (smem_base & 1023u) __trap();
    tma_prefetch_desc(&maps.in);
    for (int i = 0; i < p.num_ops; ++i) {
      tma_prefetch_desc(&maps.w0[i]);
      tma_prefetch_desc(&maps.w1[i]);
    }
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&epi_done[i], 16);
    }
    mbar_init(in_full, 1);
    mbar_init(in_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (whole warp, one elected lane issues) =====
    uint32_t pos = 0, tile_iter = 0;
    const uint32_t in_full_leader = map_to_cta(smem_u32(in_full), 0);
    const uint32_t ring_full_leader = map_to_cta(smem_u32(&ring_full[0]), 0);
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tile_iter) {
      // chain-input tile (this CTA's 128 rows), resident until the last op that reads it has completed
      mbar_wait(in_empty, (tile_iter & 1u) ^ 1u);
      const int64_t row0 = tile * 256 + (int64_t)rank * 128;
      if (elect_one_sync()) {
        if (rank == 0) mbar_arrive_expect_tx(in_full, 2u * (uint32_t)p.in_kb * kBlkBytes);
        for (int kb = 0; kb < p.in_kb; ++kb)
          tma_load_2d_pair(smem_base + kSmemIn + kb * kBlkBytes, &maps.in, in_full_leader, kb * kBK, (int)row0);
      }
      __syncwarp();
      for (int l = 0; l < p.num_ops; ++l) {
        const PairOp& L = p.op[l];
        const int nkb = L.kb_act + L.kb_in;
        for (int h = 0; h < 2; ++h) {
          const int nh = h ? tp.half[l].n1 : tp.half[l].n0;
          if (!nh) continue;
          const CUtensorMap* wm = h ? &maps.w1[l] : &maps.w0[l];
          for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t s = pos % kRing, ph = (pos / kRing) & 1u;
            mbar_wait(&ring_empty[s], ph ^ 1u);
            if (elect_one_sync()) {
              if (rank == 0) mbar_arrive_expect_tx(&ring_full[s], (uint32_t)nh * 128u);
              tma_load_2d_pair(smem_base + kSmemRing + s * kWStage, wm, ring_full_leader + 8u * s, kb * kBK,
                               h * 128 + (int)rank * (nh >> 1));
            }
            __syncwarp();
            ++pos;
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader CTA; whole warp, one elected lane issues) =====
    uint32_t pos = 0;
    uint32_t nfull0 = 0, nfull1 = 0;   // commits of each accumulator half so far
    uint32_t wdone0 = 0, wdone1 = 0;   // epilogue completions of each half already waited for
    uint32_t cur = 0;                  // A buffer the current op reads
    uint32_t tile_iter = 0;
    // wait until the epilogue has finished the first `target` uses of accumulator half hh (drained it and, for a
    // hidden op, written its half of the next A buffer)
    auto ensure = [&](int hh, uint32_t target) {
      uint32_t& wd = hh ? wdone1 : wdone0;
      if (wd < target) {
        mbar_wait_cluster(&epi_done[hh], (target - 1u) & 1u);
        tc_fence_after();
        wd = target;
      }
    };
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tile_iter) {
      bool in_ready = false;
      for (int l = 0; l < p.num_ops; ++l) {
        const PairOp& L = p.op[l];
        const int nkb = L.kb_act + L.kb_in;
        const uint32_t a_tmem = tmem_base + kABuf0 + cur * 128u;
        const uint32_t base0 = nfull0, base1 = nfull1;   // uses before this op: what its operands / accumulators depend on
        for (int h = 0; h < 2; ++h) {
          const int nh = h ? tp.half[l].n1 : tp.half[l].n0;
          if (!nh) continue;
          const uint32_t idesc = make_idesc2(nh);
          const uint32_t tmem_d = tmem_base + kAcc0 + (uint32_t)h * 128u;
          ensure(h, h ? base1 : base0);   // accumulator half drained by the epilogue of its previous use
          const uint32_t tri = (tile_iter * p.num_ops + l) * 2 + h;
          if (p.trace && blockIdx.x == 0 && tri < 128 && lane == 0) p.trace[tri * 8 + 0] = clock64();
          for (int kb = 0; kb < nkb; ++kb) {
            const bool from_act = kb < L.kb_act;
            if (from_act) {
              // A K blocks 0-1 were written by the half-0 epilogue of the previous hidden op, 2-3 by its half-1 epilogue
              const int src = kb >> 1;
              ensure(src, src ? base1 : base0);
            } else if (!in_ready) {
              mbar_wait(in_full, tile_iter & 1u);
              in_ready = true;
            }
            const uint32_t s = pos % kRing;
            mbar_wait(&ring_full[s], (pos / kRing) & 1u);
            tc_fence_after();
            const uint32_t sb = smem_base + kSmemRing + s * kWStage;
            if (elect_one_sync()) {
#pragma unroll
              for (int kk = 0; kk < kBK / kUmmaK; ++kk) {
                const uint32_t koff = kk * kUmmaK * 2;
                const uint64_t db = make_desc(sb + koff, 16, 1024);
                const uint32_t accum = (kb | kk) ? 1u : 0u;
                if (from_act) {
                  umma2_ts_bf16(tmem_d, a_tmem + (uint32_t)(kb * 32 + kk * 8), db, idesc, accum);
                } else {
                  const uint32_t sa = smem_base + kSmemIn + (kb - L.kb_act) * kBlkBytes;
                  umma2_bf16(tmem_d, make_desc(sa + koff, 16, 1024), db, idesc, accum);
                }
              }
              umma2_commit(&ring_empty[s]);
              if (kb == nkb - 1) umma2_commit(&acc_full[h]);
            }
            __syncwarp();
            ++pos;
          }
          if (h) ++nfull1; else ++nfull0;
          if (p.trace && blockIdx.x == 0 && tri < 128 && lane == 0) p.trace[tri * 8 + 1] = clock64();
        }
        if (l == tp.last_in_op) {
          if (elect_one_sync()) umma2_commit(in_empty);
          __syncwarp();
        }
        if (L.kind == 0) cur ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue warps: q = TMEM lane quadrant (32 rows), c = 64-column half of the accumulator half =====
    const int q = (warp - 4) & 3;
    const int c = (warp - 4) >> 2;
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t bias_base = smem_base + kSmemBias;   // [2 (op parity)][256] floats
    const uint32_t done_addr0 = map_to_cta(smem_u32(&epi_done[0]), 0);
    uint32_t opcount = 0, nfull0 = 0, nfull1 = 0, cur = 0;
    uint32_t has_group = 0, newer = 0;   // TMA-store bookkeeping per staging buffer (bit h)
    bool any_store = false;
    uint4 bits_next = make_uint4(0, 0, 0, 0);   // ReLU bits of this thread's row: words {2c, 2c+1} of half 0, then of half 1
    auto load_bits = [&](int64_t tile, int l) -> uint4 {
      uint4 b = make_uint4(0, 0, 0, 0);
      if (l < p.num_ops && tile < num_tiles && p.op[l].kind == 0) {
        const int64_t row = tile * 256 + (int64_t)rank * 128 + r_in_tile;
        if (row < p.m) {
          const uint2 b0 = __ldg(reinterpret_cast<const uint2*>(p.op[l].mask_bits + (size_t)row * 8) + c);
          const uint2 b1 = __ldg(reinterpret_cast<const uint2*>(p.op[l].mask_bits + (size_t)row * 8) + 2 + c);
          b = make_uint4(b0.x, b0.y, b1.x, b1.y);
        }
      }
      return b;
    };
    if (MODE == 1) bits_next = load_bits(cluster_id, 0);
    for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int64_t row = tile * 256 + (int64_t)rank * 128 + r_in_tile;
      const bool row_ok = row < p.m;
      for (int l = 0; l < p.num_ops; ++l, ++opcount) {
        const PairOp& L = p.op[l];
        const bool hidden = L.kind == 0;
        const float* bias_ptr = hidden ? (MODE == 0 ? L.bias : nullptr) : p.gepi[L.gepi].bias;
        const uint32_t bias_buf = bias_base + (opcount & 1u) * 1024u;
        // every warp stages the bias slice it will read (columns h*128 + c*64 .. +64 of both halves); warps with the
        // same c write identical values, so only __syncwarp is needed
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        const int bcol = (lane >> 4) * 128 + c * 64 + (lane & 15) * 4;   // lanes 0-15: half 0, 16-31: half 1
        if (bias_ptr && bcol < L.n) bv = __ldg(reinterpret_cast<const float4*>(bias_ptr + bcol));
        const uint4 bits_cur = bits_next;
        if (MODE == 1) bits_next = (l + 1 < p.num_ops) ? load_bits(tile, l + 1) : load_bits(tile + num_clusters, 0);
        bool bias_staged = false;
        for (int h = 0; h < 2; ++h) {
          const int nh = h ? tp.half[l].n1 : tp.half[l].n0;
          if (!nh) continue;
          const uint32_t tri = opcount * 2 + h;
          const bool tr = p.trace && blockIdx.x == 0 && warp == 4 && lane == 0 && tri < 128;
          if (tr) p.trace[tri * 8 + 2] = clock64();
          mbar_wait(&acc_full[h], (h ? nfull1 : nfull0) & 1u);
          if (h) ++nfull1; else ++nfull0;
          tc_fence_after();
          if (tr) p.trace[tri * 8 + 3] = clock64();
          if (!bias_staged && bias_ptr) {
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(bias_buf + (uint32_t)bcol * 4u), "f"(bv.x), "f"(bv.y),
                         "f"(bv.z), "f"(bv.w)
                         : "memory");
            __syncwarp();
            bias_staged = true;
          }
          const uint32_t tacc = tmem_base + kAcc0 + (uint32_t)h * 128u + lane_sel + (uint32_t)(c * 64);
          if (hidden) {
            // next op's A buffer: output column n = h*128 + c*64 + g*32 + j is K element n -> 32-bit column n / 2
            const uint32_t tnext = tmem_base + kABuf0 + (cur ^ 1u) * 128u + lane_sel + (uint32_t)(h * 64 + c * 32);
            if (L.save && (has_group >> h & 1u)) {
              if (lane == 0) {
                if (newer >> h & 1u) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
              }
              __syncwarp();
              has_group &= ~(1u << h);
            }
            const uint32_t stage_row = smem_base + kSmemStage + (uint32_t)(h * 2 * kBlkBytes + c * kBlkBytes + r_in_tile * 128);
            uint32_t bits_out[2];
            uint32_t ra[32], rb[32];
            tmem_ld32(tacc, ra);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint32_t* curr = g ? rb : ra;
              tmem_ld_wait();
              if (g == 0) tmem_ld32(tacc + 32u, rb);
              uint32_t packed[16];
              if (MODE == 0) {
                const uint32_t baddr = bias_buf + (uint32_t)(h * 128 + c * 64 + g * 32) * 4u;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 b = lds128f(baddr + 16u * i);
                  packed[2 * i] = pack_relu_bf16x2(__uint_as_float(curr[4 * i]) + b.x, __uint_as_float(curr[4 * i + 1]) + b.y);
                  packed[2 * i + 1] = pack_relu_bf16x2(__uint_as_float(curr[4 * i + 2]) + b.z, __uint_as_float(curr[4 * i + 3]) + b.w);
                }
                if (L.save_bits) bits_out[g] = relu_bits_of(packed);
              } else {
                const uint32_t bw = h ? (g ? bits_cur.w : bits_cur.z) : (g ? bits_cur.y : bits_cur.x);
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  packed[i] = apply_relu_bits(bw, i, pack_bf16x2(__uint_as_float(curr[2 * i]), __uint_as_float(curr[2 * i + 1])));
              }
              tmem_st16(tnext + (uint32_t)(g * 16), packed);
              if (L.save) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const uint32_t pos16 = (uint32_t)((g * 4 + u) ^ (r_in_tile & 7));
                  sts128(stage_row + pos16 * 16u, packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
                }
              }
            }
            tmem_st_wait();
            if (L.save) fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (tr) p.trace[tri * 8 + 4] = clock64();
            if (lane == 0) {
              mbar_arrive_cluster_addr(done_addr0 + 8u * h);
              if (L.save) {
                const int64_t row_w = tile * 256 + (int64_t)rank * 128 + q * 32;
                if (row_w < p.m)
                  tma_store_2d(&maps.save[l], smem_base + kSmemStage + (uint32_t)(h * 2 * kBlkBytes + c * kBlkBytes + q * 32 * 128),
                               h * 128 + c * 64, (int)row_w);
                tma_store_commit();
              }
            }
            if (L.save) {
              has_group |= 1u << h;
              newer &= ~(1u << h);
              if (has_group >> (h ^ 1) & 1u) newer |= 1u << (h ^ 1);
              any_store = true;
            }
            if (MODE == 0 && L.save_bits && row_ok)
              *(reinterpret_cast<uint2*>(L.save_bits + (size_t)row * 8) + h * 2 + c) = make_uint2(bits_out[0], bits_out[1]);
          } else {
            const GemmEpilogue& ge = p.gepi[L.gepi];
            const int col_lo = h * 128 + c * 64;
            const int c_end = L.n < col_lo + 64 ? L.n : col_lo + 64;
            for (int c0 = col_lo; c0 < c_end; c0 += 32) {
              uint32_t r[32];
              tmem_ld32(tacc + (uint32_t)(c0 - col_lo), r);
              tmem_ld_wait();
              if (row_ok) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  if (c0 + 16 * hh < L.n) {
                    float v[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[16 * hh + e]);
                    if (bias_ptr) {
                      const uint32_t baddr = bias_buf + (uint32_t)(c0 + 16 * hh) * 4u;
#pragma unroll
                      for (int i = 0; i < 4; ++i) {
                        const float4 b = lds128f(baddr + 16u * i);
                        v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
                      }
                    }
                    epi_global16(ge, (size_t)row, c0 + 16 * hh, v);
                  }
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_addr(done_addr0 + 8u * h);
          }
        }
        if (hidden) cur ^= 1u;
      }
    }
    if (any_store && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc2(tmem_base, kTmemCols);
}

int round32(int x) { return (x + 31) & ~31; }

}  // namespace

int launch_chain_ts(const ChainArgs& a, cudaStream_t st) {
  if (a.m <= 0) return RN_OK;
  if (a.num_ops < 1 || a.num_ops > kMaxOps) return rn_set_error(RN_ERR_ARG, "chain: 1..12 ops");
  if (a.in_cols % 64 || a.in_cols < 64 || a.in_cols > 256) return rn_set_error(RN_ERR_ARG, "chain: input tile must be 64..256 columns");
  if (a.m + 512 > 0x7fffffffLL) return rn_set_error(RN_ERR_ARG, "chain: too many rows for one launch");
  TsMaps maps;
  TsParams tp;
  memset(&tp, 0, sizeof(tp));
  PairParams& p = tp.p;
  int rc;
  int mode = -1;
  if ((rc = tc::make_map(&maps.in, a.in.hi, a.m, a.in_valid, a.in.ld, kBM))) return rc;
  p.num_ops = a.num_ops;
  p.in_kb = a.in_cols / kBK;
  p.m = a.m;
  static long long* trace_buf = nullptr;
  static int trace_left = getenv("RN_CHAIN_TRACE") ? atoi(getenv("RN_CHAIN_TRACE")) : 0;
  p.trace = nullptr;
  if (trace_left > 0) {
    if (!trace_buf) cudaMalloc(&trace_buf, 128 * 8 * sizeof(long long));
    cudaMemsetAsync(trace_buf, 0, 128 * 8 * sizeof(long long), st);
    p.trace = trace_buf;
  }
  tp.last_in_op = 0;
  for (int l = 0; l < kMaxOps; ++l) {
    memset(&maps.w0[l], 0, sizeof(CUtensorMap));
    memset(&maps.w1[l], 0, sizeof(CUtensorMap));
    memset(&maps.save[l], 0, sizeof(CUtensorMap));
    if (l >= a.num_ops) continue;
    const ChainOpArgs& L = a.op[l];
    if (L.n % 16 || L.n < 16 || L.n > 256 || (L.kind == 0 && L.n != 256)) return rn_set_error(RN_ERR_ARG, "chain: bad op width");
    if ((L.kb_act != 0 && L.kb_act != 4) || (L.kb_in != 0 && L.kb_in != p.in_kb) || L.kb_act + L.kb_in == 0 ||
        (l == 0 && L.kb_act != 0))
      return rn_set_error(RN_ERR_ARG, "chain: bad K structure");
    if (L.kb_in) tp.last_in_op = l;
    const int ktot = (L.kb_act + L.kb_in) * kBK;
    // column halves: widths rounded up to 32 (rows beyond L.n are zero-filled by TMA)
    tp.half[l].n0 = L.n >= 128 ? 128 : round32(L.n);
    tp.half[l].n1 = L.n > 128 ? round32(L.n - 128) : 0;
    if ((rc = tc::make_map(&maps.w0[l], L.w, L.n, ktot, L.w_ld, tp.half[l].n0 / 2))) return rc;
    if (tp.half[l].n1 && (rc = tc::make_map(&maps.w1[l], L.w, L.n, ktot, L.w_ld, tp.half[l].n1 / 2))) return rc;
    PairOp& o = p.op[l];
    o.n = L.n; o.kb_act = L.kb_act; o.kb_in = L.kb_in;
    o.kind = L.kind; o.gepi = L.gepi; o.bias = L.bias;
    o.mask_bits = L.mask_bits; o.save_bits = L.save_bits;
    o.save = (L.kind == 0 && L.save_hi) ? 1 : 0;
    if (o.save && (rc = tc::make_map(&maps.save[l], L.save_hi, a.m, 256, 256, 32))) return rc;
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
    if ((rc = tc::set_smem(chain_ts_kernel<0>, kSmemTotal))) return rc;
    if ((rc = tc::set_smem(chain_ts_kernel<1>, kSmemTotal))) return rc;
    smem_set = true;
  }
  const int64_t tiles = (a.m + 255) / 256;
  const int max_clusters = tc::num_sms() / 2;
  const unsigned grid = 2u * (unsigned)(tiles < max_clusters ? tiles : max_clusters);
  rn_prof_begin(RN_PROF_CHAIN_TC, st, a.algo_flops);
  if (mode == 0)
    chain_ts_kernel<0><<<grid, 384, kSmemTotal, st>>>(maps, tp);
  else
    chain_ts_kernel<1><<<grid, 384, kSmemTotal, st>>>(maps, tp);
  rn_prof_end(RN_PROF_CHAIN_TC, st);
  RN_CUDA_CHECK_LAUNCH();
  if (p.trace) {
    --trace_left;
    cudaStreamSynchronize(st);
    static long long h[128 * 8];
    cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = h[0];
    printf("chain_ts trace (mode %d, %d ops, m=%lld): per (op,half): mma_wait_done mma_issued | epi_wait_begin epi_wait_end epi_end\n", mode, a.num_ops, (long long)a.m);
    for (int i = 0; i < 56; ++i)
      printf("  op %2d half %d: %8lld %8lld | %8lld %8lld %8lld\n", i / 2, i % 2, h[i * 8] - t0, h[i * 8 + 1] - t0, h[i * 8 + 2] - t0, h[i * 8 + 3] - t0,
             h[i * 8 + 4] - t0);
  }
  return RN_OK;
}

}  // namespace rn
