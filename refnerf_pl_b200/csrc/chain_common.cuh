// Helpers shared by the fused chain kernels (chain_pair.cu: SS operands, two row tiles; chain_ts.cu: A operand from
// TMEM, column halves): cluster / cta_group::2 PTX wrappers, ReLU bit masks, the global-op epilogue and the kernel
// parameter structs.
#pragma once
#include "tc_common.cuh"

namespace rn {
namespace chain {
using namespace tc;

constexpr int kMaxOps = 12;

// Warp-uniform role code: the producer / MMA warps run their loops with all 32 lanes (so that ptxas keeps addresses,
// descriptors and counters in uniform registers) and elect one lane only for the TMA / MMA / commit instructions.
// Under a `lane == 0` branch every UTCHMMA / UTMALDG is wrapped in a divergence waterfall (ELECT + R2UR.BROADCAST
// per operand + loop), which costs ~170-200 cycles per MMA issue.
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- cluster / 2-CTA PTX ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Wait on a barrier of THIS CTA that threads / TMA / tcgen05.commit of either CTA of the pair arrive on.  Default semantics
// (acquire at CTA scope), as CUTLASS' ClusterBarrier::wait: what the waiter consumes is shared memory / TMEM read through the
// async proxy, ordered by the writers' fence.proxy.async / tcgen05.fence before their arrive.  (An `.acquire.cluster` wait
// compiles to SYNCS.PHASECHK + CCTL.IVALL -- an L1 invalidate on the MMA warp's critical path, five times per op.)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
// TMA load whose transaction bytes are credited to the barrier at `bar_cluster_addr` (the leader CTA's)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// bring a tile into L2 ahead of the TMA load that will need it
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// D[tmem, both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]^T
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at the same offset in both CTAs
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// a_f16 / b_f16: the operand is FP16 (format code 0) instead of BF16 (1).  Callers keep the two equal: a kind::f16 MMA
// with a_format != b_format raised an illegal-instruction fault on sm_100a
__device__ __forceinline__ uint32_t make_idesc2(int n, int a_f16 = 0, int b_f16 = 0) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= (a_f16 ? 0u : 1u) << 7;        // a_format
  d |= (b_f16 ? 0u : 1u) << 10;       // b_format
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(256 >> 4) << 24;    // M = 256 across the pair
  return d;
}

__device__ __forceinline__ void mbar_arrive_cluster_addr(uint32_t cluster_addr) {
  // default semantics (release at CTA scope), as CUTLASS' ClusterBarrier::arrive(cta_id): the data this signals is
  // this CTA's own shared memory / TMEM, ordered by fence.proxy.async / tcgen05.fence before the arrive
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// two floats -> bf16x2 with ReLU folded into the conversion (one F2FP instruction)
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_relu_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// activation-format packers of the chain epilogues: F16 = 0 bf16, 1 fp16
template <int F16>
__device__ __forceinline__ uint32_t pack_relu_act(float lo, float hi) {
  return F16 ? pack_relu_f16x2(lo, hi) : pack_relu_bf16x2(lo, hi);
}
template <int F16>
__device__ __forceinline__ uint32_t pack_act(float lo, float hi) {
  return F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) { tmem_ld32(taddr, r); }

// ReLU masks travel as 1 bit per activation: word w of a row covers columns [32w, 32w+32); inside a word the
// packed bf16 pair i (columns 2i, 2i+1) owns bits i and 16+i.
__device__ __forceinline__ uint32_t relu_bits_of(const uint32_t* packed /*16 bf16x2, non-negative*/) {
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) bits += __vminu2(packed[i], 0x00010001u) << i;   // 1 where the half is nonzero
  return bits;
}
// keep / zero the two halves of packed pair i according to bits i and 16+i
__device__ __forceinline__ uint32_t apply_relu_bits(uint32_t bits, int i, uint32_t packed) {
  const uint32_t x = i <= 7 ? (bits << (7 - i)) : (bits >> (i - 7));   // bit i -> 7, bit 16+i -> 23
  uint32_t m;   // prmt with selector msb set = replicate the sign bit of the selected byte (__byte_perm drops that bit)
  asm("prmt.b32 %0, %1, %2, 0xAA88;" : "=r"(m) : "r"(x), "r"(0u));
  return packed & m;
}

// epilogue for 16 consecutive columns of one row of a global (non-hidden) op; v holds the fp32 accumulators
// (bias already added by the caller)
template <int F16 = 0>
__device__ __forceinline__ void epi_global16(const GemmEpilogue& e, size_t row, int col, float* v) {
  if (e.relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (e.out.hi && col < e.out_cols) {
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = pack_act<F16>(v[2 * i], v[2 * i + 1]);
    stg256(reinterpret_cast<uint16_t*>(e.out.hi) + row * e.out.ld + col, h);
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

struct PairOp {
  int n;           // MMA N = output columns (multiple of 16)
  int kb_act;      // K blocks read from the resident activation tile (0 or 4)
  int kb_in;       // K blocks read from the chain input tile (streamed through the ring)
  int in2;         // the input K blocks of this op come from the SECOND input tensor (maps.in2)
  int kind;        // 0: hidden (result -> activation tile [+ TMA save]); 1: global epilogue
  int gepi;        // kind 1: which global epilogue
  int save;        // hidden: TMA-store the result through maps.save[op]
  int acc_half;    // chain_x3t.cu: the accumulator column half a one-half (global) op runs in
  const float* bias;       // forward hidden ops / global ops with a bias: [n] floats (padded to a multiple of 4)
  const uint32_t* mask_bits;   // backward hidden ops: ReLU bits (8 words per row, layout of relu_bits_index in gemm.cuh)
  uint32_t* save_bits;         // forward hidden ops: optional ReLU bits of the result
};
struct PairParams {
  int num_ops;
  int in_kb;
  int in2_sync_op;   // -1, or: the second input is a save of this launch; the op before its reader orders store -> load
  int a_f16, b_f16;  // operand formats of the MMAs (0 bf16, 1 fp16); the activation tile / outputs use a's
  int w_planes;      // chain_pair.cu: 2 = every weight K block arrives as a hi and a lo ring item (2 MMAs per K step)
  int split_order;   // chain_x3.cu: issue order of the (column half, K block) pairs of a 256-wide op, see res_order()
  int in_valid;      // chain_x3t.cu: valid (non-padding) columns of the chain input; K steps entirely beyond it are skipped
  int whatif;        // chain_x3.cu, RN_X3_WHATIF (TIMING EXPERIMENTS ONLY, results are garbage): bit 0 = the epilogue skips its
                     // shared-memory stores, bit 1 = the producer skips the weight loads, bit 2 = no TMA save stores
  float seed_scale;  // seed ops: vec * seed_scale
  int64_t m;
  long long* trace;   // debug timeline (RN_CHAIN_TRACE=<launches>): written by CTA 0 only
  PairOp op[kMaxOps];
  GemmEpilogue gepi[2];
};

}  // namespace chain
}  // namespace rn
