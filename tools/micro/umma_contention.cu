// Micro-benchmark: how fast does a CTA pair execute back-to-back tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16,
// SS operands, 128B swizzle -- the instruction of chain_pair.cu) while the other warps of both CTAs load the shared
// memory the way the chain kernel's epilogue does?  The chain kernel's in-kernel timeline
// (profiles/r01_chain_trace_train_fp16.txt) shows 145-240 cycles per MMA against the 128 the tensor pipe needs; this
// isolates the cause.  Modes of the 8 "epilogue" warps per CTA:
//   0 idle                      1 st.shared.v4 stream (activation write-back)
//   2 ld.shared.v4 broadcast    3 tcgen05.ld 32x32b.x32 stream (accumulator drain)
//   4 st.shared + tcgen05.ld    5 cp.async.bulk shared -> global stream (the saves)
//   6 TMA tensor loads global -> shared, 64 KB per round (the weight ring)   7 one tcgen05.commit per K block (ring release)
//   8 = 1 + 3 + 6 + 7 together
// Prints SM cycles per MMA for each mode.  Operands are zeros: timing does not depend on the data.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I refnerf_pl_b200/csrc tools/micro/umma_contention.cu \
//        -o tools/micro/umma_contention && tools/micro/umma_contention
#include <cstdio>
#include <cstdlib>

#include "chain_common.cuh"

int rn_set_error(int code, const char* msg) { fprintf(stderr, "error %d: %s\n", code, msg); return code; }
int rn_set_cuda_error(cudaError_t e, const char* f, int l) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), f, l); return -2; }
void rn_count_launch() {}
void rn_prof_begin(int, cudaStream_t, double) {}
void rn_prof_end(int, cudaStream_t) {}

using namespace rn::tc;
using namespace rn::chain;

constexpr int kAct = 65536;          // A operand: [128 rows x 256 K] bf16 per CTA (4 K blocks of 16 KB)
constexpr int kW = 65536;            // B operand: this CTA's half of a [256 x 256] weight (4 K blocks of 16 KB)
constexpr int kScratch = 65536;      // region the contention warps write / read
constexpr int kBars = kAct + kW + kScratch;
constexpr int kSmem = kBars + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
contention_kernel(const __grid_constant__ CUtensorMap wmap, int mode, int batches, long long* cycles, float* sink, unsigned char* gdst, int nn) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + kBars);   // MMA batch complete (multicast commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 2);
  volatile int* stop = reinterpret_cast<volatile int*>(done + 4);
  uint64_t* tbar = done + 6;      // TMA load completion (mode 6)
  uint64_t* dummy = done + 8;     // [4] per-K-block commits nobody waits for (mode 7)
  const bool m_sts = mode == 1 || mode == 4 || mode == 8, m_tld = mode == 3 || mode == 4 || mode == 8;
  const bool m_tma = mode == 6 || mode == 8, m_commit = mode == 7 || mode == 8;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t base = smem_u32(smem);
  for (int i = threadIdx.x; i < (kAct + kW + kScratch) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&done[0], 1);   // even batches
    mbar_init(&done[1], 1);   // odd batches: a waiter is never two phases behind its barrier
    mbar_init(tbar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&dummy[i], 1);
    *stop = 0;
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 1 && rank == 0) {
    // 16 MMAs per batch (one 256 x 256 x 256 layer on one row tile), alternating accumulators, one commit per batch;
    // the issuing warp waits for batch b-2 before issuing batch b (two batches in flight, like the chain's ping-pong)
    const uint32_t idesc = make_idesc2(nn);   // N = 256 (chain_pair / chain_x3 full-width ops) or 128 (column halves)
    const long long t0 = clock64();
    for (int b = 0; b < batches; ++b) {
      if (b >= 2) mbar_wait(&done[b & 1], (uint32_t)((b >> 1) - 1) & 1u);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t d = tmem_base + (uint32_t)(b & 1) * 256u;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t off = kb * 16384 + kk * 32;
            umma2_bf16(d, make_desc(base + off, 16, 1024), make_desc(base + kAct + off, 16, 1024), idesc, (kb | kk) ? 1u : 0u);
          }
          if (m_commit) umma2_commit(&dummy[kb]);
        }
        umma2_commit(&done[b & 1]);
      }
      __syncwarp();
    }
    for (int b = batches - 2; b < batches; ++b)
      if (b >= 0) mbar_wait(&done[b & 1], (uint32_t)(b >> 1) & 1u);
    const long long t1 = clock64();
    if (lane == 0) {
      cycles[blockIdx.x >> 1] = t1 - t0;
      *stop = 1;
      // tell the peer CTA's contention warps to stop as well
      asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(map_to_cta(smem_u32(const_cast<int*>(stop)), 1)), "r"(1) : "memory");
    }
  } else if (warp == 0 && m_tma) {
    // four 16 KB boxes per round into the scratch region, like four weight K blocks of the ring
    uint32_t round = 0;
    while (!*stop) {
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(tbar, 65536);
        for (int i = 0; i < 4; ++i) tma_load_2d(smem + kAct + kW + i * 16384, &wmap, tbar, 0, (int)(((round * 4 + i) & 31) * 128));
      }
      __syncwarp();
      mbar_wait(tbar, round & 1u);
      ++round;
    }
  } else if (warp >= 4 && mode != 0 && mode != 6 && mode != 7) {
    const int q = (warp - 4) & 3, h = (warp - 4) >> 2;
    const uint32_t row = base + kAct + kW + (uint32_t)((q * 32 + lane) * 128 + h * 32768);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 128);
    float acc = 0.f;
    uint32_t it = 0;
    while (!*stop) {
      if (m_sts) {
#pragma unroll
        for (int u = 0; u < 8; ++u) sts128(row + (uint32_t)(((u ^ (lane & 7)) * 16) + ((it & 1) * 16384)), it, it, it, it);
      }
      if (mode == 2) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += lds128f(base + kAct + kW + (uint32_t)(u * 16 + (it & 63) * 128)).x;
      }
      if (m_tld) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)((it & 3) * 32), r);
        tmem_ld_wait();
        acc += __uint_as_float(r[lane & 31]);
      }
      if (mode == 5) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst + ((size_t)blockIdx.x * 8 + (warp - 4)) * 4096),
                       "r"(base + kAct + kW + (uint32_t)((warp - 4) * 4096)), "r"(4096)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
        }
        __syncwarp();
      }
      ++it;
    }
    if (mode == 5 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 123.456f) sink[threadIdx.x] = acc;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc2(tmem_base, 512);
}

// usage: umma_contention [batches] [clusters] [N]   (clusters = 74 loads every SM of a B200: shows what the power cap does
// to the cycles per MMA when the whole chip runs tensor work)
int main(int argc, char** argv) {
  const int batches = argc > 1 ? atoi(argv[1]) : 2000;
  const int clusters = argc > 2 ? atoi(argv[2]) : 1;
  const int nn = argc > 3 ? atoi(argv[3]) : 256;
  long long* cyc;
  float* sink;
  unsigned char* gdst;
  cudaMalloc(&cyc, 8 * 128);
  cudaMalloc(&sink, 4096);
  cudaMalloc(&gdst, (size_t)2 * 128 * 8 * 4096);
  void* wsrc;
  cudaMalloc(&wsrc, (size_t)4096 * 64 * 2);
  cudaMemset(wsrc, 0, (size_t)4096 * 64 * 2);
  CUtensorMap wmap;
  if (make_map(&wmap, wsrc, 4096, 64, 64, 128)) return 1;
  cudaFuncSetAttribute(contention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  const char* names[9] = {"idle", "st.shared.v4 stream", "ld.shared.v4 broadcast", "tcgen05.ld stream", "st.shared + tcgen05.ld",
                          "bulk copy shared->global", "TMA loads global->shared", "commit per K block", "sts + tld + TMA + commits"};
  for (int mode = 0; mode < 9; ++mode) {
    contention_kernel<<<2 * clusters, 384, kSmem>>>(wmap, mode, batches, cyc, sink, gdst, nn);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("mode %d (%s): %s\n", mode, names[mode], cudaGetErrorString(e));
      return 1;
    }
    long long c[128];
    cudaMemcpy(c, cyc, 8 * clusters, cudaMemcpyDeviceToHost);
    long long lo = c[0], hi = c[0];
    double sum = 0;
    for (int i = 0; i < clusters; ++i) { lo = c[i] < lo ? c[i] : lo; hi = c[i] > hi ? c[i] : hi; sum += (double)c[i]; }
    const double per = (double)batches * 16;
    printf("mode %d  %-28s %7.1f cycles / MMA (min %.1f max %.1f over %d clusters; M256 N%d K16, cta_group::2, %d batches of 16)\n", mode,
           names[mode], sum / clusters / per, lo / per, hi / per, clusters, nn, batches);
  }
  return 0;
}
