// tc_pair_probe.cu -- ROUND-2 PREPARATION, WRITTEN WITHOUT A GPU, NEVER RUN.  Compiles for sm_100a; every claim in this header is
// what the probe is meant to ESTABLISH, not something already measured.
//
// Why: the conv kernels' tensor pipe sits at 46-55 % because a CTA must take in a full weight slice for every 256 pixels
// (36.7 B/cycle/SM needed at the MMA floor, 23.5 B/cycle/SM obtained).  tcgen05.mma.cta_group::2 lets the
// two SMs of a TPC run ONE M=256 MMA whose B operand (the weights) is split between their shared memories: each SM then
// fetches HALF of every weight slice for the same 128 pixel rows of its own.  Before touching the product kernel, this probe
// checks the mechanics in isolation, with operands written by ordinary stores (no TMA):
//
//   1. correctness: D[256 x N] = A[256 x K] * B[N x K]^T, fp16 operands, K-major NO-SWIZZLE canonical layout (the layout the conv
//      kernels use, profiles/r01_tc_probe.txt), CTA r holding A rows [128r, 128r+128) and B rows (columns of D) [N/2*r, N/2*r+N/2);
//      each CTA reads back its own 128 TMEM lanes x N columns and the host compares with a double-precision product;
//   2. rate: cycles per cta_group::2 MMA for N = 128 and 256 issued back to back by the leader (compare tools/tc_rate.cu:
//      64.1 / 128.1 cycles per cta_group::1 MMA per SM).
//
// Open questions the first run answers (each is a guess below, marked GUESS):
//   * whether BOTH CTAs execute tcgen05.alloc.cta_group::2 (as CUTLASS' Allocator2Sm is called in DeepGEMM) -- done here;
//   * that the B descriptor describes the LOCAL half (N/2 rows) at the same shared-memory offset in both CTAs and the
//     instruction descriptor carries the FULL N and M = 256;
//   * that tcgen05.commit...multicast::cluster with mask 0b11 arrives on the barrier at the same offset in both CTAs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_pair_probe tools/tc_pair_probe.cu && tools/tc_pair_probe
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;     // K-direction stride between 16-byte chunks
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;     // stride between 8-row groups
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  return d;
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// dense canonical K-major tile of R rows: element (r, k) at (r/8)*128 + (k/8)*(R*16) + (r%8)*16 + (k%8)*2
__device__ __host__ inline int kmajor_off(int r, int k, int R) { return (r / 8) * 128 + (k / 8) * (R * 16) + (r % 8) * 16 + (k % 8) * 2; }

struct Params { int N, K, iters; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
pair_probe(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, long long* __restrict__ cycles, Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar_done;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();                 // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int N = P.N, K = P.K, Nh = N / 2;
  uint8_t* sA = smem;                                       // [128 rows x K] K-major
  uint8_t* sB = smem + 128 * K * 2;                         // [N/2 rows x K] K-major: this CTA's half of B
  // ---- operands by ordinary stores (generic proxy), then made visible to the async proxy
  for (int i = tid; i < 128 * K; i += 128) {
    int r = i / K, k = i % K;
    *(__half*)(sA + kmajor_off(r, k, 128)) = A[((size_t)pair * 256 + rank * 128 + r) * K + k];
  }
  for (int i = tid; i < Nh * K; i += 128) {
    int n = i / K, k = i % K;
    *(__half*)(sB + kmajor_off(n, k, Nh)) = B[((size_t)pair * N + rank * Nh + n) * K + k];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  int ncols = 32; while (ncols < N) ncols <<= 1;
  if (warp == 0) {   // GUESS: both CTAs of the pair execute the 2-CTA allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  cluster_sync();                                           // both CTAs' operands, barriers and TMEM are ready
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;

  long long t0 = 0, t1 = 0;
  if (rank == 0 && tid == 0) {
    // instruction descriptor: D fp32 (bit 4), A/B fp16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24 with M = 256 for the pair
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t a_lbo = 128 * 16, b_lbo = (uint32_t)Nh * 16;          // K chunk stride = rows * 16 bytes (dense tile)
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    t0 = clock64();
    for (int it = 0; it < P.iters; ++it)
      for (int k0 = 0; k0 < K; k0 += 16) {
        uint64_t ad = make_desc(a0 + (k0 / 8) * a_lbo, a_lbo, 128);
        uint64_t bd = make_desc(b0 + (k0 / 8) * b_lbo, b_lbo, 128);
        uint32_t acc = (it > 0 || k0 > 0) ? 1u : 0u;
        if (P.iters > 1 && it > 0) acc = 1u;
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p; }"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    // completion to the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar_done)), "h"((uint16_t)0b11) : "memory");
  }
  {
    uint32_t done = 0, spins = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar_done)) : "memory");
      if (!done && ++spins > (1u << 24)) { if (lane == 0) printf("pair_probe: barrier timeout, block %d rank %u warp %d\n", blockIdx.x, rank, warp); __trap(); }
    }
  }
  if (rank == 0 && tid == 0) { t1 = clock64(); cycles[pair] = t1 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // each CTA reads its own 128 lanes x N columns (rows rank*128 + lane of D)
  if (P.iters == 1)
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      float* out = D + ((size_t)pair * 256 + rank * 128 + warp * 32 + lane) * N + c0;
      for (int j = 0; j < 16; ++j) out[j] = __uint_as_float(v[j]);
    }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  cluster_sync();                                           // nobody leaves while the peer may still read its shared memory / TMEM
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols));
}

static int run(int N, int K, int iters, int pairs, bool check) {
  size_t na = (size_t)pairs * 256 * K, nb = (size_t)pairs * N * K, nd = (size_t)pairs * 256 * N;
  std::vector<__half> hA(na), hB(nb);
  std::vector<float> fA(na), fB(nb);
  srand(1234 + N + K);
  for (size_t i = 0; i < na; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half_rn(v); fA[i] = __half2float(hA[i]); }
  for (size_t i = 0; i < nb; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half_rn(v); fB[i] = __half2float(hB[i]); }
  __half *dA, *dB; float* dD; long long* dC;
  CK(cudaMalloc(&dA, na * 2)); CK(cudaMalloc(&dB, nb * 2)); CK(cudaMalloc(&dD, nd * 4)); CK(cudaMalloc(&dC, pairs * sizeof(long long)));
  CK(cudaMemcpy(dA, hA.data(), na * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), nb * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, nd * 4));
  size_t smem = (size_t)128 * K * 2 + (size_t)(N / 2) * K * 2 + 1024;
  CK(cudaFuncSetAttribute(pair_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  Params P{N, K, iters};
  pair_probe<<<2 * pairs, 128, smem>>>(dA, dB, dD, dC, P);
  CK(cudaDeviceSynchronize());
  std::vector<long long> hc(pairs); CK(cudaMemcpy(hc.data(), dC, pairs * sizeof(long long), cudaMemcpyDeviceToHost));
  if (check) {
    std::vector<float> hD(nd); CK(cudaMemcpy(hD.data(), dD, nd * 4, cudaMemcpyDeviceToHost));
    double worst = 0; long bad = 0;
    for (int p = 0; p < pairs; ++p)
      for (int m = 0; m < 256; ++m)
        for (int n = 0; n < N; ++n) {
          double s = 0;
          for (int k = 0; k < K; ++k) s += (double)fA[((size_t)p * 256 + m) * K + k] * fB[((size_t)p * N + n) * K + k];
          double e = fabs(s - hD[((size_t)p * 256 + m) * N + n]);
          if (!(e <= 1e-3)) ++bad;
          if (e > worst || e != e) worst = e;
        }
    printf("correctness  M=256 (2 x 128) N=%3d K=%3d, %d pair(s): max |err| %.3e, %ld of %zu elements off by more than 1e-3  %s\n",
           N, K, pairs, worst, bad, nd, bad ? "FAIL" : "ok");
  } else {
    double sum = 0; for (long long c : hc) sum += (double)c;
    double per = sum / pairs / ((double)iters * (K / 16));
    printf("rate         cta_group::2 N=%3d: %.1f cycles per M=256 MMA issued by the leader (cta_group::1 on one SM: %d for M=128)\n", N, per, N / 2);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return 0;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  if (run(128, 64, 1, 1, true)) return 1;
  if (run(128, 64, 1, 3, true)) return 1;
  if (run(256, 64, 1, 2, true)) return 1;
  if (run(64, 32, 1, 1, true)) return 1;
  if (run(128, 64, 1000, p.multiProcessorCount / 2, false)) return 1;
  if (run(256, 64, 1000, p.multiProcessorCount / 2, false)) return 1;
  return 0;
}
