// tc_rate.cu -- how fast does one SM retire tcgen05.mma (kind::f16, M = 128, K = 16) as a function of what the conv kernels vary:
// N per instruction, operand placement (dense canonical tiles vs the halo'd patch the conv kernel addresses by shifting the
// descriptor start), and one vs two issuing warps.  No data movement at all: operands sit in shared memory, every CTA issues
// `iters` MMAs back to back and times them with clock64().  The floor from the B300 notes is M*N/256 cycles per instruction
// (64 for N = 128, 128 for N = 256); anything above that here is the tensor pipe / operand fetch, not the kernel's pipeline.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_rate tools/tc_rate.cu && tools/tc_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

struct Cfg {
  int N;            // columns per instruction
  int issuers;      // 1 or 2 warps issuing (each into its own TMEM columns, own A region, shared B)
  int a_sbo, a_lbo; // A: bytes between 8-row groups / between 16-byte K chunks
  int b_sbo, b_lbo;
  int a_walk;       // 1: walk the A start address like the conv kernel walks filter taps (5x5, row stride a_sbo) and channel chunks
  int b_walk;       // 1: same for B (the transposed formulation: pixels are the N dimension)
  int iters;
  const char* name;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__global__ void __launch_bounds__(128) rate(Cfg c, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 200 * 1024 / 16; i += 128) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  // layout of the 200 KB: [A region of issuer 0: 64 KB][A region of issuer 1: 64 KB][B: 64 KB]
  if (warp < c.issuers && lane == 0) {
    const uint32_t a_base = smem_u32(smem) + warp * 65536, b_base = smem_u32(smem) + 131072;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32, K-major both
    const uint32_t d_tmem = tmem + warp * 256;
    const uint64_t a_hi = make_desc(0, c.a_lbo, c.a_sbo) & ~0x3FFFull, b_hi = make_desc(0, c.b_lbo, c.b_sbo) & ~0x3FFFull;
    long long t0 = clock64();
    int it = 0;
    while (it < c.iters) {
      for (int ky = 0; ky < 5 && it < c.iters; ++ky)
        for (int kx = 0; kx < 5; ++kx) {
          uint32_t a_tap = c.a_walk ? (uint32_t)(ky * c.a_sbo + kx * 16) : 0u;
          uint32_t b_tap = c.b_walk ? (uint32_t)(ky * c.b_sbo + kx * 16) : 0u;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks, ++it) {   // 64 channels = 4 instructions of K = 16 (two 16-byte chunks each)
            uint64_t ad = a_hi | (uint64_t)(((a_base + a_tap + ks * 2 * c.a_lbo) >> 4) & 0x3FFF);
            uint64_t bd = b_hi | (uint64_t)(((b_base + b_tap + ks * 2 * c.b_lbo) >> 4) & 0x3FFF);
            asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                         ::"r"(d_tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(it) : "memory");
          }
        }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar[warp])) : "memory");
    long long t1 = clock64();
    cycles[blockIdx.x * 2 + warp] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long* d; CK(cudaMalloc(&d, sizeof(long long) * sms * 2));
  // dense canonical K-major tile of R rows: 8-row groups 128 B apart, K chunks R*16 B apart
  // conv patch (5x5 filter, 8-pixel-wide tile): pixel rows 12*16 = 192 B apart, channel chunks one plane = rows*192 B apart
  const int P16 = 20 * 192, P32 = 36 * 192;
  Cfg cfgs[] = {
    {128, 1, 128, 2048, 128, 2048, 0, 0, 4000, "N=128  1 issuer   A dense           B dense"},
    {256, 1, 128, 2048, 128, 4096, 0, 0, 4000, "N=256  1 issuer   A dense           B dense"},
    {128, 2, 128, 2048, 128, 2048, 0, 0, 4000, "N=128  2 issuers  A dense           B dense (shared)"},
    {128, 1, 192, P16, 128, 2048, 0, 0, 4000, "N=128  1 issuer   A patch, fixed    B dense"},
    {128, 1, 192, P16, 128, 2048, 1, 0, 4000, "N=128  1 issuer   A patch, tap walk B dense"},
    {128, 2, 192, P16, 128, 2048, 1, 0, 4000, "N=128  2 issuers  A patch, tap walk B dense (shared)   <- k_conv_tc today"},
    {256, 1, 128, 2048, 192, P32, 0, 1, 4000, "N=256  1 issuer   A dense (weights) B patch 8x32 px, tap walk   <- transposed formulation"},
    {128, 1, 128, 2048, 192, P16, 0, 1, 4000, "N=128  1 issuer   A dense (weights) B patch 8x16 px, tap walk"},
    {64,  1, 128, 2048, 128, 1024, 0, 0, 4000, "N=64   1 issuer   A dense           B dense"},
  };
  printf("%d SMs, SM clock %d MHz (nominal); cycles per tcgen05.mma (M=128, K=16, fp16), mean over CTAs [min..max]\n", sms, p.clockRate / 1000);
  for (const Cfg& c : cfgs) {
    CK(cudaMemset(d, 0, sizeof(long long) * sms * 2));
    rate<<<sms, 128, 200 * 1024>>>(c, d);
    CK(cudaDeviceSynchronize());
    rate<<<sms, 128, 200 * 1024>>>(c, d);
    CK(cudaDeviceSynchronize());
    static long long h[2048]; CK(cudaMemcpy(h, d, sizeof(long long) * sms * 2, cudaMemcpyDeviceToHost));
    double sum = 0, mn = 1e30, mx = 0; int n = 0;
    for (int i = 0; i < sms; ++i) for (int w = 0; w < c.issuers; ++w) { double v = (double)h[i * 2 + w] / c.iters; sum += v; if (v < mn) mn = v; if (v > mx) mx = v; ++n; }
    double per = sum / n;                       // cycles per instruction as seen by ONE issuer
    double per_sm = per / c.issuers;            // the SM retires `issuers` instructions in that time
    double floor_c = 128.0 * c.N / 256.0;
    printf("  %-86s %7.1f cyc/issuer  %7.1f cyc/SM  floor %5.0f  -> %5.1f%% of floor rate  [%.1f..%.1f]\n", c.name, per, per_sm, floor_c, 100.0 * floor_c / per_sm, mn, mx);
  }
  cudaFree(d);
  return 0;
}
