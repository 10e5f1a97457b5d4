// tc_probe.cu -- standalone probe of tcgen05 / UMMA shared-memory descriptor semantics on sm_100a.
// Not product code: it settles, by experiment, the facts the implicit-GEMM convolution engine is built on:
//   (1) no-swizzle ("interleave") K-major operands: which descriptor field is the K-direction stride (LBO) and
//       which is the 8-row-group stride (SBO);
//   (2) whether the 8-row groups may sit at an ARBITRARY stride (a halo'd image patch: stride = patch pitch)
//       and whether the start address may be shifted by whole rows (a filter-tap shift);
//   (3) the same for MN-major operands (weight-gradient layout) and for kind::tf32;
//   (4) TMEM accumulator readback mapping (32x32b: lane = row, column = n);
//   (5) cp.async.bulk global->shared with mbarrier complete_tx feeding an MMA.
// Each case prints max |gpu - cpu| ; "MATCH" means the hypothesis encoded in that case is right.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <unistd.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

struct Case {
  int esize;          // 2 = fp16 (kind::f16), 4 = tf32 (kind::tf32)
  int M, N, K;        // M = 128; N multiple of 16; K multiple of (32/esize)
  int a_mn_major, b_mn_major;
  // byte strides used to PLACE the data (per operand): between 8-element groups along MN and along K
  int a_mn_stride, a_k_stride, b_mn_stride, b_k_stride;
  int a_shift_rows;   // A's logical row r lives at patch row r + shift: descriptor start is advanced by the shift
  int swap_fields;    // 0: LBO = K-direction stride, SBO = MN-group stride (CUTLASS reading); 1: swapped
  int use_bulk;       // load B with cp.async.bulk instead of thread stores
  const char* name;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;            // descriptor version 1 (sm_100)
  // base_offset = 0, lbo_mode = 0, layout_type = 0 (no swizzle)
  return d;
}

// byte offset of logical element (mn, k) in the canonical no-swizzle layout.
//  K-major : core matrix = 8 mn-rows x 16 bytes of k; rows 16 B apart
//  MN-major: core matrix = 8 k-rows x 16 bytes of mn; rows 16 B apart
__device__ __host__ inline int elem_off(int mn, int k, int esize, int mn_major, int mn_stride, int k_stride) {
  int per = 16 / esize;   // elements per 16-byte chunk
  if (!mn_major) return (mn / 8) * mn_stride + (k / per) * k_stride + (mn % 8) * 16 + (k % per) * esize;
  return (mn / per) * mn_stride + (k / 8) * k_stride + (k % 8) * 16 + (mn % per) * esize;
}

__global__ void __launch_bounds__(128) probe(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B, const uint8_t* __restrict__ Bblob,
                                             float* __restrict__ D, Case c, int a_bytes, int b_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar_mma, bar_tx;
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((a_bytes + 1023) & ~1023);
  int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < a_bytes / 4; i += 128) ((uint32_t*)sA)[i] = 0x7e007e00u;   // NaN fill: reading a wrong place shows
  for (int i = tid; i < b_bytes / 4; i += 128) ((uint32_t*)sB)[i] = 0x7e007e00u;
  __syncthreads();
  // place A (logical row r at patch row r + shift)
  for (int i = tid; i < c.M * c.K; i += 128) {
    int r = i / c.K, k = i % c.K;
    int off = elem_off(r, k, c.esize, c.a_mn_major, c.a_mn_stride, c.a_k_stride) + c.a_shift_rows * 16;
    if (c.esize == 2) *(uint16_t*)(sA + off) = ((const uint16_t*)A)[i]; else *(uint32_t*)(sA + off) = ((const uint32_t*)A)[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_tx)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (!c.use_bulk) {
    for (int i = tid; i < c.N * c.K; i += 128) {
      int n = i / c.K, k = i % c.K;
      int off = elem_off(n, k, c.esize, c.b_mn_major, c.b_mn_stride, c.b_k_stride);
      if (c.esize == 2) *(uint16_t*)(sB + off) = ((const uint16_t*)B)[i]; else *(uint32_t*)(sB + off) = ((const uint32_t*)B)[i];
    }
  } else if (tid == 0) {
    // Bblob is the smem image prepared on the host: one 1-D bulk copy, completion counted on bar_tx
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_tx)), "r"(b_bytes));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sB)), "l"(Bblob), "r"(b_bytes), "r"(smem_u32(&bar_tx)) : "memory");
  }
  // TMEM allocation by warp 0 (power of two >= 32 columns)
  int ncols = 32; while (ncols < c.N) ncols <<= 1;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    if (c.use_bulk) {
      uint32_t done = 0;
      while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar_tx)) : "memory");
    }
    uint32_t afmt = c.esize == 2 ? 0u : 2u;   // F16 = 0, TF32 = 2
    uint32_t idesc = (1u << 4) | (afmt << 7) | (afmt << 10) | ((uint32_t)c.a_mn_major << 15) | ((uint32_t)c.b_mn_major << 16) |
                     ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    int kstep = 32 / c.esize;                  // K per instruction: 16 (f16) or 8 (tf32)
    int per = 16 / c.esize;
    for (int k0 = 0; k0 < c.K; k0 += kstep) {
      // advance along K: K-major -> (k0/per) chunks of k_stride; MN-major -> (k0/8) groups of k_stride
      uint32_t a_adv = c.a_mn_major ? (k0 / 8) * c.a_k_stride : (k0 / per) * c.a_k_stride;
      uint32_t b_adv = c.b_mn_major ? (k0 / 8) * c.b_k_stride : (k0 / per) * c.b_k_stride;
      uint32_t a_lbo = c.a_k_stride, a_sbo = c.a_mn_stride, b_lbo = c.b_k_stride, b_sbo = c.b_mn_stride;
      if (c.swap_fields) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
      uint64_t ad = make_desc(smem_u32(sA) + c.a_shift_rows * 16 + a_adv, a_lbo, a_sbo);
      uint64_t bd = make_desc(smem_u32(sB) + b_adv, b_lbo, b_sbo);
      uint32_t acc = k0 > 0;
      if (c.esize == 2)
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      else
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar_mma)) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // readback: warp w owns TMEM lanes 32w..32w+31; 32x32b.x16 = 16 consecutive columns per thread
  for (int c0 = 0; c0 < c.N; c0 += 16) {
    uint32_t v[16];
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * c.N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols));
}

static float h2f(uint16_t h) { __half x; memcpy(&x, &h, 2); return __half2float(x); }
static uint16_t f2h(float f) { __half x = __float2half_rn(f); uint16_t h; memcpy(&h, &x, 2); return h; }
static float tf32_rn(float f) { uint32_t u; memcpy(&u, &f, 4); u = (u + 0x1000u) & ~0x1FFFu; memcpy(&f, &u, 4); return f; }

static int run_case(const Case& c) {
  int M = c.M, N = c.N, K = c.K, es = c.esize;
  std::vector<uint8_t> hA((size_t)M * K * es), hB((size_t)N * K * es);
  std::vector<float> fA((size_t)M * K), fB((size_t)N * K), ref((size_t)M * N), got((size_t)M * N);
  srand(1234 + M + N * 3 + K * 7 + es);
  for (size_t i = 0; i < fA.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; if (es == 2) { uint16_t h = f2h(v); ((uint16_t*)hA.data())[i] = h; fA[i] = h2f(h); } else { v = tf32_rn(v); ((float*)hA.data())[i] = v; fA[i] = v; } }
  for (size_t i = 0; i < fB.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; if (es == 2) { uint16_t h = f2h(v); ((uint16_t*)hB.data())[i] = h; fB[i] = h2f(h); } else { v = tf32_rn(v); ((float*)hB.data())[i] = v; fB[i] = v; } }
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)fA[(size_t)m * K + k] * fB[(size_t)n * K + k]; ref[(size_t)m * N + n] = (float)s; }
  // operand footprints in smem
  int a_bytes = 0, b_bytes = 0;
  for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) { int o = elem_off(r, k, es, c.a_mn_major, c.a_mn_stride, c.a_k_stride) + c.a_shift_rows * 16 + es; if (o > a_bytes) a_bytes = o; }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { int o = elem_off(n, k, es, c.b_mn_major, c.b_mn_stride, c.b_k_stride) + es; if (o > b_bytes) b_bytes = o; }
  a_bytes = (a_bytes + 15) & ~15; b_bytes = (b_bytes + 15) & ~15;
  std::vector<uint8_t> blob(b_bytes, 0);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) memcpy(&blob[elem_off(n, k, es, c.b_mn_major, c.b_mn_stride, c.b_k_stride)], &hB[((size_t)n * K + k) * es], es);
  uint8_t *dA, *dB, *dBlob; float* dD;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dBlob, b_bytes)); CK(cudaMalloc(&dD, got.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBlob, blob.data(), b_bytes, cudaMemcpyHostToDevice)); CK(cudaMemset(dD, 0xFF, got.size() * 4));
  int smem = ((a_bytes + 1023) & ~1023) + b_bytes + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe<<<1, 128, smem>>>(dA, dB, dBlob, dD, c, a_bytes, b_bytes);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-58s LAUNCH/EXEC ERROR: %s\n", c.name, cudaGetErrorString(e)); return 2; }
  CK(cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; int nan = 0;
  for (size_t i = 0; i < got.size(); ++i) { if (!(got[i] == got[i])) { nan++; continue; } double d = fabs((double)got[i] - ref[i]); if (d > maxerr) maxerr = d; }
  bool ok = nan == 0 && maxerr < 1e-3 * sqrt((double)K);
  printf("%-58s max|err| %.3e  nan %5d  smem %6d B  -> %s\n", c.name, maxerr, nan, smem, ok ? "MATCH" : "differs");
  cudaFree(dA); cudaFree(dB); cudaFree(dBlob); cudaFree(dD);
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  int only = argc > 1 ? atoi(argv[1]) : -1;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  if (only < 0) printf("device %s sm_%d%d, %d SMs\n", p.name, p.major, p.minor, p.multiProcessorCount);
  // byte strides: dense K-major tile [M rows][K]: mn-group = 8 rows*16 B = 128 B when each 16-B K chunk is its own plane
  // ([k chunk][row][16 B]); k-chunk plane stride = rows * 16 B.
  const int M = 128;
  std::vector<Case> cases;
  auto kmaj = [&](int es, int N, int K, int a_rows_pitch8, int shift, int swap, int bulk, const char* nm) {
    // A planes: [k chunk][patch row][16 B]; 8-row groups a_rows_pitch8 rows apart (8 = dense); plane holds enough rows
    int a_rows = (M / 8) * a_rows_pitch8 + shift + 8;
    Case c{es, M, N, K, 0, 0, a_rows_pitch8 * 16, a_rows * 16, 128, N * 16, shift, swap, bulk, nm};
    cases.push_back(c);
  };
  kmaj(2, 64, 64, 8, 0, 0, 0, "f16 K-major dense            LBO=k SBO=mn");
  kmaj(2, 64, 64, 8, 0, 1, 0, "f16 K-major dense            LBO=mn SBO=k (swapped)");
  kmaj(2, 128, 64, 12, 0, 0, 0, "f16 K-major groups@12 rows   (halo'd patch pitch)");
  kmaj(2, 128, 64, 12, 3, 0, 0, "f16 K-major groups@12 rows +3-row tap shift");
  kmaj(2, 128, 64, 36, 37, 0, 0, "f16 K-major groups@36 rows +37-row shift (5x5 @32 wide)");
  kmaj(2, 128, 128, 8, 0, 0, 1, "f16 K-major B via cp.async.bulk + mbarrier tx");
  kmaj(4, 64, 32, 8, 0, 0, 0, "tf32 K-major dense");
  kmaj(4, 128, 32, 12, 3, 0, 0, "tf32 K-major groups@12 rows +3-row shift");
  kmaj(2, 256, 64, 8, 0, 0, 0, "f16 K-major N=256");
  // MN-major operands (weight-gradient GEMM: K = pixels): core = 8 k-rows x 16 B of mn.
  // planes: [mn chunk][k row][16 B] -> mn-chunk stride = krows*16, k-group stride = 8*16 (dense) or pitch*16
  auto mnmaj = [&](int es, int N, int K, int k_pitch8, int swap, const char* nm) {
    int krows = (K / 8) * k_pitch8 + 8;
    Case c{es, M, N, K, 1, 1, krows * 16, k_pitch8 * 16, krows * 16, k_pitch8 * 16, 0, swap, 0, nm};
    cases.push_back(c);
  };
  mnmaj(2, 64, 64, 8, 0, "f16 MN-major A,B dense       LBO=k SBO=mn");
  mnmaj(2, 64, 64, 8, 1, "f16 MN-major A,B dense       LBO=mn SBO=k (swapped)");
  mnmaj(2, 128, 64, 12, 0, "f16 MN-major k-groups@12 rows LBO=k SBO=mn");
  mnmaj(2, 128, 64, 12, 1, "f16 MN-major k-groups@12 rows (swapped)");
  mnmaj(4, 64, 32, 8, 0, "tf32 MN-major dense          LBO=k SBO=mn");
  mnmaj(4, 64, 32, 8, 1, "tf32 MN-major dense          (swapped)");
  if (only >= 0) { if (only >= (int)cases.size()) return 9; return run_case(cases[only]); }
  // a wrong hypothesis can fault and poison the CUDA context: every case gets its own process
  char self[4096]; ssize_t n = readlink("/proc/self/exe", self, sizeof(self) - 1); if (n <= 0) return 8; self[n] = 0;
  fflush(stdout);
  int bad = 0;
  for (size_t i = 0; i < cases.size(); ++i) {
    char cmd[4200]; snprintf(cmd, sizeof(cmd), "%s %zu", self, i);
    int r = system(cmd); if (r != 0) bad++;
  }
  printf("%d of %zu case(s) differ or faulted\n", bad, cases.size());
  return 0;
}
