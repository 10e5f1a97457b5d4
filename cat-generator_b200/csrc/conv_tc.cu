// conv_tc.cu -- tcgen05 implicit-GEMM convolution engine for sm_100a (stride 1, pad (k-1)/2, NHWC).
//
//   y[n,y,x,co] = bias[co] + sum_{ky,kx,ci} xq[n, y+ky-p, x+kx-p, ci] * W[(ky,kx,ci), co]
//
// Mapping onto the tensor core (facts established by tools/tc_probe.cu on B200, profiles/r01_tc_probe.txt):
//  * GEMM M = 128 output pixels = a tile 8 pixels wide x 16 rows; N = Cout (<= 256 per CTA); K = (tap, ci).
//  * A operand = the halo'd input patch, resident in shared memory for ALL k*k filter taps:
//      plane[c][row][col] of 16-byte chunks (c = 8 fp16 / 4 tf32 channels), canonical K-major NO-SWIZZLE layout:
//      core matrix = 8 consecutive pixels of one patch row x 16 B; LBO = plane stride (next 16-B channel chunk),
//      SBO = patch pitch (next image row).  A filter tap (ky,kx) is just a different descriptor start address
//      (+ (ky*pitch + kx) * 16 B): activations are fetched from L2 once per tile instead of k*k times.
//  * B operand = weights, pre-packed on the device into the exact shared-memory image of one
//      (64-channel block, tap) slice [c][co][16 B] and streamed by ONE cp.async.bulk per slice through a ring.
//  * accumulators in TMEM (fp32), read back with tcgen05.ld (lane = pixel, column = co).
//  * operands: fp16 (kind::f16) for fprop, tf32 (kind::tf32) for dgrad -- tools/precision_study.py and
//    tools/backward_precision_study.py: bf16 and unscaled fp16 gradients miss the parity targets.
// Warp roles (160 threads): warp 4 = weight-slice producer (1 lane) ; warp 5? no -- see kernel: warps 0-3 are
// the epilogue (they own the four TMEM lane quarters), warp 4 streams weights, warp 5 loads patches, warp 6 issues MMAs.
#include "ops.cuh"

namespace cg {

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
// Bounded wait: a pipeline bug must surface as a trapped kernel (CUDA error), never as a hung GPU box.
// try_wait suspends for a hardware-defined interval per call, so the bound is seconds of wall time.
#ifndef CG_MBAR_SPIN_LIMIT
#define CG_MBAR_SPIN_LIMIT (1u << 24)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0, a = smem_u32(b), spins = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (!done && ++spins > CG_MBAR_SPIN_LIMIT) { printf("catgen: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  // no-swizzle canonical layout, descriptor version 1; fields in 16-byte units (tools/tc_probe.cu case 1)
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
template <int ES>
__device__ __forceinline__ void umma(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if (ES == 2)
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

// ------------------------------------------------------------------ operand packing
// Activations: NHWC fp32 -> channel-blocked, zero-padded xq[N][Ci/PER][Hq][Wq][PER] (PER*ES = 16 bytes),
// Hq = roundup(H,16) + 2p, Wq = W + 2p, image at offset (p,p).  Rounding: RN to fp16 / RN to tf32.
template <int ES>
__global__ void k_pack_act(const float* __restrict__ x, uint8_t* __restrict__ xq, long nchunks, int H, int W, int Ci, int p, int Hq, int Wq) {
  constexpr int PER = 16 / ES;
  int Cq = Ci / PER;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nchunks; i += (long)gridDim.x * blockDim.x) {
    int xx = (int)(i % Wq); long t = i / Wq; int yy = (int)(t % Hq); t /= Hq; int c = (int)(t % Cq); long n = t / Cq;
    int iy = yy - p, ix = xx - p;
    uint4 out = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const float* s = x + ((n * H + iy) * W + ix) * Ci + c * PER;
      if (ES == 2) {
        float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
        __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w), h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
        out.x = *reinterpret_cast<uint32_t*>(&h0); out.y = *reinterpret_cast<uint32_t*>(&h1); out.z = *reinterpret_cast<uint32_t*>(&h2); out.w = *reinterpret_cast<uint32_t*>(&h3);
      } else {
        float4 a = *reinterpret_cast<const float4*>(s);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(out.x) : "f"(a.x)); asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(out.y) : "f"(a.y));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(out.z) : "f"(a.z)); asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(out.w) : "f"(a.w));
      }
    }
    reinterpret_cast<uint4*>(xq)[i] = out;
  }
}
// Weights: Wp[(tap,ci)][co] fp32 (conv_ref.cu layout) -> slices in main-loop order
//   Wq[cb][tap][sub][c (KB/PER planes)][co][PER]   with KB = 128/ES channels per slice (64 fp16 / 32 tf32),
//   cb = channel block of CB channels, sub = slice within the block.
template <int ES>
__global__ void k_pack_wslices(const float* __restrict__ Wp, uint8_t* __restrict__ Wq, long nchunks, int Ci, int Co, int kk, int CB) {
  constexpr int PER = 16 / ES, KB = 128 / ES, PL = KB / PER;   // PL = 8 planes per slice
  int nsub = CB / KB;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nchunks; i += (long)gridDim.x * blockDim.x) {
    int co = (int)(i % Co); long t = i / Co; int c = (int)(t % PL); t /= PL; int sub = (int)(t % nsub); t /= nsub; int tap = (int)(t % kk); int cb = (int)(t / kk);
    int ci0 = cb * CB + sub * KB + c * PER;
    uint4 out;
    uint32_t* o = &out.x;
    if (ES == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __half2 h = __floats2half2_rn(Wp[((long)tap * Ci + ci0 + 2 * j) * Co + co], Wp[((long)tap * Ci + ci0 + 2 * j + 1) * Co + co]);
        o[j] = *reinterpret_cast<uint32_t*>(&h);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o[j]) : "f"(Wp[((long)tap * Ci + ci0 + j) * Co + co]));
    }
    reinterpret_cast<uint4*>(Wq)[i] = out;
  }
}

// ------------------------------------------------------------------ the kernel
struct TcParams {
  const uint8_t* xq; const uint8_t* wq; const float* bias; float* y;
  int N, H, W, Ci, Co, k, p, Hq, Wq;   // Hq/Wq: padded dims of xq
  int CB, ncb;                         // channel block held in smem at once, number of blocks
  int tiles_x, tiles_y;                // tiles per image
  int NB;                              // Co columns handled by one CTA (<= 256); grid.y = Co / NB
  int S;                               // weight ring depth
  uint32_t patch_bytes, slice_bytes;   // one patch buffer, one weight slice
};

template <int ES>
__global__ void __launch_bounds__(224, 1) k_conv_tc(TcParams P) {
  constexpr int PER = 16 / ES, KB = 128 / ES, KSTEP = 32 / ES;   // channels per slice, K per MMA instruction
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_pfull[2], bar_pempty[2], bar_wfull[8], bar_wempty[8], bar_acc;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* patch0 = smem;                              // 2 patch buffers
  uint8_t* wring = smem + 2 * (size_t)P.patch_bytes;   // S weight slices

  // tile coordinates
  int tile = blockIdx.x;
  int tx = tile % P.tiles_x; tile /= P.tiles_x; int ty = tile % P.tiles_y; int n = tile / P.tiles_y;
  const int x0 = tx * 8, y0 = ty * 16;
  const int co0 = blockIdx.y * P.NB;
  const int Hp = 16 + 2 * P.p, Wp = 8 + 2 * P.p;       // patch rows / pitch (pixels)
  const int planes = P.CB / PER;                       // 16-byte channel planes per patch buffer
  const uint32_t plane_bytes = (uint32_t)Hp * Wp * 16;
  const int kk = P.k * P.k, nsub = P.CB / KB;
  const int nslices = P.ncb * kk * nsub;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_pfull[i], 1); mbar_init(&bar_pempty[i], 1); }
    for (int i = 0; i < P.S; ++i) { mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wempty[i], 1); }
    mbar_init(&bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  int ncols = 32; while (ncols < P.NB) ncols <<= 1;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;

  if (warp == 4) {
    // ===== weight-slice producer: one bulk copy per (cb, tap, sub) slice of this CTA's Co range
    if (lane == 0) {
      // slices for all Co are stored [slice][c][Co][PER]; a CTA that owns NB < Co columns copies per plane
      const int PL = KB / PER;
      for (int s = 0; s < nslices; ++s) {
        int st = s % P.S; uint32_t ph = (s / P.S) & 1;
        mbar_wait(&bar_wempty[st], ph ^ 1);
        mbar_expect_tx(&bar_wfull[st], P.slice_bytes);
        uint8_t* dst = wring + (size_t)st * P.slice_bytes;
        const uint8_t* src = P.wq + (size_t)s * PL * P.Co * 16;
        if (P.NB == P.Co) bulk_g2s(dst, src, P.slice_bytes, &bar_wfull[st]);
        else for (int c = 0; c < PL; ++c) bulk_g2s(dst + (size_t)c * P.NB * 16, src + ((size_t)c * P.Co + co0) * 16, P.NB * 16, &bar_wfull[st]);
      }
    }
  } else if (warp == 5) {
    // ===== patch producer: rows of the zero-padded, channel-blocked input; one bulk copy per (plane, row)
    for (int cb = 0; cb < P.ncb; ++cb) {
      int buf = cb & 1; uint32_t ph = (cb >> 1) & 1;
      if (lane == 0) { mbar_wait(&bar_pempty[buf], ph ^ 1); mbar_expect_tx(&bar_pfull[buf], P.patch_bytes); }
      __syncwarp();
      uint8_t* dst = patch0 + (size_t)buf * P.patch_bytes;
      const int Cq = P.Ci / PER;
      for (int r = lane; r < planes * Hp; r += 32) {
        int c = r / Hp, row = r % Hp;
        const uint8_t* src = P.xq + ((((size_t)n * Cq + (size_t)cb * planes + c) * P.Hq + (y0 + row)) * P.Wq + x0) * 16;
        bulk_g2s(dst + (size_t)c * plane_bytes + (size_t)row * Wp * 16, src, Wp * 16, &bar_pfull[buf]);
      }
    }
  } else if (warp == 6) {
    // ===== MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t fmt = ES == 2 ? 0u : 2u;   // F16 / TF32
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(P.NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t b_lbo = (uint32_t)P.NB * 16, b_sbo = 128;
      const uint32_t a_lbo = plane_bytes, a_sbo = (uint32_t)Wp * 16;
      int s = 0; uint32_t acc = 0;
      for (int cb = 0; cb < P.ncb; ++cb) {
        int buf = cb & 1;
        mbar_wait(&bar_pfull[buf], (cb >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t pbase = smem_u32(patch0 + (size_t)buf * P.patch_bytes);
        for (int tap = 0; tap < kk; ++tap) {
          const uint32_t tap_off = (uint32_t)((tap / P.k) * Wp + (tap % P.k)) * 16;
          for (int sub = 0; sub < nsub; ++sub, ++s) {
            int st = s % P.S;
            mbar_wait(&bar_wfull[st], (s / P.S) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t wbase = smem_u32(wring + (size_t)st * P.slice_bytes);
#pragma unroll
            for (int ks = 0; ks < KB / KSTEP; ++ks) {
              // each instruction consumes 2 consecutive 16-byte channel planes of A and of B
              uint64_t ad = umma_desc(pbase + (uint32_t)(sub * (KB / PER) + ks * 2) * plane_bytes + tap_off, a_lbo, a_sbo);
              uint64_t bd = umma_desc(wbase + (uint32_t)(ks * 2) * b_lbo, b_lbo, b_sbo);
              umma<ES>(tmem, ad, bd, idesc, acc);
              acc = 1;
            }
            umma_commit(&bar_wempty[st]);             // slice may be overwritten once these MMAs retire
          }
        }
        umma_commit(&bar_pempty[buf]);
      }
      umma_commit(&bar_acc);
    }
  }

  if (warp < 4) {
    // ===== epilogue: TMEM -> registers -> (+bias) -> NHWC fp32.  warp w owns TMEM lanes 32w..32w+31 = pixels.
    mbar_wait(&bar_acc, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    int m = warp * 32 + lane;
    int oy = y0 + (m >> 3), ox = x0 + (m & 7);
    bool valid = oy < P.H && ox < P.W;
    float* out = P.y + (((size_t)n * P.H + oy) * P.W + ox) * P.Co + co0;
    for (int c0 = 0; c0 < P.NB; c0 += 16) {
      uint32_t v[16];
      uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 o;
          o.x = __uint_as_float(v[j]) + (P.bias ? P.bias[co0 + c0 + j] : 0.f);
          o.y = __uint_as_float(v[j + 1]) + (P.bias ? P.bias[co0 + c0 + j + 1] : 0.f);
          o.z = __uint_as_float(v[j + 2]) + (P.bias ? P.bias[co0 + c0 + j + 2] : 0.f);
          o.w = __uint_as_float(v[j + 3]) + (P.bias ? P.bias[co0 + c0 + j + 3] : 0.f);
          *reinterpret_cast<float4*>(out + c0 + j) = o;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols));
}

// ------------------------------------------------------------------ host side
static bool tc_shape_ok(int H, int W, int Ci, int Co, int k, int ES) {
  int KB = 128 / ES;
  return (k == 3 || k == 5 || k == 7) && W % 8 == 0 && H % 8 == 0 && Ci % KB == 0 && Co % 16 == 0 && Co >= 16;
}

template <int ES>
static int conv_tc_run(const float* x, const float* Wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co, int k) {
  constexpr int PER = 16 / ES, KB = 128 / ES;
  const int p = (k - 1) / 2, kk = k * k;
  const int Hq = ((H + 15) / 16) * 16 + 2 * p, Wq = W + 2 * p;
  const int Hp = 16 + 2 * p, Wpx = 8 + 2 * p;
  TcParams P{};
  // N columns per CTA: largest multiple of 16 that divides Co and is <= 256
  int NB = Co; while (NB > 256 || Co % NB) NB -= 16;
  // channel block: as many 128-byte slices as keep 2 patch buffers + a >=3-deep weight ring under ~200 KB
  size_t slice_bytes = (size_t)(KB / PER) * NB * 16;
  int CB = KB;
  for (int cand = Ci; cand >= KB; cand -= KB) {
    if (Ci % cand) continue;
    size_t pb = (size_t)(cand / PER) * Hp * Wpx * 16;
    if (2 * pb + 3 * slice_bytes <= 200 * 1024) { CB = cand; break; }
  }
  size_t patch_bytes = (size_t)(CB / PER) * Hp * Wpx * 16;
  int S = (int)((212 * 1024 - 2 * patch_bytes) / slice_bytes); if (S > 8) S = 8;
  if (S < 2) return CG_ERR_UNSUPPORTED;
  size_t smem = 2 * patch_bytes + (size_t)S * slice_bytes;
  // operand buffers: activations then weight slices (16-byte aligned)
  size_t xq_bytes = (size_t)N * (Ci / PER) * Hq * Wq * 16, wq_bytes = (size_t)kk * Ci * Co * ES;
  uint8_t* ws = (uint8_t*)workspace2(xq_bytes + wq_bytes + 256);
  if (!ws) return CG_ERR_CUDA;
  uint8_t* xq = ws; uint8_t* wq = ws + ((xq_bytes + 255) & ~(size_t)255);
  long nx = (long)(xq_bytes / 16), nw = (long)(wq_bytes / 16);
  CG_LAUNCH(k_pack_act<ES>, grid1d(nx, 256), 256, 0, x, xq, nx, H, W, Ci, p, Hq, Wq);
  CG_LAUNCH(k_pack_wslices<ES>, grid1d(nw, 256), 256, 0, Wp, wq, nw, Ci, Co, kk, CB);
  P.xq = xq; P.wq = wq; P.bias = bias; P.y = y;
  P.N = N; P.H = H; P.W = W; P.Ci = Ci; P.Co = Co; P.k = k; P.p = p; P.Hq = Hq; P.Wq = Wq;
  P.CB = CB; P.ncb = Ci / CB; P.tiles_x = W / 8; P.tiles_y = (H + 15) / 16; P.NB = NB; P.S = S;
  P.patch_bytes = (uint32_t)patch_bytes; P.slice_bytes = (uint32_t)slice_bytes;
  static bool attr_done[2] = {false, false};
  if (!attr_done[ES == 2 ? 0 : 1]) {
    CG_CUDA(cudaFuncSetAttribute(k_conv_tc<ES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_done[ES == 2 ? 0 : 1] = true;
  }
  dim3 grid(N * P.tiles_x * P.tiles_y, Co / NB);
  ctx().next_flops = 2.0 * (double)N * H * W * Co * kk * Ci;
  ctx().next_bytes = (double)xq_bytes + (double)wq_bytes + 4.0 * (double)N * H * W * Co;
  CG_LAUNCH(k_conv_tc<ES>, grid, 224, smem, P);
  return CG_OK;
}

// precision per direction: 0 = fp16 operands (fprop), 1 = tf32 operands (dgrad / anything gradient-valued)
static thread_local int g_tc_grad_operands = 0;
void conv_tc_set_gradient_operands(int on) { g_tc_grad_operands = on; }

int conv_fwd_tc(const float* x, const float* Wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co, int k) {
  int ES = g_tc_grad_operands ? 4 : 2;
  if (!tc_shape_ok(H, W, Ci, Co, k, ES)) return CG_ERR_UNSUPPORTED;
  if ((((uintptr_t)x | (uintptr_t)y) & 15) != 0) return CG_ERR_UNSUPPORTED;
  return ES == 2 ? conv_tc_run<2>(x, Wp, bias, y, N, H, W, Ci, Co, k) : conv_tc_run<4>(x, Wp, bias, y, N, H, W, Ci, Co, k);
}
int conv_wgrad_tc(const float*, const float*, float*, int, int, int, int, int, int) { return CG_ERR_UNSUPPORTED; }   // next: MN-major operands

}  // namespace cg
