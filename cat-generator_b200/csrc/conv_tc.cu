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
//  * operands: fp16 (kind::f16) everywhere; gradient-valued operands are multiplied by a per-tensor power of two first
//    (tools/precision_study.py, tools/backward_precision_study.py: bf16 and unscaled fp16 gradients miss the parity
//    targets, scaled fp16 equals tf32).  The tf32 instantiation stays behind CATGEN_DGRAD_TF32=1.
// Warp roles (256 threads): warps 0-3 = epilogue (they own the four TMEM lane quarters; stores go through a shared-memory
// tile so that they leave as whole 128-byte lines), warp 4 streams weight slices, warp 5 loads patches (one tiled TMA per
// channel block), warps 6 and 7 issue the MMAs of tile 0 / tile 1 of the CTA (one thread retires an MMA per ~84 cycles,
// two reach the pipe's 64: tools/tc_rate.cu).
// File layout: pack kernels (activations, weight slices, the fused G producer, the one-launch model repack) ; k_conv_tc
// (fprop + dgrad) ; host planning / launch ; gradient-operand statistics and packing ; k_wgrad_tc ; conv_bwd_tc (one packed
// gradient operand feeds dgrad on the issuing stream and wgrad on its side stream).
// Experiment knobs (not for production): CATGEN_TC_RING (cap the weight ring), CATGEN_TC_ROT (per-CTA tap rotation),
// CATGEN_TC_DBG=1|2 (per-CTA clock64 timeline to stderr; 2 also drops the epilogue stores).
#include <unordered_map>
#include "ops.cuh"
#include <cuda.h>
#include <stdlib.h>   // CUtensorMap types only; the encoder is resolved through the runtime (no -lcuda)

namespace cg {

// ------------------------------------------------------------------ TMA tensor maps
// Measured (profiles/r01_launches_tc_engine.txt): fetching a patch as hundreds of 192-byte cp.async.bulk row copies
// runs at ~66 ns per copy per SM -- the copy engine is request-rate bound -- and made every tensor-core kernel
// 5-20x slower than its MMA time.  A patch is therefore fetched by ONE 4-D tiled TMA: box = [Wp pixels of 16 B]
// x [Hp rows] x [planes] x [1 image] of the channel-blocked, zero-padded tensor xq[N][Cq][Hq][Wq][16 B].
typedef CUresult (*cg_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static cg_tmap_encode_fn tmap_encoder() {
  static cg_tmap_encode_fn fn = nullptr; static bool tried = false;
  if (!tried) {
    tried = true; void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (cg_tmap_encode_fn)p;
  }
  return fn;
}
static int make_patch_tmap(CUtensorMap* tm, const void* xq, int ES, int N, int Cq, int Hq, int Wq, int Hp, int Wp, int planes) {
  cg_tmap_encode_fn enc = tmap_encoder();
  if (!enc) return set_err(CG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const int PER = 16 / ES;
  cuuint64_t dims[4] = {(cuuint64_t)Wq * PER, (cuuint64_t)Hq, (cuuint64_t)Cq, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)Wq * 16, (cuuint64_t)Hq * Wq * 16, (cuuint64_t)Cq * Hq * Wq * 16};
  cuuint32_t box[4] = {(cuuint32_t)(Wp * PER), (cuuint32_t)Hp, (cuuint32_t)planes, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, ES == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(xq), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(CG_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (dims %d x %d x %d x %d, box %d x %d x %d)", (int)r, Wq * PER, Hq, Cq, N, Wp * PER, Hp, planes);
  return CG_OK;
}

// DUO tiles (8 x 8 images): one 128-row tile = TWO images, their rows interleaved in the shared-memory patch as [plane][h][j = image of the pair][w],
// so that the tile's sixteen 8-pixel row groups (h, j) sit at a constant stride and a filter tap is still a shifted start address.  With the
// plain 8 x 16 tile an 8-row image filled half of every MMA's M rows (models.lua:206 512 -> 512 at 8 x 8: 373 TFLOP/s).  5-D tensor map
// (w, j, h, plane, pair); the j stride (one image) is larger than the h and plane strides behind it -- TMA only adds coordinate * stride.
static int make_patch_tmap_duo(CUtensorMap* tm, const void* xq, int N, int Cq, int Hq, int Wq, int Hp, int Wp, int planes) {
  cg_tmap_encode_fn enc = tmap_encoder();
  if (!enc) return set_err(CG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t img = (cuuint64_t)Cq * Hq * Wq * 16;
  cuuint64_t dims[5] = {(cuuint64_t)Wq * 8, 2, (cuuint64_t)Hq, (cuuint64_t)Cq, (cuuint64_t)N / 2};
  cuuint64_t strides[4] = {img, (cuuint64_t)Wq * 16, (cuuint64_t)Hq * Wq * 16, 2 * img};
  cuuint32_t box[5] = {(cuuint32_t)(Wp * 8), 2, (cuuint32_t)Hp, (cuuint32_t)planes, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(xq), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return CG_ERR_UNSUPPORTED;   // the caller falls back to the plain tile
  return CG_OK;
}

// Weight slices as a 2-D tensor of 8-byte elements: row = one 16-byte-per-column plane of a slice ([Cop] x 16 B), rows = planes in
// stream order.  A box of [NB columns] x [4 planes] is one 32-channel slice of a CTA's column block, landing as [c][NB][16 B].
// Measured (gpurun_out/r02_d_conv3_dbg.txt): the same slices fetched with 1-D cp.async.bulk copies arrived at ~10 B/cycle/SM (8 KB copies;
// ~14 B/cycle with round 1's 16 KB copies) whatever the ring depth -- the issuers waited 119-171k of 371k cycles for weights while the
// producer waited 167k cycles for free slots -- i.e. the 1-D bulk path, not L2 or the crossbar, was what starved the tensor pipe.
static int make_wslice_tmap(CUtensorMap* tm, const void* wq, int Cop, long planes, int NB) {
  cg_tmap_encode_fn enc = tmap_encoder();
  if (!enc) return set_err(CG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)Cop * 2, (cuuint64_t)planes};
  cuuint64_t strides[1] = {(cuuint64_t)Cop * 16};
  cuuint32_t box[2] = {(cuuint32_t)NB * 2, 4}, estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(wq), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(CG_ERR_CUDA, "cuTensorMapEncodeTiled (weight slices) failed with %d (Cop %d, planes %ld, NB %d)", (int)r, Cop, planes, NB);
  return CG_OK;
}

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ void tma_patch_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tma_patch_5d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(tm), "r"(c0), "r"(c1), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
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
// Descriptor in two words: hi = SBO | version (constant per operand), lo = start address | LBO << 16.  Between consecutive MMAs
// only the 14-bit start-address field moves, so the issuing thread advances `lo` with ONE add.  Measured with ncu's source page
// (profiles/r01_ncu_conv3.txt): building both 64-bit descriptors from scratch plus runtime divisions cost 31 instructions per
// MMA on the single issuing thread -- the kernel was issue-bound (tensor pipe 24%) while its operands were always ready.
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFF) | (1u << 14); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
template <int ES>
__device__ __forceinline__ void umma(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if (ES == 2)
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

// Fused producer of a conv layer's fp16 operand inside G:  y = PReLU(BN(x))  ->  nearest 2x upsample  ->  blocked/padded fp16
// (the k_pack_act<2> layout).  One pass over the conv output instead of four (BN apply, PReLU, upsample, pack), and the fp32
// PReLU output and its upsampled copy are never materialised: the packed operand is what both this layer's forward and its
// weight gradient read.  bn (PReLU's input, needed by PReLU's backward) is written once per source element when asked for.
// mean == nullptr: no batch norm (the stage after nn.Linear).  Arithmetic order matches k_bn_apply / k_prelu_fwd.
__global__ void k_bn_prelu_up_pack(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ pw,
                                   float* __restrict__ bn_out, uint8_t* __restrict__ xq, long nchunks, int h, int w, int C, int up, int p, int Hq, int Wq) {
  const int Cq = C / 8, H = h << up, W = w << up;
  const float a = *pw;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nchunks; i += (long)gridDim.x * blockDim.x) {
    int xx = (int)(i % Wq); long t = i / Wq; int yy = (int)(t % Hq); t /= Hq; int c = (int)(t % Cq); long n = t / Cq;
    int iy = yy - p, ix = xx - p;
    uint4 out = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      long off = ((n * h + (iy >> up)) * w + (ix >> up)) * C + c * 8;
      float4 v0 = *reinterpret_cast<const float4*>(x + off), v1 = *reinterpret_cast<const float4*>(x + off + 4);
      float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      if (mean) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { int cc = c * 8 + j; v[j] = (v[j] - mean[cc]) * invstd[cc] * gamma[cc] + beta[cc]; }
        if (bn_out && (!up || (((iy | ix) & 1) == 0))) {
          *reinterpret_cast<float4*>(bn_out + off) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(bn_out + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : a * v[j];
      __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]), h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
      out.x = *reinterpret_cast<uint32_t*>(&h0); out.y = *reinterpret_cast<uint32_t*>(&h1); out.z = *reinterpret_cast<uint32_t*>(&h2); out.w = *reinterpret_cast<uint32_t*>(&h3);
    }
    reinterpret_cast<uint4*>(xq)[i] = out;
  }
}


// ------------------------------------------------------------------ operand packing
// Activations: NHWC fp32 -> channel-blocked, zero-padded xq[N][Ci/PER][Hq][Wq][PER] (PER*ES = 16 bytes),
// Hq = roundup(H,16) + 2p, Wq = W + 2p, image at offset (p,p).  Rounding: RN to fp16 / RN to tf32.
// Ci = real channel count of x, Cip = padded count (multiple of the slice width): channels >= Ci are zero.
template <int ES>
__global__ void k_pack_act(const float* __restrict__ x, uint8_t* __restrict__ xq, long nchunks, int H, int W, int Ci, int Cip, int p, int Hq, int Wq,
                           const float* __restrict__ scale2) {
  constexpr int PER = 16 / ES;
  int Cq = Cip / PER;
  const bool fast = (Ci % PER) == 0;
  const float sc = scale2 ? scale2[0] : 1.f;   // per-tensor power of two for gradient-valued inputs (exact in fp32)
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nchunks; i += (long)gridDim.x * blockDim.x) {
    int xx = (int)(i % Wq); long t = i / Wq; int yy = (int)(t % Hq); t /= Hq; int c = (int)(t % Cq); long n = t / Cq;
    int iy = yy - p, ix = xx - p;
    uint4 out = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W && c * PER < Ci) {
      const float* s = x + ((n * H + iy) * W + ix) * Ci + c * PER;
      if (!fast) {   // ragged channel count (1, 3): scalar gather with zero fill
        float v[8];
#pragma unroll
        for (int j = 0; j < PER; ++j) v[j] = (c * PER + j < Ci) ? s[j] * sc : 0.f;
        if (ES == 2) {
          __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]), h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
          out.x = *reinterpret_cast<uint32_t*>(&h0); out.y = *reinterpret_cast<uint32_t*>(&h1); out.z = *reinterpret_cast<uint32_t*>(&h2); out.w = *reinterpret_cast<uint32_t*>(&h3);
        } else {
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(out.x) : "f"(v[0])); asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(out.y) : "f"(v[1]));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(out.z) : "f"(v[2])); asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(out.w) : "f"(v[3]));
        }
      } else if (ES == 2) {
        float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
        __half2 h0 = __floats2half2_rn(a.x * sc, a.y * sc), h1 = __floats2half2_rn(a.z * sc, a.w * sc), h2 = __floats2half2_rn(b.x * sc, b.y * sc), h3 = __floats2half2_rn(b.z * sc, b.w * sc);
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
// Ci/Co = real sizes of Wp, Cop = padded column count of the slices; entries with ci >= Ci or co >= Co are zero.
template <int ES>
__global__ void k_pack_wslices(const float* __restrict__ Wp, uint8_t* __restrict__ Wq, long nchunks, int Ci, int Co, int Cop, int kk, int CB) {
  constexpr int PER = 16 / ES, KB = 128 / ES, PL = KB / PER;   // PL = 8 planes per slice
  int nsub = CB / KB;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nchunks; i += (long)gridDim.x * blockDim.x) {
    int co = (int)(i % Cop); long t = i / Cop; int c = (int)(t % PL); t /= PL; int sub = (int)(t % nsub); t /= nsub; int tap = (int)(t % kk); int cb = (int)(t / kk);
    int ci0 = cb * CB + sub * KB + c * PER;
    float v[8];
#pragma unroll
    for (int j = 0; j < PER; ++j) v[j] = (co < Co && ci0 + j < Ci) ? Wp[((long)tap * Ci + ci0 + j) * Co + co] : 0.f;
    uint4 out;
    uint32_t* o = &out.x;
    if (ES == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]); o[j] = *reinterpret_cast<uint32_t*>(&h); }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o[j]) : "f"(v[j]));
    }
    reinterpret_cast<uint4*>(Wq)[i] = out;
  }
}

// Whole-model repack: a flat grid, each layer owning a run of blocks proportional to its weight count; phases run over disjoint outputs so no ordering is needed inside the launch.
// The fp16 slices are gathered straight from the Torch-layout weight (same values as k_pack_wslices<2> applied to Wp / Wd).
__device__ __forceinline__ void pack_slices_job(const float* __restrict__ W, uint8_t* __restrict__ Wq, const ConvSpec& s, int dgrad, int CB, long tid, long nthreads) {
  const int k = s.k, kk = k * k;
  const int Cin = dgrad ? s.Co : s.Ci, Cout = dgrad ? s.Ci : s.Co;         // the forward conv this operand feeds: Cin -> Cout
  const int Cop = ((Cout + 15) / 16) * 16;
  if (CB == 32) {   // k_conv_ps: 32-channel slices in stream order [cb][tap][4 planes][Cop][8], only the blocks that hold data
    const int ncb = (Cin + 31) / 32;
    const long nchunks = (long)kk * ncb * 4 * Cop;
    for (long i = tid; i < nchunks; i += nthreads) {
      int co = (int)(i % Cop); long t = i / Cop; int c = (int)(t % 4); t /= 4; int tap = (int)(t % kk); int cb = (int)(t / kk);
      int ci0 = cb * 32 + c * 8;
      int ky = tap / k, kx = tap % k; if (dgrad) { ky = k - 1 - ky; kx = k - 1 - kx; }
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int ci = ci0 + j;
        v[j] = (co < Cout && ci < Cin) ? W[torch_index(k, s.Ci, s.Co, s.in_hw, s.out_hw, ky, kx, dgrad ? co : ci, dgrad ? ci : co)] : 0.f;
      }
      uint4 out; uint32_t* o = &out.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) { __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]); o[j] = *reinterpret_cast<uint32_t*>(&h); }
      reinterpret_cast<uint4*>(Wq)[i] = out;
    }
    return;
  }
  const int Cip = ((Cin + 63) / 64) * 64, nsub = CB / 64;
  const long nchunks = (long)kk * Cip * Cop / 8;
  for (long i = tid; i < nchunks; i += nthreads) {
    int co = (int)(i % Cop); long t = i / Cop; int c = (int)(t % 8); t /= 8; int sub = (int)(t % nsub); t /= nsub; int tap = (int)(t % kk); int cb = (int)(t / kk);
    int ci0 = cb * CB + sub * 64 + c * 8;
    int ky = tap / k, kx = tap % k; if (dgrad) { ky = k - 1 - ky; kx = k - 1 - kx; }
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int ci = ci0 + j;
      v[j] = (co < Cout && ci < Cin) ? W[torch_index(k, s.Ci, s.Co, s.in_hw, s.out_hw, ky, kx, dgrad ? co : ci, dgrad ? ci : co)] : 0.f;
    }
    uint4 out; uint32_t* o = &out.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) { __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]); o[j] = *reinterpret_cast<uint32_t*>(&h); }
    reinterpret_cast<uint4*>(Wq)[i] = out;
  }
}
// mode bit 0: the fp32 operands Wp / Wd (only the CUDA-core fallback kernels read them) ; bit 1: bias + the fp16 weight slices of both directions
__global__ void k_repack_model(const PackJob* __restrict__ jobs, int njobs, int mode) {
  int j = 0;
  while (j + 1 < njobs && (int)blockIdx.x >= jobs[j + 1].blk0) ++j;
  const PackJob J = jobs[j];
  const ConvSpec s = J.s;
  const long tid = (blockIdx.x - J.blk0) * (long)blockDim.x + threadIdx.x, nth = (long)J.nblk * blockDim.x;
  const long n = (long)s.k * s.k * s.Ci * s.Co;
  const bool m32 = (mode & 1) || J.always32;
  if (m32) for (long i = tid; i < n; i += nth) {     // Wp[(ky,kx,ci)][co]
    int co = (int)(i % s.Co); long r = i / s.Co; int ci = (int)(r % s.Ci); int tap = (int)(r / s.Ci);
    J.Wp[i] = J.W[torch_index(s.k, s.Ci, s.Co, s.in_hw, s.out_hw, tap / s.k, tap % s.k, ci, co)];
  }
  if (m32 && J.need_dgrad) for (long i = tid; i < n; i += nth) {     // Wd[(ky,kx,co)][ci], taps flipped
    int ci = (int)(i % s.Ci); long r = i / s.Ci; int co = (int)(r % s.Co); int tap = (int)(r / s.Co);
    J.Wd[i] = J.W[torch_index(s.k, s.Ci, s.Co, s.in_hw, s.out_hw, s.k - 1 - tap / s.k, s.k - 1 - tap % s.k, ci, co)];
  }
  if (!(mode & 2)) return;
  for (long cop = tid; cop < s.Co; cop += nth) {
    int Cov = s.Co / s.out_hw;
    J.bp[cop] = J.b[s.out_hw == 1 ? (int)cop : (int)(cop % Cov) * s.out_hw + (int)(cop / Cov)];
  }
  if (J.wqf) pack_slices_job(J.W, J.wqf, s, 0, J.CBf, tid, nth);
  if (J.wqd) pack_slices_job(J.W, J.wqd, s, 1, J.CBd, tid, nth);
}
int repack_model(const PackJob* jobs_dev, int njobs, int total_blocks, int mode) {
  CG_LAUNCH(k_repack_model, total_blocks, 256, 0, jobs_dev, njobs, mode);
  return CG_OK;
}

// ------------------------------------------------------------------ the kernel
struct TcParams {
  int nostore;      // experiments (CATGEN_TC_DBG=2): the epilogue reads TMEM but does not store (timing only, wrong results)
  long long* dbg;   // experiments (CATGEN_TC_DBG=1): per-CTA clock64 timeline, 8 values
  int rot;          // 1: each CTA walks the filter taps from its own starting tap (spreads the CTAs' weight-slice reads over L2)
  const uint8_t* xq; const uint8_t* wq; const float* bias; float* y;
  const float* scale2;                 // [scale, 1/scale] of a gradient-valued input (device), or null
  int N, H, W, Ci, Co, k, p, Hq, Wq;   // Ci/Co: PADDED channel counts the kernel iterates over; Hq/Wq: padded dims of xq
  int Cor;                             // real Cout = row stride of y and bound for stores / bias
  int CB, ncb;                         // channel block held in smem at once, number of blocks
  int tiles_x, tiles_y;                // tiles per image
  int NB;                              // Co columns handled by one CTA (<= 256); grid.y = Co / NB
  int TL, ntiles;                      // tiles per CTA (1 or 2, sharing every weight slice) and total tile count
  int S;                               // weight ring depth
  uint32_t patch_bytes, slice_bytes;   // one patch buffer, one weight slice
};

template <int ES>
__global__ void __launch_bounds__(256, 1) k_conv_tc(TcParams P, const __grid_constant__ CUtensorMap tmx) {
  constexpr int PER = 16 / ES, KB = 128 / ES, KSTEP = 32 / ES;   // channels per slice, K per MMA instruction
  static_assert(KB / KSTEP == 4, "the issuer unrolls exactly four MMAs per weight slice");
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_pfull[2], bar_pempty[2], bar_wfull[8], bar_wempty[8], bar_acc;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long* dbg = P.dbg ? P.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (dbg && tid == 0) dbg[0] = clock64();
  uint8_t* patch0 = smem;                                      // patch buffers [2][TL]
  uint8_t* wring = smem + 2 * (size_t)P.TL * P.patch_bytes;    // S weight slices

  // this CTA's tiles: TL consecutive tiles share every weight slice (measured: a 128x128 tile alone needs 64 B/cycle/SM
  // of weights from L2 and the tensor pipe sat at 23%, profiles/r01_ncu_conv3.txt)
  const int tile0 = blockIdx.x * P.TL;
  const int ntl = (P.ntiles - tile0) < P.TL ? (P.ntiles - tile0) : P.TL;
  auto tile_xy = [&](int tl, int& n, int& y0, int& x0) {
    int t = tile0 + tl; int tx = t % P.tiles_x; t /= P.tiles_x; int ty = t % P.tiles_y; n = t / P.tiles_y; x0 = tx * 8; y0 = ty * 16;
  };
  const int co0 = blockIdx.y * P.NB;
  const int Hp = 16 + 2 * P.p, Wp = 8 + 2 * P.p;       // patch rows / pitch (pixels)
  const int planes = P.CB / PER;                       // 16-byte channel planes per patch buffer
  const uint32_t plane_bytes = (uint32_t)Hp * Wp * 16;
  const int kk = P.k * P.k, nsub = P.CB / KB;
  const int nslices = P.ncb * kk * nsub;

  if (tid == 0) {
    // with two tiles per CTA there are two MMA issuers (warps 6 and 7, one tile each): a buffer is free / the accumulators are
    // complete only when BOTH have committed
    const uint32_t nissue = (uint32_t)P.TL;
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_pfull[i], 1); mbar_init(&bar_pempty[i], nissue); }
    for (int i = 0; i < P.S; ++i) { mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wempty[i], nissue); }
    mbar_init(&bar_acc, nissue);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  int ncols = 32; while (ncols < P.TL * P.NB) ncols <<= 1;
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
      int st = 0; uint32_t ph = 0;
      const int per_cb = kk * nsub, rot_s = P.rot ? (int)((blockIdx.x * 7u) % (unsigned)kk) * nsub : 0;
      int cb_base = 0, within = rot_s;                  // slice order inside a channel block starts at this CTA's own tap
      for (int s = 0; s < nslices; ++s, st = (st + 1 == P.S) ? 0 : st + 1, ph ^= (st == 0) ? 1u : 0u) {
        mbar_wait(&bar_wempty[st], ph ^ 1);
        mbar_expect_tx(&bar_wfull[st], P.slice_bytes);
        uint8_t* dst = wring + (size_t)st * P.slice_bytes;
        const uint8_t* src = P.wq + (size_t)(cb_base + within) * PL * P.Co * 16;
        if (++within == per_cb) within = 0;
        if (within == rot_s) cb_base += per_cb;
        if (P.NB == P.Co) bulk_g2s(dst, src, P.slice_bytes, &bar_wfull[st]);
        else for (int c = 0; c < PL; ++c) bulk_g2s(dst + (size_t)c * P.NB * 16, src + ((size_t)c * P.Co + co0) * 16, P.NB * 16, &bar_wfull[st]);
      }
    }
  } else if (warp == 5) {
    // ===== patch producer: ONE tiled TMA per channel block: [planes][Hp rows][Wp pixels][16 B] lands as the A-operand image
    if (lane == 0) {
      for (int cb = 0; cb < P.ncb; ++cb) {
        int buf = cb & 1; uint32_t ph = (cb >> 1) & 1;
        mbar_wait(&bar_pempty[buf], ph ^ 1);
        mbar_expect_tx(&bar_pfull[buf], (uint32_t)ntl * P.patch_bytes);
        for (int tl = 0; tl < ntl; ++tl) {
          int n, y0, x0; tile_xy(tl, n, y0, x0);
          tma_patch_4d(patch0 + (size_t)(buf * P.TL + tl) * P.patch_bytes, &tmx, x0 * PER, y0, cb * planes, n, &bar_pfull[buf]);
        }
      }
    }
  } else if (warp == 6 || (warp == 7 && P.TL == 2)) {
    // ===== MMA issuers: warp 6 owns tile 0, warp 7 owns tile 1 (one thread each).  A single thread issued one fp16 MMA per ~190
    // cycles against 64 cycles of pipe time (profiles/r01_ncu_conv3.txt); two issuers on two schedulers double the issue rate.
    // Everything loop-invariant is hoisted; no division, no descriptor rebuild in the loop.
    if (lane == 0) {
      const int my_tl = warp - 6;
      const bool have_tile = my_tl < ntl;     // an odd tile count leaves the last CTA's second issuer only committing
      const uint32_t fmt = ES == 2 ? 0u : 2u;   // F16 / TF32
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(P.NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t b_lbo = (uint32_t)P.NB * 16;
      const uint32_t a_hi = desc_hi((uint32_t)Wp * 16), b_hi = desc_hi(128);          // SBO: patch pitch / 8 co rows
      const uint32_t plane16 = plane_bytes >> 4, patch16 = P.patch_bytes >> 4, slice16 = P.slice_bytes >> 4;
      const uint32_t a_kstep = 2 * plane16, b_kstep = 2 * (b_lbo >> 4), a_sub = (uint32_t)(KB / PER) * plane16;
      const uint32_t w_lo0 = desc_lo(smem_u32(wring), b_lbo);
      const int S = P.S, k = P.k, NB = P.NB;
      const int rot_tap = P.rot ? (int)((blockIdx.x * 7u) % (unsigned)kk) : 0;
      uint32_t st = 0, wph = 0, acc0 = 0;
      long long w_wait = 0, p_wait = 0, tq = 0;
      if (dbg && warp == 6) dbg[1] = clock64();
      for (int cb = 0; cb < P.ncb; ++cb) {
        const int buf = cb & 1;
        if (dbg) tq = clock64();
        mbar_wait(&bar_pfull[buf], (cb >> 1) & 1);
        if (dbg) p_wait += clock64() - tq;
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t p_lo0 = desc_lo(smem_u32(patch0 + (size_t)(buf * P.TL) * P.patch_bytes), plane_bytes);
        // filter taps in this CTA's own cyclic order (same order as the weight producer); ky/kx kept incrementally
        int ky = rot_tap / k, kx = rot_tap - ky * k;
        for (int t = 0; t < kk; ++t) {
          {
            uint32_t a_lo = p_lo0 + (uint32_t)(ky * Wp + kx);
            if (++kx == k) { kx = 0; if (++ky == k) ky = 0; }
            for (int sub = 0; sub < nsub; ++sub, a_lo += a_sub) {
              if (dbg) tq = clock64();
              mbar_wait(&bar_wfull[st], wph);
              if (dbg) w_wait += clock64() - tq;
              asm volatile("tcgen05.fence::after_thread_sync;");
              const uint32_t w_lo = w_lo0 + st * slice16;
              if (have_tile) {
                const uint32_t at = a_lo + (uint32_t)my_tl * patch16, tm = tmem + (uint32_t)(my_tl * NB);
                umma<ES>(tm, desc64(at, a_hi), desc64(w_lo, b_hi), idesc, acc0);
                umma<ES>(tm, desc64(at + a_kstep, a_hi), desc64(w_lo + b_kstep, b_hi), idesc, 1u);
                umma<ES>(tm, desc64(at + 2 * a_kstep, a_hi), desc64(w_lo + 2 * b_kstep, b_hi), idesc, 1u);
                umma<ES>(tm, desc64(at + 3 * a_kstep, a_hi), desc64(w_lo + 3 * b_kstep, b_hi), idesc, 1u);
              }
              acc0 = 1u;
              umma_commit(&bar_wempty[st]);             // slice may be overwritten once these MMAs retire
              if (++st == (uint32_t)S) { st = 0; wph ^= 1u; }
            }
          }
        }
        umma_commit(&bar_pempty[buf]);
      }
      umma_commit(&bar_acc);
      if (dbg && warp == 6) { dbg[2] = clock64(); dbg[3] = w_wait; dbg[4] = p_wait; }
    }
  }

  if (warp < 4) {
    // ===== epilogue: TMEM -> registers -> (+bias) -> NHWC fp32.  warp w owns TMEM lanes 32w..32w+31 = pixels.
    mbar_wait(&bar_acc, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (dbg && tid == 0) dbg[5] = clock64();
    const int m = warp * 32 + lane;
    const bool vec = (P.Cor & 3) == 0;
    const float inv = P.scale2 ? P.scale2[1] : 1.f;
    // Stores: a lane holds one PIXEL's columns, so storing from the TMEM layout writes 16 bytes per lane at a 4*Cout-byte stride
    // -- measured on conv3: 26.7k of a CTA's 108k cycles in this epilogue, 2.8k with the stores removed
    // (profiles/r01_conv3_cta_timeline*.txt).  Full-width column blocks therefore go through a per-warp shared-memory tile
    // (the patch buffers are idle once bar_acc has fired) and leave as whole 128-byte lines: 8 lanes per pixel, 4 pixels per store.
    float* stage = reinterpret_cast<float*>(patch0) + warp * (32 * 36);          // 32 pixel rows x (32 + 4 pad) floats
    const bool wide = vec && (P.NB & 31) == 0 && !P.nostore;
    for (int tl = 0; tl < ntl; ++tl) {
    int n, y0, x0; tile_xy(tl, n, y0, x0);
    if (wide) {
      for (int c0 = 0; c0 < P.NB; c0 += 32) {
        uint32_t v[32];
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tl * P.NB + c0);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        if (co0 + c0 + 32 <= P.Cor) {                    // warp-uniform
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = __uint_as_float(v[j]) * inv + (P.bias ? P.bias[co0 + c0 + j] : 0.f);
            o.y = __uint_as_float(v[j + 1]) * inv + (P.bias ? P.bias[co0 + c0 + j + 1] : 0.f);
            o.z = __uint_as_float(v[j + 2]) * inv + (P.bias ? P.bias[co0 + c0 + j + 2] : 0.f);
            o.w = __uint_as_float(v[j + 3]) * inv + (P.bias ? P.bias[co0 + c0 + j + 3] : 0.f);
            *reinterpret_cast<float4*>(stage + lane * 36 + j) = o;
          }
          __syncwarp();
          const int q = (lane & 7) * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int pl = i * 4 + (lane >> 3), mm = warp * 32 + pl;
            const int oy = y0 + (mm >> 3), ox = x0 + (mm & 7);
            float4 o = *reinterpret_cast<const float4*>(stage + pl * 36 + q);
            if (oy < P.H && ox < P.W) *reinterpret_cast<float4*>(P.y + (((size_t)n * P.H + oy) * P.W + ox) * P.Cor + co0 + c0 + q) = o;
          }
          __syncwarp();
        } else {                                          // padded columns past Cout: per-lane scalar stores of the real ones
          const int oy = y0 + (m >> 3), ox = x0 + (m & 7);
          if (oy < P.H && ox < P.W) {
            float* out = P.y + (((size_t)n * P.H + oy) * P.W + ox) * P.Cor + co0;
            for (int j = 0; j < 32; ++j) { int co = co0 + c0 + j; if (co < P.Cor) out[c0 + j] = __uint_as_float(v[j]) * inv + (P.bias ? P.bias[co] : 0.f); }
          }
        }
      }
      continue;
    }
    int oy = y0 + (m >> 3), ox = x0 + (m & 7);
    bool valid = oy < P.H && ox < P.W;
    float* out = P.y + (((size_t)n * P.H + oy) * P.W + ox) * P.Cor + co0;
    for (int c0 = 0; c0 < P.NB; c0 += 16) {
      uint32_t v[16];
      uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tl * P.NB + c0);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      if (P.nostore) { if (v[0] == 0x7fc12345u) out[0] = 1.f; }
      else if (valid && vec && co0 + c0 + 16 <= P.Cor) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 o;
          o.x = __uint_as_float(v[j]) * inv + (P.bias ? P.bias[co0 + c0 + j] : 0.f);
          o.y = __uint_as_float(v[j + 1]) * inv + (P.bias ? P.bias[co0 + c0 + j + 1] : 0.f);
          o.z = __uint_as_float(v[j + 2]) * inv + (P.bias ? P.bias[co0 + c0 + j + 2] : 0.f);
          o.w = __uint_as_float(v[j + 3]) * inv + (P.bias ? P.bias[co0 + c0 + j + 3] : 0.f);
          *reinterpret_cast<float4*>(out + c0 + j) = o;
        }
      } else if (valid) {   // padded / ragged Cout (1, 3): only the real columns exist in y
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          int co = co0 + c0 + j;
          if (co < P.Cor) out[c0 + j] = __uint_as_float(v[j]) * inv + (P.bias ? P.bias[co] : 0.f);
        }
      }
    }
    }   // tiles
    if (dbg && tid == 0) dbg[6] = clock64();
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols));
}

// ------------------------------------------------------------------ host side
// Any channel counts are taken: Cin is zero-padded to the slice width (64 fp16 / 32 tf32 channels), Cout to 16.
// For the 1-, 3- and 16-channel layers the padded MMA work is negligible next to the fp32 CUDA-core kernel they
// replace (profiles/r01_launches_tc_engine.txt: 4.5 ms of a 22 ms step).
static bool tc_shape_ok(int H, int W, int Ci, int Co, int k, int ES) {
  (void)Ci; (void)Co; (void)ES;
  return (k == 3 || k == 5 || k == 7) && W % 8 == 0 && H % 8 == 0;
}

// Tile plan of the forward-conv kernel for padded Ci -> Co channels (shared with the model-level weight packer).
//   NB: columns per CTA = largest multiple of 16 that divides Co and is <= 256
//   TL: tiles per CTA -- prefer TWO tiles sharing each weight slice (halves the L2->SM weight stream per MMA)
//   CB: the largest channel block that keeps 2 x TL patch buffers + a >= 4-deep weight ring (>= 3 for TL = 1) inside ~212 KB
template <int ES>
static void tc_plan(int Ci, int Co, int k, int ntiles, int* NBo, int* TLo, int* CBo) {
  constexpr int PER = 16 / ES, KB = 128 / ES;
  const int p = (k - 1) / 2, Hp = 16 + 2 * p, Wpx = 8 + 2 * p;
  int NB = Co; while (NB > 256 || Co % NB) NB -= 16;
  size_t slice_bytes = (size_t)(KB / PER) * NB * 16;
  int TL = 1, CB = 0;
  for (int tl = (ntiles >= 2 && 2 * NB <= 512) ? 2 : 1; tl >= 1 && !CB; --tl)
    for (int cand = Ci; cand >= KB; cand -= KB) {
      if (Ci % cand) continue;
      size_t pb = (size_t)(cand / PER) * Hp * Wpx * 16;
      if (2 * tl * pb + (size_t)(tl == 2 ? 4 : 3) * slice_bytes <= 212 * 1024) { CB = cand; TL = tl; break; }
    }
  *NBo = NB; *TLo = TL; *CBo = CB;
}

// fp16 weight slices packed once per parameter update by repack_model(), keyed by the fp32 operand pointer the caller passes
struct WSliceEntry { const uint8_t* wq; int CB; };
static std::unordered_map<const float*, WSliceEntry>& wslice_registry() { static std::unordered_map<const float*, WSliceEntry> r; return r; }
void conv_tc_register_wslices(const float* key, const uint8_t* wq, int CB) { wslice_registry()[key] = WSliceEntry{wq, CB}; }
void conv_tc_unregister_wslices(const float* key) { wslice_registry().erase(key); }
static bool conv_v1();
bool conv_tc_wslice_plan(int Cin, int Cout, int k, int* CB, size_t* bytes) {
  if (!conv_v1() && k == 1) { *CB = 32; *bytes = (size_t)((Cin + 31) / 32) * 32 * (((Cout + 15) / 16) * 16) * 2; return true; }   // nn.Linear
  if (!(k == 3 || k == 5 || k == 7)) return false;
  if (!conv_v1()) { *CB = 32; *bytes = (size_t)k * k * ((Cin + 31) / 32) * 32 * (((Cout + 15) / 16) * 16) * 2; return true; }
  const int Ci = ((Cin + 63) / 64) * 64, Co = ((Cout + 15) / 16) * 16;
  int NB, TL; tc_plan<2>(Ci, Co, k, 2, &NB, &TL, CB);
  *bytes = (size_t)k * k * Ci * Co * 2;
  return *CB != 0;
}

// ====================================================================================================================
// k_conv_ps -- the forward-conv engine of round 2: PERSISTENT CTAs, FOUR accumulator slots in TMEM on a STAGGERED schedule.
//
// What round 1 measured on k_conv_tc (profiles/r01_conv3_cta_timeline*.txt, r01_ncu_conv3_v3.txt): with two 128-pixel tiles per CTA the
// issue loop waits a quarter of its time for weight slices (every weight byte fetched from L2 feeds only 256 pixels: 965 MB over the
// crossbar per conv3 launch against 154 MB algorithmic), the epilogue (10.6k of 93k cycles) overlaps nothing, and 512 CTAs on 148 SMs
// run 3.46 waves.  This kernel changes all three:
//   * one CTA per SM loops over its tiles (tile t = blockIdx.x + i * gridDim.x): no wave quantisation, one prologue;
//   * FOUR tiles accumulate side by side (4 x NB <= 512 TMEM columns) and share every weight slice: a slice fetched once feeds 512
//     pixels (the round-1 TL4 experiment, 234 -> 194 us on conv3, profiles/r01_conv3_tl4_experiment.txt);
//   * the four slots do not run in lockstep: the weight stream is CYCLIC (slice g mod nslices) and a tile may start its K loop at any
//     slice -- it ends one full cycle later.  Slot j starts its k-th tile at slice j*D + k*(nslices + D), so the slots finish D slices
//     apart and each slot's epilogue (TMEM -> registers -> shared-memory staging -> 128-byte lines) drains while the other three keep
//     the tensor pipe busy; the slot rejoins the stream D slices after it ended.  The schedule is static: every role computes it.
// Roles (320 threads): warps 0-3 epilogue (TMEM lane quarters) ; warps 4-7 MMA issuers, one per slot (one thread retires an MMA per
// ~84 cycles, tools/tc_rate.cu, so N = 64 layers need more than two) ; warp 8 streams weight slices (32 channels x NB columns, one
// bulk copy) ; warp 9 loads patches (one tiled TMA per slot and 32-channel block, issued in the order the buffers free up).
// Operand layouts are k_conv_tc's (halo'd patch resident in shared memory, a filter tap = a shifted descriptor start address).
constexpr int PS_SLOTS = 4;
struct PsParams {
  const uint8_t *xq, *wq; const float *bias, *scale2; float* y;
  float inv_host;                      // epilogue factor on the accumulator (1 unless the weights were packed pre-scaled)
  int N, H, W, Co, Cor, k, p;          // Co: padded column count of the slices (multiple of 16); Cor: real Cout = row stride of y
  int ncb, kk, nslices;                // 32-channel blocks that hold data, taps, ncb * kk
  int tiles_x, tiles_y, ntiles;
  int NB, S, D;                        // columns per CTA, weight ring depth, stagger / gap in slices
  unsigned int* amax_out;              // optional: max|y| over the launch (atomicMax on float bits), the packing-scale bound of the next backward stage
  long long* dbg;                      // experiments (CATGEN_PS_DBG=1): per-CTA clock64 breakdown, 32 values per CTA
  int Z;                               // > 1: K-SPLIT mode (nn.Linear, kk = 1): gridDim.x = Z CTAs each stream their share of the slices for ALL tiles
                                       //      (<= 4) and write raw partial sums to y[z][...] (bias / reduction in k_splitk_reduce)
  uint32_t patch_bytes, slice_bytes;
  double* stats;                       // forward convolution feeding a training-mode BatchNormalization: per-CTA partial (sum, sum of squares) rows [gridDim.x][Cor][2]
  int pool2; float* y2;                // input-gradient convolution of an upsampled stage: the epilogue sums 2 x 2 blocks and writes y2 [N, H/2, W/2, Cor] instead of y
  int duo;                             // 8 x 8 images: a tile is an image PAIR (make_patch_tmap_duo); n of tile_xy is then the pair index
};
// Weights Wp[(tap,ci)][co] fp32 -> 32-channel slices in stream order Wq[cb][tap][c 0..3][Cop][8 fp16]; zero beyond Ci / Co.
__global__ void k_pack_wslices32(const float* __restrict__ Wp, uint8_t* __restrict__ Wq, long nchunks, int Ci, int Co, int Cop, int kk) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nchunks; i += (long)gridDim.x * blockDim.x) {
    int co = (int)(i % Cop); long t = i / Cop; int c = (int)(t % 4); t /= 4; int tap = (int)(t % kk); int cb = (int)(t / kk);
    int ci0 = cb * 32 + c * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (co < Co && ci0 + j < Ci) ? Wp[((long)tap * Ci + ci0 + j) * Co + co] : 0.f;
    uint4 out; uint32_t* o = &out.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) { __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]); o[j] = *reinterpret_cast<uint32_t*>(&h); }
    reinterpret_cast<uint4*>(Wq)[i] = out;
  }
}
// ---- error-compensated forward operands (cg_set_precision(1)).  x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (and the same for W); the
// convolution over the virtual channel list [hi(x) | lo(x) | hi(x)] against [hi(W) | hi(W) | lo(W)] accumulates hi*hi + lo*hi + hi*lo in the
// fp32 TMEM accumulator: the dropped lo*lo term is 2^-22 relative.  Cg = channels per group (a multiple of 32, the slice width); the kernel
// itself is unchanged -- it sees a convolution with 3*Cg input channels.  lo falls into fp16's subnormal range for |x| < ~0.1: its absolute
// step there is 6e-8, far below the fp32 rounding of the products it corrects.
__global__ void k_pack_act_split(const float* __restrict__ x, uint8_t* __restrict__ xq, long nchunks, int H, int W, int Ci, int Cg, int Cip, int p, int Hq, int Wq) {
  const int Cq = Cip / 8, Gq = Cg / 8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nchunks; i += (long)gridDim.x * blockDim.x) {
    int xx = (int)(i % Wq); long t = i / Wq; int yy = (int)(t % Hq); t /= Hq; int c = (int)(t % Cq); long n = t / Cq;
    int iy = yy - p, ix = xx - p;
    const int g = c / Gq, cc = (c % Gq) * 8;        // group 0 / 2: hi, group 1: lo; planes past the third group are padding
    uint4 out = make_uint4(0, 0, 0, 0);
    if (g < 3 && iy >= 0 && iy < H && ix >= 0 && ix < W && cc < Ci) {
      const float* s = x + ((n * H + iy) * W + ix) * Ci + cc;
      uint32_t* o = &out.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = cc + 2 * j < Ci ? s[2 * j] : 0.f, b = cc + 2 * j + 1 < Ci ? s[2 * j + 1] : 0.f;
        __half2 h = __floats2half2_rn(a, b);
        if (g == 1) h = __floats2half2_rn(a - __low2float(h), b - __high2float(h));
        o[j] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
    reinterpret_cast<uint4*>(xq)[i] = out;
  }
}
// Wp[(tap,ci)][co] fp32 -> 32-channel slices over the virtual channels [hi | hi | lo], each group ncb0 blocks wide
__global__ void k_pack_wslices32_split(const float* __restrict__ Wp, uint8_t* __restrict__ Wq, long nchunks, int Ci, int Co, int Cop, int kk, int ncb0) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nchunks; i += (long)gridDim.x * blockDim.x) {
    int co = (int)(i % Cop); long t = i / Cop; int c = (int)(t % 4); t /= 4; int tap = (int)(t % kk); int cbv = (int)(t / kk);
    const int g = cbv / ncb0, ci0 = (cbv % ncb0) * 32 + c * 8;
    uint4 out; uint32_t* o = &out.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = (co < Co && ci0 + 2 * j < Ci) ? Wp[((long)tap * Ci + ci0 + 2 * j) * Co + co] : 0.f;
      float b = (co < Co && ci0 + 2 * j + 1 < Ci) ? Wp[((long)tap * Ci + ci0 + 2 * j + 1) * Co + co] : 0.f;
      __half2 h = __floats2half2_rn(a, b);
      if (g == 2) h = __floats2half2_rn(a - __low2float(h), b - __high2float(h));
      o[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    reinterpret_cast<uint4*>(Wq)[i] = out;
  }
}
// ---- CTA-pair (cta_group::2) helpers; mechanics established by tools/tc_pair_probe.cu on B200 (profiles/r02_pair_probe.txt)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r; }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory"); }
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_pair(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
// tiled TMA whose completion is signalled on a barrier given as a shared::cluster address (the pair leader's)
__device__ __forceinline__ void tma_2d_bar(void* dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tma_patch_4d_bar(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tma_2d_mc(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar, uint16_t mask) {   // same offsets in every CTA of the mask
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(tm), "r"(c0), "r"(c1), "r"((uint32_t)__cvta_generic_to_shared(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {   // cta_group::1 commit arriving on the barrier at this offset in every CTA of the mask
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }

// PAIR = 1: the kernel runs as clusters of two CTAs (one TPC) driving ONE tcgen05.mma.cta_group::2 stream: M = 256 = the two CTAs' pixel
// tiles, N = NB with each CTA holding HALF of every weight slice (NB/2 columns) in its own shared memory.  Why: an M = 128 x N = 128 MMA reads
// 4 KB of A and 4 KB of B from shared memory per 64 cycles = 128 B/cycle, the whole shared-memory port, so every TMA write and every
// epilogue staging access steals tensor-pipe time (measured: 352k cycles per CTA against an MMA floor of 177k with weights, patches and TMEM
// all waiting on the pipe; gpurun_out/r02_e_conv3_dbg.txt).  In pair mode a CTA reads 4 KB + 2 KB per 64 cycles: a quarter of the port is
// free again.  The leader CTA (rank 0) issues every MMA and commits with multicast; both CTAs load (their TMA completions are counted on
// the LEADER's barriers), both drain their own TMEM lanes.
template <int PAIR>
__global__ void __launch_bounds__(320, 1) k_conv_ps(PsParams P, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmw) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_wfull[16], bar_wempty[16], bar_pfull[PS_SLOTS][2], bar_pempty[PS_SLOTS][2], bar_acc[PS_SLOTS], bar_tfree[PS_SLOTS];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* patch0 = smem;                                              // [slot][buffer] patches
  uint8_t* wring = smem + (size_t)2 * PS_SLOTS * P.patch_bytes;        // S weight slices
  float* stage0 = reinterpret_cast<float*>(wring + (size_t)P.S * P.slice_bytes);   // epilogue staging, 4 warps x 32 rows x 36 floats

  constexpr bool PR = PAIR == 1;      // cta_group::2 pairs: one MMA stream over both CTAs' tiles, B split between their shared memories
  constexpr bool MC = PAIR == 2;      // clusters of two independent CTAs (own tiles, own cta_group::1 MMAs) that share ONE multicast weight stream
  const uint32_t rank = (PR || MC) ? cluster_ctarank() : 0u;         // 0 = leader
  const int gx = PR ? (int)(gridDim.x >> 1) : (int)gridDim.x, bx = PR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // tile-column of the grid: CTAs, or pairs
  const int co0 = blockIdx.y * P.NB;
  const int ntu = PR ? (P.ntiles + 1) / 2 : P.ntiles;               // scheduling units: tiles, or tile pairs (2u + rank)
  long long* dbg = P.dbg ? P.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 32 : nullptr;
  const long long t_start = dbg ? clock64() : 0;
  const bool zs = P.Z > 1;                                            // K-split mode
  const int s_begin = zs ? (int)((long)P.nslices * blockIdx.x / P.Z) : 0;                       // first slice this CTA streams
  const int ns = zs ? (int)((long)P.nslices * (blockIdx.x + 1) / P.Z) - s_begin : P.nslices;    // slices per tile cycle (>= 1: Z <= nslices)
  const int nt = zs ? P.ntiles : (ntu - bx + gx - 1) / gx;            // tiles of this CTA (>= 1: the host keeps the grid <= the units)
  // MC: both CTAs of a cluster walk the SAME slice stream (its length is set by the one with more tiles)
  const int nt_s = MC ? max(nt, (ntu - (bx ^ 1) + gx - 1) / gx) : nt;
  const int D = zs ? 0 : P.D, kk = P.kk;
  const int ncb = zs ? ns : P.ncb;                                    // 32-channel blocks in this CTA's stream (kk = 1 in K-split mode)
  // slot j owns tiles j, j+4, ...; its k-th tile is accumulated over slices [j*D + k*(ns+D), +ns)
  int total_g = 0;
#pragma unroll
  for (int j = 0; j < PS_SLOTS; ++j) { int nk = (nt_s - j + PS_SLOTS - 1) / PS_SLOTS; if (nk > 0) { int e = j * D + (nk - 1) * (ns + D) + ns; total_g = e > total_g ? e : total_g; } }
  const int Hp = P.duo ? 2 * (8 + 2 * P.p) : 16 + 2 * P.p, Wp = 8 + 2 * P.p;   // patch rows per plane (duo: the two images' rows interleaved)
  const int rowstep = P.duo ? 2 * Wp : Wp;                                  // one image row down, in 16-byte pixels
  const uint32_t plane_bytes = (uint32_t)Hp * Wp * 16;

  if (tid == 0) {
    for (int i = 0; i < P.S; ++i) { mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wempty[i], (MC && rank == 0) ? 2 * PS_SLOTS : PS_SLOTS); }   // MC leader: both CTAs' issuers release a slice
    for (int j = 0; j < PS_SLOTS; ++j) {
      for (int b = 0; b < 2; ++b) { mbar_init(&bar_pfull[j][b], 1); mbar_init(&bar_pempty[j][b], 1); }
      mbar_init(&bar_acc[j], 1); mbar_init(&bar_tfree[j], PR ? 8 : 4);   // pair: the leader's issuer waits for BOTH CTAs' epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  int ncols = 32; while (ncols < PS_SLOTS * P.NB) ncols <<= 1;
  if (warp == 0) {
    if (PR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ncols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ncols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (PR || MC) cluster_sync_all();                                   // the peer's barriers (and TMEM) exist before anything reaches across
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  // the i-th tile of this CTA; in pair mode unit u = bx + i*gx holds tiles 2u (leader) and 2u+1 (peer) -- an odd tile count leaves the
  // last peer a GHOST tile: it recomputes the last real tile (so that every load and barrier stays symmetric) and stores nothing
  auto tile_id = [&](int i) { return zs ? i : (PR ? 2 * (bx + i * gx) + (int)rank : bx + i * gx); };
  auto tile_xy = [&](int i, int& n, int& y0, int& x0) {
    int t = tile_id(i); if (t >= P.ntiles) t = P.ntiles - 1; int tx = t % P.tiles_x; t /= P.tiles_x; int ty = t % P.tiles_y; n = t / P.tiles_y; x0 = tx * 8; y0 = ty * 16;
  };

  if (warp == 8) {
    // ===== weight producer: the cyclic slice stream, one bulk copy per slice (one per plane when the CTA owns a column block)
    if (lane == 0) {
      int st = 0, s = 0; uint32_t ph = 0;
      long long tw = 0, tq = 0;
      for (int g = 0; g < total_g; ++g) {
        if (dbg) tq = clock64();
        mbar_wait(&bar_wempty[st], ph ^ 1);
        if (dbg) tw += clock64() - tq;
        if (PR) {   // each CTA fetches ITS half of the slice's columns; both completions are counted on the leader's barrier
          if (rank == 0) mbar_expect_tx(&bar_wfull[st], 2 * P.slice_bytes);
          tma_2d_bar(wring + (size_t)st * P.slice_bytes, &tmw, (co0 + (int)rank * (P.NB / 2)) * 2, (s_begin + s) * 4, mapa_u32(smem_u32(&bar_wfull[st]), 0));
        } else if (MC) {   // every CTA arms its own barrier; the leader's ONE tiled TMA lands in both shared memories and completes on both barriers
          mbar_expect_tx(&bar_wfull[st], P.slice_bytes);
          if (rank == 0) tma_2d_mc(wring + (size_t)st * P.slice_bytes, &tmw, co0 * 2, (s_begin + s) * 4, &bar_wfull[st], (uint16_t)3);
        } else {
          mbar_expect_tx(&bar_wfull[st], P.slice_bytes);
          tma_2d(wring + (size_t)st * P.slice_bytes, &tmw, co0 * 2, (s_begin + s) * 4, &bar_wfull[st]);   // ONE tiled TMA per slice, column block included
        }
        if (++s == ns) s = 0;
        if (++st == P.S) { st = 0; ph ^= 1u; }
      }
      if (dbg) { dbg[1] = tw; dbg[2] = clock64() - t_start; }
    }
  } else if (warp == 9) {
    // ===== patch producer.  Per slot a stream of segments (= one 32-channel block of one tile resident in one of the slot's two
    // buffers); segment c of a slot may be loaded once segment c-2 has been consumed.  Loads are issued in the order in which
    // their buffers free up (4-way merge on that slice index), so waiting on one slot never holds back a load another slot needs first.
    if (lane == 0) {
      int kj[PS_SLOTS], qj[PS_SLOTS], cj[PS_SLOTS], gj[PS_SLOTS], e1[PS_SLOTS], e2[PS_SLOTS], nkj[PS_SLOTS];
#pragma unroll
      for (int j = 0; j < PS_SLOTS; ++j) { kj[j] = 0; qj[j] = 0; cj[j] = 0; gj[j] = j * D; e1[j] = -1; e2[j] = -1; nkj[j] = (nt - j + PS_SLOTS - 1) / PS_SLOTS; if (nkj[j] < 0) nkj[j] = 0; }
      for (;;) {
        int best = -1, bestr = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < PS_SLOTS; ++j) if (kj[j] < nkj[j] && e2[j] < bestr) { bestr = e2[j]; best = j; }
        if (best < 0) break;
        const int j = best;
        // tile kj of slot j starts at slice gstart; segment qj covers [gj, gj + len)
        const int gstart = j * D + kj[j] * (ns + D);
        const int s0 = gstart % ns, cb0 = s0 / kk, tap0 = s0 - cb0 * kk;     // position in THIS CTA's stream (block index relative to s_begin / kk)
        const int nseg = ncb == 1 ? 1 : ncb + (tap0 > 0 ? 1 : 0);
        int cb, len;
        if (ncb == 1) { cb = 0; len = ns; }
        else if (qj[j] == 0) { cb = cb0; len = kk - tap0; }
        else { cb = cb0 + qj[j]; if (cb >= ncb) cb -= ncb; len = (qj[j] == nseg - 1 && tap0 > 0) ? tap0 : kk; }
        cb += s_begin;                                                        // K-split (kk = 1): slices ARE channel blocks
        const int buf = cj[j] & 1; const uint32_t ph = (uint32_t)(cj[j] >> 1) & 1u;
        mbar_wait(&bar_pempty[j][buf], ph ^ 1);
        int n, y0, x0; tile_xy(j + PS_SLOTS * kj[j], n, y0, x0);
        if (PR) {
          if (rank == 0) mbar_expect_tx(&bar_pfull[j][buf], 2 * P.patch_bytes);
          tma_patch_4d_bar(patch0 + (size_t)(j * 2 + buf) * P.patch_bytes, &tmx, x0 * 8, y0, cb * 4, n, mapa_u32(smem_u32(&bar_pfull[j][buf]), 0));
        } else {
          mbar_expect_tx(&bar_pfull[j][buf], P.patch_bytes);
          if (P.duo) tma_patch_5d(patch0 + (size_t)(j * 2 + buf) * P.patch_bytes, &tmx, 0, 0, 0, cb * 4, n, &bar_pfull[j][buf]);
          else tma_patch_4d(patch0 + (size_t)(j * 2 + buf) * P.patch_bytes, &tmx, x0 * 8, y0, cb * 4, n, &bar_pfull[j][buf]);
        }
        e2[j] = e1[j]; e1[j] = gj[j] + len; gj[j] += len; cj[j]++;
        if (++qj[j] == nseg) { qj[j] = 0; kj[j]++; gj[j] = j * D + kj[j] * (ns + D); }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== MMA issuer of slot j = warp - 4.  EVERY issuer walks the whole slice stream and arrives on every slice's "empty"
    // barrier (count 4) -- with a commit behind its MMAs while its slot is accumulating, with a plain arrive otherwise.
    if (lane == 0 && (!PR || rank == 0)) {
      const int j = warp - 4;
      const int nk = (nt - j + PS_SLOTS - 1) / PS_SLOTS;
      const uint32_t idesc = (1u << 4) | ((uint32_t)(P.NB >> 3) << 17) | ((uint32_t)((PR ? 256 : 128) >> 4) << 24);
      const uint32_t b_lbo = (uint32_t)(PR ? P.NB / 2 : P.NB) * 16;     // pair: the descriptor describes the LOCAL half of B, the instruction the full N
      const uint32_t a_hi = desc_hi((uint32_t)Wp * 16), b_hi = desc_hi(128);
      const uint32_t plane16 = plane_bytes >> 4, slice16 = P.slice_bytes >> 4;
      const uint32_t a_kstep = 2 * plane16, b_kstep = 2 * (b_lbo >> 4);
      const uint32_t w_lo0 = desc_lo(smem_u32(wring), b_lbo);
      const uint32_t p_lo_buf0 = desc_lo(smem_u32(patch0 + (size_t)(j * 2) * P.patch_bytes), plane_bytes), patch16 = P.patch_bytes >> 4;
      const uint32_t tm = tmem + (uint32_t)(j * P.NB);
      const int S = P.S, k = P.k;
      uint32_t st = 0, wph = 0, acc = 0;
      int kt = 0, g0 = j * D, seg = 0;
      int tap = 0, ky = 0, kx = 0;
      uint32_t p_lo = p_lo_buf0;
      long long t_wf = 0, t_pf = 0, t_tf = 0, tq = 0;
      for (int g = 0; g < total_g; ++g) {
        if (dbg) tq = clock64();
        mbar_wait(&bar_wfull[st], wph);
        if (dbg) t_wf += clock64() - tq;
        if (kt < nk && g >= g0) {                           // g < g0 + ns holds by construction (the window is closed below)
          asm volatile("tcgen05.fence::after_thread_sync;");
          if (g == g0) {
            if (dbg) tq = clock64();
            if (kt > 0) { mbar_wait(&bar_tfree[j], (uint32_t)(kt - 1) & 1u); asm volatile("tcgen05.fence::after_thread_sync;"); }   // epilogue has drained this slot's previous tile
            if (dbg) t_tf += clock64() - tq;
            const int s0 = g0 % ns; const int cb0 = s0 / kk; tap = s0 - cb0 * kk; ky = tap / k; kx = tap - ky * k;
            if (dbg) tq = clock64();
            mbar_wait(&bar_pfull[j][seg & 1], (uint32_t)(seg >> 1) & 1u);
            if (dbg) t_pf += clock64() - tq;
            asm volatile("tcgen05.fence::after_thread_sync;");
            p_lo = p_lo_buf0 + (uint32_t)(seg & 1) * patch16;
            acc = 0;
          } else if (tap == 0 && ncb > 1) {                 // next 32-channel block: hand the buffer back, take the other one
            if (PR) umma_commit_pair(&bar_pempty[j][seg & 1]); else umma_commit(&bar_pempty[j][seg & 1]);
            ++seg;
            if (dbg) tq = clock64();
            mbar_wait(&bar_pfull[j][seg & 1], (uint32_t)(seg >> 1) & 1u);
            if (dbg) t_pf += clock64() - tq;
            asm volatile("tcgen05.fence::after_thread_sync;");
            p_lo = p_lo_buf0 + (uint32_t)(seg & 1) * patch16;
          }
          const uint32_t a_lo = p_lo + (uint32_t)(ky * rowstep + kx), w_lo = w_lo0 + st * slice16;
          if (PR) {
            umma_pair(tm, desc64(a_lo, a_hi), desc64(w_lo, b_hi), idesc, acc);
            umma_pair(tm, desc64(a_lo + a_kstep, a_hi), desc64(w_lo + b_kstep, b_hi), idesc, 1u);
          } else {
            umma<2>(tm, desc64(a_lo, a_hi), desc64(w_lo, b_hi), idesc, acc);
            umma<2>(tm, desc64(a_lo + a_kstep, a_hi), desc64(w_lo + b_kstep, b_hi), idesc, 1u);
          }
          acc = 1u;
          if (++kx == k) { kx = 0; ++ky; }
          if (++tap == kk) { tap = 0; ky = 0; kx = 0; }
          if (g == g0 + ns - 1) {                           // the tile's K cycle is complete
            if (PR) { umma_commit_pair(&bar_pempty[j][seg & 1]); umma_commit_pair(&bar_acc[j]); }
            else { umma_commit(&bar_pempty[j][seg & 1]); umma_commit(&bar_acc[j]); }
            ++seg; ++kt; g0 += ns + D;
          }
          if (PR) umma_commit_pair(&bar_wempty[st]);
          else if (MC && rank != 0) umma_commit_mc(&bar_wempty[st], (uint16_t)3);   // releases the slice here AND at the leader, whose producer refills both
          else umma_commit(&bar_wempty[st]);   // arrives when the MMAs reading this slice have retired
        } else {
          mbar_arrive(&bar_wempty[st]);
          if (PR) mbar_arrive_remote(mapa_u32(smem_u32(&bar_wempty[st]), 1));   // the peer's producer counts the same four arrivals per slice
          if (MC && rank != 0) mbar_arrive_remote(mapa_u32(smem_u32(&bar_wempty[st]), 0));
        }
        if (++st == (uint32_t)S) { st = 0; wph ^= 1u; }
      }
      if (dbg) { dbg[4 + j * 4] = t_wf; dbg[5 + j * 4] = t_pf; dbg[6 + j * 4] = t_tf; dbg[7 + j * 4] = clock64() - t_start; }
    }
  } else {
    // ===== epilogue (warps 0-3): tiles in completion order (k, j); warp w owns TMEM lanes 32w..32w+31 = pixels of the tile
    const int m = warp * 32 + lane;
    const bool vec = (P.Cor & 3) == 0;
    const float inv = (P.scale2 ? P.scale2[1] : 1.f) * P.inv_host;
    float* stage = stage0 + warp * (32 * 36);
    const bool wide = vec && (P.NB & 31) == 0;
    long long t_acc = 0, tq = 0;
    float amx = 0.f;
    double st_s[4] = {0.0, 0.0, 0.0, 0.0}, st_q[4] = {0.0, 0.0, 0.0, 0.0};   // P.stats: this lane's channel (lane of chunk c0/32): sum and sum of squares over the warp's pixels
    for (int i = 0; i < nt; ++i) {
      const int j = i & (PS_SLOTS - 1), kt = i >> 2;
      if (dbg) tq = clock64();
      mbar_wait(&bar_acc[j], (uint32_t)kt & 1u);
      if (dbg) t_acc += clock64() - tq;
      asm volatile("tcgen05.fence::after_thread_sync;");
      int n, y0, x0; tile_xy(i, n, y0, x0);
      if (zs) n += (int)blockIdx.x * P.N;                                   // partial sums of split z live in y[z]
      const int duo = P.duo;
      // tile row mm (0..127) -> output pixel: plain tile: row group g = mm >> 3 is image row y0 + g; duo: g = (image row, image of the pair)
      auto pix = [&](int mm, int Hlim, bool& ok) -> size_t {
        const int g = mm >> 3, ox = x0 + (mm & 7), oy = duo ? g >> 1 : y0 + g, nn = duo ? 2 * n + (g & 1) : n;
        ok = oy < Hlim && ox < P.W;
        return (((size_t)nn * P.H + oy) * P.W + ox) * P.Cor;
      };
      const bool ghost = PR && tile_id(i) >= P.ntiles;                    // odd tile count: the last peer tile duplicates a real one -- read TMEM, store nothing
      const int Hst = ghost ? 0 : P.H;                                      // (every store below is guarded by oy < Hst)
      const uint32_t tcol = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * P.NB);
      if (wide) {
        for (int c0 = 0; c0 < P.NB; c0 += 32) {
          uint32_t v[32];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                         "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                         "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                         "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                       : "r"(tcol + (uint32_t)c0));
          asm volatile("tcgen05.wait::ld.sync.aligned;");
          if (co0 + c0 + 32 <= P.Cor) {                    // warp-uniform
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
              float4 o;
              o.x = __uint_as_float(v[q]) * inv + (P.bias ? P.bias[co0 + c0 + q] : 0.f);
              o.y = __uint_as_float(v[q + 1]) * inv + (P.bias ? P.bias[co0 + c0 + q + 1] : 0.f);
              o.z = __uint_as_float(v[q + 2]) * inv + (P.bias ? P.bias[co0 + c0 + q + 2] : 0.f);
              o.w = __uint_as_float(v[q + 3]) * inv + (P.bias ? P.bias[co0 + c0 + q + 3] : 0.f);
              *reinterpret_cast<float4*>(stage + lane * 36 + q) = o;
              amx = fmaxf(amx, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));   // rows past the image are zero operands -> zero outputs
            }
            __syncwarp();
            if (P.stats) {   // batch-norm statistics of the conv output (models.lua:207,213,219) from the staged tile: lane = channel, fp32 over the warp's 32 pixels, double across tiles
              float fs = 0.f, fq = 0.f;
#pragma unroll 8
              for (int pl = 0; pl < 32; ++pl) {
                bool okp; (void)pix(warp * 32 + pl, Hst, okp);
                const float vv = okp ? stage[pl * 36 + lane] : 0.f;
                fs += vv; fq = fmaf(vv, vv, fq);
              }
              st_s[c0 >> 5] += (double)fs; st_q[c0 >> 5] += (double)fq;
            }
            if (P.pool2) {
              // nn.SpatialUpSamplingNearest(2) backward fused into the input-gradient convolution: the warp's 4 image rows x 8 pixels are two rows of
              // 2 x 2 blocks; 4 lanes per pooled pixel, 8 channels each -> the [N, H/2, W/2, C] gradient leaves the SM, a quarter of the bytes
              const int pp = lane >> 2, br = pp >> 2, bc = pp & 3, ch8 = (lane & 3) * 8;
              const float* s00 = stage + ((2 * br) * 8 + 2 * bc) * 36 + ch8;
              float4 a0 = *reinterpret_cast<const float4*>(s00), a1 = *reinterpret_cast<const float4*>(s00 + 4);
#pragma unroll
              for (int t = 1; t < 4; ++t) {
                const float* st = s00 + ((t >> 1) * 8 + (t & 1)) * 36;
                const float4 b0 = *reinterpret_cast<const float4*>(st), b1 = *reinterpret_cast<const float4*>(st + 4);
                a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w; a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
              }
              const int oy = (y0 + warp * 4 + 2 * br) >> 1, ox = (x0 + 2 * bc) >> 1, Hh = P.H >> 1, Wh = P.W >> 1;
              if (oy < (Hst >> 1) && ox < Wh) {
                float* o = P.y2 + (((size_t)n * Hh + oy) * Wh + ox) * P.Cor + co0 + c0 + ch8;
                *reinterpret_cast<float4*>(o) = a0; *reinterpret_cast<float4*>(o + 4) = a1;
              }
              __syncwarp();
              continue;
            }
            const int q4 = (lane & 7) * 4;
#pragma unroll
            for (int r = 0; r < 8; ++r) {                   // 8 lanes per pixel: whole 128-byte lines leave the SM
              const int pl = r * 4 + (lane >> 3), mm = warp * 32 + pl;
              bool ok; const size_t po = pix(mm, Hst, ok);
              float4 o = *reinterpret_cast<const float4*>(stage + pl * 36 + q4);
              if (ok) *reinterpret_cast<float4*>(P.y + po + co0 + c0 + q4) = o;
            }
            __syncwarp();
          } else {                                          // padded columns past Cout
            bool ok; const size_t po = pix(m, Hst, ok);
            if (ok) {
              float* out = P.y + po + co0;
              for (int q = 0; q < 32; ++q) { int co = co0 + c0 + q; if (co < P.Cor) out[c0 + q] = __uint_as_float(v[q]) * inv + (P.bias ? P.bias[co] : 0.f); }
            }
          }
        }
      } else {
        bool valid; const size_t po = pix(m, Hst, valid);
        float* out = P.y + po + co0;
        for (int c0 = 0; c0 < P.NB; c0 += 16) {
          uint32_t v[16];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                         "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                       : "r"(tcol + (uint32_t)c0));
          asm volatile("tcgen05.wait::ld.sync.aligned;");
          if (valid && vec && co0 + c0 + 16 <= P.Cor) {
#pragma unroll
            for (int q = 0; q < 16; q += 4) {
              float4 o;
              o.x = __uint_as_float(v[q]) * inv + (P.bias ? P.bias[co0 + c0 + q] : 0.f);
              o.y = __uint_as_float(v[q + 1]) * inv + (P.bias ? P.bias[co0 + c0 + q + 1] : 0.f);
              o.z = __uint_as_float(v[q + 2]) * inv + (P.bias ? P.bias[co0 + c0 + q + 2] : 0.f);
              o.w = __uint_as_float(v[q + 3]) * inv + (P.bias ? P.bias[co0 + c0 + q + 3] : 0.f);
              *reinterpret_cast<float4*>(out + c0 + q) = o;
              amx = fmaxf(amx, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
            }
          } else if (valid) {                               // ragged Cout (1, 3): only the real columns exist in y
#pragma unroll
            for (int q = 0; q < 16; ++q) { int co = co0 + c0 + q; if (co < P.Cor) out[c0 + q] = __uint_as_float(v[q]) * inv + (P.bias ? P.bias[co] : 0.f); }
          }
        }
      }
      // the slot's columns may be overwritten by its next tile
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) { if (PR && rank != 0) mbar_arrive_remote(mapa_u32(smem_u32(&bar_tfree[j]), 0)); else mbar_arrive(&bar_tfree[j]); }
    }
    if (P.amax_out) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, o));
      if (lane == 0 && amx > 0.f) atomicMax(P.amax_out, __float_as_uint(amx));
    }
    if (P.stats) {   // the four epilogue warps' sums in warp order -> ONE partial row per CTA, [gridDim.x][C][2] doubles (the format ops.cu's column reductions write)
      double* sd = reinterpret_cast<double*>(stage0);                       // 4 warps x NB x 2 doubles <= 8 KB of the 18 KB staging area; every tile of this CTA is done
      asm volatile("bar.sync 1, 128;");                                      // the epilogue warps only (the staging area is theirs)
#pragma unroll
      for (int q = 0; q < 4; ++q) if (q * 32 < P.NB) { sd[((warp * P.NB) + q * 32 + lane) * 2] = st_s[q]; sd[((warp * P.NB) + q * 32 + lane) * 2 + 1] = st_q[q]; }
      asm volatile("bar.sync 1, 128;");
      const int ch = warp * 32 + lane;
      if (ch < P.NB && co0 + ch < P.Cor) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int w = 0; w < 4; ++w) { a += sd[(w * P.NB + ch) * 2]; b += sd[(w * P.NB + ch) * 2 + 1]; }
        P.stats[((size_t)blockIdx.x * P.Cor + co0 + ch) * 2] = a; P.stats[((size_t)blockIdx.x * P.Cor + co0 + ch) * 2 + 1] = b;
      }
    }
    if (dbg && tid == 0) { dbg[20] = t_acc; dbg[21] = clock64() - t_start; dbg[22] = nt; dbg[23] = total_g; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (PR || MC) cluster_sync_all();                                   // nobody leaves while the peer may still be reached (remote arrives, multicast writes, the leader's MMAs reading this CTA's operands)
  if (dbg && tid == 0) dbg[0] = clock64() - t_start;
  if (warp == 0) {
    if (PR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols));
  }
}

static bool conv_v1() { static const bool v = getenv("CATGEN_CONV_V1") != nullptr; return v; }   // the round-1 kernel, kept for A/B measurements

static bool linear_shape_ok(int B);
// Zmax > 1: the caller allows a K split (nn.Linear: k = 1, at most four tiles); the partial sums go through workspace #1.
bool conv_tc_all_shapes_taken(int B) { return ctx().conv_engine == 1 && ctx().precision == 0 && !conv_v1() && getenv("CATGEN_DGRAD_TF32") == nullptr && linear_shape_ok(B); }
static int conv_ps_run(const float* x, const float* Wp, const float* bias, float* y, int N, int H, int W, int Cir, int Cor, int k, const float* scale2,
                       const uint8_t* xq_prepacked, int Zmax = 1) {
  // compensated operands (cg_set_precision(1), forward executors only): the same kernel over 3 * Cg virtual channels
  const bool split = ctx().split_fwd > 0 && !xq_prepacked && !scale2 && Zmax == 1;
  const int Cir_real = Cir, Cg = ((Cir + 31) / 32) * 32;
  if (split) Cir = 3 * Cg;
  const int Ci = ((Cir + 63) / 64) * 64, Co = ((Cor + 15) / 16) * 16;       // operand padding of the packed activations (shared with the weight gradient)
  const int p = (k - 1) / 2, kk = k * k;
  const int Hq = ((H + 15) / 16) * 16 + 2 * p, Wq = W + 2 * p, Hp = 16 + 2 * p, Wpx = 8 + 2 * p;
  int NB = Co > 128 ? 128 : Co; while (NB >= 16 && Co % NB) NB -= 16;
  if (NB < 16) return CG_ERR_UNSUPPORTED;
  const int ncb = (Cir + 31) / 32;
  // CTA pairs (cta_group::2) where the layer needs no K split: each CTA then holds half of every slice
  static const int pair_env = getenv("CATGEN_PS_PAIR") ? atoi(getenv("CATGEN_PS_PAIR")) : 0;   // off until the pair path has passed the parity suite on the GPU
  const int ntiles_pre = N * (W / 8) * ((H + 15) / 16);
  const bool pair = pair_env && Zmax == 1 && NB >= 32 && ntiles_pre >= 2 && (ctx().sm_count % 2) == 0;
  // 8 x 8 images: a tile is a PAIR of images (make_patch_tmap_duo) -- every MMA row is a real pixel
  static const int duo_env = getenv("CATGEN_PS_DUO") ? atoi(getenv("CATGEN_PS_DUO")) : 1;
  const int Hp_duo = 8 + 2 * p;
  // (only where the larger interleaved patches still fit eight times: 7 x 7 filters do not -- and declining here must NOT hand the layer to the
  //  round-1 kernel, which packs its own weight slices from the fp32 operands that are not refreshed when every shape is taken: found by
  //  test_closures_match_oracle_at_baseline_size, D outputs off by 1.2e-3)
  const bool duo = duo_env && H == 8 && W == 8 && (N % 2) == 0 && N >= 2 && Zmax == 1 && !pair &&
                   2 * PS_SLOTS * ((size_t)4 * 2 * Hp_duo * Wpx * 16) + 4 * 32 * 36 * sizeof(float) + 3 * ((size_t)4 * NB * 16) <= (size_t)224 * 1024;
  const size_t patch_bytes = duo ? (size_t)4 * 2 * Hp_duo * Wpx * 16 : (size_t)4 * Hp * Wpx * 16, slice_bytes = (size_t)4 * (pair ? NB / 2 : NB) * 16, stage_bytes = 4 * 32 * 36 * sizeof(float);
  const size_t budget = 224 * 1024;   // opt-in limit 227 KB per block minus the static part (barriers + 1 KB reserved: cuobjdump -res-usage says 1536 B)
  if (2 * PS_SLOTS * patch_bytes + stage_bytes + 3 * slice_bytes > budget) return CG_ERR_UNSUPPORTED;
  int S = (int)((budget - 2 * PS_SLOTS * patch_bytes - stage_bytes) / slice_bytes); if (S > 16) S = 16;
  const int ntiles = duo ? N / 2 : N * (W / 8) * ((H + 15) / 16);
  // clusters of two CTAs sharing one multicast weight stream (k_conv_ps<2>): every weight slice crosses L2 -> SM once per cluster
  static const int mc_env = getenv("CATGEN_PS_MC") ? atoi(getenv("CATGEN_PS_MC")) : 0;
  const bool mc = mc_env && !pair && Zmax == 1 && ntiles >= 2 && ctx().sm_count / (Co / NB) >= 2;
  size_t xq_bytes = (size_t)N * (Ci / 8) * Hq * Wq * 16, wq_bytes = (size_t)kk * ncb * 32 * Co * 2;
  const uint8_t* wq_cached = nullptr;
  if (!split) { auto it = wslice_registry().find(Wp); if (it != wslice_registry().end() && it->second.CB == 32) wq_cached = it->second.wq; }
  uint8_t* ws = (uint8_t*)workspace3((xq_prepacked ? 0 : ((xq_bytes + 255) & ~(size_t)255)) + (wq_cached ? 0 : wq_bytes) + 512);
  if (!ws) return CG_ERR_CUDA;
  const uint8_t* xq = xq_prepacked ? xq_prepacked : ws;
  const uint8_t* wq = wq_cached;
  long nx = (long)(xq_bytes / 16), nw = (long)(wq_bytes / 16);
  if (split) {
    ctx().next_bytes = 4.0 * N * H * W * Cir_real + (double)xq_bytes;
    CG_LAUNCH(k_pack_act_split, grid1d(nx, 256), 256, 0, x, ws, nx, H, W, Cir_real, Cg, Ci, p, Hq, Wq);
    uint8_t* wqb = ws + ((xq_bytes + 255) & ~(size_t)255);
    CG_LAUNCH(k_pack_wslices32_split, grid1d(nw, 256), 256, 0, Wp, wqb, nw, Cir_real, Cor, Co, kk, Cg / 32);
    wq = wqb;
  } else {
  if (!xq_prepacked) { ctx().next_bytes = 4.0 * N * H * W * Cir + (double)xq_bytes; CG_LAUNCH(k_pack_act<2>, grid1d(nx, 256), 256, 0, x, ws, nx, H, W, Cir, Ci, p, Hq, Wq, scale2); }
  if (!wq_cached) {
    uint8_t* wqb = ws + (xq_prepacked ? 0 : ((xq_bytes + 255) & ~(size_t)255));
    CG_LAUNCH(k_pack_wslices32, grid1d(nw, 256), 256, 0, Wp, wqb, nw, Cir, Cor, Co, kk);
    wq = wqb;
  }
  }
  PsParams P{};
  P.xq = xq; P.wq = wq; P.bias = bias; P.scale2 = scale2; P.y = y; P.inv_host = 1.f;
  P.N = N; P.H = H; P.W = W; P.Co = Co; P.Cor = Cor; P.k = k; P.p = p;
  P.ncb = ncb; P.kk = kk; P.nslices = ncb * kk;
  P.tiles_x = W / 8; P.tiles_y = (H + 15) / 16; P.ntiles = ntiles; P.NB = NB; P.S = S; P.duo = duo ? 1 : 0;
  // K split: only worth it when the tile grid alone leaves most SMs idle (Linear 20480 -> 256 at batch 128 is ONE tile x two column blocks)
  int Z = 1;
  if (Zmax > 1 && kk == 1 && ntiles <= PS_SLOTS) {
    Z = ctx().sm_count / (Co / NB); if (Z > Zmax) Z = Zmax; if (Z > P.nslices / 2) Z = P.nslices / 2; if (Z < 1) Z = 1;
  }
  P.Z = Z;
  P.amax_out = nullptr;
  if (Z == 1) { P.amax_out = ctx().next_amax; ctx().next_amax = nullptr; }
  // one-shot request of the caller (model.cu, G's backward): fuse the 2 x 2 sum of SpatialUpSamplingNearest's backward into this launch's epilogue.
  // Honoured only on the wide epilogue path of a plain (non-duo, non-split) launch; ctx().pool2_done tells the caller.
  P.pool2 = 0; P.y2 = nullptr; P.stats = nullptr;
  if (ctx().next_pool2_out) {
    if (Z == 1 && !duo && !pair && !bias && (Cor & 3) == 0 && (NB & 31) == 0 && Co == Cor && (H & 1) == 0 && (W & 1) == 0) { P.pool2 = 1; P.y2 = ctx().next_pool2_out; ctx().pool2_done = 1; }
    ctx().next_pool2_out = nullptr;
  }     // with a K split the final values come out of splitk_reduce, which takes it
  float* part = nullptr;
  if (Z > 1) {
    part = (float*)workspace(sizeof(float) * (size_t)Z * N * H * W * Cor + 256); if (!part) return CG_ERR_CUDA;
    P.y = part; P.bias = nullptr;
  }
  static const int d_env = getenv("CATGEN_PS_D") ? atoi(getenv("CATGEN_PS_D")) : 16;
  P.D = d_env < 0 ? 0 : d_env;
  if (P.D > P.nslices / 4) P.D = P.nslices / 4;   // short K loops: keep the four slots overlapping (stagger < a quarter cycle)
  P.patch_bytes = (uint32_t)patch_bytes; P.slice_bytes = (uint32_t)slice_bytes;
  const size_t smem = 2 * PS_SLOTS * patch_bytes + (size_t)S * slice_bytes + stage_bytes;
  static bool attr_done = false;
  if (!attr_done) {
    CG_CUDA(cudaFuncSetAttribute(k_conv_ps<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    CG_CUDA(cudaFuncSetAttribute(k_conv_ps<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    CG_CUDA(cudaFuncSetAttribute(k_conv_ps<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    attr_done = true;
  }
  const int gy = Co / NB;
  int gx = ctx().sm_count / gy; if (gx < 1) gx = 1;
  if (pair) { const int units = (ntiles + 1) / 2; int pairs = gx / 2; if (pairs < 1) pairs = 1; if (pairs > units) pairs = units; gx = 2 * pairs; }
  else
  // Fill the machine first: measured (gpurun_out/r02_e_layers.txt vs r02_d_layers_v1.txt), insisting on four tiles per CTA left most SMs
  // idle on the 8x8 / 16x16 layers (128 -> 128 7x7 at 8x8: 32 CTAs, 117 us against 52 us with 64 CTAs) -- the tensor pipe, not the
  // weight stream, bounds a CTA, so sharing slices among fewer, busier CTAs buys nothing there.
  if (gx > ntiles) gx = ntiles;
  // ... but at least CATGEN_PS_MINTILES tiles per CTA, balanced.  Alone, every layer is fastest with one tile per CTA; inside the training step the
  // opposite holds (profiles/r02_tiles_per_cta.txt: 5.46 ms at 1, 5.20 ms at 4): a k_conv_ps CTA owns its SM (224 KB of shared memory), so a
  // convolution that fills the machine serialises D's four branches and the generator forward running ahead; at four tiles per CTA the small
  // convolutions take 32-64 SMs each and run side by side.
  // The SAME value on the main stream and inside lanes by default: a tile's K-accumulation order depends on the slot it lands in, i.e. on the
  // grid, so a lane-dependent grid made "concurrent lanes" and "one stream" differ in the last fp32 bits (and, through fp16 operand rounding
  // and Adam's sign-like first steps, visibly: three schedule-equivalence tests failed with main 1 / lane 4, which was 0.02 ms faster).
  static const int mintiles_env = getenv("CATGEN_PS_MINTILES") ? atoi(getenv("CATGEN_PS_MINTILES")) : 4;
  static const int mintiles_lane = getenv("CATGEN_PS_MINTILES_LANE") ? atoi(getenv("CATGEN_PS_MINTILES_LANE")) : mintiles_env;   // inside a concurrent lane (D's branches, the G-ahead lane)
  const int mt = ctx().lane >= 0 ? mintiles_lane : mintiles_env;
  if (!pair && Z == 1 && mt > 0) {
    int per = (ntiles + gx - 1) / gx; if (per < mt) per = mt; if (per > ntiles) per = ntiles;
    gx = (ntiles + per - 1) / per;
  }
  if (mc) gx &= ~1;                                                       // whole clusters; every CTA keeps at least one tile (gx <= ntiles)
  if (Z > 1) gx = Z;
  // one-shot request (model.cu, G's forward): per-CTA batch-norm partial sums from the epilogue instead of a separate pass over the output
  if (ctx().next_stats_part) {
    if (Z == 1 && !pair && (Cor & 3) == 0 && (NB & 31) == 0 && Co == Cor && gx <= ctx().next_stats_cap) { P.stats = ctx().next_stats_part; ctx().stats_rows = gx; }
    ctx().next_stats_part = nullptr;
  }
  dim3 grid(gx, gy);
  ctx().next_flops = 2.0 * (double)N * H * W * Cor * kk * (split ? Cir_real : Cir);             // algorithmic (unpadded) work; the compensation MMAs are overhead, not work
  ctx().next_bytes = (double)xq_bytes + (double)wq_bytes + 4.0 * (double)N * H * W * Cor;
  CUtensorMap tmx, tmw;
  if (duo) { int ts = make_patch_tmap_duo(&tmx, xq, N, Ci / 8, Hq, Wq, Hp_duo, Wpx, 4); if (ts == CG_ERR_UNSUPPORTED) return set_err(CG_ERR_CUDA, "duo tensor map rejected by the driver (set CATGEN_PS_DUO=0)"); CG_TRY(ts); }
  else CG_TRY(make_patch_tmap(&tmx, xq, 2, N, Ci / 8, Hq, Wq, Hp, Wpx, 4));
  CG_TRY(make_wslice_tmap(&tmw, wq, Co, (long)P.nslices * 4, pair ? NB / 2 : NB));
  static const bool dbg_on = getenv("CATGEN_PS_DBG") != nullptr;
  static long long* dbg_buf = nullptr;
  if (dbg_on && !dbg_buf) cudaMalloc(&dbg_buf, sizeof(long long) * 32 * 1024);
  P.dbg = (dbg_on && (long)gx * gy <= 1024) ? dbg_buf : nullptr;
  if (P.dbg) cudaMemsetAsync(dbg_buf, 0, sizeof(long long) * 32 * 1024, ctx().stream);
  if (pair || mc) {   // cluster of two CTAs along x = one TPC
    cudaLaunchConfig_t cfg{}; cfg.gridDim = grid; cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem; cfg.stream = ctx().stream;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (ctx().prof_on) prof_begin(mc ? "k_conv_ps<mc>" : "k_conv_ps<pair>");
    cudaError_t le = mc ? cudaLaunchKernelEx(&cfg, k_conv_ps<2>, P, tmx, tmw) : cudaLaunchKernelEx(&cfg, k_conv_ps<1>, P, tmx, tmw);
    if (ctx().prof_on) prof_end();
    ctx().next_flops = 0; ctx().next_bytes = 0; ctx().launches++;
    if (ctx().trace_launches) fprintf(stderr, "[launch] k_conv_ps<cluster>\n");
    if (le != cudaSuccess) return set_err(CG_ERR_CUDA, "%s:%d cluster launch k_conv_ps<%d> -> %s", __FILE__, __LINE__, mc ? 2 : 1, cudaGetErrorString(le));
  } else
  CG_LAUNCH(k_conv_ps<0>, grid, 320, smem, P, tmx, tmw);
  if (P.dbg) {   // experiments only: where does a CTA's time go (cycles, mean over CTAs)
    cudaStreamSynchronize(ctx().stream);
    size_t n = (size_t)gx * gy;
    std::vector<long long> h(n * 32);
    cudaMemcpy(h.data(), dbg_buf, sizeof(long long) * 32 * n, cudaMemcpyDeviceToHost);
    double m[32] = {0}; double tmax = 0;
    for (size_t i = 0; i < n; ++i) { for (int q = 0; q < 32; ++q) m[q] += (double)h[i * 32 + q] / n; if ((double)h[i * 32] > tmax) tmax = (double)h[i * 32]; }
    const double floor_t = m[22] * P.nslices * 2.0 * (128.0 * NB / 256.0);
    fprintf(stderr, "[ps dbg] N=%d %dx%d Ci=%d Co=%d k=%d grid=%dx%d NB=%d S=%d D=%d ns=%d | per CTA: total %.0f (max %.0f) tiles %.1f slices %.0f MMA-floor %.0f | "
                    "weights: producer waits-for-empty %.0f | issuers wait weights %.0f %.0f %.0f %.0f, patches %.0f %.0f %.0f %.0f, tmem-free %.0f %.0f %.0f %.0f | epilogue waits-for-acc %.0f of %.0f\n",
            N, H, W, Cir, Cor, k, gx, gy, NB, S, P.D, P.nslices, m[0], tmax, m[22], m[23], floor_t, m[1],
            m[4], m[8], m[12], m[16], m[5], m[9], m[13], m[17], m[6], m[10], m[14], m[18], m[20], m[21]);
  }
  if (Z > 1) CG_TRY(splitk_reduce(part, Z, (long)N * H * W * Cor, Cor, bias, y));   // fixed z order; adds the bias
  return CG_OK;
}

// nn.Linear on the tensor cores (models.lua:199,697,700,852,854): y[B, out] = x[B, in] W^T + b is the 1x1 convolution of
// the "image" whose PIXELS are the batch rows -- x viewed as [B/128 images][16 rows][8 px][in channels] is the same memory as x[B][in], a
// tile of 8 x 16 pixels is exactly one M = 128 MMA tile -- so the conv engine takes it unchanged with k = 1, split along K when the
// tile grid is too small to fill the machine (k_conv_ps, K-split mode).
static bool linear_shape_ok(int B) { return B >= 8 && B % 8 == 0 && (B <= 128 || B % 128 == 0); }
static int linear_tc_run(const float* x, const float* Wp, const float* bias, float* y, int B, int Cir, int Cor, const float* scale2, const uint8_t* xq_prepacked) {
  const int Nimg = B <= 128 ? 1 : B / 128, Hh = B <= 128 ? B / 8 : 16;
  return conv_ps_run(x, Wp, bias, y, Nimg, Hh, 8, Cir, Cor, 1, scale2, xq_prepacked, 1 << 20);
}
// ====================================================================================================================

template <int ES>
static int conv_tc_run(const float* x, const float* Wp, const float* bias, float* y, int N, int H, int W, int Cir, int Cor, int k, const float* scale2 = nullptr,
                       const uint8_t* xq_prepacked = nullptr) {   // xq_prepacked: the operand already in blocked/padded form (shared gradient operand)
  if (ES == 2 && !conv_v1()) { int s4 = conv_ps_run(x, Wp, bias, y, N, H, W, Cir, Cor, k, scale2, xq_prepacked); if (s4 != CG_ERR_UNSUPPORTED) return s4; }
  if (ctx().split_fwd > 0 && !scale2 && !xq_prepacked) return CG_ERR_UNSUPPORTED;   // compensated mode: what the round-2 kernel declines goes to the fp32 kernel, never to single fp16
  constexpr int PER = 16 / ES, KB = 128 / ES;
  const int Ci = ((Cir + KB - 1) / KB) * KB, Co = ((Cor + 15) / 16) * 16;   // padded sizes the kernel iterates over
  const int p = (k - 1) / 2, kk = k * k;
  const int Hq = ((H + 15) / 16) * 16 + 2 * p, Wq = W + 2 * p;
  const int Hp = 16 + 2 * p, Wpx = 8 + 2 * p;
  TcParams P{};
  int NB, TL, CB;
  tc_plan<ES>(Ci, Co, k, N * (W / 8) * ((H + 15) / 16), &NB, &TL, &CB);
  size_t slice_bytes = (size_t)(KB / PER) * NB * 16;
  const int ntiles = N * (W / 8) * ((H + 15) / 16);
  if (!CB) return CG_ERR_UNSUPPORTED;
  size_t patch_bytes = (size_t)(CB / PER) * Hp * Wpx * 16;
  int S = (int)((216 * 1024 - 2 * TL * patch_bytes) / slice_bytes); if (S > 8) S = 8;
  static const int s_cap = getenv("CATGEN_TC_RING") ? atoi(getenv("CATGEN_TC_RING")) : 8;   // experiments: cap the weight ring depth
  if (S > s_cap && s_cap >= 2) S = s_cap;
  if (S < 2) return CG_ERR_UNSUPPORTED;
  size_t smem = 2 * TL * patch_bytes + (size_t)S * slice_bytes;
  // operand buffers: activations then weight slices (16-byte aligned)
  size_t xq_bytes = (size_t)N * (Ci / PER) * Hq * Wq * 16, wq_bytes = (size_t)kk * Ci * Co * ES;
  const uint8_t* wq_cached = nullptr;
  if (ES == 2) { auto it = wslice_registry().find(Wp); if (it != wslice_registry().end() && it->second.CB == CB) wq_cached = it->second.wq; }
  if (!wq_cached && ctx().fp32_operands_stale > 0)   // this kernel would pack its slices from an fp32 operand the model executor did not refresh
    return set_err(CG_ERR_STATE, "round-1 conv kernel asked for fp32 operands that are stale (N=%d %dx%d %d->%d k=%d)", N, H, W, Cir, Cor, k);
  uint8_t* ws = (uint8_t*)workspace3((xq_prepacked ? 0 : xq_bytes) + (wq_cached ? 0 : wq_bytes) + 512);
  if (!ws) return CG_ERR_CUDA;
  const uint8_t* xq = xq_prepacked ? xq_prepacked : ws;
  const uint8_t* wq = wq_cached;
  long nx = (long)(xq_bytes / 16), nw = (long)(wq_bytes / 16);
  if (!xq_prepacked) CG_LAUNCH(k_pack_act<ES>, grid1d(nx, 256), 256, 0, x, ws, nx, H, W, Cir, Ci, p, Hq, Wq, scale2);
  if (!wq_cached) {
    uint8_t* wqb = ws + (xq_prepacked ? 0 : ((xq_bytes + 255) & ~(size_t)255));
    CG_LAUNCH(k_pack_wslices<ES>, grid1d(nw, 256), 256, 0, Wp, wqb, nw, Cir, Cor, Co, kk, CB);
    wq = wqb;
  }
  static const bool dbg_on = getenv("CATGEN_TC_DBG") != nullptr;
  static long long* dbg_buf = nullptr;
  if (dbg_on && !dbg_buf) cudaMalloc(&dbg_buf, sizeof(long long) * 8 * 65536);
  P.dbg = dbg_on ? dbg_buf : nullptr;
  P.nostore = (dbg_on && atoi(getenv("CATGEN_TC_DBG")) == 2) ? 1 : 0;
  static const int rot_on = getenv("CATGEN_TC_ROT") ? atoi(getenv("CATGEN_TC_ROT")) : 0;
  P.rot = rot_on;
  P.xq = xq; P.wq = wq; P.bias = bias; P.y = y; P.scale2 = scale2;
  P.N = N; P.H = H; P.W = W; P.Ci = Ci; P.Co = Co; P.Cor = Cor; P.k = k; P.p = p; P.Hq = Hq; P.Wq = Wq;
  P.CB = CB; P.ncb = Ci / CB; P.tiles_x = W / 8; P.tiles_y = (H + 15) / 16; P.NB = NB; P.S = S; P.TL = TL; P.ntiles = ntiles;
  P.patch_bytes = (uint32_t)patch_bytes; P.slice_bytes = (uint32_t)slice_bytes;
  static bool attr_done[2] = {false, false};
  if (!attr_done[ES == 2 ? 0 : 1]) {
    CG_CUDA(cudaFuncSetAttribute(k_conv_tc<ES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_done[ES == 2 ? 0 : 1] = true;
  }
  dim3 grid((ntiles + TL - 1) / TL, Co / NB);
  ctx().next_flops = 2.0 * (double)N * H * W * Cor * kk * Cir;   // algorithmic (unpadded) work
  ctx().next_bytes = (double)xq_bytes + (double)wq_bytes + 4.0 * (double)N * H * W * Cor;
  CUtensorMap tmx;
  CG_TRY(make_patch_tmap(&tmx, xq, ES, N, Ci / PER, Hq, Wq, Hp, Wpx, CB / PER));
  if (ES == 2) CG_LAUNCH(k_conv_tc<2>, grid, 256, smem, P, tmx);
  else CG_LAUNCH(k_conv_tc<4>, grid, 256, smem, P, tmx);
  if (dbg_on && dbg_buf && (long)grid.x * grid.y <= 65536) {   // experiments only: where does a CTA's time go (cycles, mean over CTAs)
    cudaStreamSynchronize(ctx().stream);
    size_t n = (size_t)grid.x * grid.y;
    std::vector<long long> h(n * 8);
    cudaMemcpy(h.data(), dbg_buf, sizeof(long long) * 8 * n, cudaMemcpyDeviceToHost);
    double tot = 0, pro = 0, loop = 0, ww = 0, pw = 0, drain = 0, epi = 0;
    for (size_t i = 0; i < n; ++i) {
      const long long* d = &h[i * 8];
      tot += d[6] - d[0]; pro += d[1] - d[0]; loop += d[2] - d[1]; ww += d[3]; pw += d[4]; drain += d[5] - d[2]; epi += d[6] - d[5];
    }
    fprintf(stderr, "[tc dbg] N=%d HxW=%dx%d Ci=%d Co=%d k=%d grid=%ux%u TL=%d CB=%d NB=%d S=%d | per CTA cycles: total %.0f = prologue %.0f + issue loop %.0f (waiting: weights %.0f, patch %.0f) + drain %.0f + epilogue %.0f | MMA floor %.0f\n",
            N, H, W, Cir, Cor, k, grid.x, grid.y, TL, CB, NB, S, tot / n, pro / n, loop / n, ww / n, pw / n, drain / n, epi / n,
            (double)TL * (Ci / (128 / ES)) * kk * 4 * (128.0 * NB / 256.0));
  }
  return CG_OK;
}

// Gradient-valued inputs (dgrad: the "activation" operand is gy, |g| ~ 1e-6) are multiplied by a per-tensor power of two before
// the fp16 conversion and the result is divided by it in the epilogue.  tools/backward_precision_study.py: identical to tf32
// accuracy (1.0e-4 .. 1.9e-4 of max on G's gradients) at twice the tensor rate; UNSCALED fp16 is 100-700x worse.
// The tf32 instantiation stays available (CATGEN_DGRAD_TF32=1) as the reference it replaced.
static thread_local int g_tc_grad_operands = 0;
void conv_tc_set_gradient_operands(int on) { g_tc_grad_operands = on; }
__global__ void k_absmax(const float* __restrict__ x, long n, unsigned int* __restrict__ out);
__global__ void k_make_scale(const unsigned int* __restrict__ amax, float* __restrict__ scale2);

int absmax_into(const float* x, long n, unsigned int* amax) { CG_LAUNCH(k_absmax, grid1d(n, 256, 8), 256, 0, x, n, amax); return CG_OK; }
static float* tc_scale_scratch() {   // per lane: [scale, 1/scale, amax bits, amax bits (fused path, kept zero between uses)]; its own allocation, workspace3 may be regrown by the run
  static float* p[Ctx::kLanes + 1] = {nullptr};
  float*& q = p[ctx().lane + 1];
  if (!q) { if (cudaMalloc(&q, 16 * sizeof(float)) != cudaSuccess || cudaMemset(q, 0, 16 * sizeof(float)) != cudaSuccess) q = nullptr; }
  return q;
}

// ONE packed gradient operand per layer: gq[N][Cg/8][Hq][Wq][8] fp16, zero-padded by the filter radius, channels padded to 64,
// multiplied by the per-tensor power of two.  dgrad reads halo'd patches from it (A operand), wgrad reads 16x8-pixel tiles
// from it (B operand) through a second tensor map.  Before this, gy was read three times per layer (absmax, tf32 pack, fp16 tile pack).
struct GradOperand { const uint8_t* gq; const float* scale2; int Cg; };
// gb_acc (optional): also add the bias gradient (column sums of gy) -- it comes out of the same pass as max|gy|
static int pack_grad_operand(const float* gy, int N, int H, int W, int C, int k, GradOperand* out, float* gb_acc = nullptr) {
  const int p = (k - 1) / 2, Hq = ((H + 15) / 16) * 16 + 2 * p, Wq = W + 2 * p, Cg = ((C + 63) / 64) * 64;
  CG_TRY(side_wait());   // a weight-gradient chain on the side stream may still read the operand and scale rewritten below
  float* sc = tc_scale_scratch(); if (!sc) return set_err(CG_ERR_CUDA, "scale scratch allocation failed");
  unsigned int* amax = (unsigned int*)(sc + 2);
  size_t bytes = (size_t)N * (Cg / 8) * Hq * Wq * 16;
  uint8_t* gq = (uint8_t*)workspace4(bytes + 256); if (!gq) return CG_ERR_CUDA;
  long n = (long)N * H * W * C;
  if (gb_acc) CG_TRY(colsum_acc_absmax(gy, gb_acc, (long)N * H * W, C, amax + 1, sc));
  else {
    CG_CUDA(cudaMemsetAsync(amax, 0, sizeof(unsigned int), ctx().stream));
    CG_LAUNCH(k_absmax, grid1d(n, 256, 8), 256, 0, gy, n, amax);
    CG_LAUNCH(k_make_scale, 1, 1, 0, amax, sc);
  }
  long nq = (long)(bytes / 16);
  CG_LAUNCH(k_pack_act<2>, grid1d(nq, 256), 256, 0, gy, gq, nq, H, W, C, Cg, p, Hq, Wq, (const float*)sc);
  out->gq = gq; out->scale2 = sc; out->Cg = Cg;
  return CG_OK;
}

// ---- cached operand path (G's stages): the producer above writes the operand once, forward and weight gradient both read it.
// The forward operand (Cin padded to 64) and the weight-gradient operand (Cin 64 or a multiple of 128) must be the same buffer.
bool conv_tc_cached_ok(int H, int W, int Ci, int Co, int k) {
  static const bool dgrad_tf32 = getenv("CATGEN_DGRAD_TF32") != nullptr;
  return !dgrad_tf32 && (Ci == 64 || Ci % 128 == 0) && tc_shape_ok(H, W, Ci, Co, k, 2) && tc_shape_ok(H, W, Co, Ci, k, 2);   // implies wgrad_shape_ok
}
size_t conv_tc_operand_bytes(int N, int H, int W, int Ci, int k) {
  const int p = (k - 1) / 2;
  return (size_t)N * (Ci / 8) * (((H + 15) / 16) * 16 + 2 * p) * (W + 2 * p) * 16;
}
// x: [N,h,w,C] conv output (or the Linear output viewed NHWC); operand for a k x k conv on the (h<<up) x (w<<up) image
int bn_prelu_up_pack(const float* x, const float* gamma, const float* beta, const float* mean, const float* invstd, const float* pw,
                     float* bn_out, uint8_t* xq, int N, int h, int w, int C, int up, int k) {
  const int p = (k - 1) / 2, H = h << up, W = w << up;
  const int Hq = ((H + 15) / 16) * 16 + 2 * p, Wq = W + 2 * p;
  long n = (long)N * (C / 8) * Hq * Wq;
  ctx().next_bytes = 4.0 * N * h * w * C * (bn_out ? 2 : 1) + 16.0 * n;
  CG_LAUNCH(k_bn_prelu_up_pack, grid1d(n, 256), 256, 0, x, gamma, beta, mean, invstd, pw, bn_out, xq, n, h, w, C, up, p, Hq, Wq);
  return CG_OK;
}
int conv_fwd_tc_packed(const uint8_t* xq, const float* Wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co, int k) {
  return conv_tc_run<2>(nullptr, Wp, bias, y, N, H, W, Ci, Co, k, nullptr, xq);
}

int conv_fwd_tc(const float* x, const float* Wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co, int k) {
  static const bool dgrad_tf32 = getenv("CATGEN_DGRAD_TF32") != nullptr;
  if (k == 1 && H == 1 && W == 1 && ctx().split_fwd > 0 && !g_tc_grad_operands) return CG_ERR_UNSUPPORTED;   // compensated mode: nn.Linear forwards run on the fp32 kernel
  if (k == 1 && H == 1 && W == 1 && !conv_v1() && !dgrad_tf32 && linear_shape_ok(N) && (((uintptr_t)x | (uintptr_t)y) & 15) == 0) {   // nn.Linear
    if (!g_tc_grad_operands) return linear_tc_run(x, Wp, bias, y, N, Ci, Co, nullptr, nullptr);
    const int Nimg = N <= 128 ? 1 : N / 128, Hh = N <= 128 ? N / 8 : 16;
    GradOperand g; CG_TRY(pack_grad_operand(x, Nimg, Hh, 8, Ci, 1, &g));                    // gradient-valued input: per-tensor power-of-two scale
    return linear_tc_run(x, Wp, bias, y, N, Ci, Co, g.scale2, g.gq);
  }
  if (!tc_shape_ok(H, W, Ci, Co, k, 2)) return CG_ERR_UNSUPPORTED;
  if ((((uintptr_t)x | (uintptr_t)y) & 15) != 0) return CG_ERR_UNSUPPORTED;
  if (!g_tc_grad_operands) return conv_tc_run<2>(x, Wp, bias, y, N, H, W, Ci, Co, k);
  if (dgrad_tf32) return conv_tc_run<4>(x, Wp, bias, y, N, H, W, Ci, Co, k);
  GradOperand g; CG_TRY(pack_grad_operand(x, N, H, W, Ci, k, &g));
  return conv_tc_run<2>(x, Wp, bias, y, N, H, W, Ci, Co, k, g.scale2, g.gq);
}

// =================================================================== weight gradient on the tensor cores
//   gWp[(tap,ci)][co] = sum_{n,y,x} x[n, y+ky-p, x+kx-p, ci] * gy[n,y,x,co]
// GEMM per tap: D_tap[ci (M=128)][co (N=NB)] += A^T B over K = pixels (16 per instruction = 2 rows of 8).
// Both operands are MN-major fp16 (tools/tc_probe.cu rows 10/12): the x patch is the SAME shared-memory image the
// forward kernel uses (plane = 8 ci, 16-byte pixels; LBO = patch pitch, SBO = plane stride) and a filter tap is again
// a shifted start address; gy is pre-tiled [tile][co chunk][16 rows][8 px][8 co] so one bulk copy fetches a tile.
// gy is gradient-valued (|g| ~ 1e-6): it is multiplied by a per-tensor power of two before the fp16 conversion and
// the result is divided by it in the epilogue (tools/backward_precision_study.py: identical to tf32 accuracy).
// Up to 512/NB taps accumulate side by side in TMEM; CTAs split the pixel range and write partial sums that
// conv_ref.cu's fixed-order reduction adds up (deterministic).
__global__ void k_absmax(const float* __restrict__ x, long n, unsigned int* __restrict__ out) {
  float m = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));   // non-negative floats order like uints
}
__global__ void k_make_scale(const unsigned int* __restrict__ amax, float* __restrict__ scale2) {
  float m = __uint_as_float(*amax);
  float sc = 1.f;
  if (m > 0.f && isfinite(m)) sc = exp2f(floorf(log2f(16384.f / m)));
  if (!(sc > 0.f) || !isfinite(sc)) sc = 1.f;
  scale2[0] = sc; scale2[1] = 1.f / sc;
}
struct TcWParams {
  const uint8_t* xq; const uint8_t* gq; float* part; const float* scale2;
  int N, H, W, Ci, Co, k, p, Hq, Wq;   // Ci/Co: PADDED counts (operand addressing)
  int Cir, Cor;                        // real counts: rows per tap and row stride of the partial sums
  int tiles_x, tiles_y, tiles_total;
  int NB, TG, ntg, ncib, cim, ncob, Z;
  uint32_t patch_bytes, patch_load_bytes, g_bytes;
};

__global__ void __launch_bounds__(256, 1) k_wgrad_tc(TcWParams P, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmg) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_acc;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Hp = 16 + 2 * P.p, Wp = 8 + 2 * P.p;
  const uint32_t plane_bytes = (uint32_t)Hp * Wp * 16;
  const uint32_t stage_bytes = P.patch_bytes + P.g_bytes;
  // CTA -> (pixel split z, ci block, co block, tap group)
  const int z = blockIdx.x;
  int b = blockIdx.y; const int tg = b % P.ntg; b /= P.ntg; const int cob = b % P.ncob; const int cib = b / P.ncob;
  const int tap0 = tg * P.TG, kk = P.k * P.k;
  const int ntap = (tap0 + P.TG <= kk) ? P.TG : (kk - tap0);
  const long t0 = (long)P.tiles_total * z / P.Z, t1 = (long)P.tiles_total * (z + 1) / P.Z;

  if (tid == 0) {
    // two MMA issuers (warps 6, 7) take alternate filter taps: one thread retires an M=128 MMA per ~84 cycles against the pipe's
    // 64 (tools/tc_rate.cu, profiles/r01_tc_rate.txt); a stage is free / the accumulators are complete when BOTH have committed
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_full[i], 2); mbar_init(&bar_empty[i], 2); }
    mbar_init(&bar_acc, 2);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  int ncols = 32; while (ncols < P.TG * P.NB) ncols <<= 1;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (P.cim < 128) {   // Ci = 64: rows 64..127 of the M = 128 instruction read zero planes (written once, never reloaded)
    for (int st = 0; st < 2; ++st) {
      uint4* zp = reinterpret_cast<uint4*>(smem + (size_t)st * stage_bytes + P.patch_load_bytes);
      for (uint32_t i = tid; i < (P.patch_bytes - P.patch_load_bytes) / 16; i += blockDim.x) zp[i] = make_uint4(0, 0, 0, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;

  if (warp == 4) {
    if (lane == 0) {   // gy tiles: one tiled TMA each, [NB/8 planes][16 rows][8 px][16 B] out of the shared blocked operand
      int it = 0;
      for (long t = t0; t < t1; ++t, ++it) {
        int buf = it & 1;
        int tx = (int)(t % P.tiles_x); long q = t / P.tiles_x; int ty = (int)(q % P.tiles_y); int n = (int)(q / P.tiles_y);
        mbar_wait(&bar_empty[buf], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&bar_full[buf], P.g_bytes);
        tma_patch_4d(smem + (size_t)buf * stage_bytes + P.patch_bytes, &tmg, (tx * 8 + P.p) * 8, ty * 16 + P.p, cob * (P.NB / 8), n, &bar_full[buf]);
      }
    }
  } else if (warp == 5) {   // x patches: ONE tiled TMA per tile
    if (lane == 0) {
      int it = 0;
      for (long t = t0; t < t1; ++t, ++it) {
        int buf = it & 1;
        int tx = (int)(t % P.tiles_x); long q = t / P.tiles_x; int ty = (int)(q % P.tiles_y); int n = (int)(q / P.tiles_y);
        mbar_wait(&bar_empty[buf], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&bar_full[buf], P.patch_load_bytes);
        tma_patch_4d(smem + (size_t)buf * stage_bytes, &tmx, tx * 8 * 8, ty * 16, cib * 16, n, &bar_full[buf]);
      }
    }
  } else if (warp == 6 || warp == 7) {
    if (lane == 0) {
      // MN-major A and B (bits 15, 16), fp16, fp32 accumulate, M = 128, N = NB
      const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(P.NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t a_lbo = (uint32_t)Wp * 16;                        // k-group = next patch row; mn chunk (SBO) = next ci plane
      const uint32_t a_hi = desc_hi(plane_bytes), b_hi = desc_hi(2048); // B: k-group (LBO) = next 8-px row (128 B); mn chunk = next co plane
      const uint32_t a_kstep = 2 * (uint32_t)Wp, b_kstep = 16;          // two k-groups per instruction, in 16-byte units
      const int k = P.k, NB = P.NB, me = warp - 6;
      int ky0 = (tap0 + me) / k, kx0 = (tap0 + me) - ky0 * k;           // this issuer's first tap; the only division, once per CTA
      uint32_t acc0 = 0;
      int it = 0;
      for (long t = t0; t < t1; ++t, ++it) {
        const int buf = it & 1;
        mbar_wait(&bar_full[buf], (it >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t pbase = smem_u32(smem + (size_t)buf * stage_bytes);
        const uint32_t p_lo = desc_lo(pbase, a_lbo), g_lo = desc_lo(pbase + P.patch_bytes, 128);
        int ky = ky0, kx = kx0; uint32_t tm = tmem + (uint32_t)(me * NB);
        for (int tl = me; tl < ntap; tl += 2, tm += (uint32_t)(2 * NB)) {
          const uint32_t a_lo = p_lo + (uint32_t)(ky * Wp + kx);
          umma<2>(tm, desc64(a_lo, a_hi), desc64(g_lo, b_hi), idesc, acc0);
#pragma unroll
          for (int ks = 1; ks < 8; ++ks)
            umma<2>(tm, desc64(a_lo + (uint32_t)ks * a_kstep, a_hi), desc64(g_lo + (uint32_t)ks * b_kstep, b_hi), idesc, 1u);
          kx += 2; while (kx >= k) { kx -= k; ++ky; }
        }
        acc0 = 1u;
        umma_commit(&bar_empty[buf]);
      }
      umma_commit(&bar_acc);
    }
  }

  if (warp < 4) {
    mbar_wait(&bar_acc, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    const float inv = P.scale2[1];
    const int ci_l = warp * 32 + lane;
    const int ci = cib * 128 + ci_l;
    const bool row_ok = ci_l < P.cim && ci < P.Cir;
    const bool valid = row_ok && t1 > t0;
    const bool vec = (P.Cor & 3) == 0;
    // full-width column blocks leave through a per-warp shared-memory tile as whole 128-byte lines (see k_conv_tc's epilogue);
    // the stage buffers are idle once bar_acc has fired
    float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 36);
    const bool wide = vec && (P.NB & 31) == 0;
    for (int tl = 0; tl < ntap; ++tl) {
      float* out = P.part + ((size_t)z * kk * P.Cir + (size_t)(tap0 + tl) * P.Cir + (size_t)ci) * P.Cor + (size_t)cob * P.NB;
      if (wide) {
        for (int c0 = 0; c0 < P.NB; c0 += 32) {
          uint32_t v[32];
          uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tl * P.NB + c0);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                         "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                         "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                         "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                       : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;");
          const int cbase = cob * P.NB + c0;
          if (cbase + 32 <= P.Cor) {                     // warp-uniform
            const float sc = (t1 > t0) ? inv : 0.f;       // empty pixel range: zeros
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(stage + lane * 36 + j) = make_float4(__uint_as_float(v[j]) * sc, __uint_as_float(v[j + 1]) * sc, __uint_as_float(v[j + 2]) * sc, __uint_as_float(v[j + 3]) * sc);
            __syncwarp();
            const int q = (lane & 7) * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int pl = i * 4 + (lane >> 3), cl = warp * 32 + pl, cg_ = cib * 128 + cl;
              float4 o = *reinterpret_cast<const float4*>(stage + pl * 36 + q);
              if (cl < P.cim && cg_ < P.Cir)
                *reinterpret_cast<float4*>(P.part + ((size_t)z * kk * P.Cir + (size_t)(tap0 + tl) * P.Cir + (size_t)cg_) * P.Cor + (size_t)cbase + q) = o;
            }
            __syncwarp();
          } else if (row_ok) {
            for (int j = 0; j < 32; ++j) if (cbase + j < P.Cor) out[c0 + j] = valid ? __uint_as_float(v[j]) * inv : 0.f;
          }
        }
        continue;
      }
      for (int c0 = 0; c0 < P.NB; c0 += 16) {
        uint32_t v[16];
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tl * P.NB + c0);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        const int cbase = cob * P.NB + c0;
        if (row_ok && vec && cbase + 16 <= P.Cor) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(out + c0 + j) = valid ? make_float4(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv, __uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv)
                                                             : make_float4(0.f, 0.f, 0.f, 0.f);   // empty pixel range
        } else if (row_ok) {   // padded / ragged Cout
#pragma unroll
          for (int j = 0; j < 16; ++j) if (cbase + j < P.Cor) out[c0 + j] = valid ? __uint_as_float(v[j]) * inv : 0.f;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols));
}

__global__ void k_sum_parts(const float* __restrict__ part, int Z, long n, float* __restrict__ out) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int zz = 0; zz < Z; ++zz) s += part[(long)zz * n + i];   // fixed order: deterministic
    out[i] = s;
  }
}

// tile map of the blocked gradient operand: box = 8 px x 16 rows x planes
static int make_tile_tmap(CUtensorMap* tm, const void* gq, int N, int Cq, int Hq, int Wq, int planes) {
  cg_tmap_encode_fn enc = tmap_encoder();
  if (!enc) return set_err(CG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[4] = {(cuuint64_t)Wq * 8, (cuuint64_t)Hq, (cuuint64_t)Cq, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)Wq * 16, (cuuint64_t)Hq * Wq * 16, (cuuint64_t)Cq * Hq * Wq * 16};
  cuuint32_t box[4] = {64, 16, (cuuint32_t)planes, 1}, estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(gq), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(CG_ERR_CUDA, "cuTensorMapEncodeTiled (gradient tiles) failed with %d", (int)r);
  return CG_OK;
}
static bool wgrad_shape_ok(int H, int W, int Cir, int k) { return (k == 3 || k == 5 || k == 7) && W % 8 == 0 && H % 8 == 0 && (Cir <= 64 || Cir % 128 == 0); }

static int conv_wgrad_tc_impl(const float* x, const GradOperand& g, float* gWp_out, int N, int H, int W, int Cir, int Cor, int k, float* gW_acc, int* done,
                              const uint8_t* xq_prepacked = nullptr) {   // xq_prepacked: the forward's cached operand (bn_prelu_up_pack)
  // Cin <= 64 is zero-padded to 64 (rows 64..127 of the M = 128 instruction read zero planes); Cout is padded to 16
  const int Ci = Cir <= 64 ? 64 : Cir, Co = ((Cor + 15) / 16) * 16;
  const int p = (k - 1) / 2, kk = k * k;
  const int Hq = ((H + 15) / 16) * 16 + 2 * p, Wq = W + 2 * p, Hp = 16 + 2 * p, Wpx = 8 + 2 * p;
  TcWParams P{};
  P.N = N; P.H = H; P.W = W; P.Ci = Ci; P.Co = Co; P.Cir = Cir; P.Cor = Cor; P.k = k; P.p = p; P.Hq = Hq; P.Wq = Wq;
  P.tiles_x = W / 8; P.tiles_y = (H + 15) / 16; P.tiles_total = N * P.tiles_x * P.tiles_y;
  // N block: largest multiple of 16 dividing Co, <= 128, whose double-buffered stage (x patch + gy tile) fits shared memory
  const size_t patch_b = (size_t)16 * Hp * Wpx * 16;
  int NB = Co > 128 ? 128 : Co;
  while (NB >= 16 && (Co % NB || 2 * (patch_b + (size_t)(NB / 8) * 2048) > 216 * 1024)) NB -= 16;
  if (NB < 16) return CG_ERR_UNSUPPORTED;
  P.NB = NB; P.ncob = Co / NB;
  P.TG = 512 / NB; if (P.TG > kk) P.TG = kk;
  P.ntg = (kk + P.TG - 1) / P.TG;
  P.cim = Ci == 64 ? 64 : 128; P.ncib = Ci == 64 ? 1 : Ci / 128;
  P.patch_bytes = (uint32_t)(16 * Hp * Wpx * 16); P.patch_load_bytes = (uint32_t)((P.cim / 8) * Hp * Wpx * 16);
  P.g_bytes = (uint32_t)(NB / 8) * 2048;
  size_t smem = 2 * ((size_t)P.patch_bytes + P.g_bytes);
  if (smem > 216 * 1024) return CG_ERR_UNSUPPORTED;
  int base = P.ncib * P.ncob * P.ntg;
  // pixel splits: fill the machine -- divided by CATGEN_WG_DIV (experiments): the weight gradient runs on a side stream beside the input-gradient
  // convolution and, in D, beside three other branches; a grid that leaves SMs to them can shorten the step although the kernel alone gets slower
  static const int wg_div = getenv("CATGEN_WG_DIV") ? atoi(getenv("CATGEN_WG_DIV")) : 1;
  int Z = (ctx().sm_count / (wg_div > 0 ? wg_div : 1) + base - 1) / base; if (Z > P.tiles_total / 2) Z = P.tiles_total / 2; if (Z < 1) Z = 1;
  P.Z = Z;
  size_t xq_bytes = (size_t)N * (Ci / 8) * Hq * Wq * 16;
  size_t part_bytes = (size_t)Z * kk * Cir * Cor * sizeof(float);
  size_t o1 = xq_prepacked ? 0 : ((xq_bytes + 255) & ~(size_t)255);
  uint8_t* ws = (uint8_t*)workspace3(o1 + part_bytes + 512);
  if (!ws) return CG_ERR_CUDA;
  const uint8_t* xq = xq_prepacked ? xq_prepacked : ws; float* part = (float*)(ws + o1);
  long nx = (long)(xq_bytes / 16);
  if (!xq_prepacked) CG_LAUNCH(k_pack_act<2>, grid1d(nx, 256), 256, 0, x, ws, nx, H, W, Cir, Ci, p, Hq, Wq, (const float*)nullptr);
  P.xq = xq; P.gq = g.gq; P.part = part; P.scale2 = g.scale2;
  static bool attr_done = false;
  if (!attr_done) { CG_CUDA(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr_done = true; }
  dim3 grid(Z, base);
  ctx().next_flops = 2.0 * (double)N * H * W * Cor * kk * Cir;   // algorithmic (unpadded) work
  ctx().next_bytes = (double)xq_bytes + (double)N * (g.Cg / 8) * Hq * Wq * 16 + 4.0 * (double)kk * Cir * Cor;
  CUtensorMap tmx, tmg;
  CG_TRY(make_patch_tmap(&tmx, xq, 2, N, Ci / 8, Hq, Wq, Hp, Wpx, P.cim / 8));
  CG_TRY(make_tile_tmap(&tmg, g.gq, N, g.Cg / 8, Hq, Wq, NB / 8));
  CG_LAUNCH(k_wgrad_tc, grid, 256, smem, P, tmx, tmg);
  long nW = (long)kk * Cir * Cor;
  // Measured (gpurun_out r02_c): summing the Z partials inside k_parts_to_torch_acc (even with four load chains per thread) costs
  // 2.1 ms per step against 0.69 ms for the fully parallel element-wise sum followed by the layout change: keep the two kernels.
  CG_LAUNCH(k_sum_parts, grid1d(nW, 256, 2), 256, 0, part, Z, nW, gWp_out);          // fully parallel, fixed z order
  if (gW_acc && parts_to_torch_acc(gWp_out, 1, 0, gW_acc, Cir, Cor, kk) == CG_OK) { if (done) *done = 1; }   // layout change only
  return CG_OK;
}

int conv_wgrad_tc(const float* x, const float* gy, float* gWp_out, int N, int H, int W, int Cir, int Cor, int k, float* gW_acc, int* done) {
  if (!wgrad_shape_ok(H, W, Cir, k)) return CG_ERR_UNSUPPORTED;
  if ((((uintptr_t)x | (uintptr_t)gy | (uintptr_t)gWp_out) & 15) != 0) return CG_ERR_UNSUPPORTED;
  GradOperand g; CG_TRY(pack_grad_operand(gy, N, H, W, Cor, k, &g));
  return conv_wgrad_tc_impl(x, g, gWp_out, N, H, W, Cir, Cor, k, gW_acc, done);
}

// Whole backward of one conv layer: weight gradient AND input gradient from ONE packed gradient operand.
//   gWp_out[(tap,ci)][co] (overwritten) ; gx[N,H,W,Ci] = conv(gy, Wd)
// Backward of one conv layer from a gradient operand that is ALREADY packed (fuse_d.cu: act_bwd_pack wrote gq and its scale): weight
// gradient chain on the side stream straight into the Torch-layout gradient, input gradient on the issuing stream.  x / xq: the layer's
// input (fp32, packed here) or its cached operand.  gx may be null (no input gradient wanted); gW_acc null skips the weight gradient.
int conv_bwd_tc_gq(const float* x, const uint8_t* xq_prepacked, const uint8_t* gq, const float* scale2, const float* Wd, float* gx,
                   int N, int H, int W, int Ci, int Co, int k, float* gW_acc) {
  if (!wgrad_shape_ok(H, W, Ci, k) || !tc_shape_ok(H, W, Co, Ci, k, 2) || conv_v1()) return CG_ERR_UNSUPPORTED;
  GradOperand g; g.gq = gq; g.scale2 = scale2; g.Cg = ((Co + 63) / 64) * 64;
  if (gW_acc) {
    const int side = side_begin();
    int wdone = 0;
    float* gWp = (float*)workspace(sizeof(float) * (size_t)k * k * Ci * Co + 256);
    int wst = gWp ? conv_wgrad_tc_impl(x, g, gWp, N, H, W, Ci, Co, k, gW_acc, &wdone, xq_prepacked) : CG_ERR_CUDA;
    if (wst == CG_OK && !wdone) wst = set_err(CG_ERR_STATE, "weight gradient was not accumulated");
    if (side) { int est = side_end(); if (wst == CG_OK) wst = est; }
    CG_TRY(wst);
  }
  if (!gx) { ctx().next_amax = nullptr; return CG_OK; }
  return conv_tc_run<2>(nullptr, Wd, nullptr, gx, N, H, W, Co, Ci, k, scale2, gq);
}

int conv_bwd_tc(const float* x, const float* gy, const float* Wd, float* gWp_out, float* gx, int N, int H, int W, int Ci, int Co, int k, float* gW_acc, int* done,
                const uint8_t* xq_prepacked, float* gb_acc, int* bias_done) {
  static const bool dgrad_tf32 = getenv("CATGEN_DGRAD_TF32") != nullptr;
  if (dgrad_tf32 || !wgrad_shape_ok(H, W, Ci, k) || !tc_shape_ok(H, W, Co, Ci, k, 2)) return CG_ERR_UNSUPPORTED;
  if ((((uintptr_t)x | (uintptr_t)gy | (uintptr_t)gWp_out | (uintptr_t)gx) & 15) != 0) return CG_ERR_UNSUPPORTED;
  if (xq_prepacked && !(Ci == 64 || Ci % 128 == 0)) return CG_ERR_UNSUPPORTED;
  GradOperand g; CG_TRY(pack_grad_operand(gy, N, H, W, Co, k, &g, gb_acc));
  if (gb_acc && bias_done) *bias_done = 1;
  // weight-gradient chain (MMA kernel, split sum, layout change) on the side stream: nothing reads it before the model's
  // backward ends, while the input gradient below is on the critical path
  // (only when the result goes straight into the Torch-layout gradient: the packed form then lives in the side stream's scratch)
  const int side = gW_acc ? side_begin() : 0;
  int wst = CG_OK, wdone = 0;
  if (side) {
    float* gWp_side = (float*)workspace(sizeof(float) * (size_t)k * k * Ci * Co + 256);
    wst = gWp_side ? conv_wgrad_tc_impl(x, g, gWp_side, N, H, W, Ci, Co, k, gW_acc, &wdone, xq_prepacked) : CG_ERR_CUDA;
    if (wst == CG_OK && !wdone) wst = set_err(CG_ERR_STATE, "side-stream weight gradient was not accumulated");
    if (done) *done = wdone;
    int est = side_end(); if (wst == CG_OK) wst = est;
  } else wst = conv_wgrad_tc_impl(x, g, gWp_out, N, H, W, Ci, Co, k, gW_acc, done, xq_prepacked);
  CG_TRY(wst);
  // dgrad = forward convolution of gy (Co channels in) with the flipped weights (Ci channels out)
  return conv_tc_run<2>(gy, Wd, nullptr, gx, N, H, W, Co, Ci, k, g.scale2, g.gq);
}

}  // namespace cg
