// stn_fused.cu -- the spatial transformer of models.lua:814-906 in FOUR launches per pass instead of ~16 forward / ~28 backward.
//
//   localisation network (models.lua:843-860):  AvgPool2 -> conv3x3(ch->16) -> LeakyReLU -> conv3x3(16->16) -> LeakyReLU -> AvgPool2 ->
//       View -> Linear(16*(S/4)^2 -> 64) -> LeakyReLU -> Linear(64 -> n_theta) -> AffineTransformMatrixGenerator
//   sampler (models.lua:868-903):               AffineGridGeneratorBHWD(S,S) -> BilinearSamplerBHWD
//
// Round 1 ran every one of these modules as its own kernel (the 3- and 16-channel convolutions zero-padded to 64 channels on the tensor
// cores): D32_st3 holds four transformers and they accounted for ~350 of a step's ~740 launches while doing < 1 % of its FLOPs -- the
// discriminator's critical path was a chain of ~5 us launches.  Here ONE CTA PER IMAGE runs the whole localisation network out of
// shared memory in fp32 on the CUDA cores (1.5 MFLOP per image: nothing to gain from tensor cores, and theta -- which decides WHERE the
// sampler reads -- keeps fp32 accuracy), the grid is never materialised (the sampler computes its coordinates from the 2x3 matrix), and
// the backward mirrors it: sampler backward (scatter-add, fp32 atomics as before) -> one CTA per image for every input/activation
// gradient and the per-image parameter-gradient partials -> one fixed-order reduction over the batch into the flat Torch-layout gradient.
// Arithmetic per module follows ops.cu's single-module kernels (same formulas, fp32); the op-level C-ABI entry points still use those.
#include "model.cuh"

namespace cg {

__device__ __forceinline__ float lrelu_f(float v) { return v >= 0.f ? v : 0.333f * v; }        // LeakyReLU.lua:13-19
__device__ __forceinline__ float lrelu_g(float x, float g) { return x >= 0.f ? g : 0.333f * g; }   // :21-31 (gradient 1 at x == 0)

struct StnArgs {
  // parameters, Torch layout (flat parameter vector)
  const float *W1, *b1, *W2, *b2, *L1, *lb1, *L2, *lb2;
  const float *W1p, *W1d, *W2p, *W2d;   // packed fp32 operands of the two convolutions (conv_ref.cu layouts): Wp[(tap,ci)][co], Wd[(8-tap,co)][ci]
  int B, ch, S, rot, scl, trn, nth;
  const float* in;                     // [B,S,S,ch]
  // saved by the forward (global)
  float *pool1, *c1o, *c2o, *pool2, *l1o, *theta, *A;
  int dbg;                             // experiments (CATGEN_STN_DBG=1): block 0 records clock64 at every phase boundary
};
__device__ long long g_stn_dbg[32];
#define STN_T(k) do { if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) g_stn_dbg[k] = clock64(); } while (0)

// ---- localisation-network convolutions out of shared memory, round 2b.
// The first fused version read one activation and one weight word from shared memory per FMA (2 LDS / FMA): with 16 warps per SM the
// kernels were bound by shared-memory instruction issue and latency (CATGEN_STN_DBG=1, profiles/r02_stn_phases.txt: conv1 46k of 92k cycles
// forward; the per-image weight gradient 112k / 222k of 241k / 421k cycles backward), and three transformer branches sharing an SM slowed each
// other 3x.  Now: activations live in HALO'D buffers (no border branches), the channel index is contiguous and padded so that every operand
// is fetched with 16-byte loads, each thread keeps several pixels (forward / input gradient) or 4 output channels x 9 taps (weight gradient)
// in registers.  Shared-memory traffic per FMA drops 4-8x; the arithmetic is the same fp32 sums in a different (still fixed) order.
//
// shared-memory plan (floats).  CS1 = roundup(ch,4) + 4 and 20 = 16 + 4 are the padded channel strides (16-byte aligned, bank-staggered);
// HP = (P+2)^2 halo'd pixels.
struct StnSmem { int xh, a2h, g2h, cs, w1, w2, pool2, v64, misc, total, CS1, chp, HP; };
__host__ __device__ inline StnSmem stn_smem(int ch, int S, bool bwd) {
  const int P = S / 2, f = 16 * (S / 4) * (S / 4);
  StnSmem m; m.chp = (ch + 3) & ~3; m.CS1 = m.chp + 4; m.HP = (P + 2) * (P + 2);
  int o = 0;
  m.xh = o; o += m.HP * m.CS1;                       // pooled input, halo'd   (backward: afterwards the plain [P*P][ch] input gradient)
  m.a2h = o; o += m.HP * 20;                         // LeakyReLU(conv1), halo'd (backward: afterwards gc1, the gradient w.r.t. conv1's output)
  m.g2h = o; o += bwd ? m.HP * 20 : 0;               // backward: gradient w.r.t. conv2's output, halo'd
  m.cs = o; o += bwd ? 0 : P * P * 16;               // forward: conv2's output, plain
  m.w1 = o; o += bwd ? 9 * m.chp * 20 : 9 * 16 * m.CS1;   // forward [tap][co][CS1] (ci contiguous); backward [tap][ci][20] (co contiguous)
  m.w2 = o; o += 9 * 16 * 20;
  m.pool2 = o; o += f; m.v64 = o; o += 64; m.misc = o; o += 64; m.total = o;
  return m;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) { return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, acc)))); }
__device__ __forceinline__ void smem_zero(float* sm, int n) {   // n multiple of 4
  float4* q = reinterpret_cast<float4*>(sm);
  for (int i = threadIdx.x; i < n / 4; i += blockDim.x) q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
// Index arithmetic: the divisors (channels, P, S) are run-time values and a 32-bit division costs ~40 instructions -- the element-wise
// loops of the first version spent most of their cycles dividing (profiles/r02_stn_phases.txt, phases "w+pool1", "gin").  Powers of two shift.
struct Dv { int d, sh; };
__device__ __forceinline__ Dv mk_dv(int d) { Dv v; v.d = d; v.sh = (d & (d - 1)) == 0 ? 31 - __clz(d) : -1; return v; }
__device__ __forceinline__ int dv_div(int i, const Dv& v) { return v.sh >= 0 ? i >> v.sh : i / v.d; }
__device__ __forceinline__ int dv_mod(int i, const Dv& v) { return v.sh >= 0 ? i & (v.d - 1) : i % v.d; }
// forward layout [tap][co][CS] (ci contiguous) from the packed input-gradient operand Wd[(8-tap, co)][ci] (taps flipped): row copies
__device__ __forceinline__ void load_w_fwd(const float* __restrict__ Wd, float* __restrict__ w_s, int Ci, int CS) {
  const int n = 9 * 16 * Ci; const Dv dc = mk_dv(Ci);
  for (int i0 = threadIdx.x; i0 < n; i0 += 8 * blockDim.x) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; v[u] = i < n ? Wd[i] : 0.f; }
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; if (i < n) { const int row = dv_div(i, dc), ci = dv_mod(i, dc); w_s[((8 - (row >> 4)) * 16 + (row & 15)) * CS + ci] = v[u]; } }
  }
}
// input-gradient layout [tap][ci (nci rows)][20] (co contiguous) from the packed forward operand Wp[(tap, ci)][co]: row copies
__device__ __forceinline__ void load_w_bwd(const float* __restrict__ Wp, float* __restrict__ w_s, int Ci, int nci) {
  const int n = 9 * Ci * 16; const Dv dc = mk_dv(Ci);
  for (int i0 = threadIdx.x; i0 < n; i0 += 8 * blockDim.x) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; v[u] = i < n ? Wp[i] : 0.f; }
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; if (i < n) { const int r = i >> 4, tap = dv_div(r, dc), ci = dv_mod(r, dc); w_s[(tap * nci + ci) * 20 + (i & 15)] = v[u]; } }
  }
}
// out[p][co] = bias[co] + sum_{tap, ci} w[tap][co][ci] * xh[p + tap][ci]   (xh halo'd: pixel (y,x) sits at (y+1,x+1)).  512 threads: co = tid & 15,
// pixel group = tid >> 4; PX pixels per thread (P*P == 32*PX).  Per (tap, 4 channels): 1 + PX 16-byte loads for 4*PX FMAs.
template <int PX, typename Emit>
__device__ __forceinline__ void conv_fwd_v(const float* __restrict__ xh, const float* __restrict__ wf, const float* __restrict__ bias, int P, int CS, int C4, Emit emit) {
  const int co = threadIdx.x & 15, pg = threadIdx.x >> 4, PW = P + 2;
  float acc[PX]; int base[PX];
  const float bv = bias[co];
  const int lP = 31 - __clz(P);                                           // P is 8 or 16 (stn_fused_shape_ok)
#pragma unroll
  for (int i = 0; i < PX; ++i) { const int p = pg + 32 * i, y = p >> lP, x = p - y * P; base[i] = (y * PW + x) * CS; acc[i] = bv; }
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int off = ((tap / 3) * PW + (tap % 3)) * CS;
    const float4* wr = reinterpret_cast<const float4*>(wf + (tap * 16 + co) * CS);
#pragma unroll 4
    for (int c4 = 0; c4 < C4; ++c4) {
      const float4 w = wr[c4];
#pragma unroll
      for (int i = 0; i < PX; ++i) acc[i] = dot4(w, *reinterpret_cast<const float4*>(xh + base[i] + off + 4 * c4), acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < PX; ++i) emit(pg + 32 * i, co, acc[i]);
}

__device__ __forceinline__ void atm_eval(const float* th, int rot, int scl, int trn, float* M6);   // ops.cu formulas, restated below

__global__ void __launch_bounds__(512) k_stn_loc_fwd(StnArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, tid = threadIdx.x, ch = a.ch, S = a.S, P = S / 2, Q = S / 4, f = 16 * Q * Q, PW = P + 2;
  const StnSmem m = stn_smem(ch, S, false);
  float *xh = sm + m.xh, *a2h = sm + m.a2h, *cs = sm + m.cs, *w1 = sm + m.w1, *w2 = sm + m.w2, *pool2 = sm + m.pool2, *v64 = sm + m.v64;
  STN_T(0);
  smem_zero(sm, m.w2);                                                    // halos, channel padding and weight padding (everything below w2 is fully written)
  __syncthreads();
  load_w_fwd(a.W1d, w1, ch, m.CS1); load_w_fwd(a.W2d, w2, 16, 20);
  const float* in = a.in + (size_t)b * S * S * ch;
  float* p1g = a.pool1 + (size_t)b * P * P * ch;
  const Dv dch = mk_dv(ch), dP = mk_dv(P);
  for (int i0 = tid; i0 < P * P * ch; i0 += 4 * blockDim.x) {             // nn.SpatialAveragePooling(2,2,2,2); sixteen loads in flight per thread
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x; v[u] = 0.f;
      if (i < P * P * ch) { const int c = dv_mod(i, dch), p = dv_div(i, dch), y = dv_div(p, dP), x = p - y * P; const float* s0 = in + ((size_t)(2 * y) * S + 2 * x) * ch + c;
        v[u] = (s0[0] + s0[ch] + s0[(size_t)S * ch] + s0[(size_t)S * ch + ch]) * 0.25f; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < P * P * ch) { const int c = dv_mod(i, dch), p = dv_div(i, dch), y = dv_div(p, dP), x = p - y * P; xh[((y + 1) * PW + x + 1) * m.CS1 + c] = v[u]; p1g[i] = v[u]; }
    }
  }
  __syncthreads();
  STN_T(1);
  {
    float* c1g = a.c1o + (size_t)b * P * P * 16;
    auto emit = [&](int p, int co, float v) { const int y = dv_div(p, dP), x = p - y * P; c1g[p * 16 + co] = v; a2h[((y + 1) * PW + x + 1) * 20 + co] = lrelu_f(v); };
    if (P == 8) conv_fwd_v<2>(xh, w1, a.b1, P, m.CS1, m.chp / 4, emit); else conv_fwd_v<8>(xh, w1, a.b1, P, m.CS1, m.chp / 4, emit);
  }
  __syncthreads();
  STN_T(2);
  STN_T(3);
  {
    float* c2g = a.c2o + (size_t)b * P * P * 16;
    auto emit = [&](int p, int co, float v) { c2g[p * 16 + co] = v; cs[p * 16 + co] = v; };
    if (P == 8) conv_fwd_v<2>(a2h, w2, a.b2, P, 20, 4, emit); else conv_fwd_v<8>(a2h, w2, a.b2, P, 20, 4, emit);
  }
  __syncthreads();
  STN_T(4);
  for (int i = tid; i < f; i += blockDim.x) {                              // LeakyReLU then AvgPool2 -> [Q][Q][16]
    int c = i & 15, p = i >> 4, y = p / Q, x = p - y * Q;
    const float* s0 = cs + ((2 * y) * P + 2 * x) * 16 + c;
    float v = (lrelu_f(s0[0]) + lrelu_f(s0[16]) + lrelu_f(s0[P * 16]) + lrelu_f(s0[P * 16 + 16])) * 0.25f;
    pool2[i] = v; a.pool2[(size_t)b * f + i] = v;
  }
  __syncthreads();
  STN_T(5);
  {   // nn.View + nn.Linear(f, 64): Torch feature index ft = c*Q*Q + s for our [s][c].  One warp per four output rows, lanes over consecutive
      // ft (coalesced weight reads), four loads in flight, fixed shuffle tree
    const int warp = tid >> 5, lane = tid & 31, QQ = Q * Q, nw = (int)(blockDim.x >> 5);
    for (int o0 = warp * 4; o0 < 64; o0 += nw * 4) {
      const float* Wr = a.L1 + (size_t)o0 * f;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
      for (int ft = lane; ft < f; ft += 32) {
        const int c = ft / QQ, s2 = ft - c * QQ; const float pv = pool2[s2 * 16 + c];
        const float w0 = Wr[ft], w1v = Wr[(size_t)f + ft], w2v = Wr[(size_t)2 * f + ft], w3 = Wr[(size_t)3 * f + ft];
        acc[0] += pv * w0; acc[1] += pv * w1v; acc[2] += pv * w2v; acc[3] += pv * w3;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float sum = warp_sum(acc[r]);
        if (lane == 0) { const int o = o0 + r; float v = sum + a.lb1[o]; a.l1o[(size_t)b * 64 + o] = v; v64[o] = lrelu_f(v); }
      }
    }
  }
  __syncthreads();
  STN_T(6);
  if (tid < a.nth) {                                                       // nn.Linear(64, n_theta)
    float acc = a.lb2[tid];
    for (int i = 0; i < 64; ++i) acc += v64[i] * a.L2[tid * 64 + i];
    a.theta[(size_t)b * 4 + tid] = acc; sm[m.misc + tid] = acc;
  }
  __syncthreads();
  if (tid == 0) { float M6[6]; atm_eval(sm + m.misc, a.rot, a.scl, a.trn, M6); for (int i = 0; i < 6; ++i) a.A[(size_t)b * 6 + i] = M6[i]; }
  STN_T(7);
}

// nn.AffineTransformMatrixGenerator: I * R(alpha) * S(s) * T(tx,ty), first two rows; R = [[c,-s],[s,c]] (ops.cu k_atm_fwd)
__device__ __forceinline__ void m3(const float* x, const float* y, float* z) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) z[i * 3 + j] = x[i * 3] * y[j] + x[i * 3 + 1] * y[3 + j] + x[i * 3 + 2] * y[6 + j];
}
__device__ __forceinline__ void atm_parts(const float* th, int rot, int scl, int trn, float* R, float* Sm, float* T, int* idx) {
  int p = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) { R[i] = Sm[i] = T[i] = (i % 4 == 0) ? 1.f : 0.f; }
  idx[0] = idx[1] = idx[2] = -1;
  if (rot) { float al = th[p]; idx[0] = p++; float c = cosf(al), s = sinf(al); R[0] = c; R[1] = -s; R[3] = s; R[4] = c; }
  if (scl) { float s = th[p]; idx[1] = p++; Sm[0] = s; Sm[4] = s; }
  if (trn) { idx[2] = p; T[2] = th[p]; T[5] = th[p + 1]; }
}
__device__ __forceinline__ void atm_eval(const float* th, int rot, int scl, int trn, float* M6) {
  float R[9], Sm[9], T[9], RS[9], M[9]; int idx[3];
  atm_parts(th, rot, scl, trn, R, Sm, T, idx); m3(R, Sm, RS); m3(RS, T, M);
#pragma unroll
  for (int i = 0; i < 6; ++i) M6[i] = M[i];
}
__device__ __forceinline__ void atm_grad(const float* th, const float* G, int rot, int scl, int trn, float* gt) {   // ops.cu k_atm_bwd
  float R[9], Sm[9], T[9], tmp[9], M[9]; int idx[3];
  atm_parts(th, rot, scl, trn, R, Sm, T, idx);
  if (rot) {
    float al = th[idx[0]]; float c = cosf(al), s = sinf(al);
    float dR[9] = {-s, -c, 0, c, -s, 0, 0, 0, 0};
    m3(dR, Sm, tmp); m3(tmp, T, M);
    float v = 0; for (int i = 0; i < 6; ++i) v += G[i] * M[i];
    gt[idx[0]] = v;
  }
  if (scl) {
    float dS[9] = {1, 0, 0, 0, 1, 0, 0, 0, 0};
    m3(R, dS, tmp); m3(tmp, T, M);
    float v = 0; for (int i = 0; i < 6; ++i) v += G[i] * M[i];
    gt[idx[1]] = v;
  }
  if (trn) {
    float RS[9]; m3(R, Sm, RS);
    for (int q = 0; q < 2; ++q) {
      float dT[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; dT[q == 0 ? 2 : 5] = 1;
      m3(RS, dT, M);
      float v = 0; for (int i = 0; i < 6; ++i) v += G[i] * M[i];
      gt[idx[2] + q] = v;
    }
  }
}

// ---------------------------------------------------------------- sampler: grid coordinates from A on the fly (ops.cu k_grid_fwd + k_bil_fwd)
__device__ __forceinline__ void grid_at(const float* A6, int i, int j, int H, int W, float& gy, float& gx) {
  float yb = -1.f + 2.f * i / (H - 1), xb = -1.f + 2.f * j / (W - 1);
  gy = A6[0] * yb + A6[1] * xb + A6[2]; gx = A6[3] * yb + A6[4] * xb + A6[5];
}
// LPP lanes share a pixel and stride over its channels: 32 for the 64-channel branch transformers, 4 for the 1- / 3-channel input transformer
// (with a whole warp per pixel 29 of 32 lanes idled there: 45 us forward, 69 us backward for 1.5 MB of image, profiles/r02_timeline.txt)
template <int LPP>
__global__ void k_stn_sample_fwd(const float* __restrict__ img, const float* __restrict__ A, float* __restrict__ out, long npix, int H, int W, int C) {
  long pix = (blockIdx.x * (long)blockDim.x + threadIdx.x) / LPP; int lane = threadIdx.x % LPP;
  if (pix >= npix) return;
  long b = pix / ((long)H * W); int r = (int)(pix - b * H * W), ii = r / W, jj = r - ii * W;
  float gy_, gx_; grid_at(A + b * 6, ii, jj, H, W, gy_, gx_);
  float yc = (gy_ + 1.f) * (H - 1) / 2.f, xc = (gx_ + 1.f) * (W - 1) / 2.f;
  float fy = floorf(yc), fx = floorf(xc); int y0 = (int)fy, x0 = (int)fx;
  float wy = 1.f - (yc - fy), wx = 1.f - (xc - fx);
  bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H, vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
  const float* base = img + b * H * W * C;
  for (int c = lane; c < C; c += LPP) {
    float a00 = (vy0 && vx0) ? base[((long)y0 * W + x0) * C + c] : 0.f;
    float a01 = (vy0 && vx1) ? base[((long)y0 * W + x0 + 1) * C + c] : 0.f;
    float a10 = (vy1 && vx0) ? base[((long)(y0 + 1) * W + x0) * C + c] : 0.f;
    float a11 = (vy1 && vx1) ? base[((long)(y0 + 1) * W + x0 + 1) * C + c] : 0.f;
    out[pix * C + c] = wx * wy * a00 + (1.f - wx) * wy * a01 + wx * (1.f - wy) * a10 + (1.f - wx) * (1.f - wy) * a11;
  }
}
// backward of the sampler: scatter-add into gimg (fp32 atomics, as ops.cu k_bil_bwd and for the reason given there) and the
// per-pixel gradient w.r.t. the grid point, which the localisation backward reduces per image in fixed order
template <int LPP>
__global__ void k_stn_sample_bwd(const float* __restrict__ img, const float* __restrict__ A, const float* __restrict__ gout,
                                 float* __restrict__ gimg, float* __restrict__ ggrid, long npix, int H, int W, int C) {
  long pix = (blockIdx.x * (long)blockDim.x + threadIdx.x) / LPP; int lane = threadIdx.x % LPP;
  const bool live = pix < npix;                                            // no early return: every lane takes part in the shuffles below
  if (!live) pix = npix - 1;
  long b = pix / ((long)H * W); int r = (int)(pix - b * H * W), ii = r / W, jj = r - ii * W;
  float gy_, gx_; grid_at(A + b * 6, ii, jj, H, W, gy_, gx_);
  float yc = (gy_ + 1.f) * (H - 1) / 2.f, xc = (gx_ + 1.f) * (W - 1) / 2.f;
  float fy = floorf(yc), fx = floorf(xc); int y0 = (int)fy, x0 = (int)fx;
  float wy = 1.f - (yc - fy), wx = 1.f - (xc - fx);
  bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H, vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
  const float* base = img + b * H * W * C; float* gb = gimg + b * H * W * C;
  float d00 = 0, d01 = 0, d10 = 0, d11 = 0;
  if (live) for (int c = lane; c < C; c += LPP) {
    float gv = gout[pix * C + c];
    if (vy0 && vx0) { long o = ((long)y0 * W + x0) * C + c; atomicAdd(gb + o, wx * wy * gv); d00 += base[o] * gv; }
    if (vy0 && vx1) { long o = ((long)y0 * W + x0 + 1) * C + c; atomicAdd(gb + o, (1.f - wx) * wy * gv); d01 += base[o] * gv; }
    if (vy1 && vx0) { long o = ((long)(y0 + 1) * W + x0) * C + c; atomicAdd(gb + o, wx * (1.f - wy) * gv); d10 += base[o] * gv; }
    if (vy1 && vx1) { long o = ((long)(y0 + 1) * W + x0 + 1) * C + c; atomicAdd(gb + o, (1.f - wx) * (1.f - wy) * gv); d11 += base[o] * gv; }
  }
#pragma unroll
  for (int o = LPP / 2; o > 0; o >>= 1) {                                   // fixed tree over the LPP lanes of this pixel
    d00 += __shfl_xor_sync(0xffffffffu, d00, o); d01 += __shfl_xor_sync(0xffffffffu, d01, o); d10 += __shfl_xor_sync(0xffffffffu, d10, o); d11 += __shfl_xor_sync(0xffffffffu, d11, o);
  }
  if (live && lane == 0) {
    float gyf = -wx * d00 + wx * d10 - (1.f - wx) * d01 + (1.f - wx) * d11;
    float gxf = -wy * d00 + wy * d01 - (1.f - wy) * d10 + (1.f - wy) * d11;
    ggrid[pix * 2] = gyf * (H - 1) / 2.f;
    ggrid[pix * 2 + 1] = gxf * (W - 1) / 2.f;
  }
}

// ---------------------------------------------------------------- localisation backward, one CTA per image
struct StnBwdArgs {
  StnArgs f;
  const float* ggrid;                  // [B,S,S,2] from k_stn_sample_bwd
  float* gin;                          // [B,S,S,ch]: the sampler's input gradient; the localisation branch's is ADDED (nn.ConcatTable)
  float* gl1;                          // [B,64]  gradient w.r.t. Linear1's output (for the cross-batch weight gradient)
  float* part;                         // per-image parameter-gradient partials, Torch layout: [B][np_part] or null (parameter gradients skipped)
  unsigned int* amax_out;              // optional: max|gin| after both branches were summed (float bits)
  int np_part;                         // = 16*ch*9 + 16 + 16*16*9 + 16 + 64 + nth*64 + nth   (W1,b1,W2,b2,lb1,L2,lb2; L1's weight goes through gl1)
};

__device__ __forceinline__ float block_sum_f(float v, float* scratch) {   // fixed order; result valid in every thread
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += scratch[i];
  return r;
}
// gx[p][ci] = sum_{tap, co} wd[tap][ci][co] * gyh[p + (2-ky, 2-kx)][co]  (gyh halo'd).  Threads: ci = tid % nci, pixel group = tid / nci;
// PX pixels per thread (PX * 512 / nci == P*P).  Per (tap, 4 co): 1 + PX 16-byte loads for 4*PX FMAs.
template <int PX, typename Emit>
__device__ __forceinline__ void conv_dgrad_v(const float* __restrict__ gyh, const float* __restrict__ wd, int P, int nci, Emit emit) {
  const int ci = threadIdx.x % nci, g = threadIdx.x / nci, ng = (int)blockDim.x / nci, PW = P + 2;
  float acc[PX]; int base[PX];
#pragma unroll
  for (int i = 0; i < PX; ++i) { const int p = g + ng * i, y = p >> (31 - __clz(P)), x = p - y * P; base[i] = (y * PW + x) * 20; acc[i] = 0.f; }
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {                                     // (fully unrolled, ptxas hoisted 288 16-byte loads and spilled)
    const int off = ((2 - tap / 3) * PW + (2 - tap % 3)) * 20;
    const float4* wr = reinterpret_cast<const float4*>(wd + (tap * nci + ci) * 20);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 w = wr[q];
#pragma unroll
      for (int i = 0; i < PX; ++i) acc[i] = dot4(w, *reinterpret_cast<const float4*>(gyh + base[i] + off + 4 * q), acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < PX; ++i) emit(g + ng * i, ci, acc[i]);
}
// per-image weight-gradient partial in TORCH layout: gW[co][ci][tap] = sum_p gy[p][co] * x[p + tap - 1][ci].  A task = (ci, 2 output channels) with the
// nine taps of each in registers; SL adjacent lanes split the pixels of a task and are summed by a fixed shuffle tree.
// Per pixel: one 8-byte load of gy + 9 loads of x for 18 FMAs.
template <int SL>
__device__ __forceinline__ void conv_wgrad_v(const float* __restrict__ xh, int CSx, const float* __restrict__ gyh, float* __restrict__ gW, int P, int Ci) {
  // (two output channels per task: four -- 36 accumulators -- spilled under the 128-register cap of a 512-thread block)
  const int t = threadIdx.x / SL, sl = threadIdx.x % SL, PW = P + 2;
  const bool active = t < Ci * 8;                                           // whole SL-groups are active or idle; idle lanes still take part in the shuffles
  const int ci = active ? t % Ci : 0, cop = active ? t / Ci : 0;            // cop: pair of output channels
  float acc[2][9];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[j][q] = 0.f;
  if (active) for (int p = sl; p < P * P; p += SL) {
    const int y = p >> (31 - __clz(P)), x = p - y * P;
    const float2 g = *reinterpret_cast<const float2*>(gyh + ((y + 1) * PW + x + 1) * 20 + cop * 2);
    const float* xb = xh + (y * PW + x) * CSx + ci;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      const float xv = xb[((q / 3) * PW + q % 3) * CSx];
      acc[0][q] = fmaf(g.x, xv, acc[0][q]); acc[1][q] = fmaf(g.y, xv, acc[1][q]);
    }
  }
#pragma unroll
  for (int o = SL / 2; o > 0; o >>= 1)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int q = 0; q < 9; ++q) acc[j][q] += __shfl_xor_sync(0xffffffffu, acc[j][q], o);
  if (active && sl == 0)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int q = 0; q < 9; ++q) gW[((size_t)(cop * 2 + j) * Ci + ci) * 9 + q] = acc[j][q];
}
// gb[co] = sum_p gy[p][co]: warp co, lanes over pixels, fixed shuffle tree
__device__ __forceinline__ void bias_grad_v(const float* __restrict__ gyh, float* __restrict__ gb, int P) {
  const int co = threadIdx.x >> 5, lane = threadIdx.x & 31, PW = P + 2;
  if (co < 16) {
    float v = 0.f;
    for (int p = lane; p < P * P; p += 32) { const int y = p >> (31 - __clz(P)), x = p - y * P; v += gyh[((y + 1) * PW + x + 1) * 20 + co]; }
    v = warp_sum(v);
    if (lane == 0) gb[co] = v;
  }
}

__global__ void __launch_bounds__(512) k_stn_loc_bwd(StnBwdArgs q) {
  extern __shared__ __align__(16) float sm[];
  const StnArgs& a = q.f;
  const int b = blockIdx.x, tid = threadIdx.x, ch = a.ch, S = a.S, P = S / 2, Q = S / 4, f = 16 * Q * Q, nth = a.nth, PW = P + 2;
  const StnSmem m = stn_smem(ch, S, true);
  float *xh = sm + m.xh, *a2h = sm + m.a2h, *g2h = sm + m.g2h, *w1 = sm + m.w1, *w2 = sm + m.w2, *pool2 = sm + m.pool2, *v64 = sm + m.v64, *misc = sm + m.misc;
  const Dv dch = mk_dv(ch), dP = mk_dv(P), dS = mk_dv(S);
  STN_T(0);
  smem_zero(sm, m.pool2);
  __syncthreads();
  load_w_bwd(a.W1p, w1, ch, m.chp); load_w_bwd(a.W2p, w2, 16, 16);
  float* part = q.part ? q.part + (size_t)b * q.np_part : nullptr;
  const int oW1 = 0, ob1 = oW1 + 16 * ch * 9, oW2 = ob1 + 16, ob2 = oW2 + 16 * 16 * 9, olb1 = ob2 + 16, oL2 = olb1 + 64, olb2 = oL2 + nth * 64;
  // ---- AffineGridGeneratorBHWD backward: gA = sum_pixels ggrid^T * (y_i, x_j, 1), fixed-order block reduction (ops.cu k_grid_bwd)
  {
    float s6[6] = {0, 0, 0, 0, 0, 0};
    const float* gg = q.ggrid + (size_t)b * S * S * 2;
    for (int i = tid; i < S * S; i += blockDim.x) {
      int ii = i >> (31 - __clz(S)), j = i - ii * S;                          // S is 16 or 32
      float yb = -1.f + 2.f * ii / (S - 1), xb = -1.f + 2.f * j / (S - 1);
      float g0 = gg[i * 2], g1 = gg[i * 2 + 1];
      s6[0] += g0 * yb; s6[1] += g0 * xb; s6[2] += g0; s6[3] += g1 * yb; s6[4] += g1 * xb; s6[5] += g1;
    }
    for (int k = 0; k < 6; ++k) { float r = block_sum_f(s6[k], misc + 16); if (tid == 0) misc[k] = r; }
    __syncthreads();
    if (tid == 0) {                                                        // AffineTransformMatrixGenerator backward -> gtheta in misc[8..]
      float gt[4] = {0, 0, 0, 0};
      atm_grad(a.theta + (size_t)b * 4, misc, a.rot, a.scl, a.trn, gt);
      for (int k = 0; k < nth; ++k) misc[8 + k] = gt[k];
    }
    __syncthreads();
  }
  STN_T(1);
  // ---- Linear(64, nth) backward: v64 <- lrelu(l1o); gal1 = L2^T gtheta; partials gL2 = gtheta (x) al1, glb2 = gtheta
  const float* l1o = a.l1o + (size_t)b * 64;
  if (tid < 64) {
    const float al = lrelu_f(l1o[tid]);
    float g = 0.f;
    for (int k = 0; k < nth; ++k) { g += a.L2[k * 64 + tid] * misc[8 + k]; if (part) part[oL2 + k * 64 + tid] = misc[8 + k] * al; }
    g = lrelu_g(l1o[tid], g);                                              // LeakyReLU backward: gradient w.r.t. Linear1's output
    v64[tid] = g; q.gl1[(size_t)b * 64 + tid] = g;
    if (part) part[olb1 + tid] = g;
  }
  if (part && tid < nth) part[olb2 + tid] = misc[8 + tid];
  __syncthreads();
  STN_T(2);
  // ---- Linear(f, 64) backward (input gradient): gpool2[s][c] = sum_o L1[o][c*Q*Q + s] * gl1[o]
  {
    // consecutive threads read consecutive weights of a row; the 64 rows are split over G = blockDim / f thread groups when f < blockDim
    // (f = 256: every thread has 32 rows, 16 loads in flight), partial sums combined in group order through the (still unused) g2h region
    const int G = f < (int)blockDim.x ? (int)blockDim.x / f : 1, rows = 64 / G;
    float* scr = g2h;                                                       // G * f floats <= 1024 (HP * 20 >= 2000)
    for (int e = tid; e < f * G; e += blockDim.x) {
      const int ft = e % f, gq = e / f, o0 = gq * rows;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
      for (int o = o0; o < o0 + rows; o += 4) {
        a0 += a.L1[(size_t)o * f + ft] * v64[o]; a1 += a.L1[(size_t)(o + 1) * f + ft] * v64[o + 1];
        a2 += a.L1[(size_t)(o + 2) * f + ft] * v64[o + 2]; a3 += a.L1[(size_t)(o + 3) * f + ft] * v64[o + 3];
      }
      scr[e] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    for (int ft = tid; ft < f; ft += blockDim.x) {
      const int c = ft / (Q * Q), s2 = ft - c * Q * Q;
      float v = 0.f; for (int gq = 0; gq < G; ++gq) v += scr[gq * f + ft];
      pool2[s2 * 16 + c] = v;
    }
    __syncthreads();
    for (int e = tid; e < f * G; e += blockDim.x) scr[e] = 0.f;              // g2h must be all zero again (its halo is never written)
  }
  __syncthreads();
  STN_T(3);
  // ---- AvgPool2 backward + LeakyReLU backward: gc2 (g2h) = lrelu'(c2o) * gpool2 / 4;  a1 = lrelu(c1o) (a2h) = conv2's input; pooled input (xh)
  const float* c2o = a.c2o + (size_t)b * P * P * 16; const float* c1o = a.c1o + (size_t)b * P * P * 16;
  for (int i0 = tid; i0 < P * P * 16; i0 += 4 * blockDim.x) {
    float u2[4], u1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * blockDim.x; const bool ok = i < P * P * 16; u2[u] = ok ? c2o[i] : 0.f; u1[u] = ok ? c1o[i] : 0.f; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < P * P * 16) {
        const int c = i & 15, p = i >> 4, y = dv_div(p, dP), x = p - y * P, h = ((y + 1) * PW + x + 1) * 20 + c;
        g2h[h] = lrelu_g(u2[u], pool2[((y >> 1) * Q + (x >> 1)) * 16 + c] * 0.25f);
        a2h[h] = lrelu_f(u1[u]);
      }
    }
  }
  const float* p1g = a.pool1 + (size_t)b * P * P * ch;
  for (int i0 = tid; i0 < P * P * ch; i0 += 4 * blockDim.x) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * blockDim.x; v[u] = i < P * P * ch ? p1g[i] : 0.f; }
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * blockDim.x; if (i < P * P * ch) { const int c = dv_mod(i, dch), p = dv_div(i, dch), y = dv_div(p, dP), x = p - y * P; xh[((y + 1) * PW + x + 1) * m.CS1 + c] = v[u]; } }
  }
  __syncthreads();
  STN_T(4);
  if (part) { conv_wgrad_v<4>(a2h, 20, g2h, part + oW2, P, 16); bias_grad_v(g2h, part + ob2, P); }
  __syncthreads();
  STN_T(5);
  // ---- conv2 input gradient, LeakyReLU backward on the way: gc1 -> a2h (a1 is no longer needed once conv2's weight gradient is done)
  {
    auto emit = [&](int p, int ci, float v) { const int y = dv_div(p, dP), x = p - y * P; a2h[((y + 1) * PW + x + 1) * 20 + ci] = lrelu_g(c1o[p * 16 + ci], v); };
    if (P == 8) conv_dgrad_v<2>(g2h, w2, P, 16, emit); else conv_dgrad_v<8>(g2h, w2, P, 16, emit);
  }
  __syncthreads();
  STN_T(6);
  STN_T(7);
  if (part) {
    if (ch >= 64) conv_wgrad_v<1>(xh, m.CS1, a2h, part + oW1, P, ch); else conv_wgrad_v<16>(xh, m.CS1, a2h, part + oW1, P, ch);   // ch <= 4: 8*ch tasks x 16 lanes
    bias_grad_v(a2h, part + ob1, P);
  }
  __syncthreads();
  STN_T(8);
  // ---- conv1 input gradient (plain [p][ch], into the xh region: the pooled input was consumed by the weight gradient above)
  {
    float* gp1 = xh;
    auto emit = [&](int p, int ci, float v) { if (ci < ch) gp1[p * ch + ci] = v; };
    if (m.chp == 64) conv_dgrad_v<8>(a2h, w1, P, 64, emit);             // P == 8
    else conv_dgrad_v<2>(a2h, w1, P, 4, emit);                          // P == 16, ch <= 4
  }
  __syncthreads();
  STN_T(9);
  // ---- AvgPool2 backward ADDED into the sampler's input gradient (the two ConcatTable branches sum)
  float* gin = q.gin + (size_t)b * S * S * ch;
  float amx = 0.f;
  for (int i0 = tid; i0 < S * S * ch; i0 += 8 * blockDim.x) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; v[u] = i < S * S * ch ? gin[i] : 0.f; }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < S * S * ch) {
        const int c = dv_mod(i, dch), p = dv_div(i, dch), y = dv_div(p, dS), x = p - y * S;
        const float r = v[u] + xh[((y >> 1) * P + (x >> 1)) * ch + c] * 0.25f;
        gin[i] = r; amx = fmaxf(amx, fabsf(r));
      }
    }
  }
  if (q.amax_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, o));
    if ((tid & 31) == 0 && amx > 0.f) atomicMax(q.amax_out, __float_as_uint(amx));
  }
  STN_T(10);
}

// fixed-order sum of the per-image partials over the batch + Linear1's weight gradient (sum over the batch of gl1 (x) pool2), accumulated
// into the flat Torch-layout gradient.  Grid-stride over [np_part + 64*f] elements.
struct StnRedArgs {
  const float *part, *gl1, *pool2; int B, np_part, f, Q, ch, nth;
  float *gW1, *gb1, *gW2, *gb2, *gL1, *glb1, *gL2, *glb2;
};
__global__ void k_stn_param_reduce(StnRedArgs r) {
  const long total = (long)r.np_part + 64L * r.f;
  const int oW1 = 0, ob1 = oW1 + 16 * r.ch * 9, oW2 = ob1 + 16, ob2 = oW2 + 16 * 16 * 9, olb1 = ob2 + 16, oL2 = olb1 + 64, olb2 = oL2 + r.nth * 64;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    if (e < r.np_part) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f; int b = 0;                 // fixed association, four loads in flight
      for (; b + 4 <= r.B; b += 4) {
        a0 += r.part[(size_t)b * r.np_part + e]; a1 += r.part[(size_t)(b + 1) * r.np_part + e];
        a2 += r.part[(size_t)(b + 2) * r.np_part + e]; a3 += r.part[(size_t)(b + 3) * r.np_part + e];
      }
      for (; b < r.B; ++b) a0 += r.part[(size_t)b * r.np_part + e];
      acc = (a0 + a1) + (a2 + a3);
      float* dst; int o = (int)e;
      if (o < ob1) dst = r.gW1 + o; else if (o < oW2) dst = r.gb1 + (o - ob1); else if (o < ob2) dst = r.gW2 + (o - oW2); else if (o < olb1) dst = r.gb2 + (o - ob2);
      else if (o < oL2) dst = r.glb1 + (o - olb1); else if (o < olb2) dst = r.gL2 + (o - oL2); else dst = r.glb2 + (o - olb2);
      *dst += acc;
    } else {
      // threads enumerate (o, our feature index mine = s*16 + c): the batch loop then reads pool2 coalesced; the (few) writes scatter to
      // the Torch feature index ft = c*Q*Q + s
      const long w = e - r.np_part; const int o = (int)(w / r.f), mine = (int)(w - (long)o * r.f);
      const int c = mine & 15, s = mine >> 4, ft = c * r.Q * r.Q + s;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f; int b = 0;
      for (; b + 4 <= r.B; b += 4) {
        a0 += r.gl1[(size_t)b * 64 + o] * r.pool2[(size_t)b * r.f + mine]; a1 += r.gl1[(size_t)(b + 1) * 64 + o] * r.pool2[(size_t)(b + 1) * r.f + mine];
        a2 += r.gl1[(size_t)(b + 2) * 64 + o] * r.pool2[(size_t)(b + 2) * r.f + mine]; a3 += r.gl1[(size_t)(b + 3) * 64 + o] * r.pool2[(size_t)(b + 3) * r.f + mine];
      }
      for (; b < r.B; ++b) a0 += r.gl1[(size_t)b * 64 + o] * r.pool2[(size_t)b * r.f + mine];
      r.gL1[(size_t)o * r.f + ft] += (a0 + a1) + (a2 + a3);
    }
  }
}

// ---------------------------------------------------------------- host side
static size_t stn_smem_bytes(int ch, int S, bool bwd) { return sizeof(float) * (size_t)stn_smem(ch, S, bwd).total; }
bool stn_fused_shape_ok(int ch, int S) { return (S == 16 && ch == 64) || (S == 32 && ch >= 1 && ch <= 4); }   // the register tilings above are written for these
static bool stn_dbg_on() { static const bool on = getenv("CATGEN_STN_DBG") != nullptr; return on; }
static void stn_dbg_print(const char* what, int ch, int S, int B, int n) {
  cudaStreamSynchronize(ctx().stream);
  long long h[32]; cudaMemcpyFromSymbol(h, g_stn_dbg, sizeof(h));
  fprintf(stderr, "[stn dbg] %s ch=%d S=%d B=%d total %lld cycles; phases:", what, ch, S, B, h[n] - h[0]);
  for (int i = 1; i <= n; ++i) fprintf(stderr, " %lld", h[i] - h[i - 1]);
  fprintf(stderr, "\n");
}
static int stn_set_attr() {
  static bool done = false;
  if (done) return CG_OK;
  CG_CUDA(cudaFuncSetAttribute(k_stn_loc_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
  CG_CUDA(cudaFuncSetAttribute(k_stn_loc_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
  done = true; return CG_OK;
}
int stn_fused_forward(const StnFusedParams& p, const float* in, int B, float* pool1, float* c1o, float* c2o, float* pool2, float* l1o, float* theta, float* A, float* out) {
  if (!stn_fused_shape_ok(p.ch, p.S)) return set_err(CG_ERR_UNSUPPORTED, "fused spatial transformer: ch=%d S=%d", p.ch, p.S);
  CG_TRY(stn_set_attr());
  StnArgs a{};
  a.W1 = p.W1; a.b1 = p.b1; a.W2 = p.W2; a.b2 = p.b2; a.L1 = p.L1; a.lb1 = p.lb1; a.L2 = p.L2; a.lb2 = p.lb2; a.W1p = p.W1p; a.W1d = p.W1d; a.W2p = p.W2p; a.W2d = p.W2d;
  a.B = B; a.ch = p.ch; a.S = p.S; a.rot = p.rot; a.scl = p.scl; a.trn = p.trn; a.nth = p.nth; a.in = in;
  a.pool1 = pool1; a.c1o = c1o; a.c2o = c2o; a.pool2 = pool2; a.l1o = l1o; a.theta = theta; a.A = A;
  a.dbg = stn_dbg_on() ? 1 : 0;
  CG_LAUNCH(k_stn_loc_fwd, B, 512, stn_smem_bytes(p.ch, p.S, false), a);
  if (a.dbg) stn_dbg_print("loc_fwd [w+pool1, conv1, lrelu, conv2, pool2, linear1, linear2+atm]", p.ch, p.S, B, 7);
  long npix = (long)B * p.S * p.S;
  ctx().next_bytes = 8.0 * (double)npix * p.ch;
  if (p.ch <= 4) CG_LAUNCH(k_stn_sample_fwd<4>, cdiv(npix * 4, 256), 256, 0, in, (const float*)A, out, npix, p.S, p.S, p.ch);
  else CG_LAUNCH(k_stn_sample_fwd<32>, cdiv(npix * 32, 256), 256, 0, in, (const float*)A, out, npix, p.S, p.S, p.ch);
  return CG_OK;
}
// gout [B,S,S,ch] -> gin [B,S,S,ch] (written); parameter gradients accumulated into g* unless skip_param_grads
int stn_fused_backward(const StnFusedParams& p, const StnFusedGrads& g, const float* in, int B, const float* pool1, const float* c1o, const float* c2o, const float* pool2,
                       const float* l1o, const float* theta, const float* A, const float* gout, float* gin, float* ggrid, float* gl1, float* part, int skip_param_grads, unsigned int* amax_out) {
  CG_TRY(stn_set_attr());
  long npix = (long)B * p.S * p.S;
  CG_CUDA(cudaMemsetAsync(gin, 0, sizeof(float) * (size_t)npix * p.ch, ctx().stream));
  ctx().next_bytes = 12.0 * (double)npix * p.ch;
  if (p.ch <= 4) CG_LAUNCH(k_stn_sample_bwd<4>, cdiv(npix * 4, 256), 256, 0, in, A, gout, gin, ggrid, npix, p.S, p.S, p.ch);
  else CG_LAUNCH(k_stn_sample_bwd<32>, cdiv(npix * 32, 256), 256, 0, in, A, gout, gin, ggrid, npix, p.S, p.S, p.ch);
  StnBwdArgs q{};
  StnArgs& a = q.f;
  a.W1 = p.W1; a.b1 = p.b1; a.W2 = p.W2; a.b2 = p.b2; a.L1 = p.L1; a.lb1 = p.lb1; a.L2 = p.L2; a.lb2 = p.lb2; a.W1p = p.W1p; a.W1d = p.W1d; a.W2p = p.W2p; a.W2d = p.W2d;
  a.B = B; a.ch = p.ch; a.S = p.S; a.rot = p.rot; a.scl = p.scl; a.trn = p.trn; a.nth = p.nth; a.in = in;
  a.pool1 = const_cast<float*>(pool1); a.c1o = const_cast<float*>(c1o); a.c2o = const_cast<float*>(c2o); a.pool2 = const_cast<float*>(pool2);
  a.l1o = const_cast<float*>(l1o); a.theta = const_cast<float*>(theta); a.A = const_cast<float*>(A);
  q.ggrid = ggrid; q.gin = gin; q.gl1 = gl1; q.np_part = stn_fused_part_floats(p.ch, p.nth); q.part = skip_param_grads ? nullptr : part; q.amax_out = amax_out;
  a.dbg = stn_dbg_on() ? 1 : 0;
  CG_LAUNCH(k_stn_loc_bwd, B, 512, stn_smem_bytes(p.ch, p.S, true), q);
  if (a.dbg) stn_dbg_print(skip_param_grads ? "loc_bwd(no wgrad) [w+grid, l2, l1 dgrad, pool/lrelu, wgrad2, dgrad2, gc1, wgrad1, dgrad1, gin]" : "loc_bwd [w+grid, l2, l1 dgrad, pool/lrelu, wgrad2, dgrad2, gc1, wgrad1, dgrad1, gin]", p.ch, p.S, B, 10);
  if (!skip_param_grads) {
    StnRedArgs r{};
    const int Q = p.S / 4;
    r.part = part; r.gl1 = gl1; r.pool2 = pool2; r.B = B; r.np_part = q.np_part; r.f = 16 * Q * Q; r.Q = Q; r.ch = p.ch; r.nth = p.nth;
    r.gW1 = g.W1; r.gb1 = g.b1; r.gW2 = g.W2; r.gb2 = g.b2; r.gL1 = g.L1; r.glb1 = g.lb1; r.gL2 = g.L2; r.glb2 = g.lb2;
    long total = (long)r.np_part + 64L * r.f;
    CG_LAUNCH(k_stn_param_reduce, grid1d(total, 128), 128, 0, r);
  }
  return CG_OK;
}

}  // namespace cg
