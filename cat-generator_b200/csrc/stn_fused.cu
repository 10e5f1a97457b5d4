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
  int B, ch, S, rot, scl, trn, nth;
  const float* in;                     // [B,S,S,ch]
  // saved by the forward (global)
  float *pool1, *c1o, *c2o, *pool2, *l1o, *theta, *A;
};

// 3x3 "same" convolution of a [P x P x Ci] tile in shared memory with weights in shared memory laid out [(tap, ci)][17] (16 + 1 pad:
// conflict-free both with co across a warp's threads, here, and with ci across them, in the input gradient); thread ->
// (pixel group, co).  out = bias + sum; written to out_s [P*P][16] (and out_g when non-null).
__device__ __forceinline__ void conv3x3_16(const float* __restrict__ x_s, const float* __restrict__ w_s, const float* __restrict__ bias, float* __restrict__ out_s,
                                            float* __restrict__ out_g, int P, int Ci) {
  const int co = threadIdx.x & 15;
  for (int p = threadIdx.x >> 4; p < P * P; p += blockDim.x >> 4) {
    const int y = p / P, x = p - y * P;
    float acc = bias[co];
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1; if (yy < 0 || yy >= P) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1; if (xx < 0 || xx >= P) continue;
        const float* xr = x_s + (yy * P + xx) * Ci;
        const float* wr = w_s + ((ky * 3 + kx) * Ci) * 17 + co;
        // four independent chains (a single one ran at one FMA per shared-memory latency: the kernels were latency-bound at 8 warps per SM)
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f; int ci = 0;
        for (; ci + 4 <= Ci; ci += 4) { a0 += xr[ci] * wr[ci * 17]; a1 += xr[ci + 1] * wr[(ci + 1) * 17]; a2 += xr[ci + 2] * wr[(ci + 2) * 17]; a3 += xr[ci + 3] * wr[(ci + 3) * 17]; }
        for (; ci < Ci; ++ci) a0 += xr[ci] * wr[ci * 17];
        acc += (a0 + a1) + (a2 + a3);
      }
    }
    out_s[p * 16 + co] = acc;
    if (out_g) out_g[p * 16 + co] = acc;
  }
}

// shared-memory plan (floats): pool1 [P*P*ch] | a [P*P*16] | c [P*P*16] | w1 [9*ch*16] | w2 [9*16*16] | pool2 [f] | v64 [64] | misc [32]
struct StnSmem { int pool1, a, c, w1, w2, pool2, v64, misc, total; };
__host__ __device__ inline StnSmem stn_smem(int ch, int S) {
  const int P = S / 2, f = 16 * (S / 4) * (S / 4);
  StnSmem m; int o = 0;
  m.pool1 = o; o += P * P * ch; m.a = o; o += P * P * 16; m.c = o; o += P * P * 16;
  m.w1 = o; o += 9 * ch * 17; m.w2 = o; o += 9 * 16 * 17; m.pool2 = o; o += f; m.v64 = o; o += 64; m.misc = o; o += 32; m.total = o;
  return m;
}
__device__ __forceinline__ void load_conv_w(const float* __restrict__ W, float* __restrict__ w_s, int Ci) {   // Torch [16][Ci][3][3] -> [(tap,ci)][17]
  for (int i = threadIdx.x; i < 16 * Ci * 9; i += blockDim.x) { int tap = i % 9, r = i / 9, ci = r % Ci, co = r / Ci; w_s[(tap * Ci + ci) * 17 + co] = W[i]; }
}

__device__ __forceinline__ void atm_eval(const float* th, int rot, int scl, int trn, float* M6);   // ops.cu formulas, restated below

__global__ void __launch_bounds__(512) k_stn_loc_fwd(StnArgs a) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, tid = threadIdx.x, ch = a.ch, S = a.S, P = S / 2, Q = S / 4, f = 16 * Q * Q;
  const StnSmem m = stn_smem(ch, S);
  float *pool1 = sm + m.pool1, *as = sm + m.a, *cs = sm + m.c, *w1 = sm + m.w1, *w2 = sm + m.w2, *pool2 = sm + m.pool2, *v64 = sm + m.v64;
  load_conv_w(a.W1, w1, ch); load_conv_w(a.W2, w2, 16);
  const float* in = a.in + (size_t)b * S * S * ch;
  for (int i = tid; i < P * P * ch; i += blockDim.x) {                     // nn.SpatialAveragePooling(2,2,2,2)
    int c = i % ch, p = i / ch, y = p / P, x = p - y * P;
    const float* s0 = in + ((size_t)(2 * y) * S + 2 * x) * ch + c;
    float v = (s0[0] + s0[ch] + s0[(size_t)S * ch] + s0[(size_t)S * ch + ch]) * 0.25f;
    pool1[i] = v; a.pool1[(size_t)b * P * P * ch + i] = v;
  }
  __syncthreads();
  conv3x3_16(pool1, w1, a.b1, cs, a.c1o + (size_t)b * P * P * 16, P, ch);
  __syncthreads();
  for (int i = tid; i < P * P * 16; i += blockDim.x) as[i] = lrelu_f(cs[i]);
  __syncthreads();
  conv3x3_16(as, w2, a.b2, cs, a.c2o + (size_t)b * P * P * 16, P, 16);
  __syncthreads();
  for (int i = tid; i < f; i += blockDim.x) {                              // LeakyReLU then AvgPool2 -> [Q][Q][16]
    int c = i & 15, p = i >> 4, y = p / Q, x = p - y * Q;
    const float* s0 = cs + ((2 * y) * P + 2 * x) * 16 + c;
    float v = (lrelu_f(s0[0]) + lrelu_f(s0[16]) + lrelu_f(s0[P * 16]) + lrelu_f(s0[P * 16 + 16])) * 0.25f;
    pool2[i] = v; a.pool2[(size_t)b * f + i] = v;
  }
  __syncthreads();
  {   // nn.View + nn.Linear(f, 64): Torch feature index ft = c*Q*Q + s for our [s][c].  One warp per output row, lanes over consecutive
      // ft (coalesced weight reads), fixed shuffle tree
    // (a single dependent chain of L2 loads per warp made this loop 60 of the kernel's 85 us: four rows at a time, four loads each in flight)
    const int warp = tid >> 5, lane = tid & 31, QQ = Q * Q, nw = (int)(blockDim.x >> 5);
    for (int o0 = warp * 4; o0 < 64; o0 += nw * 4) {
      const float* Wr = a.L1 + (size_t)o0 * f;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int ft = lane; ft < f; ft += 32) {
        const int c = ft / QQ, s2 = ft - c * QQ; const float pv = pool2[s2 * 16 + c];
        const float w0 = Wr[ft], w1 = Wr[(size_t)f + ft], w2 = Wr[(size_t)2 * f + ft], w3 = Wr[(size_t)3 * f + ft];
        acc[0] += pv * w0; acc[1] += pv * w1; acc[2] += pv * w2; acc[3] += pv * w3;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float sum = warp_sum(acc[r]);
        if (lane == 0) { const int o = o0 + r; float v = sum + a.lb1[o]; a.l1o[(size_t)b * 64 + o] = v; v64[o] = lrelu_f(v); }
      }
    }
  }
  __syncthreads();
  if (tid < a.nth) {                                                       // nn.Linear(64, n_theta)
    float acc = a.lb2[tid];
    for (int i = 0; i < 64; ++i) acc += v64[i] * a.L2[tid * 64 + i];
    a.theta[(size_t)b * 4 + tid] = acc; sm[m.misc + tid] = acc;
  }
  __syncthreads();
  if (tid == 0) { float M6[6]; atm_eval(sm + m.misc, a.rot, a.scl, a.trn, M6); for (int i = 0; i < 6; ++i) a.A[(size_t)b * 6 + i] = M6[i]; }
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
__global__ void k_stn_sample_fwd(const float* __restrict__ img, const float* __restrict__ A, float* __restrict__ out, long npix, int H, int W, int C) {
  long pix = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
  if (pix >= npix) return;
  long b = pix / ((long)H * W); int r = (int)(pix - b * H * W), ii = r / W, jj = r - ii * W;
  float gy_, gx_; grid_at(A + b * 6, ii, jj, H, W, gy_, gx_);
  float yc = (gy_ + 1.f) * (H - 1) / 2.f, xc = (gx_ + 1.f) * (W - 1) / 2.f;
  float fy = floorf(yc), fx = floorf(xc); int y0 = (int)fy, x0 = (int)fx;
  float wy = 1.f - (yc - fy), wx = 1.f - (xc - fx);
  bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H, vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
  const float* base = img + b * H * W * C;
  for (int c = lane; c < C; c += 32) {
    float a00 = (vy0 && vx0) ? base[((long)y0 * W + x0) * C + c] : 0.f;
    float a01 = (vy0 && vx1) ? base[((long)y0 * W + x0 + 1) * C + c] : 0.f;
    float a10 = (vy1 && vx0) ? base[((long)(y0 + 1) * W + x0) * C + c] : 0.f;
    float a11 = (vy1 && vx1) ? base[((long)(y0 + 1) * W + x0 + 1) * C + c] : 0.f;
    out[pix * C + c] = wx * wy * a00 + (1.f - wx) * wy * a01 + wx * (1.f - wy) * a10 + (1.f - wx) * (1.f - wy) * a11;
  }
}
// backward of the sampler: scatter-add into gimg (fp32 atomics, as ops.cu k_bil_bwd and for the reason given there) and the
// per-pixel gradient w.r.t. the grid point, which the localisation backward reduces per image in fixed order
__global__ void k_stn_sample_bwd(const float* __restrict__ img, const float* __restrict__ A, const float* __restrict__ gout,
                                 float* __restrict__ gimg, float* __restrict__ ggrid, long npix, int H, int W, int C) {
  long pix = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
  if (pix >= npix) return;
  long b = pix / ((long)H * W); int r = (int)(pix - b * H * W), ii = r / W, jj = r - ii * W;
  float gy_, gx_; grid_at(A + b * 6, ii, jj, H, W, gy_, gx_);
  float yc = (gy_ + 1.f) * (H - 1) / 2.f, xc = (gx_ + 1.f) * (W - 1) / 2.f;
  float fy = floorf(yc), fx = floorf(xc); int y0 = (int)fy, x0 = (int)fx;
  float wy = 1.f - (yc - fy), wx = 1.f - (xc - fx);
  bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H, vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
  const float* base = img + b * H * W * C; float* gb = gimg + b * H * W * C;
  float d00 = 0, d01 = 0, d10 = 0, d11 = 0;
  for (int c = lane; c < C; c += 32) {
    float gv = gout[pix * C + c];
    if (vy0 && vx0) { long o = ((long)y0 * W + x0) * C + c; atomicAdd(gb + o, wx * wy * gv); d00 += base[o] * gv; }
    if (vy0 && vx1) { long o = ((long)y0 * W + x0 + 1) * C + c; atomicAdd(gb + o, (1.f - wx) * wy * gv); d01 += base[o] * gv; }
    if (vy1 && vx0) { long o = ((long)(y0 + 1) * W + x0) * C + c; atomicAdd(gb + o, wx * (1.f - wy) * gv); d10 += base[o] * gv; }
    if (vy1 && vx1) { long o = ((long)(y0 + 1) * W + x0 + 1) * C + c; atomicAdd(gb + o, (1.f - wx) * (1.f - wy) * gv); d11 += base[o] * gv; }
  }
  d00 = warp_sum(d00); d01 = warp_sum(d01); d10 = warp_sum(d10); d11 = warp_sum(d11);
  if (lane == 0) {
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
// input gradient of a 3x3 "same" conv: gx[p][ci] = sum_{tap,co} gy[p - tap + 1][co] * W[co][ci][tap]; weights in smem as [(tap,ci)][16]
__device__ __forceinline__ void conv3x3_dgrad(const float* __restrict__ gy_s, const float* __restrict__ w_s, float* __restrict__ gx_s, int P, int Ci) {
  for (int i = threadIdx.x; i < P * P * Ci; i += blockDim.x) {
    const int ci = i % Ci, p = i / Ci, y = p / P, x = p - y * P;
    float acc = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y - ky + 1; if (yy < 0 || yy >= P) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x - kx + 1; if (xx < 0 || xx >= P) continue;
        const float* g = gy_s + (yy * P + xx) * 16;
        const float* wr = w_s + ((ky * 3 + kx) * Ci + ci) * 17;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int co = 0; co < 16; co += 4) { a0 += g[co] * wr[co]; a1 += g[co + 1] * wr[co + 1]; a2 += g[co + 2] * wr[co + 2]; a3 += g[co + 3] * wr[co + 3]; }
        acc += (a0 + a1) + (a2 + a3);
      }
    }
    gx_s[i] = acc;
  }
}
// per-image weight-gradient partial of a 3x3 conv in TORCH layout: gW[co][ci][ky][kx] = sum_p gy[p][co] * x[p + tap - 1][ci]; gb[co] = sum_p gy[p][co].
// thread <-> (co, ci) pair with all nine taps in registers: a warp's threads share co (broadcast read of gy) and read consecutive ci of x.
__device__ __forceinline__ void conv3x3_wgrad(const float* __restrict__ x_s, const float* __restrict__ gy_s, float* __restrict__ gW, float* __restrict__ gb, int P, int Ci) {
  for (int e = threadIdx.x; e < 16 * Ci; e += blockDim.x) {
    const int ci = e % Ci, co = e / Ci;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    for (int y = 0; y < P; ++y)
      for (int x = 0; x < P; ++x) {
        const float g = gy_s[(y * P + x) * 16 + co];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = y + ky - 1; if (yy < 0 || yy >= P) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int xx = x + kx - 1; if (xx < 0 || xx >= P) continue;
            acc[ky * 3 + kx] += g * x_s[(yy * P + xx) * Ci + ci];
          }
        }
      }
#pragma unroll
    for (int t = 0; t < 9; ++t) gW[(size_t)e * 9 + t] = acc[t];
  }
  for (int co = threadIdx.x; co < 16; co += blockDim.x) { float acc = 0.f; for (int p = 0; p < P * P; ++p) acc += gy_s[p * 16 + co]; gb[co] = acc; }
}

__global__ void __launch_bounds__(512) k_stn_loc_bwd(StnBwdArgs q) {
  extern __shared__ float sm[];
  const StnArgs& a = q.f;
  const int b = blockIdx.x, tid = threadIdx.x, ch = a.ch, S = a.S, P = S / 2, Q = S / 4, f = 16 * Q * Q, nth = a.nth;
  const StnSmem m = stn_smem(ch, S);
  float *pool1 = sm + m.pool1, *as = sm + m.a, *cs = sm + m.c, *w1 = sm + m.w1, *w2 = sm + m.w2, *pool2 = sm + m.pool2, *v64 = sm + m.v64, *misc = sm + m.misc;
  load_conv_w(a.W1, w1, ch); load_conv_w(a.W2, w2, 16);
  float* part = q.part ? q.part + (size_t)b * q.np_part : nullptr;
  const int oW1 = 0, ob1 = oW1 + 16 * ch * 9, oW2 = ob1 + 16, ob2 = oW2 + 16 * 16 * 9, olb1 = ob2 + 16, oL2 = olb1 + 64, olb2 = oL2 + nth * 64;
  // ---- AffineGridGeneratorBHWD backward: gA = sum_pixels ggrid^T * (y_i, x_j, 1), fixed-order block reduction (ops.cu k_grid_bwd)
  {
    float s6[6] = {0, 0, 0, 0, 0, 0};
    const float* gg = q.ggrid + (size_t)b * S * S * 2;
    for (int i = tid; i < S * S; i += blockDim.x) {
      int ii = i / S, j = i - ii * S;
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
  // ---- Linear(f, 64) backward (input gradient): gpool2[s][c] = sum_o L1[o][c*Q*Q + s] * gl1[o]
  for (int ft = tid; ft < f; ft += blockDim.x) {                           // consecutive threads read consecutive weights of each row
    const int c = ft / (Q * Q), s2 = ft - c * Q * Q;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;                            // four independent chains: four L2 loads in flight per thread
#pragma unroll 4
    for (int o = 0; o < 64; o += 4) {
      a0 += a.L1[(size_t)o * f + ft] * v64[o]; a1 += a.L1[(size_t)(o + 1) * f + ft] * v64[o + 1];
      a2 += a.L1[(size_t)(o + 2) * f + ft] * v64[o + 2]; a3 += a.L1[(size_t)(o + 3) * f + ft] * v64[o + 3];
    }
    pool2[s2 * 16 + c] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  // ---- AvgPool2 backward + LeakyReLU backward: gc2 (in cs) = lrelu'(c2o) * gpool2 / 4
  const float* c2o = a.c2o + (size_t)b * P * P * 16; const float* c1o = a.c1o + (size_t)b * P * P * 16;
  for (int i = tid; i < P * P * 16; i += blockDim.x) {
    const int c = i & 15, p = i >> 4, y = p / P, x = p - y * P;
    cs[i] = lrelu_g(c2o[i], pool2[((y >> 1) * Q + (x >> 1)) * 16 + c] * 0.25f);
    as[i] = lrelu_f(c1o[i]);                                               // a1 = conv2's input
  }
  __syncthreads();
  if (part) conv3x3_wgrad(as, cs, part + oW2, part + ob2, P, 16);
  __syncthreads();
  // ---- conv2 input gradient ga1 -> `as` (a1 is no longer needed once conv2's weight gradient is done)
  conv3x3_dgrad(cs, w2, as, P, 16);
  __syncthreads();
  for (int i = tid; i < P * P * 16; i += blockDim.x) cs[i] = lrelu_g(c1o[i], as[i]);   // gc1
  const float* p1g = a.pool1 + (size_t)b * P * P * ch;
  for (int i = tid; i < P * P * ch; i += blockDim.x) pool1[i] = p1g[i];
  __syncthreads();
  if (part) conv3x3_wgrad(pool1, cs, part + oW1, part + ob1, P, ch);
  __syncthreads();
  // ---- conv1 input gradient, then AvgPool2 backward ADDED into the sampler's input gradient (the two ConcatTable branches sum)
  conv3x3_dgrad(cs, w1, pool1, P, ch);                                      // pool1 <- gpool1 (its forward values were consumed above)
  __syncthreads();
  float* gin = q.gin + (size_t)b * S * S * ch;
  float amx = 0.f;
  for (int i = tid; i < S * S * ch; i += blockDim.x) {
    const int c = i % ch, p = i / ch, y = p / S, x = p - y * S;
    const float v = gin[i] + pool1[((y >> 1) * P + (x >> 1)) * ch + c] * 0.25f;
    gin[i] = v; amx = fmaxf(amx, fabsf(v));
  }
  if (q.amax_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, o));
    if ((tid & 31) == 0 && amx > 0.f) atomicMax(q.amax_out, __float_as_uint(amx));
  }
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
static size_t stn_smem_bytes(int ch, int S) { return sizeof(float) * (size_t)stn_smem(ch, S).total; }
static int stn_set_attr() {
  static bool done = false;
  if (done) return CG_OK;
  CG_CUDA(cudaFuncSetAttribute(k_stn_loc_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CG_CUDA(cudaFuncSetAttribute(k_stn_loc_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  done = true; return CG_OK;
}
int stn_fused_forward(const StnFusedParams& p, const float* in, int B, float* pool1, float* c1o, float* c2o, float* pool2, float* l1o, float* theta, float* A, float* out) {
  CG_TRY(stn_set_attr());
  StnArgs a{};
  a.W1 = p.W1; a.b1 = p.b1; a.W2 = p.W2; a.b2 = p.b2; a.L1 = p.L1; a.lb1 = p.lb1; a.L2 = p.L2; a.lb2 = p.lb2;
  a.B = B; a.ch = p.ch; a.S = p.S; a.rot = p.rot; a.scl = p.scl; a.trn = p.trn; a.nth = p.nth; a.in = in;
  a.pool1 = pool1; a.c1o = c1o; a.c2o = c2o; a.pool2 = pool2; a.l1o = l1o; a.theta = theta; a.A = A;
  CG_LAUNCH(k_stn_loc_fwd, B, 512, stn_smem_bytes(p.ch, p.S), a);
  long npix = (long)B * p.S * p.S;
  ctx().next_bytes = 8.0 * (double)npix * p.ch;
  CG_LAUNCH(k_stn_sample_fwd, cdiv(npix * 32, 256), 256, 0, in, (const float*)A, out, npix, p.S, p.S, p.ch);
  return CG_OK;
}
// gout [B,S,S,ch] -> gin [B,S,S,ch] (written); parameter gradients accumulated into g* unless skip_param_grads
int stn_fused_backward(const StnFusedParams& p, const StnFusedGrads& g, const float* in, int B, const float* pool1, const float* c1o, const float* c2o, const float* pool2,
                       const float* l1o, const float* theta, const float* A, const float* gout, float* gin, float* ggrid, float* gl1, float* part, int skip_param_grads, unsigned int* amax_out) {
  CG_TRY(stn_set_attr());
  long npix = (long)B * p.S * p.S;
  CG_CUDA(cudaMemsetAsync(gin, 0, sizeof(float) * (size_t)npix * p.ch, ctx().stream));
  ctx().next_bytes = 12.0 * (double)npix * p.ch;
  CG_LAUNCH(k_stn_sample_bwd, cdiv(npix * 32, 256), 256, 0, in, A, gout, gin, ggrid, npix, p.S, p.S, p.ch);
  StnBwdArgs q{};
  StnArgs& a = q.f;
  a.W1 = p.W1; a.b1 = p.b1; a.W2 = p.W2; a.b2 = p.b2; a.L1 = p.L1; a.lb1 = p.lb1; a.L2 = p.L2; a.lb2 = p.lb2;
  a.B = B; a.ch = p.ch; a.S = p.S; a.rot = p.rot; a.scl = p.scl; a.trn = p.trn; a.nth = p.nth; a.in = in;
  a.pool1 = const_cast<float*>(pool1); a.c1o = const_cast<float*>(c1o); a.c2o = const_cast<float*>(c2o); a.pool2 = const_cast<float*>(pool2);
  a.l1o = const_cast<float*>(l1o); a.theta = const_cast<float*>(theta); a.A = const_cast<float*>(A);
  q.ggrid = ggrid; q.gin = gin; q.gl1 = gl1; q.np_part = stn_fused_part_floats(p.ch, p.nth); q.part = skip_param_grads ? nullptr : part; q.amax_out = amax_out;
  CG_LAUNCH(k_stn_loc_bwd, B, 512, stn_smem_bytes(p.ch, p.S), q);
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
