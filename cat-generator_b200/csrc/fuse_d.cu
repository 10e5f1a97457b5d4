// fuse_d.cu -- the element-wise chains between D32_st3's convolutions (models.lua:646-660,680-699) as ONE kernel each.
//
// Forward, per convolution output y:   nn.PReLU -> [SpatialAveragePooling | SpatialMaxPooling](2,2) -> SpatialDropout mask  and then
// whatever the consumer needs: a dense fp32 tensor (the trunk output T, read by three transformers), a channel slot of the nn.Concat
// buffer (models.lua:661-699), and / or the NEXT convolution's operand -- blocked, zero-padded fp16 (conv_tc.cu k_pack_act layout), which
// that layer's forward AND its weight gradient read, so the fp32 activation between two convolutions is never materialised.
// Round 1 ran each of these modules as its own kernel plus a pack per convolution call (profiles/r01_bench_tc_engine.json: k_pack_act 94
// launches, k_prelu_fwd 22, k_mask_channels 24, k_max/avg_fwd 26, k_copy_channels 16 per step).  Arithmetic per element is that of the
// single-module kernels in ops.cu, in the same order, so results are bit-identical to the unfused sequence.
#include "model.cuh"

namespace cg {

struct ActFwdArgs {
  const float *y, *pw, *mask;          // conv output [N,H,W,C]; PReLU slope (device scalar); dropout multipliers [N][mask_stride] (+ offset applied by the caller) or null
  int N, H, W, C, pool, mask_stride;   // pool: 0 none, 1 average 2x2, 2 max 2x2 (first maximum in scan order wins, ops.cu k_max_fwd)
  uint8_t* idx;                        // max-pool argmax [N,Ho,Wo,C], or null
  float* out; int out_stride, out_off; // fp32 result at [n,yo,xo, out_off + c] of a tensor with out_stride channels, or null
  uint8_t* xq; int p, Hq, Wq;          // fp16 operand of a following k x k convolution (p = (k-1)/2), or null
};

__device__ __forceinline__ void act8(const ActFwdArgs& a, long n, int yo, int xo, int c8, float a_slope, float* v, uint32_t* kidx) {
  const int C = a.C;
  if (a.pool == 0) {
    const float* s = a.y + ((n * a.H + yo) * a.W + xo) * C + c8;
    float4 v0 = *reinterpret_cast<const float4*>(s), v1 = *reinterpret_cast<const float4*>(s + 4);
    float t[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = t[j] > 0.f ? t[j] : a_slope * t[j];
    return;
  }
  const float* s = a.y + ((n * a.H + 2 * yo) * a.W + 2 * xo) * C + c8;
  float q[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float* sr = s + ((r >> 1) * (long)a.W + (r & 1)) * C;
    float4 v0 = *reinterpret_cast<const float4*>(sr), v1 = *reinterpret_cast<const float4*>(sr + 4);
    float t[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) q[r][j] = t[j] > 0.f ? t[j] : a_slope * t[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (a.pool == 1) v[j] = (q[0][j] + q[1][j] + q[2][j] + q[3][j]) * 0.25f;
    else {
      float b = q[0][j]; uint32_t k = 0;
      if (q[1][j] > b) { b = q[1][j]; k = 1; }
      if (q[2][j] > b) { b = q[2][j]; k = 2; }
      if (q[3][j] > b) { b = q[3][j]; k = 3; }
      v[j] = b; kidx[j] = k;
    }
  }
}

__global__ void k_act_fwd(ActFwdArgs a, long nwork) {
  const int C = a.C, Cq = C / 8, Ho = a.pool ? a.H / 2 : a.H, Wo = a.pool ? a.W / 2 : a.W;
  const float slope = *a.pw;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nwork; i += (long)gridDim.x * blockDim.x) {
    long n; int c, yo, xo; bool inside = true;
    if (a.xq) {                        // iterate over the padded operand: [N][Cq][Hq][Wq] chunks of 16 bytes
      int xx = (int)(i % a.Wq); long t = i / a.Wq; int yy = (int)(t % a.Hq); t /= a.Hq; c = (int)(t % Cq); n = t / Cq;
      yo = yy - a.p; xo = xx - a.p; inside = yo >= 0 && yo < Ho && xo >= 0 && xo < Wo;
    } else {                           // [N][Ho][Wo][Cq]
      c = (int)(i % Cq); long t = i / Cq; xo = (int)(t % Wo); t /= Wo; yo = (int)(t % Ho); n = t / Ho;
    }
    uint4 packed = make_uint4(0, 0, 0, 0);
    if (inside) {
      float v[8]; uint32_t kidx[8];
      act8(a, n, yo, xo, c * 8, slope, v, kidx);
      if (a.pool == 2 && a.idx) {
        uint8_t* d = a.idx + ((n * Ho + yo) * Wo + xo) * C + c * 8;
        uint2 w; w.x = kidx[0] | (kidx[1] << 8) | (kidx[2] << 16) | (kidx[3] << 24); w.y = kidx[4] | (kidx[5] << 8) | (kidx[6] << 16) | (kidx[7] << 24);
        *reinterpret_cast<uint2*>(d) = w;
      }
      if (a.mask) {
        const float* mk = a.mask + n * a.mask_stride + c * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= mk[j];
      }
      if (a.out) {
        float* o = a.out + ((n * Ho + yo) * Wo + xo) * a.out_stride + a.out_off + c * 8;
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
      if (a.xq) {
        __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]), h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
        packed.x = *reinterpret_cast<uint32_t*>(&h0); packed.y = *reinterpret_cast<uint32_t*>(&h1); packed.z = *reinterpret_cast<uint32_t*>(&h2); packed.w = *reinterpret_cast<uint32_t*>(&h3);
      }
    }
    if (a.xq) reinterpret_cast<uint4*>(a.xq)[i] = packed;
  }
}

// y [N,H,W,C] (C % 8 == 0; H, W even when pooling).  k: the following convolution's filter size when xq is wanted.
int act_pool_mask_pack(const float* y, const float* pw, int N, int H, int W, int C, int pool, const float* mask, int mask_stride, uint8_t* idx,
                       float* out, int out_stride, int out_off, uint8_t* xq, int k) {
  if (C % 8 || (pool && ((H | W) & 1)) || (out && ((out_stride | out_off) & 3))) return set_err(CG_ERR_ARG, "act_pool_mask_pack: unsupported shape");
  ActFwdArgs a{};
  a.y = y; a.pw = pw; a.mask = mask; a.N = N; a.H = H; a.W = W; a.C = C; a.pool = pool; a.mask_stride = mask_stride; a.idx = idx;
  a.out = out; a.out_stride = out_stride; a.out_off = out_off; a.xq = xq;
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  long nwork;
  if (xq) { a.p = (k - 1) / 2; a.Hq = ((Ho + 15) / 16) * 16 + 2 * a.p; a.Wq = Wo + 2 * a.p; nwork = (long)N * (C / 8) * a.Hq * a.Wq; }
  else nwork = (long)N * Ho * Wo * (C / 8);
  ctx().next_bytes = 4.0 * N * H * W * C + (out ? 4.0 * N * Ho * Wo * C : 0.0) + (xq ? 16.0 * nwork : 0.0);
  CG_LAUNCH(k_act_fwd, grid1d(nwork, 256), 256, 0, a, nwork);
  return CG_OK;
}

}  // namespace cg

namespace cg {

// ======================================================================================================================
// Backward.  Per convolution output y the reference chain is  (upstream gradient) -> SpatialDropout' -> pool' -> PReLU'  followed by the
// layer's accGradParameters / updateGradInput.  One kernel computes gc = dL/dy from the upstream gradient(s) and emits everything the
// rest of the layer's backward needs in the same pass:
//   * the fp16 gradient OPERAND of the two backward convolutions (blocked / zero-padded, multiplied by a per-tensor power of two);
//   * per-(image, 8-channel plane) partial sums of the bias gradient and of PReLU's weight gradient (fixed-order finalize below).
// The packing scale cannot wait for max|gc| (that would be a second pass): it is derived from max|upstream|, which the PRODUCER of the
// upstream gradient records for free in its epilogue (conv_tc.cu amax_out, stn_fused.cu) -- |gc| <= sum_i max|g_i| * max(1, |slope|)
// (* 1/4 behind an average pool), so the scale is at most a few binades more conservative than round 1's exact one: fp16 keeps 10 bits
// of mantissa either way and nothing can overflow.
struct ActBwdArgs {
  const float* g[4]; int ng, g_stride, g_off;   // upstream gradient(s) [N,Hu,Wu,g_stride] at channel offset g_off, summed in index order (nn.Concat's backward order)
  const float* mask; int mask_stride;            // dropout multipliers [N][mask_stride] (pointer already offset) or null
  int pool; const uint8_t* idx;                  // 0 none / 1 average / 2 max (argmax bytes [N,Hu,Wu,C])
  const float *y, *pw;                           // conv output = PReLU input [N,H,W,C]; slope (device scalar)
  int N, H, W, C;
  const unsigned int* amax[4];                   // max|g[i]| as float bits
  float* scale2;                                 // out: [scale, 1/scale]
  uint8_t* gq; int p, Hq, Wq;                    // packed operand for the layer's k x k backward convolutions
  double* part;                                  // [N * C/8][9]; null: parameter gradients are not wanted
};

__global__ void __launch_bounds__(256) k_act_bwd(ActBwdArgs a) {
  __shared__ double red[9][8];
  const int c = blockIdx.x, tid = threadIdx.x, C = a.C, Cq = C / 8;
  const long n = blockIdx.y;
  const float slope = *a.pw;
  float bound = 0.f;
  for (int i = 0; i < a.ng; ++i) bound += __uint_as_float(*a.amax[i]);
  bound *= fmaxf(1.f, fabsf(slope));
  if (a.pool == 1) bound *= 0.25f;
  float sc = 1.f;
  if (bound > 0.f && isfinite(bound)) sc = exp2f(floorf(log2f(16384.f / bound)));
  if (!(sc > 0.f) || !isfinite(sc)) sc = 1.f;
  if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) { a.scale2[0] = sc; a.scale2[1] = 1.f / sc; }
  const int Hu = a.pool ? a.H / 2 : a.H, Wu = a.pool ? a.W / 2 : a.W;
  float cs[8] = {0, 0, 0, 0, 0, 0, 0, 0}; float ps = 0.f;
  uint4* dst = reinterpret_cast<uint4*>(a.gq) + (n * Cq + c) * (long)a.Hq * a.Wq;
  for (int i = tid; i < a.Hq * a.Wq; i += blockDim.x) {
    const int yy = i / a.Wq, xx = i - yy * a.Wq, oy = yy - a.p, ox = xx - a.p;
    uint4 packed = make_uint4(0, 0, 0, 0);
    if (oy >= 0 && oy < a.H && ox >= 0 && ox < a.W) {
      const int yu = a.pool ? oy >> 1 : oy, xu = a.pool ? ox >> 1 : ox;
      const long uo = ((n * Hu + yu) * Wu + xu);
      float g[8];
      {
        const float* s = a.g[0] + uo * a.g_stride + a.g_off + c * 8;
        float4 v0 = *reinterpret_cast<const float4*>(s), v1 = *reinterpret_cast<const float4*>(s + 4);
        g[0] = v0.x; g[1] = v0.y; g[2] = v0.z; g[3] = v0.w; g[4] = v1.x; g[5] = v1.y; g[6] = v1.z; g[7] = v1.w;
      }
      if (a.ng > 1) {   // gT = ((((0 + p0) + p1) + p2) + p3): the order nn.Concat's backward adds the branches in
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 0.f + g[j];
        for (int q = 1; q < a.ng; ++q) {
          const float* s = a.g[q] + uo * a.g_stride + a.g_off + c * 8;
          float4 v0 = *reinterpret_cast<const float4*>(s), v1 = *reinterpret_cast<const float4*>(s + 4);
          g[0] += v0.x; g[1] += v0.y; g[2] += v0.z; g[3] += v0.w; g[4] += v1.x; g[5] += v1.y; g[6] += v1.z; g[7] += v1.w;
        }
      }
      if (a.mask) {
        const float* mk = a.mask + n * a.mask_stride + c * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] *= mk[j];
      }
      if (a.pool == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] *= 0.25f;
      } else if (a.pool == 2) {
        const uint2 w = *reinterpret_cast<const uint2*>(a.idx + uo * C + c * 8);
        const uint32_t k = (uint32_t)(((oy & 1) << 1) | (ox & 1));
#pragma unroll
        for (int j = 0; j < 8; ++j) { uint32_t kj = ((j < 4 ? w.x : w.y) >> (8 * (j & 3))) & 0xffu; if (kj != k) g[j] = 0.f; }
      }
      const float* ys = a.y + ((n * a.H + oy) * a.W + ox) * C + c * 8;
      float4 y0 = *reinterpret_cast<const float4*>(ys), y1 = *reinterpret_cast<const float4*>(ys + 4);
      const float yv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {   // nn.PReLU backward (ops.cu k_prelu_bwd): gx = x > 0 ? g : w g ; gw += sum_{x <= 0} x g
        if (!(yv[j] > 0.f)) { ps += yv[j] * g[j]; g[j] = slope * g[j]; }
        cs[j] += g[j];
      }
      __half2 h0 = __floats2half2_rn(g[0] * sc, g[1] * sc), h1 = __floats2half2_rn(g[2] * sc, g[3] * sc), h2 = __floats2half2_rn(g[4] * sc, g[5] * sc), h3 = __floats2half2_rn(g[6] * sc, g[7] * sc);
      packed.x = *reinterpret_cast<uint32_t*>(&h0); packed.y = *reinterpret_cast<uint32_t*>(&h1); packed.z = *reinterpret_cast<uint32_t*>(&h2); packed.w = *reinterpret_cast<uint32_t*>(&h3);
    }
    dst[i] = packed;
  }
  if (!a.part) return;
  // fixed-order block reduction of the nine partial sums
  double v[9];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = warp_sum_d((double)cs[j]);
  v[8] = warp_sum_d((double)ps);
  const int w = tid >> 5, l = tid & 31;
  if (l == 0) {
#pragma unroll
    for (int j = 0; j < 9; ++j) red[j][w] = v[j];
  }
  __syncthreads();
  if (tid < 9) {
    double s = 0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) s += red[tid][q];
    a.part[(n * Cq + c) * 9 + tid] = s;
  }
}
// gb[c] += sum_n part[n][c / 8][c % 8] ; gpw += sum_{n, plane} part[n][plane][8]   (fixed order)
__global__ void __launch_bounds__(256) k_act_bwd_final(const double* __restrict__ part, int N, int Cq, float* __restrict__ gb, float* __restrict__ gpw) {
  __shared__ double red[8];
  const int tid = threadIdx.x;
  for (int c = tid; c < Cq * 8; c += blockDim.x) {
    double s = 0;
    for (int n = 0; n < N; ++n) s += part[((long)n * Cq + (c >> 3)) * 9 + (c & 7)];
    gb[c] += (float)s;
  }
  double s = 0;
  for (int i = tid; i < N * Cq; i += blockDim.x) s += part[(long)i * 9 + 8];
  s = warp_sum_d(s);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) { double t = 0; for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += red[q]; *gpw += (float)t; }
}

size_t act_bwd_operand_bytes(int N, int H, int W, int C, int k) {
  const int p = (k - 1) / 2;
  return (size_t)N * (C / 8) * (((H + 15) / 16) * 16 + 2 * p) * (W + 2 * p) * 16;
}
// y [N,H,W,C] with C % 64 == 0.  g[i]: upstream gradients at the pooled resolution.  part: N * C/8 * 9 doubles, or null with gb / gpw null.
int act_bwd_pack(const float* const* g, const unsigned int* const* amax, int ng, int g_stride, int g_off, const float* mask, int mask_stride, int pool, const uint8_t* idx,
                 const float* y, const float* pw, int N, int H, int W, int C, int k, uint8_t* gq, float* scale2, double* part, float* gb_acc, float* gpw_acc) {
  if (C % 64 || ng < 1 || ng > 4 || (pool && ((H | W) & 1)) || ((g_stride | g_off) & 3)) return set_err(CG_ERR_ARG, "act_bwd_pack: unsupported shape");
  ActBwdArgs a{};
  for (int i = 0; i < ng; ++i) { a.g[i] = g[i]; a.amax[i] = amax[i]; }
  a.ng = ng; a.g_stride = g_stride; a.g_off = g_off; a.mask = mask; a.mask_stride = mask_stride; a.pool = pool; a.idx = idx; a.y = y; a.pw = pw;
  a.N = N; a.H = H; a.W = W; a.C = C; a.scale2 = scale2; a.gq = gq; a.p = (k - 1) / 2; a.Hq = ((H + 15) / 16) * 16 + 2 * a.p; a.Wq = W + 2 * a.p;
  a.part = part;
  ctx().next_bytes = 4.0 * N * H * W * C * (1.0 + (pool ? 0.25 : 1.0) * ng) + 16.0 * N * (C / 8) * a.Hq * a.Wq;
  CG_LAUNCH(k_act_bwd, dim3(C / 8, N), 256, 0, a);
  if (part) CG_LAUNCH(k_act_bwd_final, 1, 256, 0, (const double*)part, N, C / 8, gb_acc, gpw_acc);
  return CG_OK;
}


// =================================================================== the discriminator's head (models.lua:697-701)
//   Linear(20480,256) output h1o -> PReLU -> Dropout(mask) -> Linear(256,1) -> Sigmoid           (forward, one warp per sample)
//   and its mirror image                                                                          (backward, one block)
// As separate modules this was 6 launches forward and 13 backward on the step's main stream, each a few microseconds of dependent launch
// latency for 32k elements (profiles/r02_timeline.txt, 1978-2513 us).  Everything here is fp32; sums run in a fixed order.
__global__ void k_d_head_fwd(const float* __restrict__ h1o, const float* __restrict__ pw, const float* __restrict__ mask, const float* __restrict__ W2,
                             const float* __restrict__ b2, float* __restrict__ hd, float* __restrict__ h2o, float* __restrict__ hsig, int B) {
  const int w = (int)((blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= B) return;
  const float a = *pw; float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int o = lane + 32 * j; const long i = (long)w * 256 + o;
    const float v = h1o[i], act = v > 0.f ? v : a * v, d = act * mask[i];
    hd[i] = d; s = fmaf(d, W2[o], s);
  }
  s = warp_sum(s);
  if (lane == 0) { const float pre = s + b2[0]; h2o[w] = pre; hsig[w] = 1.f / (1.f + expf(-pre)); }
}
int d_head_fwd(const float* h1o, const float* pw, const float* mask, const float* W2, const float* b2, float* hd, float* h2o, float* hsig, int B) {
  ctx().next_bytes = 4.0 * B * 256 * 3;
  CG_LAUNCH(k_d_head_fwd, cdiv((long)B * 32, 256), 256, 0, h1o, pw, mask, W2, b2, hd, h2o, hsig, B); return CG_OK;
}
// gout [B] = dLoss/dSigmoid.  Writes gh1 [B,256] (gradient w.r.t. Linear1's output); with param_grads also accumulates Linear2's weight and
// bias gradient and the PReLU slope gradient.  1024 threads = 256 columns x 4 batch quarters, quarters combined in order 0..3.
__global__ void __launch_bounds__(1024) k_d_head_bwd(const float* __restrict__ gout, const float* __restrict__ hsig, const float* __restrict__ hd, const float* __restrict__ h1o,
                                                     const float* __restrict__ pw, const float* __restrict__ mask, const float* __restrict__ W2, float* __restrict__ gh1,
                                                     float* __restrict__ gW2, float* __restrict__ gb2, float* __restrict__ gpw, int B, int param_grads) {
  __shared__ float s_g2[1024];
  __shared__ float red[4][256];
  __shared__ double redd[32];
  const int tid = threadIdx.x, o = tid & 255, bq = tid >> 8;
  for (int b = tid; b < B; b += 1024) { const float y = hsig[b]; s_g2[b] = gout[b] * y * (1.f - y); }     // nn.Sigmoid backward
  __syncthreads();
  const float a = *pw, w2 = W2[o];
  float cW2 = 0.f; double cpw = 0.0;
  for (int b = bq; b < B; b += 4) {
    const float g2 = s_g2[b]; const long i = (long)b * 256 + o;
    cW2 = fmaf(g2, hd[i], cW2);                                          // Linear2: gW2[o] += gh2[b] * hd[b][o]
    const float g = g2 * w2 * mask[i];                                    // Linear2 input gradient, Dropout backward
    const float x = h1o[i];
    float r;
    if (x > 0.f) r = g; else { r = a * g; cpw += (double)(x * g); }       // PReLU backward (ops.cu k_prelu_bwd)
    gh1[i] = r;
  }
  if (!param_grads) return;
  red[bq][o] = cW2;
  __syncthreads();
  if (bq == 0) gW2[o] += (red[0][o] + red[1][o]) + (red[2][o] + red[3][o]);
  // PReLU slope gradient: warp sums, then warps in order
  cpw = warp_sum_d(cpw);
  if ((tid & 31) == 0) redd[tid >> 5] = cpw;
  __syncthreads();
  if (tid == 0) { double t = 0; for (int k = 0; k < 32; ++k) t += redd[k]; gpw[0] += (float)t; }
  if (tid < 32) {                                                        // Linear2 bias gradient = sum_b gh2[b]
    float v = 0.f; for (int b = tid; b < B; b += 32) v += s_g2[b];
    v = warp_sum(v);
    if (tid == 0) gb2[0] += v;
  }
}
int d_head_bwd(const float* gout, const float* hsig, const float* hd, const float* h1o, const float* pw, const float* mask, const float* W2, float* gh1,
               float* gW2, float* gb2, float* gpw, int B, int param_grads) {
  if (B > 1024) return set_err(CG_ERR_UNSUPPORTED, "d_head_bwd: batch %d", B);
  ctx().next_bytes = 4.0 * B * 256 * 4;
  CG_LAUNCH(k_d_head_bwd, 1, 1024, 0, gout, hsig, hd, h1o, pw, mask, W2, gh1, gW2, gb2, gpw, B, param_grads); return CG_OK;
}
}  // namespace cg
