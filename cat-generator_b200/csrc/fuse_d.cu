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
