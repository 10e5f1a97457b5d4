// ops.cu -- HBM-bound kernels of the hot path: layout changes, PReLU/LeakyReLU/Sigmoid, nearest upsample,
// pooling, dropout masks, batch norm (warp-shuffle + smem column reductions), the spatial transformer,
// BCE, penalty/clamp and Adam.  Semantics: SURVEY.md Appendix A; reference call sites cited per kernel.
#include "ops.cuh"

namespace cg {

// =================================================================== layout
// Torch NCHW boundary <-> internal NHWC.  32x32 smem tile transpose per image: both sides coalesced.
__global__ void k_transpose(const float* __restrict__ x, float* __restrict__ y, int R, int Cc) {
  // x: [n][R][Cc] -> y: [n][Cc][R]
  __shared__ float t[32][33];
  int n = blockIdx.z;
  const float* xs = x + (long)n * R * Cc;
  float* ys = y + (long)n * R * Cc;
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Cc) t[i][threadIdx.x] = xs[(long)r * Cc + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) ys[(long)c * R + r] = t[threadIdx.x][i];
  }
}
int nchw_to_nhwc(const float* x, float* y, int N, int C, int HW) {   // [n][C][HW] -> [n][HW][C]
  if (C == 1 || HW == 1) { CG_CUDA(cudaMemcpyAsync(y, x, sizeof(float) * (size_t)N * C * HW, cudaMemcpyDeviceToDevice, ctx().stream)); return CG_OK; }
  dim3 g(cdiv(HW, 32), cdiv(C, 32), N), b(32, 8);
  CG_LAUNCH(k_transpose, g, b, 0, x, y, C, HW);
  return CG_OK;
}
int nhwc_to_nchw(const float* x, float* y, int N, int C, int HW) {   // [n][HW][C] -> [n][C][HW]
  if (C == 1 || HW == 1) { CG_CUDA(cudaMemcpyAsync(y, x, sizeof(float) * (size_t)N * C * HW, cudaMemcpyDeviceToDevice, ctx().stream)); return CG_OK; }
  dim3 g(cdiv(C, 32), cdiv(HW, 32), N), b(32, 8);
  CG_LAUNCH(k_transpose, g, b, 0, x, y, HW, C);
  return CG_OK;
}

// =================================================================== reductions to a scalar (deterministic 2 stage)
__device__ __forceinline__ double block_sum_d(double v) {
  __shared__ double sh[32];
  __syncthreads();
  v = warp_sum_d(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0;
  if (w == 0) {
    r = l < (blockDim.x >> 5) ? sh[l] : 0.0;
    r = warp_sum_d(r);
  }
  return r;   // valid in thread 0
}
// out (float) : mode 0 -> = scale*sum, mode 1 -> += scale*sum
__global__ void k_final_sum(const double* __restrict__ part, int nparts, int stride, int which, float* out, double scale, int mode) {
  double s = 0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += part[(long)i * stride + which];
  s = block_sum_d(s);
  if (threadIdx.x == 0) { if (mode) *out += (float)(s * scale); else *out = (float)(s * scale); }
}

// =================================================================== pointwise
// nn.PReLU single shared slope: y = x>0 ? x : w*x
__global__ void k_prelu_fwd(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, long n) {
  float a = *w;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = x[i]; y[i] = v > 0.f ? v : a * v;
  }
}
int prelu_fwd(const float* x, const float* w, float* y, long n) {
  CG_LAUNCH(k_prelu_fwd, grid1d(n, 256, 4), 256, 0, x, w, y, n); return CG_OK;
}
// gx = x>0 ? g : w*g ; gw += sum_{x<=0} x*g   (block partials in double, fixed-order final sum)
__global__ void k_prelu_bwd(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ w,
                            float* __restrict__ gx, double* __restrict__ part, long n) {
  float a = *w; double s = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = x[i], g = gy[i];
    if (v > 0.f) { if (gx) gx[i] = g; }
    else { if (gx) gx[i] = a * g; s += (double)(v * g); }
  }
  s = block_sum_d(s);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
int prelu_bwd(const float* x, const float* gy, const float* w, float* gx, float* gw_acc, long n) {
  int g = grid1d(n, 256, 4);
  double* part = (double*)workspace(sizeof(double) * g);
  if (!part) return CG_ERR_CUDA;
  CG_LAUNCH(k_prelu_bwd, g, 256, 0, x, gy, w, gx, part, n);
  if (gw_acc) CG_LAUNCH(k_final_sum, 1, 256, 0, part, g, 1, 0, gw_acc, 1.0, 1);
  return CG_OK;
}
// nn.LeakyReLU (/root/reference/LeakyReLU.lua:13-19): max(x,0) + s*min(x,0)
__global__ void k_lrelu_fwd(const float* __restrict__ x, float s, float* __restrict__ y, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = x[i]; y[i] = v >= 0.f ? v : s * v;
  }
}
int lrelu_fwd(const float* x, float s, float* y, long n) { CG_LAUNCH(k_lrelu_fwd, grid1d(n, 256, 4), 256, 0, x, s, y, n); return CG_OK; }
// LeakyReLU.lua:21-31: gradient is gy where x >= 0 (INCLUDING x == 0), s*gy where x < 0
__global__ void k_lrelu_bwd(const float* __restrict__ x, const float* __restrict__ gy, float s, float* __restrict__ gx, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    gx[i] = x[i] >= 0.f ? gy[i] : s * gy[i];
}
int lrelu_bwd(const float* x, const float* gy, float s, float* gx, long n) { CG_LAUNCH(k_lrelu_bwd, grid1d(n, 256, 4), 256, 0, x, gy, s, gx, n); return CG_OK; }
__global__ void k_sigmoid_fwd(const float* __restrict__ x, float* __restrict__ y, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) y[i] = 1.f / (1.f + expf(-x[i]));
}
int sigmoid_fwd(const float* x, float* y, long n) { CG_LAUNCH(k_sigmoid_fwd, grid1d(n, 256, 4), 256, 0, x, y, n); return CG_OK; }
__global__ void k_sigmoid_bwd(const float* __restrict__ y, const float* __restrict__ gy, float* __restrict__ gx, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) { float v = y[i]; gx[i] = gy[i] * v * (1.f - v); }
}
int sigmoid_bwd(const float* y, const float* gy, float* gx, long n) { CG_LAUNCH(k_sigmoid_bwd, grid1d(n, 256, 4), 256, 0, y, gy, gx, n); return CG_OK; }
// nn.SoftMax: y = exp(x - max) / sum exp(x - max) per row (one thread per row; rows here are 2 wide)
__global__ void k_softmax_rows(const float* __restrict__ x, float* __restrict__ y, long rows, int C) {
  for (long r = blockIdx.x * (long)blockDim.x + threadIdx.x; r < rows; r += (long)gridDim.x * blockDim.x) {
    const float* s = x + r * C; float m = s[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, s[c]);
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(s[c] - m);
    for (int c = 0; c < C; ++c) y[r * C + c] = expf(s[c] - m) / sum;
  }
}
int softmax_rows(const float* x, float* y, long rows, int C) { CG_LAUNCH(k_softmax_rows, grid1d(rows, 128), 128, 0, x, y, rows, C); return CG_OK; }
__global__ void k_add(float* __restrict__ a, const float* __restrict__ b, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) a[i] += b[i];
}
int add_inplace(float* a, const float* b, long n) { CG_LAUNCH(k_add, grid1d(n, 256, 4), 256, 0, a, b, n); return CG_OK; }
__global__ void k_fill(float* __restrict__ a, float v, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) a[i] = v;
}
int fill(float* a, float v, long n) { CG_LAUNCH(k_fill, grid1d(n, 256, 4), 256, 0, a, v, n); return CG_OK; }
__global__ void k_scale(float* __restrict__ a, float s, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) a[i] *= s;
}
int scale_inplace(float* a, float s, long n) { CG_LAUNCH(k_scale, grid1d(n, 256, 4), 256, 0, a, s, n); return CG_OK; }
// nn.SpatialDropout (A.12): one multiplier per (n, c), NHWC
__global__ void k_mask_channels(const float* __restrict__ x, const float* __restrict__ m, float* __restrict__ y, long n, int HWC, int C) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int img = (int)(i / HWC), c = (int)(i % C);
    y[i] = x[i] * m[(long)img * C + c];
  }
}
int mask_channels(const float* x, const float* mask_nc, float* y, int N, int HW, int C) {
  long n = (long)N * HW * C;
  CG_LAUNCH(k_mask_channels, grid1d(n, 256, 4), 256, 0, x, mask_nc, y, n, HW * C, C); return CG_OK;
}
__global__ void k_mask_elems(const float* __restrict__ x, const float* __restrict__ m, float* __restrict__ y, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) y[i] = x[i] * m[i];
}
int mask_elems(const float* x, const float* mask, float* y, long n) { CG_LAUNCH(k_mask_elems, grid1d(n, 256, 4), 256, 0, x, mask, y, n); return CG_OK; }

// =================================================================== NHWC spatial
// nn.SpatialUpSamplingNearest(2) (A.5)
__global__ void k_up_fwd(const float* __restrict__ x, float* __restrict__ y, long n, int H, int W, int C) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long p = i / C; int X = (int)(p % (2 * W)); p /= (2 * W); int Y = (int)(p % (2 * H)); long img = p / (2 * H);
    y[i] = x[((img * H + (Y >> 1)) * W + (X >> 1)) * C + c];
  }
}
int upsample2x_fwd(const float* x, float* y, int N, int H, int W, int C) {
  long n = (long)N * 4 * H * W * C; CG_LAUNCH(k_up_fwd, grid1d(n, 256, 4), 256, 0, x, y, n, H, W, C); return CG_OK;
}
__global__ void k_up_bwd(const float* __restrict__ gy, float* __restrict__ gx, long n, int H, int W, int C) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long p = i / C; int xx = (int)(p % W); p /= W; int yy = (int)(p % H); long img = p / H;
    const float* g = gy + ((img * 2 * H + 2 * yy) * 2 * W + 2 * xx) * C + c;
    gx[i] = g[0] + g[C] + g[(long)2 * W * C] + g[(long)2 * W * C + C];
  }
}
int upsample2x_bwd(const float* gy, float* gx, int N, int H, int W, int C) {
  long n = (long)N * H * W * C; CG_LAUNCH(k_up_bwd, grid1d(n, 256, 4), 256, 0, gy, gx, n, H, W, C); return CG_OK;
}
// nn.SpatialAveragePooling(2,2,2,2) (A.12)
__global__ void k_avg_fwd(const float* __restrict__ x, float* __restrict__ y, long n, int H, int W, int C) {
  int Ho = H / 2, Wo = W / 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long p = i / C; int xx = (int)(p % Wo); p /= Wo; int yy = (int)(p % Ho); long img = p / Ho;
    const float* s = x + ((img * H + 2 * yy) * W + 2 * xx) * C + c;
    y[i] = (s[0] + s[C] + s[(long)W * C] + s[(long)W * C + C]) * 0.25f;
  }
}
int avgpool2_fwd(const float* x, float* y, int N, int H, int W, int C) {
  long n = (long)N * (H / 2) * (W / 2) * C; CG_LAUNCH(k_avg_fwd, grid1d(n, 256, 4), 256, 0, x, y, n, H, W, C); return CG_OK;
}
__global__ void k_avg_bwd(const float* __restrict__ gy, float* __restrict__ gx, long n, int H, int W, int C) {
  int Ho = H / 2, Wo = W / 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long p = i / C; int X = (int)(p % W); p /= W; int Y = (int)(p % H); long img = p / H;
    gx[i] = gy[((img * Ho + (Y >> 1)) * Wo + (X >> 1)) * C + c] * 0.25f;
  }
}
int avgpool2_bwd(const float* gy, float* gx, int N, int H, int W, int C) {
  long n = (long)N * H * W * C; CG_LAUNCH(k_avg_bwd, grid1d(n, 256, 4), 256, 0, gy, gx, n, H, W, C); return CG_OK;
}
// nn.SpatialMaxPooling(2,2) (A.12): first maximum wins in scan order (0,0),(0,1),(1,0),(1,1)
__global__ void k_max_fwd(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ idx, long n, int H, int W, int C) {
  int Ho = H / 2, Wo = W / 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long p = i / C; int xx = (int)(p % Wo); p /= Wo; int yy = (int)(p % Ho); long img = p / Ho;
    const float* s = x + ((img * H + 2 * yy) * W + 2 * xx) * C + c;
    float v0 = s[0], v1 = s[C], v2 = s[(long)W * C], v3 = s[(long)W * C + C];
    float b = v0; int k = 0;
    if (v1 > b) { b = v1; k = 1; }
    if (v2 > b) { b = v2; k = 2; }
    if (v3 > b) { b = v3; k = 3; }
    y[i] = b; idx[i] = (uint8_t)k;
  }
}
int maxpool2_fwd(const float* x, float* y, uint8_t* idx, int N, int H, int W, int C) {
  long n = (long)N * (H / 2) * (W / 2) * C; CG_LAUNCH(k_max_fwd, grid1d(n, 256, 4), 256, 0, x, y, idx, n, H, W, C); return CG_OK;
}
__global__ void k_max_bwd(const float* __restrict__ gy, const uint8_t* __restrict__ idx, float* __restrict__ gx, long n, int H, int W, int C) {
  int Ho = H / 2, Wo = W / 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long p = i / C; int X = (int)(p % W); p /= W; int Y = (int)(p % H); long img = p / H;
    long o = ((img * Ho + (Y >> 1)) * Wo + (X >> 1)) * C + c;
    int k = ((Y & 1) << 1) | (X & 1);
    gx[i] = idx[o] == k ? gy[o] : 0.f;
  }
}
int maxpool2_bwd(const float* gy, const uint8_t* idx, float* gx, int N, int H, int W, int C) {
  long n = (long)N * H * W * C; CG_LAUNCH(k_max_bwd, grid1d(n, 256, 4), 256, 0, gy, idx, gx, n, H, W, C); return CG_OK;
}

// =================================================================== column reductions over [M, C]
// block (32, 8): 32 consecutive channels (coalesced 128 B rows) x 8 row lanes; grid (ceil(C/32), S).
// Each thread accumulates a short fp32 run, lanes combine in double through smem; partials [S][C][2] are
// summed in fixed order by the finalize kernels => deterministic.  MODE: 0 = (x, x*x); 1 = (g, g*xhat);
// 2 = (x, 0).
template <int MODE>
__global__ void k_colreduce(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ mean,
                            const float* __restrict__ invstd, double* __restrict__ part, long M, int C, long rows_per_split) {
  __shared__ double sh[2][8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  long r0 = (long)blockIdx.y * rows_per_split, r1 = r0 + rows_per_split; if (r1 > M) r1 = M;
  double a0 = 0, a1 = 0;
  if (c < C) {
    float mu = 0.f, is = 1.f;
    if (MODE == 1) { mu = mean[c]; is = invstd[c]; }
    float f0 = 0, f1 = 0; int cnt = 0;
    for (long r = r0 + threadIdx.y; r < r1; r += 8) {
      float v = x[r * C + c];
      if (MODE == 0) { f0 += v; f1 += v * v; }
      else if (MODE == 1) { float gg = g[r * C + c]; f0 += gg; f1 += gg * ((v - mu) * is); }
      else { f0 += v; }
      if (++cnt == 64) { a0 += f0; a1 += f1; f0 = f1 = 0; cnt = 0; }
    }
    a0 += f0; a1 += f1;
  }
  sh[0][threadIdx.y][threadIdx.x] = a0; sh[1][threadIdx.y][threadIdx.x] = a1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double s0 = 0, s1 = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s0 += sh[0][i][threadIdx.x]; s1 += sh[1][i][threadIdx.x]; }
    part[((long)blockIdx.y * C + c) * 2 + 0] = s0;
    part[((long)blockIdx.y * C + c) * 2 + 1] = s1;
  }
}
static int colreduce_splits(long M, int C) {
  int colblocks = cdiv(C, 32);
  int want = (ctx().sm_count * 4 + colblocks - 1) / colblocks;
  long maxs = (M + 63) / 64; if (maxs < 1) maxs = 1;
  if (want > maxs) want = (int)maxs;
  if (want < 1) want = 1;
  return want;
}
// finalize kernels: one WARP per column; lane l sums partials l, l+32, ... then a fixed shuffle tree (deterministic)
__device__ __forceinline__ void col_partials(const double* __restrict__ part, int S, int C, int c, int lane, double& s0, double& s1) {
  s0 = 0; s1 = 0;
  for (int s = lane; s < S; s += 32) { s0 += part[((long)s * C + c) * 2]; s1 += part[((long)s * C + c) * 2 + 1]; }
  s0 = warp_sum_d(s0); s1 = warp_sum_d(s1);
}
// per-column totals of the split partials (fixed order): the quantity sync-BN all-reduces
__global__ void k_col_totals(const double* __restrict__ part, int S, int C, double* __restrict__ tot) {
  int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s0, s1; col_partials(part, S, C, c, lane, s0, s1);
  if (lane) return;
  tot[2 * c] = s0; tot[2 * c + 1] = s1;
}
// nn.SpatialBatchNormalization training forward (A.3): biased batch variance for normalisation,
// unbiased into running_var, momentum 0.1
__global__ void k_bn_stats_final(const double* __restrict__ part, int S, int C, double m, float eps, float mom,
                                 float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ run_mean, float* __restrict__ run_var) {
  int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s0, s1; col_partials(part, S, C, c, lane, s0, s1);
  if (lane) return;
  double mu = s0 / m, var = s1 / m - mu * mu; if (var < 0) var = 0;
  mean[c] = (float)mu; invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (run_mean) run_mean[c] = (1.f - mom) * run_mean[c] + mom * (float)mu;
  if (run_var) run_var[c] = (1.f - mom) * run_var[c] + mom * (float)(var * m / (m - 1.0));
}
__global__ void k_bn_apply(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                           const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ y, long n, int C) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    y[i] = (x[i] - mean[c]) * invstd[c] * gamma[c] + beta[c];
  }
}
static int bn_fwd_train_impl(const double* pre_part, int pre_S, const float* x, const float* gamma, const float* beta, float* y, float* mean, float* invstd,
                             float* run_mean, float* run_var, long M, int C, float eps, float mom) {
  int S = pre_part ? pre_S : colreduce_splits(M, C);
  long rps = (M + S - 1) / S;
  double* part = pre_part ? const_cast<double*>(pre_part) : (double*)workspace(sizeof(double) * 2 * (size_t)(S + 1) * C);
  if (!part) return CG_ERR_CUDA;
  if (!pre_part) {
    dim3 g(cdiv(C, 32), S), b(32, 8);
    ctx().next_bytes = 4.0 * (double)M * C;
    CG_LAUNCH(k_colreduce<0>, g, b, 0, x, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, part, M, C, rps);
  }
  if (ctx().sync_bn && ctx().world > 1) {
    // sync-BN: statistics over the GLOBAL batch = all-reduced per-channel (sum, sum of squares), 2*C doubles
    double* tot = part + 2 * (size_t)S * C;
    CG_LAUNCH(k_col_totals, cdiv(C, 4), 128, 0, part, S, C, tot);
    CG_TRY(dist_allreduce_sum_f64(tot, 2L * C));
    CG_LAUNCH(k_bn_stats_final, cdiv(C, 4), 128, 0, tot, 1, C, (double)M * ctx().world, eps, mom, mean, invstd, run_mean, run_var);
  } else
    CG_LAUNCH(k_bn_stats_final, cdiv(C, 4), 128, 0, part, S, C, (double)M, eps, mom, mean, invstd, run_mean, run_var);
  long n = M * C;
  if (y) CG_LAUNCH(k_bn_apply, grid1d(n, 256, 4), 256, 0, x, gamma, beta, mean, invstd, y, n, C);   // y == nullptr: statistics only (the caller fuses the apply)
  return CG_OK;
}
int bn_fwd_train(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* invstd,
                 float* run_mean, float* run_var, long M, int C, float eps, float mom) {
  return bn_fwd_train_impl(nullptr, 0, x, gamma, beta, y, mean, invstd, run_mean, run_var, M, C, eps, mom);
}
int bn_fwd_train_pre(const double* pre_part, int pre_S, const float* x, const float* gamma, const float* beta, float* y, float* mean, float* invstd,
                     float* run_mean, float* run_var, long M, int C, float eps, float mom) {
  return bn_fwd_train_impl(pre_part, pre_S, x, gamma, beta, y, mean, invstd, run_mean, run_var, M, C, eps, mom);
}
__global__ void k_bn_eval(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ rm, const float* __restrict__ rv, float* __restrict__ y, long n, int C, float eps) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    y[i] = (x[i] - rm[c]) * rsqrtf(rv[c] + eps) * gamma[c] + beta[c];
  }
}
int bn_fwd_eval(const float* x, const float* gamma, const float* beta, float* y, const float* run_mean, const float* run_var, long M, int C, float eps) {
  long n = M * C; CG_LAUNCH(k_bn_eval, grid1d(n, 256, 4), 256, 0, x, gamma, beta, run_mean, run_var, y, n, C, eps); return CG_OK;
}
// backward (A.3): ggamma += sum g*xhat, gbeta += sum g, gx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat))
__global__ void k_bn_bwd_final(const double* __restrict__ part, int S, int C, double m, float* __restrict__ mg, float* __restrict__ mgx,
                               float* __restrict__ ggamma_acc, float* __restrict__ gbeta_acc) {
  int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s0, s1; col_partials(part, S, C, c, lane, s0, s1);
  if (lane) return;
  mg[c] = (float)(s0 / m); mgx[c] = (float)(s1 / m);
  if (gbeta_acc) gbeta_acc[c] += (float)s0;
  if (ggamma_acc) ggamma_acc[c] += (float)s1;
}
__global__ void k_bn_bwd_apply(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ gamma,
                               const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ mg,
                               const float* __restrict__ mgx, float* __restrict__ gx, long n, int C) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float xh = (x[i] - mean[c]) * invstd[c];
    gx[i] = gamma[c] * invstd[c] * (gy[i] - mg[c] - xh * mgx[c]);
  }
}
// sync-BN backward: the means in gx are over the GLOBAL batch, while ggamma / gbeta accumulate this rank's LOCAL sums -- the
// parameter-gradient all-reduce (sum over ranks, scaled 1/world with the rest of the flat vector) then yields the gradient of the
// global-mean loss, exactly what one device with the whole batch computes.
__global__ void k_bn_bwd_final_sync(const double* __restrict__ tot_local, const double* __restrict__ tot_glob, int C, double m_glob,
                                    float* __restrict__ mg, float* __restrict__ mgx, float* __restrict__ ggamma_acc, float* __restrict__ gbeta_acc) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mg[c] = (float)(tot_glob[2 * c] / m_glob); mgx[c] = (float)(tot_glob[2 * c + 1] / m_glob);
  if (gbeta_acc) gbeta_acc[c] += (float)tot_local[2 * c];
  if (ggamma_acc) ggamma_acc[c] += (float)tot_local[2 * c + 1];
}
int bn_bwd(const float* x, const float* gy, const float* gamma, const float* mean, const float* invstd,
           float* gx, float* ggamma_acc, float* gbeta_acc, long M, int C) {
  int S = colreduce_splits(M, C);
  long rps = (M + S - 1) / S;
  size_t pbytes = sizeof(double) * 2 * (size_t)(S + 2) * C;
  char* wsb = (char*)workspace(pbytes + sizeof(float) * 2 * C);
  if (!wsb) return CG_ERR_CUDA;
  double* part = (double*)wsb; float* mg = (float*)(wsb + pbytes); float* mgx = mg + C;
  dim3 g(cdiv(C, 32), S), b(32, 8);
  ctx().next_bytes = 8.0 * (double)M * C;
  CG_LAUNCH(k_colreduce<1>, g, b, 0, x, gy, mean, invstd, part, M, C, rps);
  if (ctx().sync_bn && ctx().world > 1) {
    double* tot = part + 2 * (size_t)S * C; double* totg = tot + 2 * (size_t)C;
    CG_LAUNCH(k_col_totals, cdiv(C, 4), 128, 0, part, S, C, tot);
    CG_CUDA(cudaMemcpyAsync(totg, tot, sizeof(double) * 2 * C, cudaMemcpyDeviceToDevice, ctx().stream));
    CG_TRY(dist_allreduce_sum_f64(totg, 2L * C));
    CG_LAUNCH(k_bn_bwd_final_sync, cdiv(C, 128), 128, 0, tot, totg, C, (double)M * ctx().world, mg, mgx, ggamma_acc, gbeta_acc);
  } else
    CG_LAUNCH(k_bn_bwd_final, cdiv(C, 4), 128, 0, part, S, C, (double)M, mg, mgx, ggamma_acc, gbeta_acc);
  if (gx) { long n = M * C; CG_LAUNCH(k_bn_bwd_apply, grid1d(n, 256, 4), 256, 0, x, gy, gamma, mean, invstd, mg, mgx, gx, n, C); }
  return CG_OK;
}
__global__ void k_colsum_final(const double* __restrict__ part, int S, int C, float* __restrict__ out_acc) {
  int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s0, s1; col_partials(part, S, C, c, lane, s0, s1);
  if (lane) return;
  out_acc[c] += (float)s0;
}
int colsum_acc(const float* x, float* out_acc, long M, int C) {
  int S = colreduce_splits(M, C);
  long rps = (M + S - 1) / S;
  double* part = (double*)workspace(sizeof(double) * 2 * (size_t)S * C);
  if (!part) return CG_ERR_CUDA;
  dim3 g(cdiv(C, 32), S), b(32, 8);
  CG_LAUNCH(k_colreduce<2>, g, b, 0, x, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, part, M, C, rps);
  CG_LAUNCH(k_colsum_final, cdiv(C, 4), 128, 0, part, S, C, out_acc);
  return CG_OK;
}

// colsum_acc fused with the tensor-core engine's gradient-operand statistics: one pass over gy yields the bias gradient
// (same partials, same order as colsum_acc) AND max|gy|; the finalize kernel turns the maximum into the power-of-two scale
// [scale, 1/scale] the fp16 gradient operand is packed with (conv_tc.cu) and re-zeroes the maximum for the next layer.
__global__ void k_colreduce_absmax(const float* __restrict__ x, double* __restrict__ part, long M, int C, long rows_per_split, unsigned int* __restrict__ amax) {
  __shared__ double sh[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  long r0 = (long)blockIdx.y * rows_per_split, r1 = r0 + rows_per_split; if (r1 > M) r1 = M;
  double a0 = 0; float mx = 0.f;
  if (c < C) {
    float f0 = 0; int cnt = 0;
    for (long r = r0 + threadIdx.y; r < r1; r += 8) {
      float v = x[r * C + c];
      f0 += v; mx = fmaxf(mx, fabsf(v));
      if (++cnt == 64) { a0 += f0; f0 = 0; cnt = 0; }
    }
    a0 += f0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (threadIdx.x == 0 && mx > 0.f) atomicMax(amax, __float_as_uint(mx));   // non-negative floats order like uints
  sh[threadIdx.y][threadIdx.x] = a0;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double s0 = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s0 += sh[i][threadIdx.x];
    part[((long)blockIdx.y * C + c) * 2 + 0] = s0;
    part[((long)blockIdx.y * C + c) * 2 + 1] = 0.0;
  }
}
__global__ void k_colsum_final_scale(const double* __restrict__ part, int S, int C, float* __restrict__ out_acc, unsigned int* __restrict__ amax, float* __restrict__ scale2) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float m = __uint_as_float(*amax);
    float sc = 1.f;
    if (m > 0.f && isfinite(m)) sc = exp2f(floorf(log2f(16384.f / m)));
    if (!(sc > 0.f) || !isfinite(sc)) sc = 1.f;
    scale2[0] = sc; scale2[1] = 1.f / sc;
    *amax = 0u;
  }
  int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s0, s1; col_partials(part, S, C, c, lane, s0, s1);
  if (lane) return;
  out_acc[c] += (float)s0;
}
int colsum_acc_absmax(const float* x, float* out_acc, long M, int C, unsigned int* amax_zeroed, float* scale2) {
  int S = colreduce_splits(M, C);
  long rps = (M + S - 1) / S;
  double* part = (double*)workspace(sizeof(double) * 2 * (size_t)S * C);
  if (!part) return CG_ERR_CUDA;
  dim3 g(cdiv(C, 32), S), b(32, 8);
  CG_LAUNCH(k_colreduce_absmax, g, b, 0, x, part, M, C, rps, amax_zeroed);
  CG_LAUNCH(k_colsum_final_scale, cdiv(C, 4), 128, 0, part, S, C, out_acc, amax_zeroed, scale2);
  return CG_OK;
}

// =================================================================== spatial transformer (A.11, [upstream] stn)
// nn.AffineTransformMatrixGenerator: I * R(alpha) * S(s) * T(tx,ty), first two rows; R = [[c,-s],[s,c]]
__device__ __forceinline__ void mat3mul(const float* a, const float* b, float* c) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}
__device__ __forceinline__ void atm_factors(const float* th, int rot, int scl, int trn, float* R, float* S, float* T, int* idx) {
  int p = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) { R[i] = S[i] = T[i] = (i % 4 == 0) ? 1.f : 0.f; }
  idx[0] = idx[1] = idx[2] = -1;
  if (rot) { float a = th[p]; idx[0] = p++; float cs = cosf(a), sn = sinf(a); R[0] = cs; R[1] = -sn; R[3] = sn; R[4] = cs; }
  if (scl) { float s = th[p]; idx[1] = p++; S[0] = s; S[4] = s; }
  if (trn) { idx[2] = p; T[2] = th[p]; T[5] = th[p + 1]; }
}
__global__ void k_atm_fwd(const float* __restrict__ theta, float* __restrict__ A, int B, int rot, int scl, int trn, int nth) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float R[9], S[9], T[9], RS[9], M[9]; int idx[3];
  atm_factors(theta + (long)b * nth, rot, scl, trn, R, S, T, idx);
  mat3mul(R, S, RS); mat3mul(RS, T, M);
#pragma unroll
  for (int i = 0; i < 6; ++i) A[(long)b * 6 + i] = M[i];
}
int affine_matrix_fwd(const float* theta, float* A, int B, int rot, int scl, int trn) {
  int nth = (rot ? 1 : 0) + (scl ? 1 : 0) + (trn ? 2 : 0);
  CG_LAUNCH(k_atm_fwd, cdiv(B, 128), 128, 0, theta, A, B, rot, scl, trn, nth); return CG_OK;
}
__global__ void k_atm_bwd(const float* __restrict__ theta, const float* __restrict__ gA, float* __restrict__ gth, int B, int rot, int scl, int trn, int nth) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* th = theta + (long)b * nth;
  float R[9], S[9], T[9], tmp[9], M[9]; int idx[3];
  atm_factors(th, rot, scl, trn, R, S, T, idx);
  const float* G = gA + (long)b * 6;
  float* gt = gth + (long)b * nth;
  if (rot) {
    float a = th[idx[0]]; float cs = cosf(a), sn = sinf(a);
    float dR[9] = {-sn, -cs, 0, cs, -sn, 0, 0, 0, 0};
    mat3mul(dR, S, tmp); mat3mul(tmp, T, M);
    float s = 0; for (int i = 0; i < 6; ++i) s += G[i] * M[i];
    gt[idx[0]] = s;
  }
  if (scl) {
    float dS[9] = {1, 0, 0, 0, 1, 0, 0, 0, 0};
    mat3mul(R, dS, tmp); mat3mul(tmp, T, M);
    float s = 0; for (int i = 0; i < 6; ++i) s += G[i] * M[i];
    gt[idx[1]] = s;
  }
  if (trn) {
    float RS[9]; mat3mul(R, S, RS);
    for (int q = 0; q < 2; ++q) {
      float dT[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; dT[q == 0 ? 2 : 5] = 1;
      mat3mul(RS, dT, M);
      float s = 0; for (int i = 0; i < 6; ++i) s += G[i] * M[i];
      gt[idx[2] + q] = s;
    }
  }
}
int affine_matrix_bwd(const float* theta, const float* gA, float* gtheta, int B, int rot, int scl, int trn) {
  int nth = (rot ? 1 : 0) + (scl ? 1 : 0) + (trn ? 2 : 0);
  CG_LAUNCH(k_atm_bwd, cdiv(B, 128), 128, 0, theta, gA, gtheta, B, rot, scl, trn, nth); return CG_OK;
}
// nn.AffineGridGeneratorBHWD: grid[b,i,j,:] = A[b] * (y_i, x_j, 1); channel 0 = y, channel 1 = x
__global__ void k_grid_fwd(const float* __restrict__ A, float* __restrict__ grid, long n, int H, int W) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int j = (int)(i % W); long p = i / W; int ii = (int)(p % H); long b = p / H;
    const float* a = A + b * 6;
    float yb = -1.f + 2.f * ii / (H - 1), xb = -1.f + 2.f * j / (W - 1);
    grid[i * 2 + 0] = a[0] * yb + a[1] * xb + a[2];
    grid[i * 2 + 1] = a[3] * yb + a[4] * xb + a[5];
  }
}
int affine_grid_fwd(const float* A, float* grid, int B, int H, int W) {
  long n = (long)B * H * W; CG_LAUNCH(k_grid_fwd, grid1d(n, 256), 256, 0, A, grid, n, H, W); return CG_OK;
}
// gA[b] = sum_pixels ggrid^T * base : one block per image, fixed-order block reduction (deterministic)
__global__ void k_grid_bwd(const float* __restrict__ gg, float* __restrict__ gA, int H, int W) {
  int b = blockIdx.x;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    int ii = i / W, j = i % W;
    double yb = -1.0 + 2.0 * ii / (H - 1), xb = -1.0 + 2.0 * j / (W - 1);
    const float* g = gg + ((long)b * H * W + i) * 2;
    s[0] += g[0] * yb; s[1] += g[0] * xb; s[2] += g[0];
    s[3] += g[1] * yb; s[4] += g[1] * xb; s[5] += g[1];
  }
  for (int q = 0; q < 6; ++q) {
    double r = block_sum_d(s[q]);
    if (threadIdx.x == 0) gA[(long)b * 6 + q] = (float)r;
  }
}
int affine_grid_bwd(const float* ggrid, float* gA, int B, int H, int W) { CG_LAUNCH(k_grid_bwd, B, 256, 0, ggrid, gA, H, W); return CG_OK; }
// nn.BilinearSamplerBHWD forward: one warp per output pixel, lanes stride over channels (coalesced NHWC)
__global__ void k_bil_fwd(const float* __restrict__ img, const float* __restrict__ grid, float* __restrict__ out, long npix, int H, int W, int C) {
  long pix = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
  if (pix >= npix) return;
  long b = pix / ((long)H * W);
  float gy_ = grid[pix * 2], gx_ = grid[pix * 2 + 1];
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
int bilinear_fwd(const float* img, const float* grid, float* out, int B, int H, int W, int C) {
  long npix = (long)B * H * W; CG_LAUNCH(k_bil_fwd, cdiv(npix * 32, 256), 256, 0, img, grid, out, npix, H, W, C); return CG_OK;
}
// backward: scatter-add into gimg with fp32 atomics (summation order not fixed: results are reproducible only
// to fp32 rounding -- the reference pinned this op to the CPU for that reason, models.lua:889-893); ggrid from
// corner dot products reduced across the warp.
__global__ void k_bil_bwd(const float* __restrict__ img, const float* __restrict__ grid, const float* __restrict__ gout,
                          float* __restrict__ gimg, float* __restrict__ ggrid, long npix, int H, int W, int C) {
  long pix = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
  if (pix >= npix) return;
  long b = pix / ((long)H * W);
  float gy_ = grid[pix * 2], gx_ = grid[pix * 2 + 1];
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
int bilinear_bwd(const float* img, const float* grid, const float* gout, float* gimg, float* ggrid, int B, int H, int W, int C) {
  long npix = (long)B * H * W;
  CG_CUDA(cudaMemsetAsync(gimg, 0, sizeof(float) * (size_t)npix * C, ctx().stream));
  CG_LAUNCH(k_bil_bwd, cdiv(npix * 32, 256), 256, 0, img, grid, gout, gimg, ggrid, npix, H, W, C); return CG_OK;
}

// =================================================================== criterion / optimiser / rng
// nn.BCECriterion (A.7): eps = 1e-12 inside the logs, mean over n; single block (n = batch size)
__global__ void k_bce(const float* __restrict__ p, const float* __restrict__ t, int n, float* __restrict__ loss, float* __restrict__ g) {
  const double eps = 1e-12; double s = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double pi = p[i], ti = t[i];
    s += ti * log(pi + eps) + (1.0 - ti) * log(1.0 - pi + eps);
    if (g) g[i] = (float)(-(1.0 / n) * (ti - pi) / ((1.0 - pi + eps) * (pi + eps)));
  }
  s = block_sum_d(s);
  if (threadIdx.x == 0 && loss) *loss = (float)(-s / n);
}
int bce(const float* p, const float* t, int n, float* loss_dev, float* g) { CG_LAUNCH(k_bce, 1, 256, 0, p, t, n, loss_dev, g); return CG_OK; }
// adversarial.lua:92-98,110-112 / :201-212
__global__ void k_penalty_clamp(float* __restrict__ g, const float* __restrict__ p, long n, float l1sign, float l2, float clampv,
                                int want_norms, double* __restrict__ part) {
  double n1 = 0, n2 = 0;
  bool pen = (l1sign != 0.f) || (l2 != 0.f);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float pv = p[i], gv = g[i];
    if (want_norms) { n1 += fabs((double)pv); n2 += (double)pv * pv; }
    if (pen) gv += (pv > 0.f ? 1.f : (pv < 0.f ? -1.f : 0.f)) * l1sign + pv * l2;
    if (clampv != 0.f) gv = fminf(fmaxf(gv, -clampv), clampv);
    g[i] = gv;
  }
  if (want_norms) {
    n1 = block_sum_d(n1); n2 = block_sum_d(n2);
    if (threadIdx.x == 0) { part[blockIdx.x * 2] = n1; part[blockIdx.x * 2 + 1] = n2; }
  }
}
__global__ void k_penalty_final(const double* __restrict__ part, int nparts, float l1, float l2, float* out) {
  double a = 0, b = 0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) { a += part[i * 2]; b += part[i * 2 + 1]; }
  a = block_sum_d(a); b = block_sum_d(b);
  if (threadIdx.x == 0) *out = (float)(l1 * a + l2 * b / 2.0);
}
int penalty_clamp(float* g, const float* p, long n, float l1, float l1sign, float l2, float clampv, float* loss_add_dev) {
  bool pen = (l1 != 0.f) || (l2 != 0.f);
  if (!pen) { l1sign = 0.f; l2 = 0.f; }   // the Lua `if OPT.X_L1 ~= 0 or OPT.X_L2 ~= 0` guard
  int gsz = grid1d(n, 256, 4);
  int want = (loss_add_dev && pen) ? 1 : 0;
  double* part = (double*)workspace(sizeof(double) * 2 * gsz);
  if (!part) return CG_ERR_CUDA;
  CG_LAUNCH(k_penalty_clamp, gsz, 256, 0, g, p, n, l1sign, l2, clampv, want, part);
  if (loss_add_dev) {
    if (want) CG_LAUNCH(k_penalty_final, 1, 256, 0, part, gsz, l1, l2, loss_add_dev);
    else CG_CUDA(cudaMemsetAsync(loss_add_dev, 0, sizeof(float), ctx().stream));
  }
  return CG_OK;
}
// optim.adam (A.8): eps added to sqrt(v) BEFORE the bias correction is applied through stepSize
// Step count and RNG offset live in DEVICE memory so that a captured CUDA graph of the training step replays correctly
// (a host value passed as a kernel argument would be frozen at capture time).
__global__ void k_inc_i32(int* p) { *p += 1; }
__global__ void k_add_u64(unsigned long long* p, unsigned long long v) { *p += v; }
__global__ void k_adam(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
                       const int* __restrict__ t_dev, float lr, float b1, float b2, float eps) {
  __shared__ float step_s;
  if (threadIdx.x == 0) {   // optim.adam: stepSize = lr * sqrt(1 - beta2^t) / (1 - beta1^t), t already incremented
    int t = *t_dev;
    step_s = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
  }
  __syncthreads();
  const float step = step_s;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float gv = g[i];
    float mv = b1 * m[i] + (1.f - b1) * gv;
    float vv = b2 * v[i] + (1.f - b2) * gv * gv;
    m[i] = mv; v[i] = vv;
    x[i] -= step * mv / (sqrtf(vv) + eps);
  }
}
// t_dev: device step counter of this parameter vector; it is incremented here (t <- t + 1) and then used
int adam(float* x, const float* g, float* m, float* v, long n, int* t_dev, float lr, float b1, float b2, float eps) {
  CG_LAUNCH(k_inc_i32, 1, 1, 0, t_dev);
  CG_LAUNCH(k_adam, grid1d(n, 256, 4), 256, 0, x, g, m, v, n, (const int*)t_dev, lr, b1, b2, eps); return CG_OK;
}
// penalty + clamp + Adam fused (see ops.cuh).  The norms are taken from p BEFORE the update, as adversarial.lua:92-98 does.
__global__ void k_penalty_clamp_adam(float* __restrict__ g, float* __restrict__ x, float* __restrict__ m, float* __restrict__ v, long n, float gscale,
                                     float l1sign, float l2, float clampv, int want_norms, double* __restrict__ part,
                                     const int* __restrict__ t_dev, float lr, float b1, float b2, float eps) {
  __shared__ float step_s;
  if (threadIdx.x == 0) {
    int t = *t_dev;
    step_s = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
  }
  __syncthreads();
  const float step = step_s;
  const bool pen = (l1sign != 0.f) || (l2 != 0.f);
  double n1 = 0, n2 = 0;
  auto one = [&](float pv, float gv, float mi, float vi, float& go, float& mo, float& vo, float& xo) {
    if (gscale != 1.f) gv *= gscale;
    if (pen) gv += (pv > 0.f ? 1.f : (pv < 0.f ? -1.f : 0.f)) * l1sign + pv * l2;
    if (clampv != 0.f) gv = fminf(fmaxf(gv, -clampv), clampv);
    go = gv;
    mo = b1 * mi + (1.f - b1) * gv;
    vo = b2 * vi + (1.f - b2) * gv * gv;
    xo = pv - step * mo / (sqrtf(vo) + eps);
  };
  // 16-byte accesses over the aligned body (the four vectors come from cudaMalloc), scalar tail; the norms: fp32 over a thread's four
  // elements, double across iterations (four 8-byte conversions per element made this kernel 2.8 TB/s)
  const bool al16 = ((((uintptr_t)x) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0;
  const long n4 = al16 ? (n >> 2) : 0;
  float4* x4 = reinterpret_cast<float4*>(x); float4* g4 = reinterpret_cast<float4*>(g); float4* m4 = reinterpret_cast<float4*>(m); float4* v4 = reinterpret_cast<float4*>(v);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 pv = x4[i], gv = g4[i], mi = m4[i], vi = v4[i];
    float4 go, mo, vo, xo;
    one(pv.x, gv.x, mi.x, vi.x, go.x, mo.x, vo.x, xo.x); one(pv.y, gv.y, mi.y, vi.y, go.y, mo.y, vo.y, xo.y);
    one(pv.z, gv.z, mi.z, vi.z, go.z, mo.z, vo.z, xo.z); one(pv.w, gv.w, mi.w, vi.w, go.w, mo.w, vo.w, xo.w);
    if (want_norms) { n1 += (double)((fabsf(pv.x) + fabsf(pv.y)) + (fabsf(pv.z) + fabsf(pv.w))); n2 += (double)((pv.x * pv.x + pv.y * pv.y) + (pv.z * pv.z + pv.w * pv.w)); }
    g4[i] = go; m4[i] = mo; v4[i] = vo; x4[i] = xo;
  }
  for (long i = (n4 << 2) + blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float pv = x[i];
    float go, mo, vo, xo; one(pv, g[i], m[i], v[i], go, mo, vo, xo);
    if (want_norms) { n1 += fabs((double)pv); n2 += (double)pv * pv; }
    g[i] = go; m[i] = mo; v[i] = vo; x[i] = xo;
  }
  if (want_norms) {
    n1 = block_sum_d(n1); n2 = block_sum_d(n2);
    if (threadIdx.x == 0) { part[blockIdx.x * 2] = n1; part[blockIdx.x * 2 + 1] = n2; }
  }
}
int penalty_clamp_adam(float* g, float* p, float* m, float* v, long n, float gscale, float l1, float l1sign, float l2, float clampv, float* loss_add_dev,
                       int* t_dev, float lr, float b1, float b2, float eps) {
  bool pen = (l1 != 0.f) || (l2 != 0.f);
  if (!pen) { l1sign = 0.f; l2 = 0.f; }
  int gsz = grid1d(n, 256, 4);
  int want = (loss_add_dev && pen) ? 1 : 0;
  double* part = (double*)workspace(sizeof(double) * 2 * gsz);
  if (!part) return CG_ERR_CUDA;
  CG_LAUNCH(k_inc_i32, 1, 1, 0, t_dev);
  ctx().next_bytes = 32.0 * (double)n;   // read g, p, m, v ; write g, p, m, v
  CG_LAUNCH(k_penalty_clamp_adam, gsz, 256, 0, g, p, m, v, n, gscale, l1sign, l2, clampv, want, part, (const int*)t_dev, lr, b1, b2, eps);
  if (loss_add_dev) {
    if (want) CG_LAUNCH(k_penalty_final, 1, 256, 0, part, gsz, l1, l2, loss_add_dev);
    else CG_CUDA(cudaMemsetAsync(loss_add_dev, 0, sizeof(float), ctx().stream));
  }
  return CG_OK;
}
__global__ void k_uniform(float* __restrict__ dst, long n, float lo, float hi, uint32_t k0, uint32_t k1, uint64_t offset) {
  long nq = (n + 3) / 4;
  for (long q = blockIdx.x * (long)blockDim.x + threadIdx.x; q < nq; q += (long)gridDim.x * blockDim.x) {
    uint64_t ctr = offset + (uint64_t)q;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    philox4x32(c, k0, k1);
#pragma unroll
    for (int j = 0; j < 4; ++j) { long i = q * 4 + j; if (i < n) dst[i] = lo + (hi - lo) * u01(c[j]); }
  }
}
int uniform(float* dst, long n, float lo, float hi, uint64_t seed, uint64_t offset) {
  CG_LAUNCH(k_uniform, grid1d((n + 3) / 4, 256), 256, 0, dst, n, lo, hi, (uint32_t)seed, (uint32_t)(seed >> 32), offset); return CG_OK;
}
__global__ void k_bernoulli(float* __restrict__ dst, long n, float p_drop, float keep, uint32_t k0, uint32_t k1,
                            const unsigned long long* __restrict__ offset_dev, unsigned long long rel) {
  const uint64_t offset = *offset_dev + rel;
  long nq = (n + 3) / 4;
  for (long q = blockIdx.x * (long)blockDim.x + threadIdx.x; q < nq; q += (long)gridDim.x * blockDim.x) {
    uint64_t ctr = offset + (uint64_t)q;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 1u, 0u};
    philox4x32(c, k0, k1);
#pragma unroll
    for (int j = 0; j < 4; ++j) { long i = q * 4 + j; if (i < n) dst[i] = u01(c[j]) >= p_drop ? keep : 0.f; }
  }
}
// offset = *offset_dev + rel (Philox counter); the caller advances *offset_dev with rng_advance after a group of masks
int bernoulli_mask(float* dst, long n, float p_drop, float keep_value, uint64_t seed, const unsigned long long* offset_dev, uint64_t rel) {
  CG_LAUNCH(k_bernoulli, grid1d((n + 3) / 4, 256), 256, 0, dst, n, p_drop, keep_value, (uint32_t)seed, (uint32_t)(seed >> 32), offset_dev, (unsigned long long)rel); return CG_OK;
}
int rng_advance(unsigned long long* offset_dev, uint64_t by) { CG_LAUNCH(k_add_u64, 1, 1, 0, offset_dev, (unsigned long long)by); return CG_OK; }

}  // namespace cg
