// conv_ref.cu -- generic fp32 CUDA-core implicit-GEMM convolution over NHWC (stride 1, pad (k-1)/2) with
// deterministic split-K, used for every shape the tcgen05 engine does not take (Cin/Cout of 1, 3, 16; Linear
// layers run through it as 1x1 convolutions) and as the second party in the tensor-core parity tests.
// Semantics: SURVEY.md A.1 / A.2.  Reference call sites: models.lua:145-154,199-222,646-700,844-854.
#include "ops.cuh"

namespace cg {

// ------------------------------------------------------------------ parameter packing
__global__ void k_pack_fprop(const float* __restrict__ W, float* __restrict__ Wp, long n, ConvSpec s) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int co = (int)(i % s.Co); long r = i / s.Co; int ci = (int)(r % s.Ci); int tap = (int)(r / s.Ci);
    Wp[i] = W[torch_index(s.k, s.Ci, s.Co, s.in_hw, s.out_hw, tap / s.k, tap % s.k, ci, co)];
  }
}
int pack_fprop(const float* W, float* Wp, const ConvSpec& s) {
  long n = (long)s.k * s.k * s.Ci * s.Co; CG_LAUNCH(k_pack_fprop, grid1d(n, 256), 256, 0, W, Wp, n, s); return CG_OK;
}
// dgrad as a forward conv of gy: Wd[(ky,kx,co)][ci] = W[co][ci][k-1-ky][k-1-kx]
__global__ void k_pack_dgrad(const float* __restrict__ W, float* __restrict__ Wd, long n, ConvSpec s) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int ci = (int)(i % s.Ci); long r = i / s.Ci; int co = (int)(r % s.Co); int tap = (int)(r / s.Co);
    int ky = s.k - 1 - tap / s.k, kx = s.k - 1 - tap % s.k;
    Wd[i] = W[torch_index(s.k, s.Ci, s.Co, s.in_hw, s.out_hw, ky, kx, ci, co)];
  }
}
int pack_dgrad(const float* W, float* Wd, const ConvSpec& s) {
  long n = (long)s.k * s.k * s.Ci * s.Co; CG_LAUNCH(k_pack_dgrad, grid1d(n, 256), 256, 0, W, Wd, n, s); return CG_OK;
}
__global__ void k_pack_bias(const float* __restrict__ b, float* __restrict__ bp, ConvSpec s, int dir, int acc) {
  int cop = blockIdx.x * blockDim.x + threadIdx.x;
  if (cop >= s.Co) return;
  int Cov = s.Co / s.out_hw;
  int fo = s.out_hw == 1 ? cop : (cop % Cov) * s.out_hw + cop / Cov;
  if (dir == 0) bp[cop] = b[fo];
  else { if (acc) bp[fo] += b[cop]; else bp[fo] = b[cop]; }
}
int pack_bias(const float* b, float* bp, const ConvSpec& s) { CG_LAUNCH(k_pack_bias, cdiv(s.Co, 128), 128, 0, b, bp, s, 0, 0); return CG_OK; }
int unpack_bias_acc(const float* gbp, float* gb_acc, const ConvSpec& s) { CG_LAUNCH(k_pack_bias, cdiv(s.Co, 128), 128, 0, gbp, gb_acc, s, 1, 1); return CG_OK; }
__global__ void k_unpack_wgrad(const float* __restrict__ gWp, float* __restrict__ gW, long n, ConvSpec s) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int co = (int)(i % s.Co); long r = i / s.Co; int ci = (int)(r % s.Ci); int tap = (int)(r / s.Ci);
    gW[torch_index(s.k, s.Ci, s.Co, s.in_hw, s.out_hw, tap / s.k, tap % s.k, ci, co)] += gWp[i];   // accGradParameters adds
  }
}
// nn.Linear beside an nn.View (k = 1): gW[fo][fi] += gWp[cip][cop] with both axes permuted (torch_index).  A 32 x 32 tile through shared
// memory: rows = the packed input features that are 32 CONSECUTIVE Torch features (so every output row is written as one 128-byte run),
// columns = 32 consecutive packed outputs (read as 128-byte runs).  The element-per-thread kernel above wrote with a stride of Ci floats
// (D's Linear 20480 -> 256: 114 us for 21 MB, profiles/r02_timeline.txt).
__global__ void __launch_bounds__(256) k_unpack_linear_view(const float* __restrict__ gWp, float* __restrict__ gW, ConvSpec s) {
  __shared__ float t[32][33];
  const int Civ = s.Ci / s.in_hw, Cov = s.Co / s.out_hw;
  const int fi0 = blockIdx.x * 32, cop0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int fi = fi0 + r, cop = cop0 + tx;
    float v = 0.f;
    if (fi < s.Ci && cop < s.Co) { const int cip = (fi % s.in_hw) * Civ + fi / s.in_hw; v = gWp[(long)cip * s.Co + cop]; }
    t[r][tx] = v;
  }
  __syncthreads();
  for (int c = ty; c < 32; c += 8) {
    const int cop = cop0 + c, fi = fi0 + tx;
    if (cop < s.Co && fi < s.Ci) { const int fo = (cop % Cov) * s.out_hw + cop / Cov; gW[(long)fo * s.Ci + fi] += t[tx][c]; }
  }
}
int unpack_wgrad_acc(const float* gWp, float* gW_acc, const ConvSpec& s) {
  long n = (long)s.k * s.k * s.Ci * s.Co;
  if (s.k == 1 && !(s.in_hw == 1 && s.out_hw == 1)) {
    ctx().next_bytes = 12.0 * (double)n;
    CG_LAUNCH(k_unpack_linear_view, dim3(cdiv(s.Ci, 32), cdiv(s.Co, 32)), 256, 0, gWp, gW_acc, s); return CG_OK;
  }
  CG_LAUNCH(k_unpack_wgrad, grid1d(n, 256), 256, 0, gWp, gW_acc, n, s); return CG_OK;
}

// ------------------------------------------------------------------ split-K reduction (fixed order)
__global__ void k_splitk_reduce(const float* __restrict__ part, int S, long n, int ncol, const float* __restrict__ bias, float* __restrict__ out, unsigned int* __restrict__ amax) {
  float mx = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float s = bias ? bias[i % ncol] : 0.f;
    for (int z = 0; z < S; ++z) s += part[(long)z * n + i];
    out[i] = s; mx = fmaxf(mx, fabsf(s));
  }
  if (amax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(amax, __float_as_uint(mx));
  }
}

int splitk_reduce(const float* part, int S, long n, int ncol, const float* bias, float* out) {
  ctx().next_bytes = 4.0 * (double)n * (S + 1);
  unsigned int* amax = ctx().next_amax; ctx().next_amax = nullptr;
  CG_LAUNCH(k_splitk_reduce, grid1d(n, 256, 2), 256, 0, part, S, n, ncol, bias, out, amax); return CG_OK;
}

// ------------------------------------------------------------------ partial sums -> Torch-layout gradient, fused
// gW[co][ci][ky][kx] += sum_z part[z][(tap,ci)][co].  A block takes all taps of CIB input channels x 32 output channels: the packed
// partials are read coalesced along co and summed in fixed z order (deterministic), transposed through shared memory, and each
// output channel's [ci][tap] run is written contiguously.  Replaces k_sum_parts / k_splitk_reduce + k_unpack_wgrad, whose
// strided read-modify-write (stride k*k floats) was 1.5 ms of a 10 ms step (profiles/r01_bench_tc_engine.json).
__global__ void __launch_bounds__(256) k_parts_to_torch_acc(const float* __restrict__ part, int Z, long zstride, float* __restrict__ gW,
                                                            int Ci, int Co, int kk, int CIB) {
  __shared__ float t[6600];                       // kk * CIB * 33 <= 25*8*33 = 6600 (k = 7 uses CIB = 4: 49*4*33 = 6468)
  const int ci0 = blockIdx.x * CIB, co0 = blockIdx.y * 32, tid = threadIdx.x;
  const int nload = kk * CIB * 32;
  for (int i = tid; i < nload; i += 256) {
    int col = i & 31, r = i >> 5, cil = r % CIB, tap = r / CIB;
    int ci = ci0 + cil, co = co0 + col;
    float sum = 0.f;
    if (ci < Ci && co < Co) {
      const float* src = part + ((long)tap * Ci + ci) * Co + co;
      // four independent chains keep four loads in flight per thread (a single chain walked Z loads megabytes apart serially:
      // 5.2 ms in round 1); the association ((s0+s1)+(s2+s3)) is fixed, so the result is deterministic
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int z = 0;
      for (; z + 4 <= Z; z += 4) {
        s0 += src[(long)z * zstride]; s1 += src[(long)(z + 1) * zstride]; s2 += src[(long)(z + 2) * zstride]; s3 += src[(long)(z + 3) * zstride];
      }
      for (; z < Z; ++z) s0 += src[(long)z * zstride];
      sum = (s0 + s1) + (s2 + s3);
    }
    t[(tap * CIB + cil) * 33 + col] = sum;
  }
  __syncthreads();
  const int run = CIB * kk;                        // contiguous floats per output channel in the Torch layout
  for (int i = tid; i < 32 * run; i += 256) {
    int col = i / run, r = i - col * run, cil = r / kk, tap = r - cil * kk;
    int ci = ci0 + cil, co = co0 + col;
    if (ci < Ci && co < Co) gW[((long)co * Ci + ci) * kk + tap] += t[(tap * CIB + cil) * 33 + col];   // accGradParameters adds
  }
}
// Z > 1: the split partials are summed here as well (round 1 needed a separate k_sum_parts pass first because a single dependent
// chain per thread made this kernel 5.2 ms; see the loop above).
int parts_to_torch_acc(const float* part, int Z, long zstride, float* gW_acc, int Ci, int Co, int kk) {
  int CIB = kk <= 9 ? 8 : (kk <= 25 ? 4 : 2);     // smaller channel blocks for big filters: more blocks, runs of >= 98 floats
  if (kk * CIB * 33 > 6600) return CG_ERR_UNSUPPORTED;
  dim3 g(cdiv(Ci, CIB), cdiv(Co, 32));
  CG_LAUNCH(k_parts_to_torch_acc, g, 256, 0, part, Z, zstride, gW_acc, Ci, Co, kk, CIB);
  return CG_OK;
}

// ------------------------------------------------------------------ forward / dgrad: C[M=pixels, Co] = A[M, K=(tap,ci)] * Wp[K, Co]
constexpr int BM = 64, BN = 64, BK = 16;
template <bool VEC>
__global__ void __launch_bounds__(256) k_conv_fwd(const float* __restrict__ x, const float* __restrict__ Wp, const float* __restrict__ bias,
                                                  float* __restrict__ y, long M, int H, int W, int Ci, int Co, int k, int Ktot, int Kper, int S) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  int tid = threadIdx.x;
  long m0 = (long)blockIdx.x * BM; int n0 = blockIdx.y * BN;
  int kbeg = blockIdx.z * Kper, kend = kbeg + Kper; if (kend > Ktot) kend = Ktot;
  int p = (k - 1) / 2;
  // A loader: row am, 4 consecutive kk starting at akq
  int am = tid >> 2, akq = (tid & 3) * 4;
  long mrow = m0 + am; bool mval = mrow < M;
  int px = 0, py = 0; long pimg = 0;
  if (mval) { px = (int)(mrow % W); long t = mrow / W; py = (int)(t % H); pimg = t / H; }
  // B loader: row bk, 4 consecutive co at bc
  int bk = tid >> 4, bc = (tid & 15) * 4;
  int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    int kk = k0 + akq;
    if (mval) {
      if (VEC) {   // Ci % 4 == 0: the 4 kk share one tap and are contiguous in memory
        if (kk < kend) {
          int tap = kk / Ci, ci = kk - tap * Ci;
          int iy = py + tap / k - p, ix = px + tap % k - p;
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
            float4 v = *reinterpret_cast<const float4*>(x + ((pimg * H + iy) * W + ix) * Ci + ci);
            av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int kj = kk + j;
          if (kj < kend) {
            int tap = kj / Ci, ci = kj - tap * Ci;
            int iy = py + tap / k - p, ix = px + tap % k - p;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) av[j] = x[((pimg * H + iy) * W + ix) * Ci + ci];
          }
        }
      }
    }
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    int kb = k0 + bk;
    if (kb < kend) {
      const float* wr = Wp + (long)kb * Co + n0 + bc;
      if (VEC && n0 + bc + 3 < Co) { float4 v = *reinterpret_cast<const float4*>(wr); bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w; }
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (n0 + bc + j < Co) bv[j] = wr[j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) As[akq + j][am] = av[j];
    *reinterpret_cast<float4*>(&Bs[bk][bc]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < BK; ++q) {
      float4 a = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[q][tx * 4]);
      float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  float* out = S > 1 ? y + (long)blockIdx.z * M * Co : y;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = n0 + tx * 4 + j;
      if (c < Co) out[m * Co + c] = acc[i][j] + ((S == 1 && bias) ? bias[c] : 0.f);
    }
  }
}

static int conv_fwd_ref(const float* x, const float* Wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co, int k) {
  long M = (long)N * H * W; int Ktot = k * k * Ci;
  int tiles = cdiv(M, BM) * cdiv(Co, BN);
  int S = 1;
  int target = ctx().sm_count * 2;
  if (tiles < target) { S = target / tiles; int maxs = Ktot / (4 * BK); if (S > maxs) S = maxs; if (S < 1) S = 1; if (S > 64) S = 64; }
  int Kper = cdiv(cdiv(Ktot, S), BK) * BK; S = cdiv(Ktot, Kper);
  dim3 g(cdiv(M, BM), cdiv(Co, BN), S);
  bool vec = (Ci % 4 == 0) && (Co % 4 == 0) && (((uintptr_t)x & 15) == 0) && (((uintptr_t)Wp & 15) == 0);
  float* dst = y;
  if (S > 1) { dst = (float*)workspace(sizeof(float) * (size_t)S * M * Co); if (!dst) return CG_ERR_CUDA; }
  double fl = 2.0 * (double)M * Co * Ktot, by = 4.0 * ((double)M * Ci + (double)Ktot * Co + (double)M * Co);
  ctx().next_flops = fl; ctx().next_bytes = by;
  if (vec) CG_LAUNCH(k_conv_fwd<true>, g, 256, 0, x, Wp, bias, dst, M, H, W, Ci, Co, k, Ktot, Kper, S);
  else CG_LAUNCH(k_conv_fwd<false>, g, 256, 0, x, Wp, bias, dst, M, H, W, Ci, Co, k, Ktot, Kper, S);
  if (S > 1) { long n = M * Co; CG_LAUNCH(k_splitk_reduce, grid1d(n, 256, 2), 256, 0, dst, S, n, Co, bias, y, (unsigned int*)nullptr); }
  return CG_OK;
}

// ------------------------------------------------------------------ wgrad: gWp[K=(tap,ci), Co] = sum_m A[m, K] * gy[m, Co]
template <bool VEC>
__global__ void __launch_bounds__(256) k_conv_wgrad(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ out,
                                                    long M, int H, int W, int Ci, int Co, int k, int Ktot, long Mper) {
  __shared__ __align__(16) float As[BK][BM];   // [m chunk][kk]
  __shared__ __align__(16) float Gs[BK][BN];   // [m chunk][co]
  int tid = threadIdx.x;
  int kk0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  long mbeg = (long)blockIdx.z * Mper, mend = mbeg + Mper; if (mend > M) mend = M;
  int p = (k - 1) / 2;
  int lm = tid >> 4, lq = (tid & 15) * 4;   // loaders: row lm of the chunk, 4 consecutive kk / co
  int ty = tid >> 4, tx = tid & 15;
  // per-thread tap decode for its 4 kk (fixed for the whole kernel)
  int tap4[4], ci4[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { int kj = kk0 + lq + j; tap4[j] = kj < Ktot ? kj / Ci : -1; ci4[j] = kj < Ktot ? kj % Ci : 0; }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long mc = mbeg; mc < mend; mc += BK) {
    long m = mc + lm;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < mend) {
      int px = (int)(m % W); long t = m / W; int py = (int)(t % H); long pimg = t / H;
      if (VEC) {
        if (tap4[0] >= 0) {
          int iy = py + tap4[0] / k - p, ix = px + tap4[0] % k - p;
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
            float4 v = *reinterpret_cast<const float4*>(x + ((pimg * H + iy) * W + ix) * Ci + ci4[0]);
            av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (tap4[j] >= 0) {
          int iy = py + tap4[j] / k - p, ix = px + tap4[j] % k - p;
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) av[j] = x[((pimg * H + iy) * W + ix) * Ci + ci4[j]];
        }
      }
      const float* gr = gy + m * Co + n0 + lq;
      if (VEC && n0 + lq + 3 < Co) { float4 v = *reinterpret_cast<const float4*>(gr); gv[0] = v.x; gv[1] = v.y; gv[2] = v.z; gv[3] = v.w; }
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (n0 + lq + j < Co) gv[j] = gr[j];
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lm][lq]) = make_float4(av[0], av[1], av[2], av[3]);
    *reinterpret_cast<float4*>(&Gs[lm][lq]) = make_float4(gv[0], gv[1], gv[2], gv[3]);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < BK; ++q) {
      float4 a = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Gs[q][tx * 4]);
      float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  float* o = out + (long)blockIdx.z * Ktot * Co;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int kk = kk0 + ty * 4 + i;
    if (kk >= Ktot) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) { int c = n0 + tx * 4 + j; if (c < Co) o[(long)kk * Co + c] = acc[i][j]; }
  }
}

static int conv_wgrad_ref(const float* x, const float* gy, float* gWp_out, int N, int H, int W, int Ci, int Co, int k, float* gW_acc, int* done) {
  long M = (long)N * H * W; int Ktot = k * k * Ci;
  int tiles = cdiv(Ktot, BM) * cdiv(Co, BN);
  int target = ctx().sm_count * 3;
  int S = cdiv(target, tiles); long maxs = (M + 4 * BK - 1) / (4 * BK); if (S > maxs) S = (int)maxs; if (S < 1) S = 1; if (S > 256) S = 256;
  long Mper = ((M + S - 1) / S + BK - 1) / BK * BK; S = (int)((M + Mper - 1) / Mper);
  dim3 g(cdiv(Ktot, BM), cdiv(Co, BN), S);
  bool vec = (Ci % 4 == 0) && (Co % 4 == 0) && (((uintptr_t)x & 15) == 0) && (((uintptr_t)gy & 15) == 0);
  float* dst = gWp_out;
  if (S > 1) { dst = (float*)workspace(sizeof(float) * (size_t)S * Ktot * Co); if (!dst) return CG_ERR_CUDA; }
  ctx().next_flops = 2.0 * (double)M * Co * Ktot; ctx().next_bytes = 4.0 * ((double)M * Ci + (double)M * Co + (double)Ktot * Co);
  if (vec) CG_LAUNCH(k_conv_wgrad<true>, g, 256, 0, x, gy, dst, M, H, W, Ci, Co, k, Ktot, Mper);
  else CG_LAUNCH(k_conv_wgrad<false>, g, 256, 0, x, gy, dst, M, H, W, Ci, Co, k, Ktot, Mper);
  if (S > 1) { long n = (long)Ktot * Co; CG_LAUNCH(k_splitk_reduce, grid1d(n, 256, 2), 256, 0, dst, S, n, Co, (const float*)nullptr, gWp_out, (unsigned int*)nullptr); }
  if (gW_acc && parts_to_torch_acc(gWp_out, 1, 0, gW_acc, Ci, Co, k * k) == CG_OK) { if (done) *done = 1; }
  return CG_OK;
}

// ------------------------------------------------------------------ engine dispatch
// conv_tc.cu provides these; they return CG_ERR_UNSUPPORTED for shapes the tensor-core path does not take.
int conv_fwd_tc(const float* x, const float* Wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co, int k);
int conv_wgrad_tc(const float* x, const float* gy, float* gWp_out, int N, int H, int W, int Ci, int Co, int k, float* gW_acc, int* done);
int conv_bwd_tc(const float* x, const float* gy, const float* Wd, float* gWp_out, float* gx, int N, int H, int W, int Ci, int Co, int k, float* gW_acc, int* done,
                const uint8_t* xq_prepacked, float* gb_acc, int* bias_done);
void conv_tc_set_gradient_operands(int on);   // tf32 operands for gradient-valued inputs (tools/backward_precision_study.py)

int conv_fwd(const float* x, const float* Wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co, int k) {
  if (ctx().conv_engine == 1) { int s = conv_fwd_tc(x, Wp, bias, y, N, H, W, Ci, Co, k); if (s != CG_ERR_UNSUPPORTED) return s; }
  if (ctx().fp32_operands_stale)
    return set_err(CG_ERR_STATE, "conv %dx%d %d->%d k=%d at batch %d fell back to the fp32 kernel, whose operands were not refreshed (model_repack need32)", H, W, Ci, Co, k, N);
  return conv_fwd_ref(x, Wp, bias, y, N, H, W, Ci, Co, k);
}
int conv_dgrad(const float* gy, const float* Wd, float* gx, int N, int H, int W, int Ci, int Co, int k) {
  // dgrad is a forward convolution of gy (Co channels) with the flipped/transposed weights
  conv_tc_set_gradient_operands(1);
  int s = conv_fwd(gy, Wd, nullptr, gx, N, H, W, Co, Ci, k);
  conv_tc_set_gradient_operands(0);
  return s;
}
// gW_acc (optional): Torch-layout gradient [Co][Ci][k][k] of a plain conv; when the engine could add into it directly *done = 1
// and gWp_out is left untouched, otherwise gWp_out holds the packed gradient and the caller unpacks.
int conv_wgrad(const float* x, const float* gy, float* gWp_out, int N, int H, int W, int Ci, int Co, int k, float* gW_acc, int* done) {
  if (done) *done = 0;
  if (ctx().conv_engine == 1) { int s = conv_wgrad_tc(x, gy, gWp_out, N, H, W, Ci, Co, k, gW_acc, done); if (s != CG_ERR_UNSUPPORTED) return s; }
  return conv_wgrad_ref(x, gy, gWp_out, N, H, W, Ci, Co, k, gW_acc, done);
}
// weight gradient + input gradient of one layer; the tensor-core engine packs the gradient operand once for both
// xq (optional): the forward's cached fp16 operand for x; then x may be null and only the tensor-core engine can serve the call
int conv_backward(const float* x, const float* gy, const float* Wd, float* gWp_out, float* gx, int N, int H, int W, int Ci, int Co, int k, float* gW_acc, int* done,
                  const uint8_t* xq, float* gb_acc, int* bias_done) {
  if (done) *done = 0;
  if (bias_done) *bias_done = 0;
  if (ctx().conv_engine == 1) { int s = conv_bwd_tc(x, gy, Wd, gWp_out, gx, N, H, W, Ci, Co, k, gW_acc, done, xq, gb_acc, bias_done); if (s != CG_ERR_UNSUPPORTED) return s; }
  if (!x) return set_err(CG_ERR_STATE, "this forward cached only the tensor-core operand of the layer input; backward needs the same conv engine");
  CG_TRY(conv_wgrad(x, gy, gWp_out, N, H, W, Ci, Co, k, gW_acc, done));
  return conv_dgrad(gy, Wd, gx, N, H, W, Ci, Co, k);
}

}  // namespace cg
