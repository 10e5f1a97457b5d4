/*
 * catgen_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See catgen_oracle.h.
 * PARITY UNPINNED (no reference tests / golden vectors exist).
 *
 * Algorithms follow the Torch7-era CPU path: per-sample im2col + SGEMM convolutions
 * (THNN SpatialConvolutionMM), fp32 storage; reductions (BN statistics,
 * PReLU slope gradient, BCE) accumulate in double so the oracle is the most accurate party.
 */
#include "catgen_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 0;
void og_set_threads(int n) {
  g_threads = n;
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}
int og_get_threads(void) {
#ifdef _OPENMP
  return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
  return 1;
#endif
}

/* threads for a region with `work` independent items: never more threads than items (on a 128-core host the
   first version forked 128 threads for 8 samples and every one of them allocated full-size scratch) */
static int team(long work) {
  int t = og_get_threads();
  if (work < 1) work = 1;
  return (long)t < work ? t : (int)work;
}
#define OG_SMALL 32768   /* pointwise loops shorter than this run serially */

/* ------------------------------------------------------------------ GEMM helpers (serial) */
/* C[M,N] += A[M,K] * B[K,N]  (row major) */
static void gemm_nn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc) {
  const int KB = 128, NB = 512;
  for (int k0 = 0; k0 < K; k0 += KB) {
    int k1 = k0 + KB < K ? k0 + KB : K;
    for (int j0 = 0; j0 < N; j0 += NB) {
      int j1 = j0 + NB < N ? j0 + NB : N;
      int i = 0;
      for (; i + 4 <= M; i += 4) {
        float* c0 = C + (long)(i + 0) * ldc; float* c1 = C + (long)(i + 1) * ldc;
        float* c2 = C + (long)(i + 2) * ldc; float* c3 = C + (long)(i + 3) * ldc;
        for (int k = k0; k < k1; ++k) {
          float a0 = A[(long)(i + 0) * lda + k], a1 = A[(long)(i + 1) * lda + k];
          float a2 = A[(long)(i + 2) * lda + k], a3 = A[(long)(i + 3) * lda + k];
          const float* b = B + (long)k * ldb;
          for (int j = j0; j < j1; ++j) {
            float bv = b[j];
            c0[j] += a0 * bv; c1[j] += a1 * bv; c2[j] += a2 * bv; c3[j] += a3 * bv;
          }
        }
      }
      for (; i < M; ++i) {
        float* c0 = C + (long)i * ldc;
        for (int k = k0; k < k1; ++k) {
          float a0 = A[(long)i * lda + k];
          const float* b = B + (long)k * ldb;
          for (int j = j0; j < j1; ++j) c0[j] += a0 * b[j];
        }
      }
    }
  }
}
/* C[M,N] += A^T * B, A stored [K,M], B [K,N] */
static void gemm_tn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc) {
  const int NB = 512;
  for (int j0 = 0; j0 < N; j0 += NB) {
    int j1 = j0 + NB < N ? j0 + NB : N;
    for (int k = 0; k < K; ++k) {
      const float* a = A + (long)k * lda;
      const float* b = B + (long)k * ldb;
      for (int i = 0; i < M; ++i) {
        float av = a[i];
        float* c = C + (long)i * ldc;
        for (int j = j0; j < j1; ++j) c[j] += av * b[j];
      }
    }
  }
}
/* C[M,N] += A * B^T, A [M,K], B [N,K] */
static void gemm_nt(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc) {
  for (int i = 0; i < M; ++i) {
    const float* a = A + (long)i * lda;
    int j = 0;
    for (; j + 4 <= N; j += 4) {
      const float* b0 = B + (long)(j + 0) * ldb; const float* b1 = B + (long)(j + 1) * ldb;
      const float* b2 = B + (long)(j + 2) * ldb; const float* b3 = B + (long)(j + 3) * ldb;
      float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      for (int k = 0; k < K; ++k) {
        float av = a[k];
        s0 += av * b0[k]; s1 += av * b1[k]; s2 += av * b2[k]; s3 += av * b3[k];
      }
      float* c = C + (long)i * ldc + j;
      c[0] += s0; c[1] += s1; c[2] += s2; c[3] += s3;
    }
    for (; j < N; ++j) {
      const float* b0 = B + (long)j * ldb;
      float s0 = 0;
      for (int k = 0; k < K; ++k) s0 += a[k] * b0[k];
      C[(long)i * ldc + j] += s0;
    }
  }
}

/* ------------------------------------------------------------------ convolution (A.1) */
static void im2col(const float* x, float* col, int Ci, int H, int W, int k) {
  int p = (k - 1) / 2;
  for (int c = 0; c < Ci; ++c)
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) {
        float* dst = col + ((long)(c * k + ky) * k + kx) * H * W;
        for (int y = 0; y < H; ++y) {
          int iy = y + ky - p;
          if (iy < 0 || iy >= H) { memset(dst + (long)y * W, 0, sizeof(float) * W); continue; }
          const float* src = x + ((long)c * H + iy) * W;
          for (int xx = 0; xx < W; ++xx) {
            int ix = xx + kx - p;
            dst[(long)y * W + xx] = (ix >= 0 && ix < W) ? src[ix] : 0.f;
          }
        }
      }
}
static void col2im_add(const float* col, float* x, int Ci, int H, int W, int k) {
  int p = (k - 1) / 2;
  for (int c = 0; c < Ci; ++c)
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) {
        const float* src = col + ((long)(c * k + ky) * k + kx) * H * W;
        for (int y = 0; y < H; ++y) {
          int iy = y + ky - p;
          if (iy < 0 || iy >= H) continue;
          float* dst = x + ((long)c * H + iy) * W;
          for (int xx = 0; xx < W; ++xx) {
            int ix = xx + kx - p;
            if (ix >= 0 && ix < W) dst[ix] += src[(long)y * W + xx];
          }
        }
      }
}

void og_conv2d_fwd(const float* x, const float* W, const float* b, float* y,
                   int N, int Ci, int H, int Wd, int Co, int k) {
  long HW = (long)H * Wd; int Kc = Ci * k * k;
#pragma omp parallel num_threads(team(N))
  {
    float* col = (float*)malloc(sizeof(float) * Kc * HW);
#pragma omp for schedule(static)
    for (int n = 0; n < N; ++n) {
      im2col(x + (long)n * Ci * HW, col, Ci, H, Wd, k);
      float* yn = y + (long)n * Co * HW;
      for (int o = 0; o < Co; ++o) {
        float bv = b ? b[o] : 0.f;
        for (long i = 0; i < HW; ++i) yn[o * HW + i] = bv;
      }
      gemm_nn(Co, (int)HW, Kc, W, Kc, col, (int)HW, yn, (int)HW);
    }
    free(col);
  }
}

void og_conv2d_bwd_data(const float* gy, const float* W, float* gx,
                        int N, int Ci, int H, int Wd, int Co, int k) {
  long HW = (long)H * Wd; int Kc = Ci * k * k;
#pragma omp parallel num_threads(team(N))
  {
    float* col = (float*)malloc(sizeof(float) * Kc * HW);
#pragma omp for schedule(static)
    for (int n = 0; n < N; ++n) {
      memset(col, 0, sizeof(float) * Kc * HW);
      /* col[Kc,HW] = W^T[Kc,Co] * gy[Co,HW] */
      gemm_tn(Kc, (int)HW, Co, W, Kc, gy + (long)n * Co * HW, (int)HW, col, (int)HW);
      float* gxn = gx + (long)n * Ci * HW;
      memset(gxn, 0, sizeof(float) * Ci * HW);
      col2im_add(col, gxn, Ci, H, Wd, k);
    }
    free(col);
  }
}

void og_conv2d_bwd_filter(const float* x, const float* gy, float* gW, float* gb,
                          int N, int Ci, int H, int Wd, int Co, int k) {
  long HW = (long)H * Wd; int Kc = Ci * k * k;
  long nW = (long)Co * Kc, na = nW + Co;
  /* private accumulators per thread, bounded to ~512 MB in total, allocated once and first-touched by their
     owner; the cross-thread sum is a parallel loop over elements in fixed thread order (no critical section) */
  int T = team(N);
  long cap = (512L << 20) / (long)(sizeof(float) * na); if (cap < 1) cap = 1;
  if (T > cap) T = (int)cap;
  float* accs = (float*)malloc(sizeof(float) * na * T);
#pragma omp parallel num_threads(T)
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num();
#else
    int tid = 0;
#endif
    /* num_threads(T) is an upper bound: the runtime may deliver a smaller team (thread limits, cgroup quota).
       Only slots of threads that exist are zeroed and summed. */
#ifdef _OPENMP
    int nt = omp_get_num_threads();
#else
    int nt = 1;
#endif
    float* acc = accs + (long)tid * na;
    memset(acc, 0, sizeof(float) * na);
    float* col = (float*)malloc(sizeof(float) * Kc * HW);
#pragma omp for schedule(static)
    for (int n = 0; n < N; ++n) {
      im2col(x + (long)n * Ci * HW, col, Ci, H, Wd, k);
      const float* gyn = gy + (long)n * Co * HW;
      /* gW[Co,Kc] += gy[Co,HW] * col^T[HW,Kc] */
      gemm_nt(Co, Kc, (int)HW, gyn, (int)HW, col, (int)HW, acc, Kc);
      for (int o = 0; o < Co; ++o) {
        double sb = 0;
        for (long i = 0; i < HW; ++i) sb += gyn[o * HW + i];
        acc[nW + o] += (float)sb;
      }
    }
    free(col);
#pragma omp for schedule(static)
    for (long i = 0; i < na; ++i) {
      float sum = 0;
      for (int t = 0; t < nt; ++t) sum += accs[(long)t * na + i];
      if (i < nW) gW[i] += sum; else if (gb) gb[i - nW] += sum;
    }
  }
  free(accs);
}

void og_conv_upsample_fwd(const float* x, const float* W, const float* b, float* y,
                          int N, int Ci, int H, int Wd, int nOut, int k, int f) {
  /* layers/SpatialConvolutionUpsample.lua:13,16-28: parent conv to nOut*f*f planes, then a contiguous :view */
  og_conv2d_fwd(x, W, b, y, N, Ci, H, Wd, nOut * f * f, k);
}

/* ------------------------------------------------------------------ linear (A.2) */
void og_linear_fwd(const float* x, const float* W, const float* b, float* y, int N, int in, int out) {
#pragma omp parallel for schedule(static) num_threads(team(out))
  for (int j = 0; j < out; ++j) {
    const float* w = W + (long)j * in;
    for (int n = 0; n < N; ++n) {
      const float* xn = x + (long)n * in;
      float s = 0;
      for (int k = 0; k < in; ++k) s += xn[k] * w[k];
      y[(long)n * out + j] = s + (b ? b[j] : 0.f);
    }
  }
}
void og_linear_bwd(const float* x, const float* gy, const float* W, float* gx, float* gW, float* gb,
                   int N, int in, int out) {
  if (gx) {
#pragma omp parallel for schedule(static) num_threads(team(N))
    for (int n = 0; n < N; ++n) {
      float* g = gx + (long)n * in;
      memset(g, 0, sizeof(float) * in);
      for (int j = 0; j < out; ++j) {
        float gv = gy[(long)n * out + j];
        const float* w = W + (long)j * in;
        for (int k = 0; k < in; ++k) g[k] += gv * w[k];
      }
    }
  }
  if (gW) {
#pragma omp parallel for schedule(static) num_threads(team(out))
    for (int j = 0; j < out; ++j) {
      float* gw = gW + (long)j * in;
      double sb = 0;
      for (int n = 0; n < N; ++n) {
        float gv = gy[(long)n * out + j];
        const float* xn = x + (long)n * in;
        for (int k = 0; k < in; ++k) gw[k] += gv * xn[k];
        sb += gv;
      }
      if (gb) gb[j] += (float)sb;
    }
  }
}

/* ------------------------------------------------------------------ batch norm (A.3) */
void og_bn_fwd_train(const float* x, const float* gamma, const float* beta, float* y,
                     float* save_mean, float* save_invstd, float* run_mean, float* run_var,
                     int N, int C, int HW, float eps, float momentum) {
  double m = (double)N * HW;
#pragma omp parallel for schedule(static) num_threads(team(C))
  for (int c = 0; c < C; ++c) {
    double s = 0;
    for (int n = 0; n < N; ++n) {
      const float* p = x + ((long)n * C + c) * HW;
      for (int i = 0; i < HW; ++i) s += p[i];
    }
    double mean = s / m, v = 0;
    for (int n = 0; n < N; ++n) {
      const float* p = x + ((long)n * C + c) * HW;
      for (int i = 0; i < HW; ++i) { double d = p[i] - mean; v += d * d; }
    }
    double var = v / m;
    float invstd = (float)(1.0 / sqrt(var + (double)eps));
    float fm = (float)mean;
    save_mean[c] = fm; save_invstd[c] = invstd;
    if (run_mean) run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * fm;
    if (run_var) run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)(v / (m - 1.0));
    float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    for (int n = 0; n < N; ++n) {
      const float* p = x + ((long)n * C + c) * HW;
      float* q = y + ((long)n * C + c) * HW;
      for (int i = 0; i < HW; ++i) q[i] = (p[i] - fm) * invstd * g + bt;
    }
  }
}
void og_bn_fwd_eval(const float* x, const float* gamma, const float* beta, float* y,
                    const float* run_mean, const float* run_var, int N, int C, int HW, float eps) {
#pragma omp parallel for schedule(static) num_threads(team(C))
  for (int c = 0; c < C; ++c) {
    float invstd = 1.f / sqrtf(run_var[c] + eps);
    for (int n = 0; n < N; ++n) {
      const float* p = x + ((long)n * C + c) * HW;
      float* q = y + ((long)n * C + c) * HW;
      for (int i = 0; i < HW; ++i) q[i] = (p[i] - run_mean[c]) * invstd * gamma[c] + beta[c];
    }
  }
}
void og_bn_bwd_train(const float* x, const float* gy, const float* gamma,
                     const float* save_mean, const float* save_invstd,
                     float* gx, float* ggamma, float* gbeta, int N, int C, int HW) {
  double m = (double)N * HW;
#pragma omp parallel for schedule(static) num_threads(team(C))
  for (int c = 0; c < C; ++c) {
    double sg = 0, sgx = 0;
    float mean = save_mean[c], invstd = save_invstd[c];
    for (int n = 0; n < N; ++n) {
      const float* p = x + ((long)n * C + c) * HW;
      const float* g = gy + ((long)n * C + c) * HW;
      for (int i = 0; i < HW; ++i) { sg += g[i]; sgx += (double)g[i] * ((p[i] - mean) * invstd); }
    }
    if (gbeta) gbeta[c] += (float)sg;
    if (ggamma) ggamma[c] += (float)sgx;
    float mg = (float)(sg / m), mgx = (float)(sgx / m);
    float gi = gamma[c] * invstd;
    if (gx)
      for (int n = 0; n < N; ++n) {
        const float* p = x + ((long)n * C + c) * HW;
        const float* g = gy + ((long)n * C + c) * HW;
        float* q = gx + ((long)n * C + c) * HW;
        for (int i = 0; i < HW; ++i) q[i] = gi * (g[i] - mg - (p[i] - mean) * invstd * mgx);
      }
  }
}

/* ------------------------------------------------------------------ pointwise */
void og_prelu_fwd(const float* x, float w, float* y, long n) {
#pragma omp parallel for schedule(static) if (n > OG_SMALL)
  for (long i = 0; i < n; ++i) y[i] = x[i] > 0 ? x[i] : w * x[i];
}
void og_prelu_bwd(const float* x, const float* gy, float w, float* gx, float* gw, long n) {
  double s = 0;
#pragma omp parallel for schedule(static) reduction(+ : s) if (n > OG_SMALL)
  for (long i = 0; i < n; ++i) {
    if (x[i] > 0) { if (gx) gx[i] = gy[i]; }
    else { if (gx) gx[i] = w * gy[i]; s += (double)x[i] * gy[i]; }
  }
  if (gw) *gw += (float)s;
}
void og_leakyrelu_fwd(const float* x, float s, float* y, long n) {
  /* LeakyReLU.lua:13-19: (|x|+x)/2 + (|x|-x)*(-0.5*s) */
#pragma omp parallel for schedule(static) if (n > OG_SMALL)
  for (long i = 0; i < n; ++i) {
    float a = fabsf(x[i]);
    y[i] = (a + x[i]) * 0.5f + (a - x[i]) * (-0.5f * s);
  }
}
void og_leakyrelu_bwd(const float* x, const float* gy, float s, float* gx, long n) {
  /* LeakyReLU.lua:21-31: negative buffer = (|x|-x)*(-0.5 s) <= 0; sign()+1 is 1 where x>=0 and 0 where x<0 */
#pragma omp parallel for schedule(static) if (n > OG_SMALL)
  for (long i = 0; i < n; ++i) gx[i] = x[i] >= 0 ? gy[i] : s * gy[i];
}
void og_sigmoid_fwd(const float* x, float* y, long n) {
#pragma omp parallel for schedule(static) if (n > OG_SMALL)
  for (long i = 0; i < n; ++i) y[i] = 1.f / (1.f + expf(-x[i]));
}
void og_sigmoid_bwd(const float* y, const float* gy, float* gx, long n) {
#pragma omp parallel for schedule(static) if (n > OG_SMALL)
  for (long i = 0; i < n; ++i) gx[i] = gy[i] * y[i] * (1.f - y[i]);
}
void og_upsample2x_fwd(const float* x, float* y, int NC, int H, int W) {
#pragma omp parallel for schedule(static) num_threads(team(NC))
  for (int c = 0; c < NC; ++c)
    for (int Y = 0; Y < 2 * H; ++Y)
      for (int X = 0; X < 2 * W; ++X)
        y[((long)c * 2 * H + Y) * 2 * W + X] = x[((long)c * H + Y / 2) * W + X / 2];
}
void og_upsample2x_bwd(const float* gy, float* gx, int NC, int H, int W) {
#pragma omp parallel for schedule(static) num_threads(team(NC))
  for (int c = 0; c < NC; ++c)
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) {
        const float* g = gy + ((long)c * 2 * H + 2 * i) * 2 * W + 2 * j;
        gx[((long)c * H + i) * W + j] = g[0] + g[1] + g[2 * W] + g[2 * W + 1];
      }
}
void og_avgpool2_fwd(const float* x, float* y, int NC, int H, int W) {
  int Ho = H / 2, Wo = W / 2;
#pragma omp parallel for schedule(static) num_threads(team(NC))
  for (int c = 0; c < NC; ++c)
    for (int i = 0; i < Ho; ++i)
      for (int j = 0; j < Wo; ++j) {
        const float* p = x + ((long)c * H + 2 * i) * W + 2 * j;
        y[((long)c * Ho + i) * Wo + j] = (p[0] + p[1] + p[W] + p[W + 1]) * 0.25f;
      }
}
void og_avgpool2_bwd(const float* gy, float* gx, int NC, int H, int W) {
  int Ho = H / 2, Wo = W / 2;
#pragma omp parallel for schedule(static) num_threads(team(NC))
  for (int c = 0; c < NC; ++c)
    for (int i = 0; i < Ho; ++i)
      for (int j = 0; j < Wo; ++j) {
        float g = gy[((long)c * Ho + i) * Wo + j] * 0.25f;
        float* p = gx + ((long)c * H + 2 * i) * W + 2 * j;
        p[0] = g; p[1] = g; p[W] = g; p[W + 1] = g;
      }
}
void og_maxpool2_fwd(const float* x, float* y, int* idx, int NC, int H, int W) {
  int Ho = H / 2, Wo = W / 2;
#pragma omp parallel for schedule(static) num_threads(team(NC))
  for (int c = 0; c < NC; ++c)
    for (int i = 0; i < Ho; ++i)
      for (int j = 0; j < Wo; ++j) {
        const float* p = x + ((long)c * H + 2 * i) * W + 2 * j;
        int off[4] = {0, 1, W, W + 1};
        int best = 0; float bv = p[0];
        for (int t = 1; t < 4; ++t) if (p[off[t]] > bv) { bv = p[off[t]]; best = t; }
        y[((long)c * Ho + i) * Wo + j] = bv;
        idx[((long)c * Ho + i) * Wo + j] = best;
      }
}
void og_maxpool2_bwd(const float* gy, const int* idx, float* gx, int NC, int H, int W) {
  int Ho = H / 2, Wo = W / 2;
#pragma omp parallel for schedule(static) num_threads(team(NC))
  for (int c = 0; c < NC; ++c)
    for (int i = 0; i < Ho; ++i)
      for (int j = 0; j < Wo; ++j) {
        float* p = gx + ((long)c * H + 2 * i) * W + 2 * j;
        int off[4] = {0, 1, W, W + 1};
        p[0] = 0; p[1] = 0; p[W] = 0; p[W + 1] = 0;
        p[off[idx[((long)c * Ho + i) * Wo + j]]] = gy[((long)c * Ho + i) * Wo + j];
      }
}
void og_mask_channels(const float* x, const float* mask_nc, float* y, int NC, int HW) {
#pragma omp parallel for schedule(static) num_threads(team(NC))
  for (int c = 0; c < NC; ++c) {
    float mv = mask_nc[c];
    for (int i = 0; i < HW; ++i) y[(long)c * HW + i] = x[(long)c * HW + i] * mv;
  }
}
void og_mask_elems(const float* x, const float* mask, float* y, long n) {
  for (long i = 0; i < n; ++i) y[i] = x[i] * mask[i];
}

/* ------------------------------------------------------------------ spatial transformer (A.11) */
static void mat3_mul(const double* a, const double* b, double* c) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += a[i * 3 + k] * b[k * 3 + j];
      c[i * 3 + j] = s;
    }
}
/* factors right-multiplied in order R, S, T starting from I (stn AffineTransformMatrixGenerator [upstream]) */
static void atm_factors(const float* th, int rot, int scl, int trn, double* R, double* S, double* T, int* idx) {
  int p = 0;
  for (int i = 0; i < 9; ++i) { R[i] = S[i] = T[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  idx[0] = idx[1] = idx[2] = -1;
  if (rot) { double a = th[p]; idx[0] = p++; R[0] = cos(a); R[1] = -sin(a); R[3] = sin(a); R[4] = cos(a); }
  if (scl) { double s = th[p]; idx[1] = p++; S[0] = s; S[4] = s; }
  if (trn) { idx[2] = p; T[2] = th[p]; T[5] = th[p + 1]; p += 2; }
}
void og_affine_matrix_fwd(const float* theta, float* A, int B, int rot, int scl, int trn) {
  int nth = (rot ? 1 : 0) + (scl ? 1 : 0) + (trn ? 2 : 0);
  for (int b = 0; b < B; ++b) {
    double R[9], S[9], T[9], RS[9], M[9]; int idx[3];
    atm_factors(theta + (long)b * nth, rot, scl, trn, R, S, T, idx);
    mat3_mul(R, S, RS); mat3_mul(RS, T, M);
    for (int i = 0; i < 6; ++i) A[(long)b * 6 + i] = (float)M[i];
  }
}
void og_affine_matrix_bwd(const float* theta, const float* gA, float* gtheta, int B, int rot, int scl, int trn) {
  int nth = (rot ? 1 : 0) + (scl ? 1 : 0) + (trn ? 2 : 0);
  for (int b = 0; b < B; ++b) {
    const float* th = theta + (long)b * nth;
    double R[9], S[9], T[9], tmp[9], M[9]; int idx[3];
    atm_factors(th, rot, scl, trn, R, S, T, idx);
    double G[9];
    for (int i = 0; i < 6; ++i) G[i] = gA[(long)b * 6 + i];
    G[6] = G[7] = G[8] = 0;
    float* gt = gtheta + (long)b * nth;
    if (rot) {
      double a = th[idx[0]];
      double dR[9] = {-sin(a), -cos(a), 0, cos(a), -sin(a), 0, 0, 0, 0};
      mat3_mul(dR, S, tmp); mat3_mul(tmp, T, M);
      double s = 0; for (int i = 0; i < 6; ++i) s += G[i] * M[i];
      gt[idx[0]] = (float)s;
    }
    if (scl) {
      double dS[9] = {1, 0, 0, 0, 1, 0, 0, 0, 0};
      mat3_mul(R, dS, tmp); mat3_mul(tmp, T, M);
      double s = 0; for (int i = 0; i < 6; ++i) s += G[i] * M[i];
      gt[idx[1]] = (float)s;
    }
    if (trn) {
      double RS[9]; mat3_mul(R, S, RS);
      for (int q = 0; q < 2; ++q) {
        double dT[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        dT[q == 0 ? 2 : 5] = 1;
        mat3_mul(RS, dT, M);
        double s = 0; for (int i = 0; i < 6; ++i) s += G[i] * M[i];
        gt[idx[2] + q] = (float)s;
      }
    }
  }
}
void og_affine_grid_fwd(const float* A, float* grid, int B, int H, int W) {
  for (int b = 0; b < B; ++b) {
    const float* a = A + (long)b * 6;
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) {
        float yb = -1.f + 2.f * i / (H - 1), xb = -1.f + 2.f * j / (W - 1);
        float* g = grid + (((long)b * H + i) * W + j) * 2;
        g[0] = a[0] * yb + a[1] * xb + a[2];
        g[1] = a[3] * yb + a[4] * xb + a[5];
      }
  }
}
void og_affine_grid_bwd(const float* ggrid, float* gA, int B, int H, int W) {
  for (int b = 0; b < B; ++b) {
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) {
        double yb = -1.0 + 2.0 * i / (H - 1), xb = -1.0 + 2.0 * j / (W - 1);
        const float* g = ggrid + (((long)b * H + i) * W + j) * 2;
        s[0] += g[0] * yb; s[1] += g[0] * xb; s[2] += g[0];
        s[3] += g[1] * yb; s[4] += g[1] * xb; s[5] += g[1];
      }
    for (int q = 0; q < 6; ++q) gA[(long)b * 6 + q] = (float)s[q];
  }
}
void og_bilinear_fwd(const float* img, const float* grid, float* out, int B, int H, int W, int C) {
#pragma omp parallel for schedule(static) num_threads(team(B))
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) {
        const float* g = grid + (((long)b * H + i) * W + j) * 2;
        float yc = (g[0] + 1.f) * (H - 1) / 2.f, xc = (g[1] + 1.f) * (W - 1) / 2.f;
        float fy = floorf(yc), fx = floorf(xc);
        int y0 = (int)fy, x0 = (int)fx;
        float wy = 1.f - (yc - fy), wx = 1.f - (xc - fx);
        float* o = out + (((long)b * H + i) * W + j) * C;
        int v00 = y0 >= 0 && y0 < H && x0 >= 0 && x0 < W;
        int v01 = y0 >= 0 && y0 < H && x0 + 1 >= 0 && x0 + 1 < W;
        int v10 = y0 + 1 >= 0 && y0 + 1 < H && x0 >= 0 && x0 < W;
        int v11 = y0 + 1 >= 0 && y0 + 1 < H && x0 + 1 >= 0 && x0 + 1 < W;
        const float* base = img + (long)b * H * W * C;
        for (int c = 0; c < C; ++c) {
          float a00 = v00 ? base[((long)y0 * W + x0) * C + c] : 0.f;
          float a01 = v01 ? base[((long)y0 * W + x0 + 1) * C + c] : 0.f;
          float a10 = v10 ? base[((long)(y0 + 1) * W + x0) * C + c] : 0.f;
          float a11 = v11 ? base[((long)(y0 + 1) * W + x0 + 1) * C + c] : 0.f;
          o[c] = wx * wy * a00 + (1.f - wx) * wy * a01 + wx * (1.f - wy) * a10 + (1.f - wx) * (1.f - wy) * a11;
        }
      }
}
void og_bilinear_bwd(const float* img, const float* grid, const float* gout,
                     float* gimg, float* ggrid, int B, int H, int W, int C) {
  memset(gimg, 0, sizeof(float) * (long)B * H * W * C);
#pragma omp parallel for schedule(static) num_threads(team(B))
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) {
        const float* g = grid + (((long)b * H + i) * W + j) * 2;
        float yc = (g[0] + 1.f) * (H - 1) / 2.f, xc = (g[1] + 1.f) * (W - 1) / 2.f;
        float fy = floorf(yc), fx = floorf(xc);
        int y0 = (int)fy, x0 = (int)fx;
        float wy = 1.f - (yc - fy), wx = 1.f - (xc - fx);
        const float* go = gout + (((long)b * H + i) * W + j) * C;
        int v00 = y0 >= 0 && y0 < H && x0 >= 0 && x0 < W;
        int v01 = y0 >= 0 && y0 < H && x0 + 1 >= 0 && x0 + 1 < W;
        int v10 = y0 + 1 >= 0 && y0 + 1 < H && x0 >= 0 && x0 < W;
        int v11 = y0 + 1 >= 0 && y0 + 1 < H && x0 + 1 >= 0 && x0 + 1 < W;
        const float* base = img + (long)b * H * W * C;
        float* gbase = gimg + (long)b * H * W * C;
        double d00 = 0, d01 = 0, d10 = 0, d11 = 0;
        for (int c = 0; c < C; ++c) {
          float gv = go[c];
          if (v00) { gbase[((long)y0 * W + x0) * C + c] += wx * wy * gv; d00 += (double)base[((long)y0 * W + x0) * C + c] * gv; }
          if (v01) { gbase[((long)y0 * W + x0 + 1) * C + c] += (1.f - wx) * wy * gv; d01 += (double)base[((long)y0 * W + x0 + 1) * C + c] * gv; }
          if (v10) { gbase[((long)(y0 + 1) * W + x0) * C + c] += wx * (1.f - wy) * gv; d10 += (double)base[((long)(y0 + 1) * W + x0) * C + c] * gv; }
          if (v11) { gbase[((long)(y0 + 1) * W + x0 + 1) * C + c] += (1.f - wx) * (1.f - wy) * gv; d11 += (double)base[((long)(y0 + 1) * W + x0 + 1) * C + c] * gv; }
        }
        double gyf = -wx * d00 + wx * d10 - (1.0 - wx) * d01 + (1.0 - wx) * d11;
        double gxf = -wy * d00 + wy * d01 - (1.0 - wy) * d10 + (1.0 - wy) * d11;
        float* gg = ggrid + (((long)b * H + i) * W + j) * 2;
        gg[0] = (float)(gyf * (H - 1) / 2.0);
        gg[1] = (float)(gxf * (W - 1) / 2.0);
      }
}
void og_nchw_to_nhwc(const float* x, float* y, int N, int C, int H, int W) {
#pragma omp parallel for schedule(static) num_threads(team(N))
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      for (int i = 0; i < H * W; ++i) y[((long)n * H * W + i) * C + c] = x[((long)n * C + c) * H * W + i];
}
void og_nhwc_to_nchw(const float* x, float* y, int N, int C, int H, int W) {
#pragma omp parallel for schedule(static) num_threads(team(N))
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      for (int i = 0; i < H * W; ++i) y[((long)n * C + c) * H * W + i] = x[((long)n * H * W + i) * C + c];
}

/* ------------------------------------------------------------------ criterion / optimiser */
float og_bce_fwd(const float* p, const float* t, int n) {
  const double eps = 1e-12; double s = 0;
  for (int i = 0; i < n; ++i) s += t[i] * log(p[i] + eps) + (1.0 - t[i]) * log(1.0 - p[i] + eps);
  return (float)(-s / n);
}
void og_bce_bwd(const float* p, const float* t, float* g, int n) {
  const double eps = 1e-12;
  for (int i = 0; i < n; ++i)
    g[i] = (float)(-(1.0 / n) * ((double)t[i] - p[i]) / ((1.0 - p[i] + eps) * (p[i] + eps)));
}
void og_adam_step(float* x, const float* g, float* m, float* v, long n, int t,
                  float lr, float b1, float b2, float eps) {
  double bc1 = 1.0 - pow((double)b1, t), bc2 = 1.0 - pow((double)b2, t);
  float step = (float)(lr * sqrt(bc2) / bc1);
#pragma omp parallel for schedule(static) if (n > OG_SMALL)
  for (long i = 0; i < n; ++i) {
    m[i] = b1 * m[i] + (1.f - b1) * g[i];
    v[i] = b2 * v[i] + (1.f - b2) * g[i] * g[i];
    x[i] -= step * m[i] / (sqrtf(v[i]) + eps);
  }
}

/* ================================================================== models */
#define OG_MAXBUF 512
typedef struct {
  int up, Ci, Co, k, bn;   /* bn=1: BN+PReLU after conv; bn=0: Sigmoid */
  long oW, ob, og, obt, opw; /* param offsets */
} g_stage;

typedef struct {
  int ch, S, rot, scl, trn, nth;
  long c1W, c1b, c2W, c2b, l1W, l1b, l2W, l2b;
  /* saved */
  const float* in; float *pool1, *c1, *a1, *c2, *a2, *pool2, *l1, *al1, *theta, *A, *grid, *img, *samp, *out;
} stn_t;

struct og_model {
  int kind, C, nz;
  long np;
  float *p, *g;
  void* bufs[OG_MAXBUF]; int nbuf;
  int B, train;
  /* G */
  int C0, s0, nst; g_stage st[4];
  long oLW, oLb, oLpw;
  float* run; long nrun;
  const float* z;
  float *lin, *act0;
  float *sin_[4], *sup[4], *sconv[4], *sbn[4], *sact[4], *smean[4], *sinv[4];
  /* D */
  stn_t stn[4];
  long t1W, t1b, t1pw, t2W, t2b, t2pw;
  long bW1[4], bb1[4], bpw1[4], bW2[4], bb2[4], bpw2[4];
  long hW1, hb1, hpw, hW2, hb2;
  float *tc1, *ta1, *tc2, *ta2, *tpool, *T;
  float *bc1[4], *ba1[4], *bmp[4], *bdr[4], *bc2[4], *cat, *catd, *h1, *ha1, *hd, *h2, *hsig;
  int* bidx[4];
  float* masks;
};

static float* BUF(og_model* m, long n) {
  if (m->nbuf >= OG_MAXBUF) { fprintf(stderr, "og: too many buffers\n"); abort(); }
  float* p = (float*)malloc(sizeof(float) * (n > 0 ? n : 1));
  m->bufs[m->nbuf++] = p;
  return p;
}
static void free_bufs(og_model* m) {
  for (int i = 0; i < m->nbuf; ++i) free(m->bufs[i]);
  m->nbuf = 0;
}

static long stn_params(stn_t* s, long o, int ch, int S, int rot, int scl, int trn) {
  s->ch = ch; s->S = S; s->rot = rot; s->scl = scl; s->trn = trn;
  s->nth = (rot ? 1 : 0) + (scl ? 1 : 0) + (trn ? 2 : 0);
  int f = 16 * (S / 4) * (S / 4);
  s->c1W = o; o += 16L * ch * 9; s->c1b = o; o += 16;
  s->c2W = o; o += 16L * 16 * 9; s->c2b = o; o += 16;
  s->l1W = o; o += 64L * f; s->l1b = o; o += 64;
  s->l2W = o; o += (long)s->nth * 64; s->l2b = o; o += s->nth;
  return o;
}

og_model* og_model_create(int kind, int C, int nz) {
  og_model* m = (og_model*)calloc(1, sizeof(og_model));
  m->kind = kind; m->C = C; m->nz = nz;
  long o = 0;
  if (kind == OG_G32UP || kind == OG_G32UPC) {
    /* models.lua:138-160 (G32up) and :196-228 (G32up-c) */
    if (kind == OG_G32UPC) {
      m->C0 = 512; m->s0 = 4; m->nst = 4;
      g_stage s[4] = {{1, 512, 512, 3, 1}, {1, 512, 256, 3, 1}, {1, 256, 128, 5, 1}, {0, 128, C, 3, 0}};
      memcpy(m->st, s, sizeof(s));
    } else {
      m->C0 = 128; m->s0 = 8; m->nst = 3;
      g_stage s[3] = {{1, 128, 256, 5, 1}, {1, 256, 128, 5, 1}, {0, 128, C, 3, 0}};
      memcpy(m->st, s, sizeof(s));
    }
    long F0 = (long)m->C0 * m->s0 * m->s0;
    m->oLW = o; o += F0 * nz; m->oLb = o; o += F0; m->oLpw = o; o += 1;
    m->nrun = 0;
    for (int i = 0; i < m->nst; ++i) {
      g_stage* s = &m->st[i];
      s->oW = o; o += (long)s->Co * s->Ci * s->k * s->k; s->ob = o; o += s->Co;
      if (s->bn) { s->og = o; o += s->Co; s->obt = o; o += s->Co; s->opw = o; o += 1; m->nrun += 2 * s->Co; }
    }
    m->run = (float*)calloc(m->nrun, sizeof(float));
    long r = 0;
    for (int i = 0; i < m->nst; ++i) if (m->st[i].bn) {
      for (int c = 0; c < m->st[i].Co; ++c) m->run[r + m->st[i].Co + c] = 1.f;
      r += 2 * m->st[i].Co;
    }
  } else {
    /* models.lua:640-711 create_D32_st3; depth-first getParameters order */
    o = stn_params(&m->stn[0], o, C, 32, 1, 0, 0);
    m->t1W = o; o += 64L * C * 9; m->t1b = o; o += 64; m->t1pw = o; o += 1;
    m->t2W = o; o += 64L * 64 * 9; m->t2b = o; o += 64; m->t2pw = o; o += 1;
    for (int b = 0; b < 3; ++b) {
      o = stn_params(&m->stn[b + 1], o, 64, 16, 1, 1, 1);
      m->bW1[b] = o; o += 64L * 64 * 9; m->bb1[b] = o; o += 64; m->bpw1[b] = o; o += 1;
      m->bW2[b] = o; o += 64L * 64 * 9; m->bb2[b] = o; o += 64; m->bpw2[b] = o; o += 1;
    }
    m->bW1[3] = o; o += 128L * 64 * 25; m->bb1[3] = o; o += 128; m->bpw1[3] = o; o += 1;
    m->bW2[3] = o; o += 128L * 128 * 49; m->bb2[3] = o; o += 128; m->bpw2[3] = o; o += 1;
    m->hW1 = o; o += 256L * 20480; m->hb1 = o; o += 256; m->hpw = o; o += 1;
    m->hW2 = o; o += 256; m->hb2 = o; o += 1;
  }
  m->np = o;
  m->p = (float*)calloc(o, sizeof(float));
  m->g = (float*)calloc(o, sizeof(float));
  return m;
}
void og_model_free(og_model* m) {
  if (!m) return;
  free_bufs(m); free(m->p); free(m->g); free(m->run); free(m->masks); free(m);
}
long og_model_nparams(const og_model* m) { return m->np; }
float* og_model_params(og_model* m) { return m->p; }
float* og_model_grads(og_model* m) { return m->g; }
float* og_model_bn_running(og_model* m, long* n) { if (n) *n = m->nrun; return m->run; }
void og_model_zero_grads(og_model* m) { memset(m->g, 0, sizeof(float) * m->np); }

static unsigned long long lcg_state;
static double lcg_u01(void) {
  lcg_state = lcg_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)(lcg_state >> 11) * (1.0 / 9007199254740992.0);
}
static void fill_uniform(float* p, long n, double a, double b) { for (long i = 0; i < n; ++i) p[i] = (float)(a + (b - a) * lcg_u01()); }
static void fill_const(float* p, long n, float v) { for (long i = 0; i < n; ++i) p[i] = v; }

static void stn_init(og_model* m, stn_t* s) {
  /* models.lua:843-860: weight-init heuristic on the loc-net (top-level children), zero bias, last Linear W=0, b=identity params */
  int f = 16 * (s->S / 4) * (s->S / 4);
  double sd;
  sd = 1.0 / sqrt(9.0 * s->ch); fill_uniform(m->p + s->c1W, 16L * s->ch * 9, -sd, sd); fill_const(m->p + s->c1b, 16, 0);
  sd = 1.0 / sqrt(9.0 * 16); fill_uniform(m->p + s->c2W, 16L * 16 * 9, -sd, sd); fill_const(m->p + s->c2b, 16, 0);
  sd = 1.0 / sqrt((double)f); fill_uniform(m->p + s->l1W, 64L * f, -sd, sd); fill_const(m->p + s->l1b, 64, 0);
  fill_const(m->p + s->l2W, (long)s->nth * 64, 0);
  int q = 0;
  if (s->rot) m->p[s->l2b + q++] = 0;
  if (s->scl) m->p[s->l2b + q++] = 1;
  if (s->trn) { m->p[s->l2b + q++] = 0; m->p[s->l2b + q++] = 0; }
}

void og_model_init(og_model* m, unsigned long long seed) {
  lcg_state = seed * 2654435761ULL + 12345ULL;
  for (int i = 0; i < 4; ++i) lcg_u01();
  if (m->kind != OG_D32_ST3) {
    long F0 = (long)m->C0 * m->s0 * m->s0;
    double sd = 1.0 / sqrt((double)m->nz);
    fill_uniform(m->p + m->oLW, F0 * m->nz, -sd, sd); fill_const(m->p + m->oLb, F0, 0); m->p[m->oLpw] = 0.25f;
    for (int i = 0; i < m->nst; ++i) {
      g_stage* s = &m->st[i];
      sd = 1.0 / sqrt((double)s->Ci * s->k * s->k);
      fill_uniform(m->p + s->oW, (long)s->Co * s->Ci * s->k * s->k, -sd, sd);
      fill_const(m->p + s->ob, s->Co, 0); /* weight-init.lua:70-72 zeroes every top-level bias */
      if (s->bn) { fill_uniform(m->p + s->og, s->Co, 0, 1); fill_const(m->p + s->obt, s->Co, 0); m->p[s->opw] = 0.25f; }
    }
  } else {
    int C = m->C; double sd;
    stn_init(m, &m->stn[0]);
    sd = 1.0 / sqrt(9.0 * C); fill_uniform(m->p + m->t1W, 64L * C * 9, -sd, sd); fill_const(m->p + m->t1b, 64, 0); m->p[m->t1pw] = 0.25f;
    sd = 1.0 / sqrt(9.0 * 64); fill_uniform(m->p + m->t2W, 64L * 64 * 9, -sd, sd); fill_const(m->p + m->t2b, 64, 0); m->p[m->t2pw] = 0.25f;
    for (int b = 0; b < 4; ++b) {
      if (b < 3) stn_init(m, &m->stn[b + 1]);
      int Ci1 = 64, Co1 = b < 3 ? 64 : 128, k1 = b < 3 ? 3 : 5, Ci2 = Co1, Co2 = Co1, k2 = b < 3 ? 3 : 7;
      /* nested in nn.Concat: not touched by weight-init => default reset(): W,b ~ U(+-1/sqrt(fan_in)) */
      sd = 1.0 / sqrt((double)Ci1 * k1 * k1);
      fill_uniform(m->p + m->bW1[b], (long)Co1 * Ci1 * k1 * k1, -sd, sd); fill_uniform(m->p + m->bb1[b], Co1, -sd, sd); m->p[m->bpw1[b]] = 0.25f;
      sd = 1.0 / sqrt((double)Ci2 * k2 * k2);
      fill_uniform(m->p + m->bW2[b], (long)Co2 * Ci2 * k2 * k2, -sd, sd); fill_uniform(m->p + m->bb2[b], Co2, -sd, sd); m->p[m->bpw2[b]] = 0.25f;
    }
    sd = 1.0 / sqrt(20480.0); fill_uniform(m->p + m->hW1, 256L * 20480, -sd, sd); fill_const(m->p + m->hb1, 256, 0); m->p[m->hpw] = 0.25f;
    sd = 1.0 / sqrt(256.0); fill_uniform(m->p + m->hW2, 256, -sd, sd); fill_const(m->p + m->hb2, 1, 0);
  }
}

/* ------------------------------------------------------------------ G */
void og_G_forward(og_model* g, const float* z, int B, float* out, int train) {
  free_bufs(g);
  g->B = B; g->train = train;
  long F0 = (long)g->C0 * g->s0 * g->s0;
  float* zc = BUF(g, (long)B * g->nz); memcpy(zc, z, sizeof(float) * B * g->nz); g->z = zc;
  g->lin = BUF(g, B * F0); g->act0 = BUF(g, B * F0);
  og_linear_fwd(zc, g->p + g->oLW, g->p + g->oLb, g->lin, B, g->nz, (int)F0);
  og_prelu_fwd(g->lin, g->p[g->oLpw], g->act0, B * F0);
  const float* cur = g->act0; int h = g->s0; long r = 0;
  for (int i = 0; i < g->nst; ++i) {
    g_stage* s = &g->st[i];
    g->sin_[i] = (float*)cur;
    if (s->up) {
      g->sup[i] = BUF(g, (long)B * s->Ci * 4 * h * h);
      og_upsample2x_fwd(cur, g->sup[i], B * s->Ci, h, h);
      h *= 2; cur = g->sup[i];
    } else g->sup[i] = (float*)cur;
    long no = (long)B * s->Co * h * h;
    g->sconv[i] = BUF(g, no);
    og_conv2d_fwd(cur, g->p + s->oW, g->p + s->ob, g->sconv[i], B, s->Ci, h, h, s->Co, s->k);
    if (s->bn) {
      g->sbn[i] = BUF(g, no); g->sact[i] = BUF(g, no);
      g->smean[i] = BUF(g, s->Co); g->sinv[i] = BUF(g, s->Co);
      if (train)
        og_bn_fwd_train(g->sconv[i], g->p + s->og, g->p + s->obt, g->sbn[i], g->smean[i], g->sinv[i],
                        g->run + r, g->run + r + s->Co, B, s->Co, h * h, 1e-5f, 0.1f);
      else
        og_bn_fwd_eval(g->sconv[i], g->p + s->og, g->p + s->obt, g->sbn[i], g->run + r, g->run + r + s->Co, B, s->Co, h * h, 1e-5f);
      r += 2 * s->Co;
      og_prelu_fwd(g->sbn[i], g->p[s->opw], g->sact[i], no);
      cur = g->sact[i];
    } else {
      g->sact[i] = BUF(g, no);
      og_sigmoid_fwd(g->sconv[i], g->sact[i], no);
      cur = g->sact[i];
    }
  }
  memcpy(out, cur, sizeof(float) * (long)B * g->C * 32 * 32);
}

/* test/diagnostic access to G's saved forward tensors of stage i (valid until the next forward): 0 conv output, 1 BN output, 2 activation */
const float* og_G_saved(const og_model* g, int stage, int which) {
  if (stage < 0 || stage >= g->nst) return 0;
  return which == 0 ? g->sconv[stage] : which == 1 ? g->sbn[stage] : g->sact[stage];
}

void og_G_backward(og_model* g, const float* gout, float* gz) {
  int B = g->B; int h = 32;
  long no = (long)B * g->C * h * h;
  float* gcur = (float*)malloc(sizeof(float) * no);
  memcpy(gcur, gout, sizeof(float) * no);
  for (int i = g->nst - 1; i >= 0; --i) {
    g_stage* s = &g->st[i];
    no = (long)B * s->Co * h * h;
    float* gconv = (float*)malloc(sizeof(float) * no);
    if (s->bn) {
      float* gbn = (float*)malloc(sizeof(float) * no);
      og_prelu_bwd(g->sbn[i], gcur, g->p[s->opw], gbn, g->g + s->opw, no);
      og_bn_bwd_train(g->sconv[i], gbn, g->p + s->og, g->smean[i], g->sinv[i], gconv, g->g + s->og, g->g + s->obt, B, s->Co, h * h);
      free(gbn);
    } else {
      og_sigmoid_bwd(g->sact[i], gcur, gconv, no);
    }
    free(gcur);
    og_conv2d_bwd_filter(g->sup[i], gconv, g->g + s->oW, g->g + s->ob, B, s->Ci, h, h, s->Co, s->k);
    float* gin = (float*)malloc(sizeof(float) * (long)B * s->Ci * h * h);
    og_conv2d_bwd_data(gconv, g->p + s->oW, gin, B, s->Ci, h, h, s->Co, s->k);
    free(gconv);
    if (s->up) {
      h /= 2;
      float* gs = (float*)malloc(sizeof(float) * (long)B * s->Ci * h * h);
      og_upsample2x_bwd(gin, gs, B * s->Ci, h, h);
      free(gin); gin = gs;
    }
    gcur = gin;
  }
  long F0 = (long)g->C0 * g->s0 * g->s0;
  float* glin = (float*)malloc(sizeof(float) * B * F0);
  og_prelu_bwd(g->lin, gcur, g->p[g->oLpw], glin, g->g + g->oLpw, B * F0);
  free(gcur);
  float* gzz = gz ? gz : (float*)malloc(sizeof(float) * (long)B * g->nz);
  og_linear_bwd(g->z, glin, g->p + g->oLW, gzz, g->g + g->oLW, g->g + g->oLb, B, g->nz, (int)F0);
  if (!gz) free(gzz);
  free(glin);
}

/* ------------------------------------------------------------------ D */
long og_D_mask_floats(int B) { return (long)B * (64 * 4 + 128 + 320 + 256); }

static void stn_forward(og_model* m, stn_t* s, const float* in, int B) {
  int ch = s->ch, S = s->S, S2 = S / 2, S4 = S / 4, f = 16 * S4 * S4;
  const float* p = m->p;
  s->in = in;
  s->pool1 = BUF(m, (long)B * ch * S2 * S2); og_avgpool2_fwd(in, s->pool1, B * ch, S, S);
  s->c1 = BUF(m, (long)B * 16 * S2 * S2); og_conv2d_fwd(s->pool1, p + s->c1W, p + s->c1b, s->c1, B, ch, S2, S2, 16, 3);
  s->a1 = BUF(m, (long)B * 16 * S2 * S2); og_leakyrelu_fwd(s->c1, 0.333f, s->a1, (long)B * 16 * S2 * S2);
  s->c2 = BUF(m, (long)B * 16 * S2 * S2); og_conv2d_fwd(s->a1, p + s->c2W, p + s->c2b, s->c2, B, 16, S2, S2, 16, 3);
  s->a2 = BUF(m, (long)B * 16 * S2 * S2); og_leakyrelu_fwd(s->c2, 0.333f, s->a2, (long)B * 16 * S2 * S2);
  s->pool2 = BUF(m, (long)B * f); og_avgpool2_fwd(s->a2, s->pool2, B * 16, S2, S2);
  s->l1 = BUF(m, (long)B * 64); og_linear_fwd(s->pool2, p + s->l1W, p + s->l1b, s->l1, B, f, 64);
  s->al1 = BUF(m, (long)B * 64); og_leakyrelu_fwd(s->l1, 0.333f, s->al1, (long)B * 64);
  s->theta = BUF(m, (long)B * s->nth); og_linear_fwd(s->al1, p + s->l2W, p + s->l2b, s->theta, B, 64, s->nth);
  s->A = BUF(m, (long)B * 6); og_affine_matrix_fwd(s->theta, s->A, B, s->rot, s->scl, s->trn);
  s->grid = BUF(m, (long)B * S * S * 2); og_affine_grid_fwd(s->A, s->grid, B, S, S);
  s->img = BUF(m, (long)B * S * S * ch); og_nchw_to_nhwc(in, s->img, B, ch, S, S);
  s->samp = BUF(m, (long)B * S * S * ch); og_bilinear_fwd(s->img, s->grid, s->samp, B, S, S, ch);
  s->out = BUF(m, (long)B * S * S * ch); og_nhwc_to_nchw(s->samp, s->out, B, ch, S, S);
}
/* gout [B,ch,S,S] -> gin [B,ch,S,S] (written), accumulates loc-net param grads */
static void stn_backward(og_model* m, stn_t* s, const float* gout, float* gin, int B) {
  int ch = s->ch, S = s->S, S2 = S / 2, S4 = S / 4, f = 16 * S4 * S4;
  const float* p = m->p; float* g = m->g;
  long nimg = (long)B * S * S * ch;
  float* go = (float*)malloc(sizeof(float) * nimg); og_nchw_to_nhwc(gout, go, B, ch, S, S);
  float* gimg = (float*)malloc(sizeof(float) * nimg);
  float* ggrid = (float*)malloc(sizeof(float) * (long)B * S * S * 2);
  og_bilinear_bwd(s->img, s->grid, go, gimg, ggrid, B, S, S, ch);
  og_nhwc_to_nchw(gimg, gin, B, ch, S, S);
  free(go); free(gimg);
  float* gA = (float*)malloc(sizeof(float) * B * 6); og_affine_grid_bwd(ggrid, gA, B, S, S); free(ggrid);
  float* gth = (float*)malloc(sizeof(float) * B * s->nth); og_affine_matrix_bwd(s->theta, gA, gth, B, s->rot, s->scl, s->trn); free(gA);
  float* gal1 = (float*)malloc(sizeof(float) * B * 64);
  og_linear_bwd(s->al1, gth, p + s->l2W, gal1, g + s->l2W, g + s->l2b, B, 64, s->nth); free(gth);
  float* gl1 = (float*)malloc(sizeof(float) * B * 64); og_leakyrelu_bwd(s->l1, gal1, 0.333f, gl1, (long)B * 64); free(gal1);
  float* gp2 = (float*)malloc(sizeof(float) * (long)B * f);
  og_linear_bwd(s->pool2, gl1, p + s->l1W, gp2, g + s->l1W, g + s->l1b, B, f, 64); free(gl1);
  long n2 = (long)B * 16 * S2 * S2;
  float* ga2 = (float*)malloc(sizeof(float) * n2); og_avgpool2_bwd(gp2, ga2, B * 16, S2, S2); free(gp2);
  float* gc2 = (float*)malloc(sizeof(float) * n2); og_leakyrelu_bwd(s->c2, ga2, 0.333f, gc2, n2); free(ga2);
  og_conv2d_bwd_filter(s->a1, gc2, g + s->c2W, g + s->c2b, B, 16, S2, S2, 16, 3);
  float* ga1 = (float*)malloc(sizeof(float) * n2); og_conv2d_bwd_data(gc2, p + s->c2W, ga1, B, 16, S2, S2, 16, 3); free(gc2);
  float* gc1 = (float*)malloc(sizeof(float) * n2); og_leakyrelu_bwd(s->c1, ga1, 0.333f, gc1, n2); free(ga1);
  og_conv2d_bwd_filter(s->pool1, gc1, g + s->c1W, g + s->c1b, B, ch, S2, S2, 16, 3);
  float* gp1 = (float*)malloc(sizeof(float) * (long)B * ch * S2 * S2);
  og_conv2d_bwd_data(gc1, p + s->c1W, gp1, B, ch, S2, S2, 16, 3); free(gc1);
  float* gin2 = (float*)malloc(sizeof(float) * nimg); og_avgpool2_bwd(gp1, gin2, B * ch, S, S); free(gp1);
  for (long i = 0; i < nimg; ++i) gin[i] += gin2[i];   /* ConcatTable sums the two branches' gradInput */
  free(gin2);
}

void og_D_forward(og_model* d, const float* x, int B, const float* masks, float* out_sig, float* out_pre) {
  free_bufs(d);
  d->B = B; int C = d->C; const float* p = d->p;
  free(d->masks);
  long nm = og_D_mask_floats(B);
  d->masks = (float*)malloc(sizeof(float) * nm);
  if (masks) memcpy(d->masks, masks, sizeof(float) * nm);
  else {
    /* evaluate(): SpatialDropout(p) scales by (1-p); Dropout (v2) is the identity */
    long o = 0;
    for (long i = 0; i < (long)B * (64 * 4 + 128); ++i) d->masks[o++] = 0.8f;
    for (long i = 0; i < (long)B * 320; ++i) d->masks[o++] = 0.5f;
    for (long i = 0; i < (long)B * 256; ++i) d->masks[o++] = 1.f;
  }
  const float* mk = d->masks;
  float* xc = BUF(d, (long)B * C * 1024); memcpy(xc, x, sizeof(float) * (long)B * C * 1024);
  stn_forward(d, &d->stn[0], xc, B);
  long n64 = (long)B * 64 * 1024;
  d->tc1 = BUF(d, n64); og_conv2d_fwd(d->stn[0].out, p + d->t1W, p + d->t1b, d->tc1, B, C, 32, 32, 64, 3);
  d->ta1 = BUF(d, n64); og_prelu_fwd(d->tc1, p[d->t1pw], d->ta1, n64);
  d->tc2 = BUF(d, n64); og_conv2d_fwd(d->ta1, p + d->t2W, p + d->t2b, d->tc2, B, 64, 32, 32, 64, 3);
  d->ta2 = BUF(d, n64); og_prelu_fwd(d->tc2, p[d->t2pw], d->ta2, n64);
  d->tpool = BUF(d, n64 / 4); og_avgpool2_fwd(d->ta2, d->tpool, B * 64, 32, 32);
  d->T = BUF(d, n64 / 4); og_mask_channels(d->tpool, mk, d->T, B * 64, 256); mk += (long)B * 64;
  d->cat = BUF(d, (long)B * 320 * 64);
  for (int b = 0; b < 4; ++b) {
    int Co = b < 3 ? 64 : 128, k1 = b < 3 ? 3 : 5, k2 = b < 3 ? 3 : 7;
    const float* bin = d->T;
    if (b < 3) { stn_forward(d, &d->stn[b + 1], d->T, B); bin = d->stn[b + 1].out; }
    long n1 = (long)B * Co * 256;
    d->bc1[b] = BUF(d, n1); og_conv2d_fwd(bin, p + d->bW1[b], p + d->bb1[b], d->bc1[b], B, 64, 16, 16, Co, k1);
    d->ba1[b] = BUF(d, n1); og_prelu_fwd(d->bc1[b], p[d->bpw1[b]], d->ba1[b], n1);
    d->bmp[b] = BUF(d, n1 / 4); d->bidx[b] = (int*)BUF(d, n1 / 4);
    og_maxpool2_fwd(d->ba1[b], d->bmp[b], d->bidx[b], B * Co, 16, 16);
    d->bdr[b] = BUF(d, n1 / 4); og_mask_channels(d->bmp[b], mk, d->bdr[b], B * Co, 64); mk += (long)B * Co;
    d->bc2[b] = BUF(d, n1 / 4); og_conv2d_fwd(d->bdr[b], p + d->bW2[b], p + d->bb2[b], d->bc2[b], B, Co, 8, 8, Co, k2);
    /* PReLU output goes straight into the Concat(2) slot */
    int coff = b * 64;
    float* tmp = BUF(d, n1 / 4); og_prelu_fwd(d->bc2[b], p[d->bpw2[b]], tmp, n1 / 4);
    for (int n = 0; n < B; ++n)
      memcpy(d->cat + ((long)n * 320 + coff) * 64, tmp + (long)n * Co * 64, sizeof(float) * Co * 64);
  }
  d->catd = BUF(d, (long)B * 20480); og_mask_channels(d->cat, mk, d->catd, B * 320, 64); mk += (long)B * 320;
  d->h1 = BUF(d, (long)B * 256); og_linear_fwd(d->catd, p + d->hW1, p + d->hb1, d->h1, B, 20480, 256);
  d->ha1 = BUF(d, (long)B * 256); og_prelu_fwd(d->h1, p[d->hpw], d->ha1, (long)B * 256);
  d->hd = BUF(d, (long)B * 256); og_mask_elems(d->ha1, mk, d->hd, (long)B * 256);
  d->h2 = BUF(d, B); og_linear_fwd(d->hd, p + d->hW2, p + d->hb2, d->h2, B, 256, 1);
  d->hsig = BUF(d, B); og_sigmoid_fwd(d->h2, d->hsig, B);
  if (out_sig) memcpy(out_sig, d->hsig, sizeof(float) * B);
  if (out_pre) memcpy(out_pre, d->h2, sizeof(float) * B);
}

void og_D_backward(og_model* d, const float* gout, float* gx) {
  int B = d->B, C = d->C; const float* p = d->p; float* g = d->g;
  const float* mk_trunk = d->masks;
  const float* mk_br = d->masks + (long)B * 64;
  const float* mk_head = d->masks + (long)B * (64 * 4 + 128);
  const float* mk_fc = mk_head + (long)B * 320;
  float* gh2 = (float*)malloc(sizeof(float) * B); og_sigmoid_bwd(d->hsig, gout, gh2, B);
  float* ghd = (float*)malloc(sizeof(float) * B * 256);
  og_linear_bwd(d->hd, gh2, p + d->hW2, ghd, g + d->hW2, g + d->hb2, B, 256, 1); free(gh2);
  float* gha1 = (float*)malloc(sizeof(float) * B * 256); og_mask_elems(ghd, mk_fc, gha1, (long)B * 256); free(ghd);
  float* gh1 = (float*)malloc(sizeof(float) * B * 256); og_prelu_bwd(d->h1, gha1, p[d->hpw], gh1, g + d->hpw, (long)B * 256); free(gha1);
  float* gcatd = (float*)malloc(sizeof(float) * (long)B * 20480);
  og_linear_bwd(d->catd, gh1, p + d->hW1, gcatd, g + d->hW1, g + d->hb1, B, 20480, 256); free(gh1);
  float* gcat = (float*)malloc(sizeof(float) * (long)B * 20480); og_mask_channels(gcatd, mk_head, gcat, B * 320, 64); free(gcatd);
  long nT = (long)B * 64 * 256;
  float* gT = (float*)calloc(nT, sizeof(float));
  const float* mk = mk_br;
  for (int b = 0; b < 4; ++b) {
    int Co = b < 3 ? 64 : 128, k1 = b < 3 ? 3 : 5, k2 = b < 3 ? 3 : 7;
    long n2 = (long)B * Co * 64, n1 = n2 * 4;
    float* go = (float*)malloc(sizeof(float) * n2);
    for (int n = 0; n < B; ++n) memcpy(go + (long)n * Co * 64, gcat + ((long)n * 320 + b * 64) * 64, sizeof(float) * Co * 64);
    float* gc2 = (float*)malloc(sizeof(float) * n2); og_prelu_bwd(d->bc2[b], go, p[d->bpw2[b]], gc2, g + d->bpw2[b], n2); free(go);
    og_conv2d_bwd_filter(d->bdr[b], gc2, g + d->bW2[b], g + d->bb2[b], B, Co, 8, 8, Co, k2);
    float* gdr = (float*)malloc(sizeof(float) * n2); og_conv2d_bwd_data(gc2, p + d->bW2[b], gdr, B, Co, 8, 8, Co, k2); free(gc2);
    float* gmp = (float*)malloc(sizeof(float) * n2); og_mask_channels(gdr, mk, gmp, B * Co, 64); mk += (long)B * Co; free(gdr);
    float* ga1 = (float*)malloc(sizeof(float) * n1); og_maxpool2_bwd(gmp, d->bidx[b], ga1, B * Co, 16, 16); free(gmp);
    float* gc1 = (float*)malloc(sizeof(float) * n1); og_prelu_bwd(d->bc1[b], ga1, p[d->bpw1[b]], gc1, g + d->bpw1[b], n1); free(ga1);
    const float* bin = b < 3 ? d->stn[b + 1].out : d->T;
    og_conv2d_bwd_filter(bin, gc1, g + d->bW1[b], g + d->bb1[b], B, 64, 16, 16, Co, k1);
    float* gbin = (float*)malloc(sizeof(float) * nT); og_conv2d_bwd_data(gc1, p + d->bW1[b], gbin, B, 64, 16, 16, Co, k1); free(gc1);
    if (b < 3) {
      float* gs = (float*)malloc(sizeof(float) * nT);
      stn_backward(d, &d->stn[b + 1], gbin, gs, B);
      for (long i = 0; i < nT; ++i) gT[i] += gs[i];
      free(gs);
    } else
      for (long i = 0; i < nT; ++i) gT[i] += gbin[i];
    free(gbin);
  }
  free(gcat);
  float* gtp = (float*)malloc(sizeof(float) * nT); og_mask_channels(gT, mk_trunk, gtp, B * 64, 256); free(gT);
  long n64 = (long)B * 64 * 1024;
  float* gta2 = (float*)malloc(sizeof(float) * n64); og_avgpool2_bwd(gtp, gta2, B * 64, 32, 32); free(gtp);
  float* gtc2 = (float*)malloc(sizeof(float) * n64); og_prelu_bwd(d->tc2, gta2, p[d->t2pw], gtc2, g + d->t2pw, n64); free(gta2);
  og_conv2d_bwd_filter(d->ta1, gtc2, g + d->t2W, g + d->t2b, B, 64, 32, 32, 64, 3);
  float* gta1 = (float*)malloc(sizeof(float) * n64); og_conv2d_bwd_data(gtc2, p + d->t2W, gta1, B, 64, 32, 32, 64, 3); free(gtc2);
  float* gtc1 = (float*)malloc(sizeof(float) * n64); og_prelu_bwd(d->tc1, gta1, p[d->t1pw], gtc1, g + d->t1pw, n64); free(gta1);
  og_conv2d_bwd_filter(d->stn[0].out, gtc1, g + d->t1W, g + d->t1b, B, C, 32, 32, 64, 3);
  float* gs0 = (float*)malloc(sizeof(float) * (long)B * C * 1024);
  og_conv2d_bwd_data(gtc1, p + d->t1W, gs0, B, C, 32, 32, 64, 3); free(gtc1);
  float* gin = gx ? gx : (float*)malloc(sizeof(float) * (long)B * C * 1024);
  stn_backward(d, &d->stn[0], gs0, gin, B);
  free(gs0);
  if (!gx) free(gin);
}

/* ------------------------------------------------------------------ training step (adversarial.lua) */
og_trainer* og_trainer_create(og_model* G, og_model* D) {
  og_trainer* t = (og_trainer*)calloc(1, sizeof(og_trainer));
  t->G = G; t->D = D;
  t->mD = (float*)calloc(D->np, sizeof(float)); t->vD = (float*)calloc(D->np, sizeof(float));
  t->mG = (float*)calloc(G->np, sizeof(float)); t->vG = (float*)calloc(G->np, sizeof(float));
  return t;
}
void og_trainer_free(og_trainer* t) { if (!t) return; free(t->mD); free(t->vD); free(t->mG); free(t->vG); free(t); }

static float penalty_and_clamp(og_model* m, float l1, float l2sign, float l2, float clampv) {
  /* adversarial.lua:92-98,110-112 (D) and :201-212 (G; NB the G sign term uses G_L2, line 206) */
  double add = 0;
  long n = m->np;
  if (l1 != 0 || l2 != 0) {
    double n1 = 0, n2 = 0;
    for (long i = 0; i < n; ++i) { n1 += fabs((double)m->p[i]); n2 += (double)m->p[i] * m->p[i]; }
    add = l1 * n1 + l2 * n2 / 2.0;
    for (long i = 0; i < n; ++i) {
      float sg = m->p[i] > 0 ? 1.f : (m->p[i] < 0 ? -1.f : 0.f);
      m->g[i] += sg * l2sign + m->p[i] * l2;
    }
  }
  if (clampv != 0)
    for (long i = 0; i < n; ++i) { float v = m->g[i]; m->g[i] = v < -clampv ? -clampv : (v > clampv ? clampv : v); }
  return (float)add;
}

float og_fevalD(og_trainer* t, const og_step_cfg* cfg, const float* inputs, const float* targets,
                const float* masks, float* d_out) {
  int B = cfg->B;
  og_model_zero_grads(t->D);                                     /* adversarial.lua:81 */
  float* out = (float*)malloc(sizeof(float) * B);
  og_D_forward(t->D, inputs, B, masks, out, NULL);               /* :84 */
  float f = og_bce_fwd(out, targets, B);                         /* :85 */
  float* df = (float*)malloc(sizeof(float) * B);
  og_bce_bwd(out, targets, df, B);                               /* :88 */
  og_D_backward(t->D, df, NULL);                                 /* :89 */
  f += penalty_and_clamp(t->D, cfg->D_L1, cfg->D_L1, cfg->D_L2, cfg->D_clamp);  /* :92-112 */
  if (d_out) memcpy(d_out, out, sizeof(float) * B);
  free(out); free(df);
  return f;
}

float og_fevalG_on_D(og_trainer* t, const og_step_cfg* cfg, const float* z, const float* masks) {
  int B = cfg->B; int C = t->G->C;
  og_model_zero_grads(t->G);                                     /* :177 */
  float* samples = (float*)malloc(sizeof(float) * (long)B * C * 1024);
  og_G_forward(t->G, z, B, samples, 1);                          /* :185 */
  float* out = (float*)malloc(sizeof(float) * B);
  og_D_forward(t->D, samples, B, masks, out, NULL);              /* :187 */
  float* tg = (float*)malloc(sizeof(float) * B);
  for (int i = 0; i < B; ++i) tg[i] = 1.f;                       /* :255 targets:fill(Y_NOT_GENERATOR) */
  float f = og_bce_fwd(out, tg, B);                              /* :188 */
  float* df = (float*)malloc(sizeof(float) * B);
  og_bce_bwd(out, tg, df, B);                                    /* :191 */
  float* gimg = (float*)malloc(sizeof(float) * (long)B * C * 1024);
  og_D_backward(t->D, df, gimg);                                 /* :192-193 (also accumulates into gradD; zeroed by next fevalD) */
  og_G_backward(t->G, gimg, NULL);                               /* :197 */
  f += penalty_and_clamp(t->G, cfg->G_L1, cfg->G_L2, cfg->G_L2, cfg->G_clamp);  /* :201-212 */
  free(samples); free(out); free(tg); free(df); free(gimg);
  return f;
}

void og_train_step(og_trainer* t, const og_step_cfg* cfg, const float* real, const float* zD,
                   const float* zG, const float* masks, float* lossD, float* lossG, float* d_out) {
  int B = cfg->B, hB = B / 2, C = t->G->C, nz = t->G->nz;
  long img = (long)C * 1024, nm = og_D_mask_floats(B);
  float* inputs = (float*)malloc(sizeof(float) * B * img);
  float* targets = (float*)malloc(sizeof(float) * B);
  int mi = 0;
  for (int k = 0; k < cfg->d_iters; ++k) {
    memcpy(inputs, real + (long)k * hB * img, sizeof(float) * hB * img);              /* :225-230 */
    og_G_forward(t->G, zD + (long)k * hB * nz, hB, inputs + hB * img, 1);             /* :233 (train-mode BN over B/2) */
    for (int i = 0; i < B; ++i) targets[i] = i < hB ? 1.f : 0.f;
    float f = og_fevalD(t, cfg, inputs, targets, masks ? masks + (long)(mi++) * nm : NULL, d_out);
    if (lossD) lossD[k] = f;
    t->tD += 1;
    og_adam_step(t->D->p, t->D->g, t->mD, t->vD, t->D->np, t->tD, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps);  /* :245 */
  }
  for (int k = 0; k < cfg->g_iters; ++k) {
    float f = og_fevalG_on_D(t, cfg, zG + (long)k * B * nz, masks ? masks + (long)(mi++) * nm : NULL);
    if (lossG) lossG[k] = f;
    t->tG += 1;
    og_adam_step(t->G->p, t->G->g, t->mG, t->vG, t->G->np, t->tG, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps);  /* :262 */
  }
  free(inputs); free(targets);
}
