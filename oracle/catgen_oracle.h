/*
 * catgen_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * A plain-C fp32 restatement of the Torch7 `nn` CPU semantics that
 * aleju/cat-generator's DCGAN training hot path executes.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library; the product path (cat-generator_b200/csrc) never does.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 * and its numerics live in un-vendored luarocks (nn, cudnn, stn, optim) that
 * cannot run in this environment.  The semantics below
 * are restated from the reference call sites and SURVEY.md Appendix A and are
 * cross-checked numerically against PyTorch-CPU autograd (tests/test_oracle_vs_torch.py) and,
 * compiled in float64 (OG_F64 below), reproduce that restatement's float64 golden vectors
 * (tests/golden) to 4e-12 on G and 1e-6 on D (tests/test_golden.py).
 *
 * All tensors are contiguous fp32, Torch7 layout (NCHW, weights [Cout,Cin,kH,kW],
 * Linear [out,in]).  Convolutions: stride 1, pad (k-1)/2 (every call site in
 * /root/reference/models.lua:138-228,640-711,843-846 uses exactly that).
 */
#ifndef CATGEN_ORACLE_H
#define CATGEN_ORACLE_H
#ifdef OG_F64
/* libcatgen_oracle_f64.so: THE SAME SOURCE with every `float` widened to double (storage, arithmetic, entry points).  Two fp32
 * implementations of these networks disagree by ~1e-3 of max|gradient| whenever one PReLU / max-pool decision lands on the
 * other side of zero (one element of 524288 did in tests/golden/G32upc_rgb_b4); in float64 that noise is gone and the oracle's
 * SEMANTICS can be compared with the float64 PyTorch restatement to ~1e-7 (tests/test_golden.py).  Model-level entry points only:
 * og_step_cfg changes layout in this build and the trainer is not bound from Python.
 * The system headers must be seen before `float` is redefined. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#define float double
#define floorf floor
#define sqrtf sqrt
#define fabsf fabs
#define expf exp
#endif
#ifdef __cplusplus
extern "C" {
#endif

void og_set_threads(int n);
int  og_get_threads(void);

/* ---- op level ---- */
/* A.1 nn.SpatialConvolution / cudnn.SpatialConvolution (models.lua:206,212,218,222,646,...) */
void og_conv2d_fwd(const float* x, const float* W, const float* b, float* y,
                   int N, int Ci, int H, int Wd, int Co, int k);
void og_conv2d_bwd_data(const float* gy, const float* W, float* gx,
                        int N, int Ci, int H, int Wd, int Co, int k);
/* accumulates (+=) into gW, gb like accGradParameters */
void og_conv2d_bwd_filter(const float* x, const float* gy, float* gW, float* gb,
                          int N, int Ci, int H, int Wd, int Co, int k);
/* A.2 nn.Linear (models.lua:199,697,700,852,854) */
void og_linear_fwd(const float* x, const float* W, const float* b, float* y, int N, int in, int out);
void og_linear_bwd(const float* x, const float* gy, const float* W, float* gx, float* gW, float* gb,
                   int N, int in, int out);
/* A.3 nn.SpatialBatchNormalization, training mode (models.lua:207,213,219) */
void og_bn_fwd_train(const float* x, const float* gamma, const float* beta, float* y,
                     float* save_mean, float* save_invstd, float* run_mean, float* run_var,
                     int N, int C, int HW, float eps, float momentum);
void og_bn_fwd_eval(const float* x, const float* gamma, const float* beta, float* y,
                    const float* run_mean, const float* run_var, int N, int C, int HW, float eps);
void og_bn_bwd_train(const float* x, const float* gy, const float* gamma,
                     const float* save_mean, const float* save_invstd,
                     float* gx, float* ggamma, float* gbeta, int N, int C, int HW);
/* A.4 nn.PReLU() with one shared slope */
void og_prelu_fwd(const float* x, float w, float* y, long n);
void og_prelu_bwd(const float* x, const float* gy, float w, float* gx, float* gw, long n);
/* nn.LeakyReLU, /root/reference/LeakyReLU.lua:13-31 (grad at x==0 is 1) */
void og_leakyrelu_fwd(const float* x, float s, float* y, long n);
void og_leakyrelu_bwd(const float* x, const float* gy, float s, float* gx, long n);
/* A.5 nn.SpatialUpSamplingNearest(2) */
void og_upsample2x_fwd(const float* x, float* y, int NC, int H, int Wd);
void og_upsample2x_bwd(const float* gy, float* gx, int NC, int H, int Wd);
/* A.6 nn.Sigmoid */
void og_sigmoid_fwd(const float* x, float* y, long n);
void og_sigmoid_bwd(const float* y, const float* gy, float* gx, long n);
/* A.12 pooling 2x2/2 */
void og_avgpool2_fwd(const float* x, float* y, int NC, int H, int Wd);
void og_avgpool2_bwd(const float* gy, float* gx, int NC, int H, int Wd);
void og_maxpool2_fwd(const float* x, float* y, int* idx, int NC, int H, int Wd);
void og_maxpool2_bwd(const float* gy, const int* idx, float* gx, int NC, int H, int Wd);
/* A.12 dropout with explicit masks (mask value already includes any rescale) */
void og_mask_channels(const float* x, const float* mask_nc, float* y, int NC, int HW);
void og_mask_elems(const float* x, const float* mask, float* y, long n);
/* A.11 spatial transformer pieces (stn: qassemoquab/stnbhwd, [upstream]) */
void og_affine_matrix_fwd(const float* theta, float* A, int B, int rot, int scl, int trn);
void og_affine_matrix_bwd(const float* theta, const float* gA, float* gtheta, int B, int rot, int scl, int trn);
void og_affine_grid_fwd(const float* A, float* grid, int B, int H, int Wd);
void og_affine_grid_bwd(const float* ggrid, float* gA, int B, int H, int Wd);
/* img, out: NHWC (BHWD); grid: [B,H,W,2] channel0=y channel1=x */
void og_bilinear_fwd(const float* img, const float* grid, float* out, int B, int H, int Wd, int C);
void og_bilinear_bwd(const float* img, const float* grid, const float* gout,
                     float* gimg, float* ggrid, int B, int H, int Wd, int C);
void og_nchw_to_nhwc(const float* x, float* y, int N, int C, int H, int Wd);
void og_nhwc_to_nchw(const float* x, float* y, int N, int C, int H, int Wd);
/* A.7 nn.BCECriterion (train.lua:181) */
float og_bce_fwd(const float* p, const float* t, int n);
void  og_bce_bwd(const float* p, const float* t, float* g, int n);
/* A.8 optim.adam (adversarial.lua:245,262); t is the 1-based step AFTER increment */
void og_adam_step(float* x, const float* g, float* m, float* v, long n, int t,
                  float lr, float b1, float b2, float eps);
/* layers/SpatialConvolutionUpsample.lua:16-28 -- conv to nOut*f*f planes; the :view is a no-op on memory */
void og_conv_upsample_fwd(const float* x, const float* W, const float* b, float* y,
                          int N, int Ci, int H, int Wd, int nOut, int k, int f);

/* ---- model level ---- */
enum { OG_G32UP = 0, OG_G32UPC = 1, OG_D32_ST3 = 2 };
typedef struct og_model og_model;

og_model* og_model_create(int kind, int C, int nz);
void      og_model_free(og_model* m);
long      og_model_nparams(const og_model* m);
float*    og_model_params(og_model* m);   /* flat, nn getParameters() order */
float*    og_model_grads(og_model* m);    /* flat, same order */
float*    og_model_bn_running(og_model* m, long* n); /* G only: [mean0,var0,mean1,var1,...] */
/* weight-init per SURVEY.md A.9 using a deterministic LCG (distribution parity only) */
void      og_model_init(og_model* m, unsigned long long seed);
void      og_model_zero_grads(og_model* m);
/* dropout masks of D in forward order: 5x SpatialDropout(0.2) [B*64 ... ], SpatialDropout(0.5) [B*320], Dropout(0.5) [B*256].
   layout: trunk[B*64], br1[B*64], br2[B*64], br3[B*64], br4[B*128], head[B*320], fc[B*256]; values are the
   multipliers (0/1 for SpatialDropout, 0/2 for Dropout).  NULL => eval mode (SpatialDropout scales by 1-p). */
long      og_D_mask_floats(int B);

/* G: z [B,nz] -> images [B,C,32,32].  train!=0: batch-stat BN + running update */
void og_G_forward(og_model* g, const float* z, int B, float* out, int train);
/* gout [B,C,32,32]; accumulates param grads; gz may be NULL */
void og_G_backward(og_model* g, const float* gout, float* gz);
/* diagnostic: G's saved forward tensor of a stage (0 conv output, 1 BN output, 2 activation), valid until the next forward */
const float* og_G_saved(const og_model* g, int stage, int which);
/* D: x [B,C,32,32] -> out_sig [B], out_pre [B] (either may be NULL) */
void og_D_forward(og_model* d, const float* x, int B, const float* masks, float* out_sig, float* out_pre);
/* gout: d loss / d sigmoid-output [B]; gx [B,C,32,32] may be NULL */
void og_D_backward(og_model* d, const float* gout, float* gx);

/* one adversarial.train step body (adversarial.lua:221-266) */
typedef struct {
  int   B, d_iters, g_iters;
  float D_L1, D_L2, G_L1, G_L2, D_clamp, G_clamp;
  float lr, beta1, beta2, eps;
} og_step_cfg;
typedef struct {
  og_model *G, *D;
  float *mD, *vD, *mG, *vG;
  int tD, tG;
} og_trainer;
og_trainer* og_trainer_create(og_model* G, og_model* D);
void        og_trainer_free(og_trainer* t);
/* real: [d_iters][B/2,C,32,32]; zD: [d_iters][B/2,nz]; zG: [g_iters][B,nz];
   masks: [(d_iters+g_iters)][og_D_mask_floats(B)] or NULL; outputs: lossD[d_iters], lossG[g_iters],
   d_out: last D-phase sigmoid outputs [B] (may be NULL) */
void og_train_step(og_trainer* t, const og_step_cfg* cfg, const float* real, const float* zD,
                   const float* zG, const float* masks, float* lossD, float* lossG, float* d_out);
/* the two closures separately (adversarial.lua:72-167 and :171-215); return loss incl. penalties */
float og_fevalD(og_trainer* t, const og_step_cfg* cfg, const float* inputs, const float* targets,
                const float* masks, float* d_out);
float og_fevalG_on_D(og_trainer* t, const og_step_cfg* cfg, const float* z, const float* masks);

#ifdef __cplusplus
}
#endif
#endif
