/*
 * catgen.h -- C-ABI of libcatgen.so: hand-written sm_100a CUDA kernels for the DCGAN training hot path of
 * aleju/cat-generator.  This is the drop-in boundary: the reference's Lua modules
 * (models.lua, adversarial.lua, LeakyReLU.lua and the files under layers/) bind these symbols through LuaJIT FFI
 * (cat-generator_b200/lua/catgen_ffi.lua carries the same declarations as an ffi.cdef); Python binds them
 * through ctypes (cat-generator_b200/catgen/lib.py).  See INTEGRATION.md.
 *
 * Conventions (they are the reference's):
 *  - tensors at the boundary are contiguous fp32 in Torch7 layout: NCHW images, conv weights [Cout,Cin,kH,kW],
 *    Linear weights [out,in]; model parameters and gradients are ONE flat vector each, in nn getParameters()
 *    order (train.lua:184-185).
 *  - every `const float*` / `float*` argument is a HOST pointer unless the function name ends in `_dev`;
 *    host memory is borrowed for the duration of the call only.
 *  - all functions return CG_OK (0) or a negative cg_status; cg_last_error() gives the message.  The Lua shim
 *    turns non-zero into error(cg_last_error()), matching the reference's assert/error convention
 *    (layers/SpatialConvolutionUpsample.lua:5-7).
 *  - there is NO CPU fallback: without a CUDA device of compute capability 10.x cg_init fails.
 *  - convolutions are stride 1, pad (k-1)/2, k odd -- the only form the reference instantiates
 *    (models.lua:145-154,206-222,646-685,844-846).
 *  - gradients ACCUMULATE into the flat gradient vector (nn accGradParameters); cg_model_zero_grads is the
 *    caller's job, as in adversarial.lua:81,177.
 */
#ifndef CATGEN_H
#define CATGEN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CG_OK = 0, CG_ERR_CUDA = -1, CG_ERR_ARG = -2, CG_ERR_STATE = -3, CG_ERR_NODEVICE = -4, CG_ERR_NCCL = -5,
  CG_ERR_UNSUPPORTED = -6
} cg_status;

/* models.create_G / models.create_D dispatch targets (models.lua:234-240, :268-277) */
typedef enum {
  CG_G32UP = 0,    /* models.lua:138-160 create_G_decoder_upsampling32  */
  CG_G32UPC = 1,   /* models.lua:196-228 create_G_decoder_upsampling32c (default G at 32x32) */
  CG_D32_ST3 = 2,  /* models.lua:640-711 create_D32_st3 (default D at 32x32) */
  CG_V32 = 3       /* models.lua:765-804 create_V32: the validator network train.lua loads (train.lua:119-123) and NN_UTILS.rateWithV scores with;
                      forward only, evaluate() mode (train.lua:123) */
} cg_model_kind;

typedef struct cg_model cg_model;     /* opaque; owns device parameters, gradients, activations */
typedef struct cg_trainer cg_trainer; /* opaque; owns Adam state for one (G, D) pair */

/* ------------------------------------------------------------------ library */
/* device: CUDA ordinal (train.lua:109 cutorch.setDevice(OPT.gpu+1) -> OPT.gpu).  Idempotent per device. */
int         cg_init(int device);
void        cg_shutdown(void);
const char* cg_last_error(void);
const char* cg_version(void);
int         cg_sync(void);                     /* cutorch.synchronize() */
/* number of kernels of THIS library launched since the last cg_reset_launch_count (bench.py gpu_launches) */
int64_t     cg_launch_count(void);
void        cg_reset_launch_count(void);
/* timing on the library's own CUDA stream (CUDA events): replaces the sys.clock() bracket of adversarial.lua:34,278-280 */
int         cg_timer_start(void);
int         cg_timer_stop(float* ms);
/* per-launch event timing of every kernel + its algorithmic flops/bytes; report is a JSON array written to out */
int         cg_profile_enable(int on);
int         cg_profile_report(char* out, int cap);
/* JSON array, one row per launch since cg_profile_enable(1): ["kernel", lane, start_us, end_us] (lane -1 = the main stream) */
int         cg_profile_timeline(char* out, int cap);
/* 1 (default): cg_train_step* replay the step as a CUDA graph once a configuration has run twice eagerly; 0: always eager.
   Results are the same either way (Adam's step count and the dropout RNG offset live in device memory). */
int         cg_set_graph_mode(int on);
int         cg_get_graph_mode(void);
/* conv engine for shapes the tensor-core path supports: 0 = fp32 CUDA-core fallback only, 1 = tcgen05 (default) */
int         cg_set_conv_engine(int engine);
int         cg_get_conv_engine(void);
/* Forward operand precision of the tensor-core engine.  0 (default): activations and weights are rounded to fp16 (fp32 accumulate); generated
   pixels stay within 1e-3 of the fp32 oracle, but rounding moves PReLU / max-pool decisions and the gradients carry ~1e-2 relative noise
   (profiles/r01_grad_diag.txt).  1: every forward convolution of G and D splits both operands into hi + lo fp16 halves and accumulates
   hi*hi + lo*hi + hi*lo on the tensor cores (3x the forward MMA work, ~2^-21 relative operand error); nn.Linear layers and the fused
   forward producers fall back to fp32 / unfused kernels.  Backward passes are unchanged (scaled fp16 gradients, already ~5e-4). */
int         cg_set_precision(int mode);
int         cg_get_precision(void);
/* cg_train_step*: inside fevalG the reference's MODEL_D:backward (adversarial.lua:193) also accumulates D's parameter gradients,
   which the next fevalD zeroes unread (adversarial.lua:78).  1 (default): that dead accumulation is skipped -- parameters,
   optimiser state, losses and d_out are bit-identical either way; 0: keep it, so D's gradient vector after a step holds
   what Torch's gradParameters would. */
int         cg_set_dead_grad_elim(int on);
/* 1 (default): independent parts of a step are issued on concurrent streams -- D32_st3's four transformer branches, each layer's
   weight-gradient chain beside its input-gradient conv, fevalG's generator forward beside fevalD -- joined by events (which graph
   capture records as parallel paths).  0: everything on one stream, in program order (per-kernel timing, debugging).
   Environment: CATGEN_LANES=0 / CATGEN_SIDE=0 set the initial state of the two mechanisms separately. */
int         cg_set_concurrency(int on);

/* ------------------------------------------------------------------ models */
/* replaces models.create_G(dimensions, noiseDim) / models.create_D(dimensions, cuda); H=W=32.
   Parameters are initialised per weight-init.lua:40-75 + nn defaults from `seed`. */
int cg_model_create(cg_model** out, int kind, int C, int nz, uint64_t seed);
int cg_model_free(cg_model* m);
/* replaces MODEL:getParameters() (train.lua:184-185): length and host copies of the flat vectors */
int cg_model_nparams(const cg_model* m, int64_t* n);
int cg_model_get_params(cg_model* m, float* host);
int cg_model_set_params(cg_model* m, const float* host);
int cg_model_get_grads(cg_model* m, float* host);
int cg_model_zero_grads(cg_model* m);                         /* GRAD_PARAMETERS:zero(), adversarial.lua:81,177 */
/* G only: SpatialBatchNormalization running mean/var, [mean_l, var_l] per BN layer (nn [upstream], A.3) */
int cg_model_bn_running_len(const cg_model* m, int64_t* n);
int cg_model_get_bn_running(cg_model* m, float* host);
int cg_model_set_bn_running(cg_model* m, const float* host);
/* replaces MODEL:training() / MODEL:evaluate() (utils/nn_utils.lua:334-349): 1 = training */
int cg_model_set_mode(cg_model* m, int training);

/* replaces MODEL_G:forward(noise) (utils/nn_utils.lua:52, adversarial.lua:185): z [B,nz] -> out [B,C,32,32] */
int cg_G_forward(cg_model* g, const float* z, int B, float* out);
/* replaces MODEL_G:backward(noise, df_do) (adversarial.lua:197): gout [B,C,32,32]; gz [B,nz] or NULL */
int cg_G_backward(cg_model* g, const float* gout, float* gz);
/* replaces MODEL_D:forward(inputs) (adversarial.lua:84,187): x [B,C,32,32] -> sigmoid outputs [B] and the
   pre-sigmoid Linear(256,1) outputs [B] (either may be NULL) */
int cg_D_forward(cg_model* d, const float* x, int B, float* out_sig, float* out_pre);
/* replaces MODEL_D:backward(inputs, df_do) and MODEL_D.modules[1].gradInput (adversarial.lua:89,192-193):
   gout [B] = d loss / d sigmoid output; gx [B,C,32,32] or NULL */
int cg_D_backward(cg_model* d, const float* gout, float* gx);
/* dropout multipliers of D's 7 dropout layers for the LAST forward, in the oracle's layout
   (trunk[B*64], br1..3[B*64], br4[B*128], head[B*320], fc[B*256]); nn.SpatialDropout / nn.Dropout
   (models.lua:651,658,667,676,684,695,699).  set_ queues `count` mask sets ([count][cg_D_mask_floats(B)]); the
   next `count` forwards of D consume them in order (parity tests replay a whole training step this way);
   otherwise masks come from a Philox stream seeded by cg_model_create's seed. */
/* replaces MODEL_V:forward(images) in NN_UTILS.rateWithV (utils/nn_utils.lua:686-711): x [B,C,32,32] -> SoftMax outputs [B,2]; column 0 is
   P(fake).  evaluate() mode only: BatchNormalization uses the running statistics (cg_model_set_bn_running: [mean_l, var_l] for V's four BN
   layers, 128 + 256 + 1024 + 1024 channels), Dropout is the identity and SpatialDropout scales by 1 - p = 0.5. */
int cg_V_forward(cg_model* v, const float* x, int B, float* out);
int cg_D_mask_floats(int B, int64_t* n);
int cg_D_get_masks(cg_model* d, float* host);
int cg_D_set_masks(cg_model* d, const float* host, int B, int count);

/* ------------------------------------------------------------------ criterion / optimiser / step */
/* nn.BCECriterion forward+backward (train.lua:181; adversarial.lua:85,88,188,191): p,t [n] -> *loss, g [n] */
int cg_bce(const float* p, const float* t, int n, float* loss, float* g);
/* L1/L2 penalty + clamp on the flat gradient (adversarial.lua:92-98,110-112 for D; :201-212 for G, whose
   sign term is scaled by l2sign = G_L2, reproducing line 206).  *loss_add gets L1*|p|_1 + L2*|p|_2^2/2. */
int cg_penalty_clamp(cg_model* m, float l1, float l2sign, float l2, float clampv, float* loss_add);

typedef struct {
  int   B, d_iters, g_iters;               /* OPT.batchSize, OPT.D_iterations, OPT.G_iterations (train.lua:26,33-34) */
  float D_L1, D_L2, G_L1, G_L2;            /* train.lua:28-31 */
  float D_clamp, G_clamp;                  /* train.lua:35-36 */
  float lr, beta1, beta2, eps;             /* optim.adam defaults */
} cg_step_cfg;

int cg_trainer_create(cg_trainer** out, cg_model* G, cg_model* D);
int cg_trainer_free(cg_trainer* t);
/* optim.adam(feval, PARAMETERS, OPTSTATE.adam.X) update part (adversarial.lua:245,262), on the model's
   current flat gradient; which: 0 = D, 1 = G */
int cg_adam_step(cg_trainer* t, int which, const cg_step_cfg* cfg);
/* one iteration of the adversarial.train loop body (adversarial.lua:221-266) with the data the Lua loop
   gathers per step: real [d_iters][B/2,C,32,32], zD [d_iters][B/2,nz], zG [g_iters][B,nz];
   outputs lossD[d_iters], lossG[g_iters], d_out [B] = D's outputs of the last D update (confusion matrix,
   adversarial.lua:101-106); any output may be NULL.  With >1 rank (cg_dist_init) each rank passes its shard. */
int cg_train_step(cg_trainer* t, const cg_step_cfg* cfg, const float* real, const float* zD, const float* zG,
                  float* lossD, float* lossG, float* d_out);
/* same step with inputs already resident in device memory (bench.py `value`); outputs stay on device
   except the scalars */
int cg_train_step_dev(cg_trainer* t, const cg_step_cfg* cfg, const float* real_dev, const float* zD_dev,
                      const float* zG_dev, float* lossD, float* lossG);
void* cg_dev_alloc(int64_t bytes);            /* helpers for callers that keep inputs in HBM */
int   cg_dev_free(void* p);
int   cg_dev_upload(void* dst_dev, const void* src_host, int64_t bytes);
int   cg_dev_download(void* dst_host, const void* src_dev, int64_t bytes);
/* page-locked host memory (cudaMallocHost): what a caller stages its batches in so that the H2D copies of cg_train_step run at
   link speed and asynchronously (bench.py's end-to-end region) */
void* cg_host_alloc(int64_t bytes);
int   cg_host_free(void* p);
/* U(lo,hi) noise on the device: NN_UTILS.createNoiseInputs (utils/nn_utils.lua:35-39) */
int   cg_uniform_dev(float* dst_dev, int64_t n, float lo, float hi, uint64_t seed, uint64_t offset);

/* ------------------------------------------------------------------ data parallel (new) */
/* one process per GPU; nccl_id: 128-byte ncclUniqueId from rank 0 (cg_dist_unique_id).  After this, the
   cg_train_step* calls all-reduce D's then G's flat gradients (sum, scaled 1/world) before penalty+clamp+Adam. */
int cg_dist_unique_id(char id_out[128]);
int cg_dist_init(int rank, int world, const char id[128]);
int cg_dist_allreduce_grads(cg_model* m);
int cg_dist_world(void);
/* Batch norm under data parallelism (SURVEY.md section 8e "BN caveat").  0 (default): G's three SpatialBatchNormalization layers use
   the statistics of each rank's LOCAL batch -- no extra collective, but not what one device with the whole batch computes.
   1: sync-BN -- per-channel (sum, sum of squares) and the two backward sums are all-reduced (2*C doubles each), so an R-rank step
   equals the single-device step on the concatenated batch. */
int cg_dist_set_sync_bn(int on);
int cg_dist_get_sync_bn(void);

/* ------------------------------------------------------------------ op level (one call per nn.Module method) */
/* nn.SpatialConvolution / cudnn.SpatialConvolution updateOutput, updateGradInput, accGradParameters */
int cg_conv2d_fprop(const float* x, const float* W, const float* b, float* y, int N, int Ci, int H, int Wd, int Co, int k);
int cg_conv2d_dgrad(const float* gy, const float* W, float* gx, int N, int Ci, int H, int Wd, int Co, int k);
int cg_conv2d_wgrad(const float* x, const float* gy, float* gW, float* gb, int N, int Ci, int H, int Wd, int Co, int k);
/* nn.SpatialConvolutionUpsample:updateOutput (layers/SpatialConvolutionUpsample.lua:16-28) and its cudnn twin
   (layers/cudnnSpatialConvolutionUpsample.lua): conv to nOut*f*f planes; the :view is a no-op on memory, so
   y is [N, nOut, H*f, W*f] contiguous.  Errors (CG_ERR_ARG) when k is even, like the Lua asserts (:5-7). */
int cg_conv_upsample_fwd(const float* x, const float* W, const float* b, float* y, int N, int Ci, int H, int Wd, int nOut, int k, int f);
int cg_conv_upsample_bwd(const float* x, const float* gy, const float* W, float* gx, float* gW, float* gb, int N, int Ci, int H, int Wd, int nOut, int k, int f);
/* nn.Linear */
int cg_linear_fwd(const float* x, const float* W, const float* b, float* y, int N, int in, int out);
int cg_linear_bwd(const float* x, const float* gy, const float* W, float* gx, float* gW, float* gb, int N, int in, int out);
/* nn.SpatialBatchNormalization, training mode (eps 1e-5, momentum 0.1) */
int cg_bn2d_fwd(const float* x, const float* gamma, const float* beta, float* y, float* save_mean, float* save_invstd,
                float* run_mean, float* run_var, int N, int C, int HW);
int cg_bn2d_bwd(const float* x, const float* gy, const float* gamma, const float* save_mean, const float* save_invstd,
                float* gx, float* ggamma, float* gbeta, int N, int C, int HW);
/* nn.PReLU() single shared slope */
int cg_prelu_fwd(const float* x, float w, float* y, int64_t n);
int cg_prelu_bwd(const float* x, const float* gy, float w, float* gx, float* gw, int64_t n);
/* nn.LeakyReLU:updateOutput / updateGradInput (LeakyReLU.lua:13-19, :21-31); gradient at x==0 is gy */
int cg_leakyrelu_fwd(const float* x, float slope, float* y, int64_t n);
int cg_leakyrelu_bwd(const float* x, const float* gy, float slope, float* gx, int64_t n);
/* nn.SpatialUpSamplingNearest(2), nn.Sigmoid, nn.SpatialAveragePooling(2,2,2,2), nn.SpatialMaxPooling(2,2) */
int cg_upsample2x_fwd(const float* x, float* y, int NC, int H, int Wd);
int cg_upsample2x_bwd(const float* gy, float* gx, int NC, int H, int Wd);
int cg_sigmoid_fwd(const float* x, float* y, int64_t n);
int cg_sigmoid_bwd(const float* y, const float* gy, float* gx, int64_t n);
int cg_avgpool2_fwd(const float* x, float* y, int NC, int H, int Wd);
int cg_avgpool2_bwd(const float* gy, float* gx, int NC, int H, int Wd);
int cg_maxpool2_fwd(const float* x, float* y, int32_t* idx, int NC, int H, int Wd);
int cg_maxpool2_bwd(const float* gy, const int32_t* idx, float* gx, int NC, int H, int Wd);
/* stn: nn.AffineTransformMatrixGenerator, nn.AffineGridGeneratorBHWD, nn.BilinearSamplerBHWD (models.lua:877-888) */
int cg_affine_matrix_fwd(const float* theta, float* A, int B, int rot, int scl, int trn);
int cg_affine_matrix_bwd(const float* theta, const float* gA, float* gtheta, int B, int rot, int scl, int trn);
int cg_affine_grid_fwd(const float* A, float* grid, int B, int H, int Wd);
int cg_affine_grid_bwd(const float* ggrid, float* gA, int B, int H, int Wd);
int cg_bilinear_fwd(const float* img_nhwc, const float* grid, float* out_nhwc, int B, int H, int Wd, int C);
int cg_bilinear_bwd(const float* img_nhwc, const float* grid, const float* gout_nhwc, float* gimg_nhwc, float* ggrid,
                    int B, int H, int Wd, int C);

#ifdef __cplusplus
}
#endif
#endif
