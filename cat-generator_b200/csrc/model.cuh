// model.cuh -- device-side executors for G32up / G32up-c (models.lua:138-160,196-228) and D32_st3
// (models.lua:640-711, STN factory :814-906).  Parameters/gradients: one flat device vector each, in
// nn getParameters() order.  Activations: NHWC fp32.
#pragma once
#include "ops.cuh"

struct cg_layer {            // a conv or a Linear (as 1x1 conv) with its packed operands
  cg::ConvSpec s;
  long oW = 0, ob = 0;       // offsets into the flat parameter vector
  float *Wp = nullptr, *Wd = nullptr, *bp = nullptr;
  bool need_dgrad = true;
};

struct cg_stn {
  int ch, S, rot, scl, trn, nth;
  int c1, c2, l1, l2;        // layer indices
  // saved activations
  const float* in;
  float *pool1, *c1o, *a1, *c2o, *a2, *pool2, *l1o, *al1, *theta, *A, *grid, *out;
  bool fused = false;        // the last forward ran stn_fused.cu (theta then has stride 4; a1, a2, al1, grid are not materialised)
};

struct cg_gstage { int up, Ci, Co, k, bn, layer; long og, obt, opw; };

struct cg_model {
  int kind, C, nz;
  long np = 0;
  float *P = nullptr, *G = nullptr;        // flat params / grads (device)
  std::vector<cg_layer> layers;
  float* packed = nullptr; size_t packed_floats = 0;
  uint8_t* wq = nullptr; cg::PackJob* jobs_dev = nullptr; int njobs = 0, repack_blocks = 0;   // fp16 weight slices + the one-launch repack table
  bool dirty = true;                        // packed operands stale w.r.t. P
  bool dirty32 = true;                      // ... the fp32 ones specifically (skipped while every shape runs on the tensor cores)
  int training = 1;
  int skip_param_grads = 0;                 // backward computes input gradients only (D inside fevalG: its parameter gradients are never read)
  uint64_t seed = 0, rng_offset = 0;          // rng_offset: host counter used only while initialising parameters
  unsigned long long* rng_dev = nullptr;      // Philox offset of the dropout masks, in device memory (graph replay)
  std::vector<cg::DBuf> fw, bw; int nfw = 0, nbw = 0;
  cg::DBuf gwp;                              // packed wgrad scratch (main stream)
  cg::DBuf gwp_lane[cg::Ctx::kLanes];        // ... and one per lane (concurrent branches)
  int B = 0;
  // ---- G
  int C0 = 0, s0 = 0, nst = 0; cg_gstage st[4];
  int lin_layer = -1; long oLpw = 0;
  float* run = nullptr; long nrun = 0;
  const float* z = nullptr; float *lin = nullptr, *act0 = nullptr;
  float *sup[4], *sconv[4], *sbn[4], *sact[4], *smean[4], *sinv[4];
  const uint8_t* sxq[4]; bool sfused[4];      // stage input cached as the tensor-core operand (then sup[i] is null)
  // ---- V (create_V32, forward only)
  int vconv[4], vlin[3]; long vbn[4];         // layer indices; offsets of the four BatchNormalization gamma vectors (beta follows)
  // ---- D
  cg_stn stn[4];
  int t1, t2, b1[4], b2[4], h1, h2;          // layer indices
  long t1pw, t2pw, bpw1[4], bpw2[4], hpw;
  float *xin, *tc1, *ta1, *tc2, *ta2, *tpool, *T, *bc1[4], *ba1[4], *bmp[4], *bdr[4], *bc2[4], *cat, *catd, *h1o, *ha1, *hd, *h2o, *hsig;
  uint8_t* bidx[4];
  bool head_fused = false;                   // the last D forward ran the head (PReLU, Dropout, Linear(256,1), Sigmoid) as one kernel (fuse_d.cu d_head_fwd)
  bool dfused = false;                       // the last D forward ran the fused chains (fuse_d.cu): ta1, ta2, tpool, ba1, bmp, bdr, cat are not materialised
  const uint8_t *xq_t2 = nullptr, *xq_b4 = nullptr, *xq_b2[4] = {nullptr, nullptr, nullptr, nullptr};   // cached conv operands (forward + weight gradient)
  unsigned int* amax = nullptr;              // 16 words: max|gradient| recorded by producers for the next stage's fp16 packing scale (fuse_d.cu)
  float* masks = nullptr; long masks_n = 0; int masks_B = 0;
  float* mq = nullptr; int mq_count = 0, mq_next = 0, mq_B = 0;   // queued user masks (cg_D_set_masks)
};

struct cg_trainer {
  cg_model *G, *D;
  float *mD, *vD, *mG, *vG;
  int* t_dev = nullptr;                       // [tD, tG]: optim.adam step counters, in device memory (graph replay)
  cg::DBuf inputs, targets, samples, dout, df, gimg, scal, stage;
  // CUDA-graph replay of the step (capi.cu): fixed input buffers + one instantiated graph per step configuration
  cg::DBuf gin;
  struct StepGraph { cg_step_cfg cfg; int engine = 0, elim = 0, lanes = 0; int warm = 0; bool failed = false; cudaGraphExec_t exec = nullptr; int64_t launches = 0; uint64_t gen = 0; bool end_dirty_G = true, end_dirty_D = true, end_dirty32_G = true, end_dirty32_D = true; };
  std::vector<StepGraph> graphs;
};

namespace cg {
// fused spatial transformer (stn_fused.cu): parameter / gradient pointers into the flat Torch-layout vectors
struct StnFusedParams { const float *W1, *b1, *W2, *b2, *L1, *lb1, *L2, *lb2; int ch, S, rot, scl, trn, nth;
                        const float *W1p, *W1d, *W2p, *W2d; };   // the convolutions' packed fp32 operands: Wp[(tap,ci)][co], Wd[(flipped tap,co)][ci]
struct StnFusedGrads { float *W1, *b1, *W2, *b2, *L1, *lb1, *L2, *lb2; };
bool stn_fused_shape_ok(int ch, int S);
inline int stn_fused_part_floats(int ch, int nth) { return 16 * ch * 9 + 16 + 16 * 16 * 9 + 16 + 64 + nth * 64 + nth; }
int stn_fused_forward(const StnFusedParams& p, const float* in, int B, float* pool1, float* c1o, float* c2o, float* pool2, float* l1o, float* theta, float* A, float* out);
int stn_fused_backward(const StnFusedParams& p, const StnFusedGrads& g, const float* in, int B, const float* pool1, const float* c1o, const float* c2o, const float* pool2,
                       const float* l1o, const float* theta, const float* A, const float* gout, float* gin, float* ggrid, float* gl1, float* part, int skip_param_grads,
                       unsigned int* amax_out = nullptr);   // amax_out: also record max|gin| (float bits, atomicMax)
// fuse_d.cu: the head of D32_st3 behind Linear(20480,256) in one launch per direction
int d_head_fwd(const float* h1o, const float* pw, const float* mask, const float* W2, const float* b2, float* hd, float* h2o, float* hsig, int B);
int d_head_bwd(const float* gout, const float* hsig, const float* hd, const float* h1o, const float* pw, const float* mask, const float* W2, float* gh1,
               float* gW2, float* gb2, float* gpw, int B, int param_grads);
// fuse_d.cu: PReLU -> [2x2 pool] -> dropout mask -> {dense fp32 / Concat slot / next conv's fp16 operand} in one pass over a conv output
int act_pool_mask_pack(const float* y, const float* pw, int N, int H, int W, int C, int pool, const float* mask, int mask_stride, uint8_t* idx,
                       float* out, int out_stride, int out_off, uint8_t* xq, int k);
size_t act_bwd_operand_bytes(int N, int H, int W, int C, int k);
int act_bwd_pack(const float* const* g, const unsigned int* const* amax, int ng, int g_stride, int g_off, const float* mask, int mask_stride, int pool, const uint8_t* idx,
                 const float* y, const float* pw, int N, int H, int W, int C, int k, uint8_t* gq, float* scale2, double* part, float* gb_acc, float* gpw_acc);
int absmax_into(const float* x, long n, unsigned int* amax);   // conv_tc.cu: *amax = max(*amax, max|x|) as float bits
int model_build(cg_model* m);                       // layout + allocation + init
int model_repack(cg_model* m, int need32 = 1);      // refresh packed operands if stale; need32 = 0: the fp32 operands may stay stale
long D_mask_floats(int B);
// device-pointer executors; boundary tensors are Torch NCHW
int G_forward_dev(cg_model* g, const float* z_dev, int B, float* out_nchw_dev);
int G_backward_dev(cg_model* g, const float* gout_nchw_dev, float* gz_dev);
int D_forward_dev(cg_model* d, const float* x_nchw_dev, int B, float* sig_dev, float* pre_dev);
int D_backward_dev(cg_model* d, const float* gout_dev, float* gx_nchw_dev);
int V_forward_dev(cg_model* v, const float* x_nchw_dev, int B, float* out_dev);
}  // namespace cg
