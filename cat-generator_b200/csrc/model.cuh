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
  // ---- D
  cg_stn stn[4];
  int t1, t2, b1[4], b2[4], h1, h2;          // layer indices
  long t1pw, t2pw, bpw1[4], bpw2[4], hpw;
  float *xin, *tc1, *ta1, *tc2, *ta2, *tpool, *T, *bc1[4], *ba1[4], *bmp[4], *bdr[4], *bc2[4], *cat, *catd, *h1o, *ha1, *hd, *h2o, *hsig;
  uint8_t* bidx[4];
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
  struct StepGraph { cg_step_cfg cfg; int engine = 0, elim = 0, lanes = 0; int warm = 0; bool failed = false; cudaGraphExec_t exec = nullptr; int64_t launches = 0; uint64_t gen = 0; bool end_dirty_G = true, end_dirty_D = true; };
  std::vector<StepGraph> graphs;
};

namespace cg {
int model_build(cg_model* m);                       // layout + allocation + init
int model_repack(cg_model* m);                      // refresh packed operands if dirty
long D_mask_floats(int B);
// device-pointer executors; boundary tensors are Torch NCHW
int G_forward_dev(cg_model* g, const float* z_dev, int B, float* out_nchw_dev);
int G_backward_dev(cg_model* g, const float* gout_nchw_dev, float* gz_dev);
int D_forward_dev(cg_model* d, const float* x_nchw_dev, int B, float* sig_dev, float* pre_dev);
int D_backward_dev(cg_model* d, const float* gout_dev, float* gx_nchw_dev);
}  // namespace cg
