// model.cu -- G / D executors.  Layer graphs follow models.lua:138-160 (G32up), :196-228 (G32up-c),
// :640-711 (D32_st3) and :814-906 (spatial transformer); initialisation follows weight-init.lua:40-75 and the
// nn defaults.
#include "model.cuh"
#include <math.h>

namespace cg {

static float* FW(cg_model* m, size_t n) {
  if ((int)m->fw.size() <= m->nfw) m->fw.resize(m->nfw + 16);
  DBuf& b = m->fw[m->nfw++];
  return b.ensure(n) == CG_OK ? b.p : nullptr;
}
static float* BW(cg_model* m, size_t n) {
  if ((int)m->bw.size() <= m->nbw) m->bw.resize(m->nbw + 16);
  DBuf& b = m->bw[m->nbw++];
  return b.ensure(n) == CG_OK ? b.p : nullptr;
}
#define NN(p) do { if (!(p)) return cg::set_err(CG_ERR_CUDA, "%s:%d device allocation failed", __FILE__, __LINE__); } while (0)

static int add_layer(cg_model* m, int Ci, int Co, int k, long& o, int in_hw = 1, int out_hw = 1, bool need_dgrad = true) {
  cg_layer L; L.s.Ci = Ci; L.s.Co = Co; L.s.k = k; L.s.in_hw = in_hw; L.s.out_hw = out_hw; L.need_dgrad = need_dgrad;
  L.oW = o; o += (long)Co * Ci * k * k; L.ob = o; o += Co;
  m->layers.push_back(L);
  return (int)m->layers.size() - 1;
}
static long stn_layout(cg_model* m, cg_stn* s, long o, int ch, int S, int rot, int scl, int trn, bool first) {
  s->ch = ch; s->S = S; s->rot = rot; s->scl = scl; s->trn = trn; s->nth = (rot ? 1 : 0) + (scl ? 1 : 0) + (trn ? 2 : 0);
  int S4 = S / 4;
  // nn has no "needs input grad" switch: dgrad is computed for every layer incl. the first
  (void)first;
  s->c1 = add_layer(m, ch, 16, 3, o);
  s->c2 = add_layer(m, 16, 16, 3, o);
  s->l1 = add_layer(m, 16 * S4 * S4, 64, 1, o, S4 * S4, 1);
  s->l2 = add_layer(m, 64, s->nth, 1, o);
  return o;
}

long D_mask_floats(int B) { return (long)B * (64 * 4 + 128 + 320 + 256); }

static int init_uniform(cg_model* m, long off, long n, float lo, float hi) {
  int s = uniform(m->P + off, n, lo, hi, m->seed, m->rng_offset);
  m->rng_offset += (uint64_t)(n + 3) / 4 + 1;
  return s;
}
static int init_layer(cg_model* m, int li, bool random_bias) {
  cg_layer& L = m->layers[li];
  long nW = (long)L.s.Co * L.s.Ci * L.s.k * L.s.k;
  float sd = 1.f / sqrtf((float)L.s.Ci * L.s.k * L.s.k);
  CG_TRY(init_uniform(m, L.oW, nW, -sd, sd));
  if (random_bias) CG_TRY(init_uniform(m, L.ob, L.s.Co, -sd, sd));
  else CG_TRY(fill(m->P + L.ob, 0.f, L.s.Co));
  return CG_OK;
}
static int init_stn(cg_model* m, cg_stn* s) {   // models.lua:843-860
  CG_TRY(init_layer(m, s->c1, false)); CG_TRY(init_layer(m, s->c2, false)); CG_TRY(init_layer(m, s->l1, false));
  cg_layer& L = m->layers[s->l2];
  CG_TRY(fill(m->P + L.oW, 0.f, (long)s->nth * 64));
  float b[4] = {0, 0, 0, 0}; int q = 0;
  if (s->rot) b[q++] = 0.f;
  if (s->scl) b[q++] = 1.f;
  if (s->trn) { b[q++] = 0.f; b[q++] = 0.f; }
  CG_CUDA(cudaMemcpyAsync(m->P + L.ob, b, sizeof(float) * s->nth, cudaMemcpyHostToDevice, ctx().stream));
  CG_CUDA(cudaStreamSynchronize(ctx().stream));   // b is a stack array
  return CG_OK;
}

int model_build(cg_model* m) {
  long o = 0; int C = m->C;
  if (m->kind == CG_G32UP || m->kind == CG_G32UPC) {
    if (m->kind == CG_G32UPC) {
      m->C0 = 512; m->s0 = 4; m->nst = 4;
      cg_gstage s[4] = {{1, 512, 512, 3, 1}, {1, 512, 256, 3, 1}, {1, 256, 128, 5, 1}, {0, 128, C, 3, 0}};
      memcpy(m->st, s, sizeof(s));
    } else {
      m->C0 = 128; m->s0 = 8; m->nst = 3;
      cg_gstage s[3] = {{1, 128, 256, 5, 1}, {1, 256, 128, 5, 1}, {0, 128, C, 3, 0}};
      memcpy(m->st, s, sizeof(s));
    }
    int hw0 = m->s0 * m->s0;
    m->lin_layer = add_layer(m, m->nz, m->C0 * hw0, 1, o, 1, hw0);
    m->oLpw = o; o += 1;
    m->nrun = 0;
    for (int i = 0; i < m->nst; ++i) {
      cg_gstage& s = m->st[i];
      s.layer = add_layer(m, s.Ci, s.Co, s.k, o);
      if (s.bn) { s.og = o; o += s.Co; s.obt = o; o += s.Co; s.opw = o; o += 1; m->nrun += 2 * s.Co; }
    }
  } else if (m->kind == CG_D32_ST3) {
    o = stn_layout(m, &m->stn[0], o, C, 32, 1, 0, 0, true);
    m->t1 = add_layer(m, C, 64, 3, o); m->t1pw = o; o += 1;
    m->t2 = add_layer(m, 64, 64, 3, o); m->t2pw = o; o += 1;
    for (int b = 0; b < 3; ++b) {
      o = stn_layout(m, &m->stn[b + 1], o, 64, 16, 1, 1, 1, false);
      m->b1[b] = add_layer(m, 64, 64, 3, o); m->bpw1[b] = o; o += 1;
      m->b2[b] = add_layer(m, 64, 64, 3, o); m->bpw2[b] = o; o += 1;
    }
    m->b1[3] = add_layer(m, 64, 128, 5, o); m->bpw1[3] = o; o += 1;
    m->b2[3] = add_layer(m, 128, 128, 7, o); m->bpw2[3] = o; o += 1;
    m->h1 = add_layer(m, 20480, 256, 1, o, 64, 1); m->hpw = o; o += 1;
    m->h2 = add_layer(m, 256, 1, 1, o);
  } else if (m->kind == CG_V32) {   // models.lua:765-804
    const int vc[4][2] = {{C, 128}, {128, 128}, {128, 256}, {256, 256}};
    m->nrun = 0;
    for (int i = 0; i < 4; ++i) {
      m->vconv[i] = add_layer(m, vc[i][0], vc[i][1], 3, o, 1, 1, false);
      if (i == 1 || i == 3) { m->vbn[i / 2] = o; o += 2 * vc[i][1]; m->nrun += 2 * vc[i][1]; }          // SpatialBatchNormalization after conv 2 and conv 4
    }
    m->vlin[0] = add_layer(m, 4096, 1024, 1, o, 16, 1, false); m->vbn[2] = o; o += 2048; m->nrun += 2048;   // View(256*4*4) -> Linear -> BatchNormalization(1024)
    m->vlin[1] = add_layer(m, 1024, 1024, 1, o, 1, 1, false); m->vbn[3] = o; o += 2048; m->nrun += 2048;
    m->vlin[2] = add_layer(m, 1024, 2, 1, o, 1, 1, false);
  } else return set_err(CG_ERR_ARG, "unknown model kind %d", m->kind);
  m->np = o;
  if (m->kind == CG_D32_ST3) { CG_CUDA(cudaMalloc(&m->amax, sizeof(unsigned int) * 16)); CG_CUDA(cudaMemsetAsync(m->amax, 0, sizeof(unsigned int) * 16, ctx().stream)); }
  CG_CUDA(cudaMalloc(&m->P, sizeof(float) * o));
  CG_CUDA(cudaMalloc(&m->G, sizeof(float) * o));
  CG_CUDA(cudaMemsetAsync(m->G, 0, sizeof(float) * o, ctx().stream));
  // packed operands: Wp, Wd, bp per layer, each 16-byte aligned
  size_t pf = 0;
  for (auto& L : m->layers) { size_t nW = (size_t)L.s.Co * L.s.Ci * L.s.k * L.s.k; pf += 2 * ((nW + 3) & ~(size_t)3) + (((size_t)L.s.Co + 3) & ~(size_t)3); }
  m->packed_floats = pf;
  CG_CUDA(cudaMalloc(&m->packed, sizeof(float) * pf));
  size_t q = 0; size_t maxW = 0;
  for (auto& L : m->layers) {
    size_t nW = (size_t)L.s.Co * L.s.Ci * L.s.k * L.s.k, nWa = (nW + 3) & ~(size_t)3;
    L.Wp = m->packed + q; q += nWa; L.Wd = m->packed + q; q += nWa; L.bp = m->packed + q; q += ((size_t)L.s.Co + 3) & ~(size_t)3;
    if (nW > maxW) maxW = nW;
  }
  CG_TRY(m->gwp.ensure(maxW + 32768));
  // one-launch repack: a job per layer, with fp16 weight slices for the layers/directions the tensor-core engine takes
  {
    std::vector<cg::PackJob> jobs(m->layers.size());
    size_t wq_total = 0;
    for (size_t i = 0; i < m->layers.size(); ++i) {
      cg_layer& L = m->layers[i]; cg::PackJob& J = jobs[i];
      J.W = m->P + L.oW; J.b = m->P + L.ob; J.Wp = L.Wp; J.Wd = L.Wd; J.bp = L.bp; J.s = L.s; J.need_dgrad = L.need_dgrad ? 1 : 0;
      J.wqf = J.wqd = nullptr; J.CBf = J.CBd = 0; J.always32 = 0;
      { long nW = (long)L.s.Co * L.s.Ci * L.s.k * L.s.k; J.blk0 = m->repack_blocks; J.nblk = (int)((nW + 2047) / 2048); m->repack_blocks += J.nblk; }
      size_t bf = 0, bd = 0;
      const bool plain = L.s.in_hw == 1 && L.s.out_hw == 1;     // Linear layers beside an nn.View (k = 1) are taken by the round-2 engine only (CB == 32)
      if (conv_tc_wslice_plan(L.s.Ci, L.s.Co, L.s.k, &J.CBf, &bf) && (plain || J.CBf == 32)) { J.wqf = (uint8_t*)(uintptr_t)(wq_total + 1); wq_total += (bf + 255) & ~(size_t)255; }
      if (L.need_dgrad && conv_tc_wslice_plan(L.s.Co, L.s.Ci, L.s.k, &J.CBd, &bd) && (plain || J.CBd == 32)) { J.wqd = (uint8_t*)(uintptr_t)(wq_total + 1); wq_total += (bd + 255) & ~(size_t)255; }
    }
    if (m->kind == CG_D32_ST3) for (int q = 0; q < 4; ++q) { jobs[m->stn[q].c1].always32 = 1; jobs[m->stn[q].c2].always32 = 1; }   // read by stn_fused.cu in fp32
    if (wq_total) CG_CUDA(cudaMalloc(&m->wq, wq_total));
    for (size_t i = 0; i < jobs.size(); ++i) {   // offsets (+1 so that offset 0 is distinguishable from "none") -> pointers
      cg::PackJob& J = jobs[i];
      if (J.wqf) { J.wqf = m->wq + ((uintptr_t)J.wqf - 1); conv_tc_register_wslices(J.Wp, J.wqf, J.CBf); }
      if (J.wqd) { J.wqd = m->wq + ((uintptr_t)J.wqd - 1); conv_tc_register_wslices(J.Wd, J.wqd, J.CBd); }
    }
    m->njobs = (int)jobs.size();
    CG_CUDA(cudaMalloc(&m->jobs_dev, sizeof(cg::PackJob) * jobs.size()));
    CG_CUDA(cudaMemcpy(m->jobs_dev, jobs.data(), sizeof(cg::PackJob) * jobs.size(), cudaMemcpyHostToDevice));
  }
  if (m->nrun) {
    CG_CUDA(cudaMalloc(&m->run, sizeof(float) * m->nrun));
    long r = 0;
    for (int i = 0; i < m->nst; ++i) if (m->st[i].bn) {
      CG_TRY(fill(m->run + r, 0.f, m->st[i].Co)); CG_TRY(fill(m->run + r + m->st[i].Co, 1.f, m->st[i].Co)); r += 2 * m->st[i].Co;
    }
    if (m->kind == CG_V32) { const int bc[4] = {128, 256, 1024, 1024}; for (int i = 0; i < 4; ++i) { CG_TRY(fill(m->run + r, 0.f, bc[i])); CG_TRY(fill(m->run + r + bc[i], 1.f, bc[i])); r += 2 * bc[i]; } }
  }
  // ---- initialisation (distribution parity only; parity tests exchange parameters as data)
  const float quarter = 0.25f;
  auto set_scalar = [&](long off, float v) -> int { return fill(m->P + off, v, 1); };
  if (m->kind == CG_V32) {   // weight-init 'heuristic' on the top-level modules: W ~ U(+-1/sqrt(fan_in)), zero bias; BN gamma ~ U(0,1), beta 0
    const int bc[4] = {128, 256, 1024, 1024};
    for (int i = 0; i < 4; ++i) CG_TRY(init_layer(m, m->vconv[i], false));
    for (int i = 0; i < 3; ++i) CG_TRY(init_layer(m, m->vlin[i], false));
    for (int i = 0; i < 4; ++i) { CG_TRY(init_uniform(m, m->vbn[i], bc[i], 0.f, 1.f)); CG_TRY(fill(m->P + m->vbn[i] + bc[i], 0.f, bc[i])); }
  } else if (m->kind != CG_D32_ST3) {
    CG_TRY(init_layer(m, m->lin_layer, false)); CG_TRY(set_scalar(m->oLpw, quarter));
    for (int i = 0; i < m->nst; ++i) {
      cg_gstage& s = m->st[i];
      CG_TRY(init_layer(m, s.layer, false));   // weight-init.lua:70-72 zeroes the bias of every top-level module
      if (s.bn) { CG_TRY(init_uniform(m, s.og, s.Co, 0.f, 1.f)); CG_TRY(fill(m->P + s.obt, 0.f, s.Co)); CG_TRY(set_scalar(s.opw, quarter)); }
    }
  } else {
    CG_TRY(init_stn(m, &m->stn[0]));
    CG_TRY(init_layer(m, m->t1, false)); CG_TRY(set_scalar(m->t1pw, quarter));
    CG_TRY(init_layer(m, m->t2, false)); CG_TRY(set_scalar(m->t2pw, quarter));
    for (int b = 0; b < 4; ++b) {
      if (b < 3) CG_TRY(init_stn(m, &m->stn[b + 1]));
      // nested inside nn.Concat => untouched by weight-init: nn default reset(), random bias
      CG_TRY(init_layer(m, m->b1[b], true)); CG_TRY(set_scalar(m->bpw1[b], quarter));
      CG_TRY(init_layer(m, m->b2[b], true)); CG_TRY(set_scalar(m->bpw2[b], quarter));
    }
    CG_TRY(init_layer(m, m->h1, false)); CG_TRY(set_scalar(m->hpw, quarter));
    CG_TRY(init_layer(m, m->h2, false));
  }
  // dropout masks continue the Philox stream after the initialisation draws
  CG_CUDA(cudaMalloc(&m->rng_dev, sizeof(unsigned long long)));
  { unsigned long long v = m->rng_offset; CG_CUDA(cudaMemcpyAsync(m->rng_dev, &v, sizeof(v), cudaMemcpyHostToDevice, ctx().stream)); CG_CUDA(cudaStreamSynchronize(ctx().stream)); }
  m->dirty = true;
  return CG_OK;
}

// Measured (profiles/r02 bench): the full repack was 143 us per network, twice per step on the critical path, two thirds of it the fp32
// operands that only the CUDA-core fallback reads.  They are refreshed only when a caller may take that fallback.
int model_repack(cg_model* m, int need32) {
  if (m->dirty) {
    CG_TRY(repack_model(m->jobs_dev, m->njobs, m->repack_blocks, need32 ? 3 : 2));   // one launch
    m->dirty = false; m->dirty32 = !need32;
  } else if (need32 && m->dirty32) {
    CG_TRY(repack_model(m->jobs_dev, m->njobs, m->repack_blocks, 1));
    m->dirty32 = false;
  }
  return CG_OK;
}

// x: [N,H,W,Ci] -> y: [N,H,W,Co]
// wgrad scratch of the stream this call issues on (branches running in lanes must not share one)
static float* gw_scratch(cg_model* m) {
  int lane = ctx().lane;
  if (lane < 0) return m->gwp.p;
  if (m->gwp_lane[lane].ensure(m->gwp.n) != CG_OK) return nullptr;
  return m->gwp_lane[lane].p;
}
struct LaneScope {   // routes the enclosed launches to lane b; leaves the lane on every exit path
  int status; bool in;
  explicit LaneScope(int b) : status(lane_enter(b)), in(status == CG_OK) {}
  ~LaneScope() { if (in) lane_exit(); }
};
static int layer_fwd(cg_model* m, int li, const float* x, float* y, int N, int H, int W) {
  cg_layer& L = m->layers[li];
  return conv_fwd(x, L.Wp, L.bp, y, N, H, W, L.s.Ci, L.s.Co, L.s.k);
}
// accumulates dW, db into the flat gradient; gx may be null
static int layer_bwd(cg_model* m, int li, const float* x, const float* gy, float* gx, int N, int H, int W, const uint8_t* xq = nullptr) {   // xq: cached operand for x
  cg_layer& L = m->layers[li];
  if (m->skip_param_grads) return gx ? conv_dgrad(gy, L.Wd, gx, N, H, W, L.s.Ci, L.s.Co, L.s.k) : CG_OK;   // same dgrad kernel and operand as below
  long M = (long)N * H * W;
  float* gwp = gw_scratch(m); NN(gwp);
  // plain convolutions let the engine add straight into the Torch-layout gradient; Linear layers beside an nn.View need the permuting unpack
  float* gW_direct = (L.s.in_hw == 1 && L.s.out_hw == 1) ? m->G + L.oW : nullptr;
  int direct = 0;
  int bias_done = 0;
  if (gx) CG_TRY(conv_backward(x, gy, L.Wd, gwp, gx, N, H, W, L.s.Ci, L.s.Co, L.s.k, gW_direct, &direct, xq, L.s.out_hw == 1 ? m->G + L.ob : nullptr, &bias_done));
  else CG_TRY(conv_wgrad(x, gy, gwp, N, H, W, L.s.Ci, L.s.Co, L.s.k, gW_direct, &direct));
  if (!direct) CG_TRY(unpack_wgrad_acc(gwp, m->G + L.oW, L.s));
  if (bias_done) {}   // the engine's single pass over gy produced it
  else if (L.s.out_hw == 1) CG_TRY(colsum_acc(gy, m->G + L.ob, M, L.s.Co));
  else {
    float* tmp = gwp;   // wgrad scratch is free again (stream ordered)
    CG_TRY(fill(tmp, 0.f, L.s.Co)); CG_TRY(colsum_acc(gy, tmp, M, L.s.Co)); CG_TRY(unpack_bias_acc(tmp, m->G + L.ob, L.s));
  }
  return CG_OK;
}

// =================================================================== G
// A stage whose input can be produced straight into the tensor-core operand: PReLU(BN(.)) -> upsample -> fp16 pack in one pass
// (conv_tc.cu: bn_prelu_up_pack); forward and weight gradient then share that buffer and no fp32 copy of the input exists.
static bool g_stage_fused(const cg_model* g, int i, int h_in) {
  if (i >= g->nst || ctx().conv_engine != 1 || !g->training || ctx().precision) return false;   // compensated operands are packed by the conv call itself
  const cg_gstage& s = g->st[i]; int H = s.up ? 2 * h_in : h_in;
  return conv_tc_cached_ok(H, H, s.Ci, s.Co, s.k);
}

// cg_set_precision(1): forward executors run their convolutions with compensated operands (conv_tc.cu, k_pack_act_split)
struct SplitScope { bool on; SplitScope() : on(ctx().precision == 1 && ctx().conv_engine == 1) { if (on) ctx().split_fwd++; } ~SplitScope() { if (on) ctx().split_fwd--; } };
struct Fp32StaleScope { bool on; explicit Fp32StaleScope(bool stale) : on(stale) { if (on) ctx().fp32_operands_stale++; } ~Fp32StaleScope() { if (on) ctx().fp32_operands_stale--; } };
int G_forward_dev(cg_model* g, const float* z_dev, int B, float* out_nchw) {
  SplitScope split_scope;
  CG_TRY(model_repack(g, !conv_tc_all_shapes_taken(B)));
  Fp32StaleScope stale_scope(g->dirty32);
  g->nfw = 0; g->B = B;
  long F0 = (long)g->C0 * g->s0 * g->s0;
  float* zc = FW(g, (size_t)B * g->nz); NN(zc);
  CG_CUDA(cudaMemcpyAsync(zc, z_dev, sizeof(float) * (size_t)B * g->nz, cudaMemcpyDeviceToDevice, ctx().stream));
  g->z = zc;
  g->lin = FW(g, B * F0); NN(g->lin);
  CG_TRY(layer_fwd(g, g->lin_layer, zc, g->lin, B, 1, 1));            // nn.Linear(nz, C0*s0*s0); output already NHWC
  int h = g->s0; long r = 0;
  bool fuse = g_stage_fused(g, 0, h);
  const float* cur = nullptr;          // fp32 input of the next stage (unfused path)
  const float* src = g->lin; int src_bn = -1;   // fused path: the tensor whose (BN +) PReLU is the next stage's input
  if (!fuse) { g->act0 = FW(g, B * F0); NN(g->act0); CG_TRY(prelu_fwd(g->lin, g->P + g->oLpw, g->act0, B * F0)); cur = g->act0; }
  for (int i = 0; i < g->nst; ++i) {
    double* stats_part = nullptr; int stats_S = 0;   // set when the conv epilogue produced this stage's batch-norm partial sums
    cg_gstage& s = g->st[i];
    const int hin = h; if (s.up) h *= 2;
    long M = (long)B * h * h, no = M * s.Co;
    g->sconv[i] = FW(g, no); NN(g->sconv[i]);
    g->sfused[i] = fuse; g->sxq[i] = nullptr; g->sup[i] = nullptr;
    if (fuse) {
      uint8_t* xq = (uint8_t*)FW(g, conv_tc_operand_bytes(B, h, h, s.Ci, s.k) / 4); NN(xq);
      float* bn_out = nullptr; const float *ga = nullptr, *be = nullptr, *me = nullptr, *iv = nullptr, *pw = g->P + g->oLpw;
      if (src_bn >= 0) {
        cg_gstage& q = g->st[src_bn];
        bn_out = g->sbn[src_bn] = FW(g, (size_t)B * hin * hin * s.Ci); NN(bn_out);
        ga = g->P + q.og; be = g->P + q.obt; me = g->smean[src_bn]; iv = g->sinv[src_bn]; pw = g->P + q.opw;
      }
      CG_TRY(bn_prelu_up_pack(src, ga, be, me, iv, pw, bn_out, xq, B, hin, hin, s.Ci, s.up, s.k));
      g->sxq[i] = xq;
      cg_layer& L = g->layers[s.layer];
      // training-mode batch norm behind this convolution: its (sum, sum of squares) come out of the conv epilogue (conv_tc.cu P.stats)
      static const bool stats_off = getenv("CATGEN_EPI_STATS_OFF") != nullptr;
      stats_part = nullptr; stats_S = 0;
      if (s.bn && g->training && !stats_off) {
        const int cap = ctx().sm_count;
        stats_part = (double*)FW(g, (size_t)4 * (cap + 1) * s.Co + 4); NN(stats_part);
        stats_part = (double*)(((uintptr_t)stats_part + 7) & ~(uintptr_t)7);
        ctx().next_stats_part = stats_part; ctx().next_stats_cap = cap; ctx().stats_rows = 0;
      }
      int cst = conv_fwd_tc_packed(xq, L.Wp, L.bp, g->sconv[i], B, h, h, s.Ci, s.Co, s.k);
      ctx().next_stats_part = nullptr;
      CG_TRY(cst);
      stats_S = ctx().stats_rows; ctx().stats_rows = 0;
    } else {
      if (s.up) {
        g->sup[i] = FW(g, (size_t)M * s.Ci); NN(g->sup[i]);
        CG_TRY(upsample2x_fwd(cur, g->sup[i], B, hin, hin, s.Ci));
        cur = g->sup[i];
      } else g->sup[i] = (float*)cur;
      CG_TRY(layer_fwd(g, s.layer, cur, g->sconv[i], B, h, h));
    }
    const bool fuse_next = s.bn && g_stage_fused(g, i + 1, h);
    g->sact[i] = nullptr; g->sbn[i] = nullptr;
    if (s.bn) {
      g->smean[i] = FW(g, s.Co); g->sinv[i] = FW(g, s.Co); NN(g->smean[i]); NN(g->sinv[i]);
      if (fuse_next) {   // statistics only; the next stage's producer applies them (g_stage_fused implies training mode)
        CG_TRY(bn_fwd_train_pre(stats_S > 0 ? stats_part : nullptr, stats_S, g->sconv[i], g->P + s.og, g->P + s.obt, nullptr, g->smean[i], g->sinv[i], g->run + r, g->run + r + s.Co, M, s.Co, 1e-5f, 0.1f));
        src = g->sconv[i]; src_bn = i;
      } else {
        g->sbn[i] = FW(g, no); g->sact[i] = FW(g, no); NN(g->sbn[i]); NN(g->sact[i]);
        if (g->training)   // adversarial.train never switches G to evaluate(): batch statistics
          CG_TRY(bn_fwd_train_pre(stats_S > 0 ? stats_part : nullptr, stats_S, g->sconv[i], g->P + s.og, g->P + s.obt, g->sbn[i], g->smean[i], g->sinv[i], g->run + r, g->run + r + s.Co, M, s.Co, 1e-5f, 0.1f));
        else
          CG_TRY(bn_fwd_eval(g->sconv[i], g->P + s.og, g->P + s.obt, g->sbn[i], g->run + r, g->run + r + s.Co, M, s.Co, 1e-5f));
        CG_TRY(prelu_fwd(g->sbn[i], g->P + s.opw, g->sact[i], no));
        cur = g->sact[i];
      }
      r += 2 * s.Co;
    } else {
      g->sact[i] = FW(g, no); NN(g->sact[i]);
      CG_TRY(sigmoid_fwd(g->sconv[i], g->sact[i], no));
      cur = g->sact[i];
    }
    fuse = fuse_next;
  }
  CG_TRY(nhwc_to_nchw(cur, out_nchw, B, g->C, 32 * 32));
  return CG_OK;
}
int G_backward_dev(cg_model* g, const float* gout_nchw, float* gz_dev) {
  if (!g->B) return set_err(CG_ERR_STATE, "G backward before forward");
  if (!g->training) return set_err(CG_ERR_STATE, "G backward needs a training-mode forward (batch-stat BN)");
  CG_TRY(model_repack(g, !conv_tc_all_shapes_taken(g->B)));   // the engine may have been switched since the forward: the operands it reads must be fresh
  Fp32StaleScope stale_scope(g->dirty32);
  g->nbw = 0; int B = g->B, h = 32;
  float* gcur = BW(g, (size_t)B * g->C * 1024); NN(gcur);
  CG_TRY(nchw_to_nhwc(gout_nchw, gcur, B, g->C, 1024));
  for (int i = g->nst - 1; i >= 0; --i) {
    cg_gstage& s = g->st[i];
    long M = (long)B * h * h, no = M * s.Co;
    float* gconv = BW(g, no); NN(gconv);
    if (s.bn) {
      float* gbn = BW(g, no); NN(gbn);
      CG_TRY(prelu_bwd(g->sbn[i], gcur, g->P + s.opw, gbn, g->G + s.opw, no));
      CG_TRY(bn_bwd(g->sconv[i], gbn, g->P + s.og, g->smean[i], g->sinv[i], gconv, g->G + s.og, g->G + s.obt, M, s.Co));
    } else CG_TRY(sigmoid_bwd(g->sact[i], gcur, gconv, no));
    float* gin = BW(g, (size_t)M * s.Ci); NN(gin);
    float* gs = nullptr;
    if (s.up) {   // ask the input-gradient convolution to sum the 2 x 2 blocks in its epilogue (conv_tc.cu P.pool2): the full-resolution gradient is then never written
      gs = BW(g, (size_t)B * (h / 2) * (h / 2) * s.Ci); NN(gs);
      static const bool off = getenv("CATGEN_POOL2_OFF") != nullptr;
      ctx().pool2_done = 0; ctx().next_pool2_out = (off || ctx().conv_engine != 1) ? nullptr : gs;
    }
    int lst = layer_bwd(g, s.layer, g->sup[i], gconv, gin, B, h, h, g->sfused[i] ? g->sxq[i] : nullptr);
    ctx().next_pool2_out = nullptr;
    CG_TRY(lst);
    if (s.up) {
      h /= 2;
      if (!ctx().pool2_done) CG_TRY(upsample2x_bwd(gin, gs, B, h, h, s.Ci));
      ctx().pool2_done = 0;
      gin = gs;
    }
    gcur = gin;
  }
  long F0 = (long)g->C0 * g->s0 * g->s0;
  float* glin = BW(g, B * F0); NN(glin);
  CG_TRY(prelu_bwd(g->lin, gcur, g->P + g->oLpw, glin, g->G + g->oLpw, B * F0));
  float* gz = gz_dev ? gz_dev : BW(g, (size_t)B * g->nz); NN(gz);
  CG_TRY(layer_bwd(g, g->lin_layer, g->z, glin, gz, B, 1, 1));
  CG_TRY(side_wait_all());   // parameter gradients are complete when this returns (stream order)
  return CG_OK;
}

// =================================================================== D
__global__ void k_copy_channels(const float* __restrict__ src, float* __restrict__ dst, long n, int Cs, int Cd, int coff, int dir) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    long r = i / Cs; int c = (int)(i % Cs);
    if (dir == 0) dst[r * Cd + coff + c] = src[i];        // nn.Concat(2) forward: into the channel slot
    else dst[i] = src[r * Cd + coff + c];                  // backward: slice of gradOutput
  }
}
static int copy_channels(const float* src, float* dst, long rows, int Cs, int Cd, int coff, int dir) {
  long n = rows * Cs; CG_LAUNCH(k_copy_channels, grid1d(n, 256, 4), 256, 0, src, dst, n, Cs, Cd, coff, dir); return CG_OK;
}

static bool d_head_use_fused() { static const bool off = getenv("CATGEN_DHEAD_UNFUSED") != nullptr; return !off; }
static bool stn_use_fused() { static const bool off = getenv("CATGEN_STN_UNFUSED") != nullptr; return ctx().conv_engine == 1 && !off; }
static void stn_params(cg_model* m, cg_stn* s, StnFusedParams* p, StnFusedGrads* g) {
  const cg_layer &c1 = m->layers[s->c1], &c2 = m->layers[s->c2], &l1 = m->layers[s->l1], &l2 = m->layers[s->l2];
  p->W1 = m->P + c1.oW; p->b1 = m->P + c1.ob; p->W2 = m->P + c2.oW; p->b2 = m->P + c2.ob;
  p->W1p = c1.Wp; p->W1d = c1.Wd; p->W2p = c2.Wp; p->W2d = c2.Wd;   // packed fp32 operands (k_repack_model, always32)
  p->L1 = m->P + l1.oW; p->lb1 = m->P + l1.ob; p->L2 = m->P + l2.oW; p->lb2 = m->P + l2.ob;
  p->ch = s->ch; p->S = s->S; p->rot = s->rot; p->scl = s->scl; p->trn = s->trn; p->nth = s->nth;
  if (g) { g->W1 = m->G + c1.oW; g->b1 = m->G + c1.ob; g->W2 = m->G + c2.oW; g->b2 = m->G + c2.ob; g->L1 = m->G + l1.oW; g->lb1 = m->G + l1.ob; g->L2 = m->G + l2.oW; g->lb2 = m->G + l2.ob; }
}
static int stn_forward(cg_model* m, cg_stn* s, const float* in, int B) {
  int ch = s->ch, S = s->S, S2 = S / 2, S4 = S / 4, f = 16 * S4 * S4;
  s->in = in;
  s->fused = stn_use_fused() && stn_fused_shape_ok(ch, S);
  if (s->fused) {   // stn_fused.cu: localisation network in one launch (one CTA per image, fp32), sampler in one launch
    long n2 = (long)B * S2 * S2 * 16;
    s->pool1 = FW(m, (size_t)B * S2 * S2 * ch); s->c1o = FW(m, n2); s->c2o = FW(m, n2); s->pool2 = FW(m, (size_t)B * f); s->l1o = FW(m, (size_t)B * 64);
    s->theta = FW(m, (size_t)B * 4); s->A = FW(m, (size_t)B * 6); s->out = FW(m, (size_t)B * S * S * ch);
    NN(s->pool1); NN(s->c1o); NN(s->c2o); NN(s->pool2); NN(s->l1o); NN(s->theta); NN(s->A); NN(s->out);
    StnFusedParams p; stn_params(m, s, &p, nullptr);
    return stn_fused_forward(p, in, B, s->pool1, s->c1o, s->c2o, s->pool2, s->l1o, s->theta, s->A, s->out);
  }
  s->pool1 = FW(m, (size_t)B * S2 * S2 * ch); NN(s->pool1); CG_TRY(avgpool2_fwd(in, s->pool1, B, S, S, ch));
  long n2 = (long)B * S2 * S2 * 16;
  s->c1o = FW(m, n2); s->a1 = FW(m, n2); s->c2o = FW(m, n2); s->a2 = FW(m, n2); NN(s->c1o); NN(s->a1); NN(s->c2o); NN(s->a2);
  CG_TRY(layer_fwd(m, s->c1, s->pool1, s->c1o, B, S2, S2)); CG_TRY(lrelu_fwd(s->c1o, 0.333f, s->a1, n2));
  CG_TRY(layer_fwd(m, s->c2, s->a1, s->c2o, B, S2, S2)); CG_TRY(lrelu_fwd(s->c2o, 0.333f, s->a2, n2));
  s->pool2 = FW(m, (size_t)B * f); NN(s->pool2); CG_TRY(avgpool2_fwd(s->a2, s->pool2, B, S2, S2, 16));
  s->l1o = FW(m, (size_t)B * 64); s->al1 = FW(m, (size_t)B * 64); NN(s->l1o); NN(s->al1);
  CG_TRY(layer_fwd(m, s->l1, s->pool2, s->l1o, B, 1, 1)); CG_TRY(lrelu_fwd(s->l1o, 0.333f, s->al1, (long)B * 64));
  s->theta = FW(m, (size_t)B * 4); NN(s->theta); CG_TRY(layer_fwd(m, s->l2, s->al1, s->theta, B, 1, 1));
  s->A = FW(m, (size_t)B * 6); NN(s->A); CG_TRY(affine_matrix_fwd(s->theta, s->A, B, s->rot, s->scl, s->trn));
  s->grid = FW(m, (size_t)B * S * S * 2); NN(s->grid); CG_TRY(affine_grid_fwd(s->A, s->grid, B, S, S));
  s->out = FW(m, (size_t)B * S * S * ch); NN(s->out); CG_TRY(bilinear_fwd(in, s->grid, s->out, B, S, S, ch));
  return CG_OK;
}
// gout, gin: [B,S,S,ch]; gin is written (sum of the sampler branch and the localisation branch, nn.ConcatTable)
static int stn_backward(cg_model* m, cg_stn* s, const float* gout, float* gin, int B, unsigned int* amax_out = nullptr) {
  int ch = s->ch, S = s->S, S2 = S / 2, S4 = S / 4, f = 16 * S4 * S4;
  if (s->fused) {
    float* ggrid = BW(m, (size_t)B * S * S * 2); float* gl1 = BW(m, (size_t)B * 64); NN(ggrid); NN(gl1);
    float* part = nullptr;
    if (!m->skip_param_grads) { part = BW(m, (size_t)B * stn_fused_part_floats(ch, s->nth)); NN(part); }
    StnFusedParams p; StnFusedGrads g; stn_params(m, s, &p, &g);
    return stn_fused_backward(p, g, s->in, B, s->pool1, s->c1o, s->c2o, s->pool2, s->l1o, s->theta, s->A, gout, gin, ggrid, gl1, part, m->skip_param_grads, amax_out);
  }
  float* ggrid = BW(m, (size_t)B * S * S * 2); NN(ggrid);
  CG_TRY(bilinear_bwd(s->in, s->grid, gout, gin, ggrid, B, S, S, ch));
  float* gA = BW(m, (size_t)B * 6); NN(gA); CG_TRY(affine_grid_bwd(ggrid, gA, B, S, S));
  float* gth = BW(m, (size_t)B * 4); NN(gth); CG_TRY(affine_matrix_bwd(s->theta, gA, gth, B, s->rot, s->scl, s->trn));
  float* gal1 = BW(m, (size_t)B * 64); NN(gal1); CG_TRY(layer_bwd(m, s->l2, s->al1, gth, gal1, B, 1, 1));
  float* gl1 = BW(m, (size_t)B * 64); NN(gl1); CG_TRY(lrelu_bwd(s->l1o, gal1, 0.333f, gl1, (long)B * 64));
  float* gp2 = BW(m, (size_t)B * f); NN(gp2); CG_TRY(layer_bwd(m, s->l1, s->pool2, gl1, gp2, B, 1, 1));
  long n2 = (long)B * S2 * S2 * 16;
  float* ga2 = BW(m, n2); NN(ga2); CG_TRY(avgpool2_bwd(gp2, ga2, B, S2, S2, 16));
  float* gc2 = BW(m, n2); NN(gc2); CG_TRY(lrelu_bwd(s->c2o, ga2, 0.333f, gc2, n2));
  float* ga1 = BW(m, n2); NN(ga1); CG_TRY(layer_bwd(m, s->c2, s->a1, gc2, ga1, B, S2, S2));
  float* gc1 = BW(m, n2); NN(gc1); CG_TRY(lrelu_bwd(s->c1o, ga1, 0.333f, gc1, n2));
  float* gp1 = BW(m, (size_t)B * S2 * S2 * ch); NN(gp1); CG_TRY(layer_bwd(m, s->c1, s->pool1, gc1, gp1, B, S2, S2));
  float* gin2 = BW(m, (size_t)B * S * S * ch); NN(gin2); CG_TRY(avgpool2_bwd(gp1, gin2, B, S, S, ch));
  CG_TRY(add_inplace(gin, gin2, (long)B * S * S * ch));
  return CG_OK;
}

static int ensure_masks(cg_model* d, int B) {
  long n = D_mask_floats(B);
  if (d->masks_n < n) { if (d->masks) { cudaStreamSynchronize(ctx().stream); cudaFree(d->masks); } ctx().alloc_gen++; CG_CUDA(cudaMalloc(&d->masks, sizeof(float) * n)); d->masks_n = n; }
  d->masks_B = B;
  if (d->mq && d->mq_next < d->mq_count) {   // masks queued by cg_D_set_masks: one set per forward, in order
    if (d->mq_B != B) return set_err(CG_ERR_ARG, "queued dropout masks are for batch %d, forward has batch %d", d->mq_B, B);
    CG_CUDA(cudaMemcpyAsync(d->masks, d->mq + (size_t)n * d->mq_next, sizeof(float) * n, cudaMemcpyDeviceToDevice, ctx().stream));
    d->mq_next++;
    return CG_OK;
  }
  long nsp = (long)B * (64 * 4 + 128), nh = (long)B * 320, nf = (long)B * 256;
  if (d->training) {   // nn.SpatialDropout(0.2) x5 (no rescale), SpatialDropout(0.5), nn.Dropout(0.5) v2 (x2) -- SURVEY.md A.12
    uint64_t r1 = (uint64_t)(nsp + 3) / 4 + 1, r2 = r1 + (uint64_t)(nh + 3) / 4 + 1, r3 = r2 + (uint64_t)(nf + 3) / 4 + 1;
    CG_TRY(bernoulli_mask(d->masks, nsp, 0.2f, 1.f, d->seed, d->rng_dev, 0));
    CG_TRY(bernoulli_mask(d->masks + nsp, nh, 0.5f, 1.f, d->seed, d->rng_dev, r1));
    CG_TRY(bernoulli_mask(d->masks + nsp + nh, nf, 0.5f, 2.f, d->seed, d->rng_dev, r2));
    CG_TRY(rng_advance(d->rng_dev, r3));
  } else {             // evaluate(): SpatialDropout scales by (1-p); Dropout v2 is the identity
    CG_TRY(fill(d->masks, 0.8f, nsp)); CG_TRY(fill(d->masks + nsp, 0.5f, nh)); CG_TRY(fill(d->masks + nsp + nh, 1.f, nf));
  }
  return CG_OK;
}

// The same network with every element-wise chain between two convolutions as ONE kernel (fuse_d.cu) and the conv operands cached:
// conv -> [PReLU -> pool -> SpatialDropout -> fp16 operand of the next conv]; that operand also feeds the next layer's weight gradient.
static uint8_t* FWQ(cg_model* m, int N, int H, int W, int C, int k) { return (uint8_t*)FW(m, conv_tc_operand_bytes(N, H, W, C, k) / 4); }
static int conv_packed(cg_model* m, int li, const uint8_t* xq, float* y, int N, int H) {
  cg_layer& L = m->layers[li];
  return conv_fwd_tc_packed(xq, L.Wp, L.bp, y, N, H, H, L.s.Ci, L.s.Co, L.s.k);
}
static int D_forward_fused(cg_model* d, int B, const float* mk, float* sig_dev, float* pre_dev) {
  long n64 = (long)B * 1024 * 64;
  d->tc1 = FW(d, n64); d->tc2 = FW(d, n64); NN(d->tc1); NN(d->tc2);
  CG_TRY(layer_fwd(d, d->t1, d->stn[0].out, d->tc1, B, 32, 32));                                   // conv C -> 64 (models.lua:646)
  uint8_t* xq_t2 = FWQ(d, B, 32, 32, 64, 3); NN(xq_t2); d->xq_t2 = xq_t2;
  CG_TRY(act_pool_mask_pack(d->tc1, d->P + d->t1pw, B, 32, 32, 64, 0, nullptr, 0, nullptr, nullptr, 0, 0, xq_t2, 3));   // PReLU -> operand of conv 64 -> 64 (:647-648)
  CG_TRY(conv_packed(d, d->t2, xq_t2, d->tc2, B, 32));
  d->T = FW(d, n64 / 4); NN(d->T);
  uint8_t* xq_b4 = FWQ(d, B, 16, 16, 64, 5); NN(xq_b4); d->xq_b4 = xq_b4;
  // PReLU -> AvgPool(2) -> SpatialDropout(0.2) (:649-651): T in fp32 for the three transformers, and as the operand of branch 4's 5x5 conv
  CG_TRY(act_pool_mask_pack(d->tc2, d->P + d->t2pw, B, 32, 32, 64, 1, mk, 64, nullptr, d->T, 64, 0, xq_b4, 5)); mk += (long)B * 64;
  d->catd = FW(d, (size_t)B * 20480); NN(d->catd);
  const float* mk_head = mk + (long)B * (64 * 3 + 128);                                            // SpatialDropout(0.5) after nn.Concat (:695), applied as each branch writes its slot
  CG_TRY(lanes_fork());
  for (int b = 0; b < 4; ++b) {
    LaneScope lane(b); CG_TRY(lane.status);
    const int Co = b < 3 ? 64 : 128, k2 = b < 3 ? 3 : 7;
    const long n1 = (long)B * 256 * Co;
    d->bc1[b] = FW(d, n1); NN(d->bc1[b]);
    if (b < 3) { CG_TRY(stn_forward(d, &d->stn[b + 1], d->T, B)); CG_TRY(layer_fwd(d, d->b1[b], d->stn[b + 1].out, d->bc1[b], B, 16, 16)); }
    else CG_TRY(conv_packed(d, d->b1[b], xq_b4, d->bc1[b], B, 16));
    d->bidx[b] = (uint8_t*)FW(d, n1 / 16 + 4); d->bc2[b] = FW(d, n1 / 4); NN(d->bidx[b]); NN(d->bc2[b]);
    uint8_t* xq2 = FWQ(d, B, 8, 8, Co, k2); NN(xq2); d->xq_b2[b] = xq2;
    // PReLU -> MaxPool(2) -> SpatialDropout(0.2) -> operand of the branch's second conv (:656-659, :681-685)
    CG_TRY(act_pool_mask_pack(d->bc1[b], d->P + d->bpw1[b], B, 16, 16, Co, 2, mk + (long)B * 64 * b, Co, d->bidx[b], nullptr, 0, 0, xq2, k2));
    CG_TRY(conv_packed(d, d->b2[b], xq2, d->bc2[b], B, 8));
    // PReLU -> this branch's channel slot of the Concat buffer, SpatialDropout(0.5) of the head applied on the way (:660, :695)
    CG_TRY(act_pool_mask_pack(d->bc2[b], d->P + d->bpw2[b], B, 8, 8, Co, 0, mk_head + b * 64, 320, nullptr, d->catd, 320, b * 64, nullptr, 0));
  }
  CG_TRY(lanes_join());
  mk = mk_head + (long)B * 320;
  d->h1o = FW(d, (size_t)B * 256); d->ha1 = FW(d, (size_t)B * 256); d->hd = FW(d, (size_t)B * 256); NN(d->h1o); NN(d->ha1); NN(d->hd);
  CG_TRY(layer_fwd(d, d->h1, d->catd, d->h1o, B, 1, 1));
  d->h2o = FW(d, B); d->hsig = FW(d, B); NN(d->h2o); NN(d->hsig);
  d->head_fused = d_head_use_fused();
  if (d->head_fused) CG_TRY(d_head_fwd(d->h1o, d->P + d->hpw, mk, d->P + d->layers[d->h2].oW, d->P + d->layers[d->h2].ob, d->hd, d->h2o, d->hsig, B));   // fuse_d.cu
  else {
    CG_TRY(prelu_fwd(d->h1o, d->P + d->hpw, d->ha1, (long)B * 256));
    CG_TRY(mask_elems(d->ha1, mk, d->hd, (long)B * 256));
    CG_TRY(layer_fwd(d, d->h2, d->hd, d->h2o, B, 1, 1));
    CG_TRY(sigmoid_fwd(d->h2o, d->hsig, B));
  }
  if (sig_dev) CG_CUDA(cudaMemcpyAsync(sig_dev, d->hsig, sizeof(float) * B, cudaMemcpyDeviceToDevice, ctx().stream));
  if (pre_dev) CG_CUDA(cudaMemcpyAsync(pre_dev, d->h2o, sizeof(float) * B, cudaMemcpyDeviceToDevice, ctx().stream));
  return CG_OK;
}

// Diagnosis knob (tests/test_gpu_configs.py reads its effect): CATGEN_D_FWD_FP32=1 runs D's FORWARD on the fp32 CUDA-core engine while the
// backward stays on the tensor-core engine -- separates the forward's fp16 operand rounding from the backward's in the gradient error.
struct EngineScope {
  int saved; bool on;
  explicit EngineScope(bool on_) : saved(ctx().conv_engine), on(on_) { if (on) ctx().conv_engine = 0; }
  ~EngineScope() { if (on) ctx().conv_engine = saved; }
};
int D_forward_dev(cg_model* d, const float* x_nchw, int B, float* sig_dev, float* pre_dev) {
  static const bool fwd_fp32 = getenv("CATGEN_D_FWD_FP32") != nullptr;
  EngineScope engine_scope(fwd_fp32 && ctx().conv_engine == 1);
  SplitScope split_scope;
  CG_TRY(model_repack(d, !conv_tc_all_shapes_taken(B)));
  Fp32StaleScope stale_scope(d->dirty32);
  CG_TRY(ensure_masks(d, B));
  d->nfw = 0; d->B = B; int C = d->C;
  const float* mk = d->masks;
  d->xin = FW(d, (size_t)B * 1024 * C); NN(d->xin);
  CG_TRY(nchw_to_nhwc(x_nchw, d->xin, B, C, 1024));                      // nn.Copy + the STN's nn.Transpose (models.lua:643,870)
  CG_TRY(stn_forward(d, &d->stn[0], d->xin, B));
  long n64 = (long)B * 1024 * 64;
  d->dfused = ctx().conv_engine == 1 && ctx().precision == 0 && getenv("CATGEN_D_UNFUSED") == nullptr && conv_tc_cached_ok(32, 32, 64, 64, 3);
  if (d->dfused) return D_forward_fused(d, B, mk, sig_dev, pre_dev);
  d->tc1 = FW(d, n64); d->ta1 = FW(d, n64); d->tc2 = FW(d, n64); d->ta2 = FW(d, n64); NN(d->tc1); NN(d->ta1); NN(d->tc2); NN(d->ta2);
  CG_TRY(layer_fwd(d, d->t1, d->stn[0].out, d->tc1, B, 32, 32)); CG_TRY(prelu_fwd(d->tc1, d->P + d->t1pw, d->ta1, n64));
  CG_TRY(layer_fwd(d, d->t2, d->ta1, d->tc2, B, 32, 32)); CG_TRY(prelu_fwd(d->tc2, d->P + d->t2pw, d->ta2, n64));
  d->tpool = FW(d, n64 / 4); d->T = FW(d, n64 / 4); NN(d->tpool); NN(d->T);
  CG_TRY(avgpool2_fwd(d->ta2, d->tpool, B, 32, 32, 64));
  CG_TRY(mask_channels(d->tpool, mk, d->T, B, 256, 64)); mk += (long)B * 64;
  d->cat = FW(d, (size_t)B * 64 * 320); NN(d->cat);
  // the four branches only read T and write disjoint channel ranges of cat: one lane each (models.lua:661-699, nn.Concat)
  CG_TRY(lanes_fork());
  for (int b = 0; b < 4; ++b) {
    LaneScope lane(b); CG_TRY(lane.status);
    int Co = b < 3 ? 64 : 128;
    const float* bin = d->T;
    if (b < 3) { CG_TRY(stn_forward(d, &d->stn[b + 1], d->T, B)); bin = d->stn[b + 1].out; }
    long n1 = (long)B * 256 * Co;
    d->bc1[b] = FW(d, n1); d->ba1[b] = FW(d, n1); NN(d->bc1[b]); NN(d->ba1[b]);
    CG_TRY(layer_fwd(d, d->b1[b], bin, d->bc1[b], B, 16, 16)); CG_TRY(prelu_fwd(d->bc1[b], d->P + d->bpw1[b], d->ba1[b], n1));
    d->bmp[b] = FW(d, n1 / 4); d->bidx[b] = (uint8_t*)FW(d, n1 / 16 + 4); d->bdr[b] = FW(d, n1 / 4); d->bc2[b] = FW(d, n1 / 4);
    NN(d->bmp[b]); NN(d->bidx[b]); NN(d->bdr[b]); NN(d->bc2[b]);
    CG_TRY(maxpool2_fwd(d->ba1[b], d->bmp[b], d->bidx[b], B, 16, 16, Co));
    CG_TRY(mask_channels(d->bmp[b], mk, d->bdr[b], B, 64, Co)); mk += (long)B * Co;
    CG_TRY(layer_fwd(d, d->b2[b], d->bdr[b], d->bc2[b], B, 8, 8));
    float* tmp = FW(d, n1 / 4); NN(tmp);
    CG_TRY(prelu_fwd(d->bc2[b], d->P + d->bpw2[b], tmp, n1 / 4));
    CG_TRY(copy_channels(tmp, d->cat, (long)B * 64, Co, 320, b * 64, 0));
  }
  CG_TRY(lanes_join());
  d->catd = FW(d, (size_t)B * 20480); NN(d->catd);
  CG_TRY(mask_channels(d->cat, mk, d->catd, B, 64, 320)); mk += (long)B * 320;
  d->h1o = FW(d, (size_t)B * 256); d->ha1 = FW(d, (size_t)B * 256); d->hd = FW(d, (size_t)B * 256); NN(d->h1o); NN(d->ha1); NN(d->hd);
  d->head_fused = false;
  CG_TRY(layer_fwd(d, d->h1, d->catd, d->h1o, B, 1, 1));
  CG_TRY(prelu_fwd(d->h1o, d->P + d->hpw, d->ha1, (long)B * 256));
  CG_TRY(mask_elems(d->ha1, mk, d->hd, (long)B * 256));
  d->h2o = FW(d, B); d->hsig = FW(d, B); NN(d->h2o); NN(d->hsig);
  CG_TRY(layer_fwd(d, d->h2, d->hd, d->h2o, B, 1, 1));
  CG_TRY(sigmoid_fwd(d->h2o, d->hsig, B));
  if (sig_dev) CG_CUDA(cudaMemcpyAsync(sig_dev, d->hsig, sizeof(float) * B, cudaMemcpyDeviceToDevice, ctx().stream));
  if (pre_dev) CG_CUDA(cudaMemcpyAsync(pre_dev, d->h2o, sizeof(float) * B, cudaMemcpyDeviceToDevice, ctx().stream));
  return CG_OK;
}

// where a PReLU weight gradient goes (nowhere when parameter gradients are skipped)
static inline float* PG(cg_model* m, long off) { return m->skip_param_grads ? nullptr : m->G + off; }

// The conv part of D's backward with every chain (dropout', pool', PReLU', bias / PReLU gradients, fp16 gradient operand) as ONE kernel
// per convolution (fuse_d.cu act_bwd_pack) and the four branch gradients summed inside the trunk's kernel.  Entered with gcatd = the
// gradient of the masked Concat buffer (the head's Linear layers run as before).  amax slots: 0 gcatd ; 1-4 each branch's second-conv
// input gradient ; 5-7 the transformers' input gradients ; 8 branch 4's first-conv input gradient ; 9 trunk conv 2's input gradient.
static int conv_layer_bwd_gq(cg_model* m, int li, const float* x, const uint8_t* xq, const uint8_t* gq, const float* sc, float* gx, int N, int H, unsigned int* amax_next) {
  cg_layer& L = m->layers[li];
  ctx().next_amax = gx ? amax_next : nullptr;
  int s = conv_bwd_tc_gq(x, xq, gq, sc, L.Wd, gx, N, H, H, L.s.Ci, L.s.Co, L.s.k, m->skip_param_grads ? nullptr : m->G + L.oW);
  ctx().next_amax = nullptr;
  return s;
}
struct ActBwd { uint8_t* gq; float* sc; double* part; };
static int act_bwd_alloc(cg_model* d, ActBwd* r, int N, int H, int C, int k) {
  r->gq = (uint8_t*)BW(d, act_bwd_operand_bytes(N, H, H, C, k) / 4); r->sc = BW(d, 4);
  r->part = d->skip_param_grads ? nullptr : (double*)BW(d, (size_t)N * (C / 8) * 9 * 2);
  if (!r->gq || !r->sc || (!d->skip_param_grads && !r->part)) return set_err(CG_ERR_CUDA, "device allocation failed");
  return CG_OK;
}
static int D_backward_fused(cg_model* d, const float* gcatd, float* gx_nchw) {
  const int B = d->B, C = d->C;
  const float* mk_trunk = d->masks; const float* mk_br = d->masks + (long)B * 64; const float* mk_head = d->masks + (long)B * (64 * 4 + 128);
  unsigned int* am = d->amax;
  CG_CUDA(cudaMemsetAsync(am, 0, sizeof(unsigned int) * 16, ctx().stream));
  CG_TRY(absmax_into(gcatd, (long)B * 20480, am + 0));
  const long nT = (long)B * 256 * 64;
  const float* gT_part[4];
  CG_TRY(lanes_fork());
  for (int b = 0; b < 4; ++b) {
    LaneScope lane(b); CG_TRY(lane.status);
    const int Co = b < 3 ? 64 : 128, k1 = b < 3 ? 3 : 5, k2 = b < 3 ? 3 : 7;
    // second conv of the branch: upstream = this branch's slot of the Concat gradient, head dropout mask applied on the way (models.lua:695)
    ActBwd r2; CG_TRY(act_bwd_alloc(d, &r2, B, 8, Co, k2));
    { const float* g[1] = {gcatd}; const unsigned int* a[1] = {am + 0};
      CG_TRY(act_bwd_pack(g, a, 1, 320, b * 64, mk_head + b * 64, 320, 0, nullptr, d->bc2[b], d->P + d->bpw2[b], B, 8, 8, Co, k2, r2.gq, r2.sc, r2.part,
                          d->skip_param_grads ? nullptr : d->G + d->layers[d->b2[b]].ob, PG(d, d->bpw2[b]))); }
    float* gdr = BW(d, (size_t)B * 64 * Co); NN(gdr);
    CG_TRY(conv_layer_bwd_gq(d, d->b2[b], nullptr, d->xq_b2[b], r2.gq, r2.sc, gdr, B, 8, am + 1 + b));
    // first conv: SpatialDropout' -> MaxPool' -> PReLU'
    ActBwd r1; CG_TRY(act_bwd_alloc(d, &r1, B, 16, Co, k1));
    { const float* g[1] = {gdr}; const unsigned int* a[1] = {am + 1 + b};
      CG_TRY(act_bwd_pack(g, a, 1, Co, 0, mk_br + (long)B * 64 * b, Co, 2, d->bidx[b], d->bc1[b], d->P + d->bpw1[b], B, 16, 16, Co, k1, r1.gq, r1.sc, r1.part,
                          d->skip_param_grads ? nullptr : d->G + d->layers[d->b1[b]].ob, PG(d, d->bpw1[b]))); }
    float* gbin = BW(d, nT); NN(gbin);
    CG_TRY(conv_layer_bwd_gq(d, d->b1[b], b < 3 ? d->stn[b + 1].out : nullptr, b == 3 ? d->xq_b4 : nullptr, r1.gq, r1.sc, gbin, B, 16, b == 3 ? am + 8 : nullptr));
    if (b < 3) {
      float* gs = BW(d, nT); NN(gs);
      CG_TRY(stn_backward(d, &d->stn[b + 1], gbin, gs, B, am + 5 + b));
      gT_part[b] = gs;
    } else gT_part[b] = gbin;
  }
  CG_TRY(lanes_join());
  // trunk conv 2: the four branch gradients are summed here (nn.Concat backward), then SpatialDropout' -> AvgPool' -> PReLU'
  ActBwd rt2; CG_TRY(act_bwd_alloc(d, &rt2, B, 32, 64, 3));
  { const unsigned int* a[4] = {am + 5, am + 6, am + 7, am + 8};
    CG_TRY(act_bwd_pack(gT_part, a, 4, 64, 0, mk_trunk, 64, 1, nullptr, d->tc2, d->P + d->t2pw, B, 32, 32, 64, 3, rt2.gq, rt2.sc, rt2.part,
                        d->skip_param_grads ? nullptr : d->G + d->layers[d->t2].ob, PG(d, d->t2pw))); }
  const long n64 = (long)B * 1024 * 64;
  float* gta1 = BW(d, n64); NN(gta1);
  CG_TRY(conv_layer_bwd_gq(d, d->t2, nullptr, d->xq_t2, rt2.gq, rt2.sc, gta1, B, 32, am + 9));
  ActBwd rt1; CG_TRY(act_bwd_alloc(d, &rt1, B, 32, 64, 3));
  { const float* g[1] = {gta1}; const unsigned int* a[1] = {am + 9};
    CG_TRY(act_bwd_pack(g, a, 1, 64, 0, nullptr, 0, 0, nullptr, d->tc1, d->P + d->t1pw, B, 32, 32, 64, 3, rt1.gq, rt1.sc, rt1.part,
                        d->skip_param_grads ? nullptr : d->G + d->layers[d->t1].ob, PG(d, d->t1pw))); }
  float* gs0 = BW(d, (size_t)B * 1024 * C); NN(gs0);
  CG_TRY(conv_layer_bwd_gq(d, d->t1, d->stn[0].out, nullptr, rt1.gq, rt1.sc, gs0, B, 32, nullptr));
  float* gin = BW(d, (size_t)B * 1024 * C); NN(gin);
  CG_TRY(stn_backward(d, &d->stn[0], gs0, gin, B));
  if (gx_nchw) CG_TRY(nhwc_to_nchw(gin, gx_nchw, B, C, 1024));        // MODEL_D.modules[1].gradInput (adversarial.lua:193)
  CG_TRY(side_wait_all());
  return CG_OK;
}

int D_backward_dev(cg_model* d, const float* gout_dev, float* gx_nchw) {
  if (!d->B) return set_err(CG_ERR_STATE, "D backward before forward");
  CG_TRY(model_repack(d, !conv_tc_all_shapes_taken(d->B)));
  Fp32StaleScope stale_scope(d->dirty32);
  d->nbw = 0; int B = d->B, C = d->C;
  const float* mk_trunk = d->masks;
  const float* mk_br = d->masks + (long)B * 64;
  const float* mk_head = d->masks + (long)B * (64 * 4 + 128);
  const float* mk_fc = mk_head + (long)B * 320;
  float* gh1 = BW(d, (size_t)B * 256); NN(gh1);
  if (d->head_fused && B <= 1024) {
    CG_TRY(d_head_bwd(gout_dev, d->hsig, d->hd, d->h1o, d->P + d->hpw, mk_fc, d->P + d->layers[d->h2].oW, gh1, d->G + d->layers[d->h2].oW, d->G + d->layers[d->h2].ob,
                      d->G + d->hpw, B, d->skip_param_grads ? 0 : 1));
  } else {
    float* gh2 = BW(d, B); NN(gh2); CG_TRY(sigmoid_bwd(d->hsig, gout_dev, gh2, B));
    float* ghd = BW(d, (size_t)B * 256); NN(ghd); CG_TRY(layer_bwd(d, d->h2, d->hd, gh2, ghd, B, 1, 1));
    float* gha1 = BW(d, (size_t)B * 256); NN(gha1); CG_TRY(mask_elems(ghd, mk_fc, gha1, (long)B * 256));
    CG_TRY(prelu_bwd(d->h1o, gha1, d->P + d->hpw, gh1, PG(d, d->hpw), (long)B * 256));
  }
  float* gcatd = BW(d, (size_t)B * 20480); NN(gcatd); CG_TRY(layer_bwd(d, d->h1, d->catd, gh1, gcatd, B, 1, 1));
  if (d->dfused && d->stn[0].fused && getenv("CATGEN_DBWD_UNFUSED") == nullptr) return D_backward_fused(d, gcatd, gx_nchw);
  float* gcat = BW(d, (size_t)B * 20480); NN(gcat); CG_TRY(mask_channels(gcatd, mk_head, gcat, B, 64, 320));
  long nT = (long)B * 256 * 64;
  float* gT = BW(d, nT); NN(gT); CG_TRY(fill(gT, 0.f, nT));
  const float* mk = mk_br;
  const float* gT_part[4];      // each branch's gradient w.r.t. T, summed after the join in branch order (as nn.Concat does)
  CG_TRY(lanes_fork());
  for (int b = 0; b < 4; ++b) {
    LaneScope lane(b); CG_TRY(lane.status);
    int Co = b < 3 ? 64 : 128;
    long n2 = (long)B * 64 * Co, n1 = n2 * 4;
    float* go = BW(d, n2); NN(go); CG_TRY(copy_channels(gcat, go, (long)B * 64, Co, 320, b * 64, 1));
    float* gc2 = BW(d, n2); NN(gc2); CG_TRY(prelu_bwd(d->bc2[b], go, d->P + d->bpw2[b], gc2, PG(d, d->bpw2[b]), n2));
    float* gdr = BW(d, n2); NN(gdr); CG_TRY(layer_bwd(d, d->b2[b], d->dfused ? nullptr : d->bdr[b], gc2, gdr, B, 8, 8, d->dfused ? d->xq_b2[b] : nullptr));
    float* gmp = BW(d, n2); NN(gmp); CG_TRY(mask_channels(gdr, mk, gmp, B, 64, Co)); mk += (long)B * Co;
    float* ga1 = BW(d, n1); NN(ga1); CG_TRY(maxpool2_bwd(gmp, d->bidx[b], ga1, B, 16, 16, Co));
    float* gc1 = BW(d, n1); NN(gc1); CG_TRY(prelu_bwd(d->bc1[b], ga1, d->P + d->bpw1[b], gc1, PG(d, d->bpw1[b]), n1));
    const float* bin = b < 3 ? d->stn[b + 1].out : d->T;
    float* gbin = BW(d, nT); NN(gbin); CG_TRY(layer_bwd(d, d->b1[b], bin, gc1, gbin, B, 16, 16, (d->dfused && b == 3) ? d->xq_b4 : nullptr));
    if (b < 3) {
      float* gs = BW(d, nT); NN(gs);
      CG_TRY(stn_backward(d, &d->stn[b + 1], gbin, gs, B));
      gT_part[b] = gs;
    } else gT_part[b] = gbin;
  }
  CG_TRY(lanes_join());
  for (int b = 0; b < 4; ++b) CG_TRY(add_inplace(gT, gT_part[b], nT));
  float* gtp = BW(d, nT); NN(gtp); CG_TRY(mask_channels(gT, mk_trunk, gtp, B, 256, 64));
  long n64 = (long)B * 1024 * 64;
  float* gta2 = BW(d, n64); NN(gta2); CG_TRY(avgpool2_bwd(gtp, gta2, B, 32, 32, 64));
  float* gtc2 = BW(d, n64); NN(gtc2); CG_TRY(prelu_bwd(d->tc2, gta2, d->P + d->t2pw, gtc2, PG(d, d->t2pw), n64));
  float* gta1 = BW(d, n64); NN(gta1); CG_TRY(layer_bwd(d, d->t2, d->dfused ? nullptr : d->ta1, gtc2, gta1, B, 32, 32, d->dfused ? d->xq_t2 : nullptr));
  float* gtc1 = BW(d, n64); NN(gtc1); CG_TRY(prelu_bwd(d->tc1, gta1, d->P + d->t1pw, gtc1, PG(d, d->t1pw), n64));
  float* gs0 = BW(d, (size_t)B * 1024 * C); NN(gs0); CG_TRY(layer_bwd(d, d->t1, d->stn[0].out, gtc1, gs0, B, 32, 32));
  float* gin = BW(d, (size_t)B * 1024 * C); NN(gin);
  CG_TRY(stn_backward(d, &d->stn[0], gs0, gin, B));
  if (gx_nchw) CG_TRY(nhwc_to_nchw(gin, gx_nchw, B, C, 1024));        // MODEL_D.modules[1].gradInput (adversarial.lua:193)
  CG_TRY(side_wait_all());   // parameter gradients are complete when this returns (stream order)
  return CG_OK;
}

}  // namespace cg

namespace cg {
// =================================================================== V (models.lua:765-804), evaluate() mode forward
// activation = nn.LeakyReLU with no argument = the repository's own module (LeakyReLU.lua:5-10, negative_scale 0.333),
// the same one D's localisation networks use.  (A newer upstream nn that already defines nn.LeakyReLU -- default 1/100 -- would win: LeakyReLU.lua:2-4.)
int V_forward_dev(cg_model* v, const float* x_nchw, int B, float* out_dev) {
  SplitScope split_scope;
  CG_TRY(model_repack(v, 1));   // V's shapes are not in conv_tc_all_shapes_taken's list: keep the fp32 operands fresh for any layer the engine declines
  Fp32StaleScope stale_scope(v->dirty32);
  v->nfw = 0; v->B = B;
  const int C = v->C;
  float* x = FW(v, (size_t)B * 1024 * C); NN(x);
  CG_TRY(nchw_to_nhwc(x_nchw, x, B, C, 1024));
  const int hw[4] = {32, 16, 8, 8}, co[4] = {128, 128, 256, 256};
  const float* cur = x; long r = 0;
  for (int i = 0; i < 4; ++i) {
    const long n = (long)B * hw[i] * hw[i] * co[i];
    float* y = FW(v, n); NN(y);
    CG_TRY(layer_fwd(v, v->vconv[i], cur, y, B, hw[i], hw[i]));
    if (i == 1 || i == 3) {   // SpatialBatchNormalization, running statistics
      float* b = FW(v, n); NN(b);
      CG_TRY(bn_fwd_eval(y, v->P + v->vbn[i / 2], v->P + v->vbn[i / 2] + co[i], b, v->run + r, v->run + r + co[i], (long)B * hw[i] * hw[i], co[i], 1e-5f));
      r += 2 * co[i]; y = b;
    }
    float* a = FW(v, n); NN(a);
    CG_TRY(lrelu_fwd(y, 0.333f, a, n));
    cur = a;
    if (i != 2) {             // SpatialMaxPooling(2,2) after conv 1, 2 and 4
      float* p = FW(v, n / 4); uint8_t* idx = (uint8_t*)FW(v, n / 16 + 4); NN(p); NN(idx);
      CG_TRY(maxpool2_fwd(a, p, idx, B, hw[i], hw[i], co[i]));
      cur = p;
    }
  }
  // nn.SpatialDropout() in evaluate(): y = (1 - p) x with p = 0.5; nn.Dropout() is the identity
  const long nf = (long)B * 4096;
  float* f = FW(v, nf); NN(f);
  CG_CUDA(cudaMemcpyAsync(f, cur, sizeof(float) * nf, cudaMemcpyDeviceToDevice, ctx().stream));
  CG_TRY(scale_inplace(f, 0.5f, nf));
  cur = f;
  for (int i = 0; i < 2; ++i) {
    float* y = FW(v, (size_t)B * 1024); float* b = FW(v, (size_t)B * 1024); float* a = FW(v, (size_t)B * 1024); NN(y); NN(b); NN(a);
    CG_TRY(layer_fwd(v, v->vlin[i], cur, y, B, 1, 1));
    CG_TRY(bn_fwd_eval(y, v->P + v->vbn[2 + i], v->P + v->vbn[2 + i] + 1024, b, v->run + r, v->run + r + 1024, B, 1024, 1e-5f)); r += 2048;
    CG_TRY(lrelu_fwd(b, 0.333f, a, (long)B * 1024));
    cur = a;
  }
  float* logit = FW(v, (size_t)B * 2); NN(logit);
  CG_TRY(layer_fwd(v, v->vlin[2], cur, logit, B, 1, 1));
  return softmax_rows(logit, out_dev, B, 2);
}
}  // namespace cg
