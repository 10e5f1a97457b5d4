// capi.cu -- extern "C" surface declared in include/catgen.h.  Host pointers in, host pointers out; all
// compute happens in this library's CUDA kernels (no CPU fallback: cg_init fails without an sm_100 device).
#include "model.cuh"
#include <math.h>
#include <new>
#ifdef CG_WITH_NCCL
#include <nccl.h>
#include <dlfcn.h>
// NCCL is bound at run time, not link time: a process that also imports torch already holds torch's own
// libnccl.so.2, and a second copy pulled in by this library's DT_NEEDED broke torch's symbol resolution
// (undefined ncclDevCommCreate, first GPU run).  Prefer the copy already loaded; otherwise load the system one.
namespace {
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi& nccl_api() {
  static NcclApi a; static bool tried = false;
  if (tried) return a;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) return a;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
  a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
  a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy && a.GetErrorString;
  return a;
}
}  // namespace
#define CG_NCCL_API() NcclApi& N = nccl_api(); if (!N.ok) return cg::set_err(CG_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "missing symbols")
#endif

namespace cg {
Ctx& ctx() { static Ctx c; return c; }
int set_err(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(ctx().err, sizeof(ctx().err), fmt, ap); va_end(ap);
  return code;
}
static void* grow(void** p, size_t* cap, size_t bytes) {
  if (*cap >= bytes && *p) return *p;
  cudaStreamSynchronize(ctx().stream);
  if (*p) cudaFree(*p);
  size_t want = bytes + bytes / 4 + 4096;
  ctx().alloc_gen++;   // captured step graphs hold raw pointers into these buffers (train_step_run re-captures)
  if (cudaMalloc(p, want) != cudaSuccess) { *p = nullptr; *cap = 0; set_err(CG_ERR_CUDA, "cudaMalloc(%zu) failed", want); return nullptr; }
  *cap = want;
  return *p;
}
void prof_begin(const char* name) {
  Ctx::ProfRec r; r.name = name; r.flops = ctx().next_flops; r.bytes = ctx().next_bytes; r.lane = ctx().lane;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b); cudaEventRecord(r.a, ctx().stream); ctx().prof.push_back(r);
}
void prof_end() { cudaEventRecord(ctx().prof.back().b, ctx().stream); }
void* workspace(size_t bytes) { return grow(&ctx().ws, &ctx().ws_bytes, bytes); }
void* workspace2(size_t bytes) { return grow(&ctx().ws2, &ctx().ws2_bytes, bytes); }
void* workspace3(size_t bytes) { return grow(&ctx().ws3, &ctx().ws3_bytes, bytes); }
void* workspace4(size_t bytes) { return grow(&ctx().ws4, &ctx().ws4_bytes, bytes); }
int lanes_fork(int first, int n) {
  Ctx& c = ctx();
  if (!c.lanes_on) return CG_OK;
  if (c.lane != -1) return set_err(CG_ERR_STATE, "lanes_fork inside a lane");
  if (!c.fork_ev) {
    CG_CUDA(cudaEventCreateWithFlags(&c.fork_ev, cudaEventDisableTiming));
    for (auto& L : c.lanes) { CG_CUDA(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking)); CG_CUDA(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming)); }
  }
  CG_CUDA(cudaEventRecord(c.fork_ev, c.stream));
  for (int b = first; b < first + n; ++b) CG_CUDA(cudaStreamWaitEvent(c.lanes[b].stream, c.fork_ev, 0));
  return CG_OK;
}
int lane_enter(int b) {
  Ctx& c = ctx();
  if (!c.lanes_on) return CG_OK;
  if (c.lane != -1 || b < 0 || b >= Ctx::kLanes || !c.fork_ev) return set_err(CG_ERR_STATE, "lane_enter(%d) in lane %d", b, c.lane);
  Ctx::Lane& L = c.lanes[b];
  c.saved.stream = c.stream; c.saved.ws = c.ws; c.saved.ws_bytes = c.ws_bytes; c.saved.ws3 = c.ws3; c.saved.ws3_bytes = c.ws3_bytes; c.saved.ws4 = c.ws4; c.saved.ws4_bytes = c.ws4_bytes;
  c.stream = L.stream; c.ws = L.ws; c.ws_bytes = L.ws_bytes; c.ws3 = L.ws3; c.ws3_bytes = L.ws3_bytes; c.ws4 = L.ws4; c.ws4_bytes = L.ws4_bytes;
  c.lane = b;
  return CG_OK;
}
int lane_exit() {
  Ctx& c = ctx();
  if (!c.lanes_on) return CG_OK;
  if (c.lane < 0) return set_err(CG_ERR_STATE, "lane_exit outside a lane");
  Ctx::Lane& L = c.lanes[c.lane];
  L.ws = c.ws; L.ws_bytes = c.ws_bytes; L.ws3 = c.ws3; L.ws3_bytes = c.ws3_bytes; L.ws4 = c.ws4; L.ws4_bytes = c.ws4_bytes;   // scratch may have grown
  c.stream = c.saved.stream; c.ws = c.saved.ws; c.ws_bytes = c.saved.ws_bytes; c.ws3 = c.saved.ws3; c.ws3_bytes = c.saved.ws3_bytes; c.ws4 = c.saved.ws4; c.ws4_bytes = c.saved.ws4_bytes;
  c.lane = -1;
  return CG_OK;
}
int lanes_join(int first, int n) {
  Ctx& c = ctx();
  if (!c.lanes_on) return CG_OK;
  if (c.lane != -1) return set_err(CG_ERR_STATE, "lanes_join inside a lane");
  for (int b = first; b < first + n; ++b) { Ctx::Lane& L = c.lanes[b]; CG_CUDA(cudaEventRecord(L.done, L.stream)); CG_CUDA(cudaStreamWaitEvent(c.stream, L.done, 0)); }
  for (int b = first + 1; b <= first + n; ++b) if (c.side[b].pending) { CG_CUDA(cudaStreamWaitEvent(c.stream, c.side[b].done, 0)); c.side[b].pending = false; }
  return CG_OK;
}
int side_begin() {
  Ctx& c = ctx();
  if (!c.side_on || c.in_side) return 0;
  Ctx::Side& W = c.side[c.lane + 1];
  if (!W.stream) {
    if (cudaStreamCreateWithFlags(&W.stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&W.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&W.done, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); W.stream = nullptr; return 0; }
  }
  if (cudaEventRecord(W.fork, c.stream) != cudaSuccess || cudaStreamWaitEvent(W.stream, W.fork, 0) != cudaSuccess) { cudaGetLastError(); return 0; }
  c.side_saved.stream = c.stream; c.side_saved.ws = c.ws; c.side_saved.ws_bytes = c.ws_bytes; c.side_saved.ws3 = c.ws3; c.side_saved.ws3_bytes = c.ws3_bytes;
  c.stream = W.stream; c.ws = W.ws; c.ws_bytes = W.ws_bytes; c.ws3 = W.ws3; c.ws3_bytes = W.ws3_bytes;
  c.in_side = true;
  return 1;
}
int side_end() {
  Ctx& c = ctx();
  if (!c.in_side) return CG_OK;
  Ctx::Side& W = c.side[c.lane + 1];
  cudaError_t e = cudaEventRecord(W.done, W.stream);
  W.pending = true;
  W.ws = c.ws; W.ws_bytes = c.ws_bytes; W.ws3 = c.ws3; W.ws3_bytes = c.ws3_bytes;
  c.stream = c.side_saved.stream; c.ws = c.side_saved.ws; c.ws_bytes = c.side_saved.ws_bytes; c.ws3 = c.side_saved.ws3; c.ws3_bytes = c.side_saved.ws3_bytes;
  c.in_side = false;
  if (e != cudaSuccess) return set_err(CG_ERR_CUDA, "side stream record failed: %s", cudaGetErrorString(e));
  return CG_OK;
}
int side_wait() {
  Ctx& c = ctx();
  if (c.in_side) return set_err(CG_ERR_STATE, "side_wait on a side stream");
  Ctx::Side& W = c.side[c.lane + 1];
  if (W.pending) { CG_CUDA(cudaStreamWaitEvent(c.stream, W.done, 0)); W.pending = false; }
  return CG_OK;
}
int side_wait_all() {
  Ctx& c = ctx();
  if (c.in_side || c.lane != -1) return set_err(CG_ERR_STATE, "side_wait_all outside the main stream");
  for (auto& W : c.side) if (W.pending) { CG_CUDA(cudaStreamWaitEvent(c.stream, W.done, 0)); W.pending = false; }
  return CG_OK;
}
void* pinned(size_t bytes) {
  Ctx& c = ctx();
  if (c.pinned_bytes >= bytes && c.pinned) return c.pinned;
  cudaStreamSynchronize(c.stream);
  if (c.pinned) cudaFreeHost(c.pinned);
  if (cudaMallocHost(&c.pinned, bytes + 4096) != cudaSuccess) { c.pinned = nullptr; c.pinned_bytes = 0; return nullptr; }
  c.pinned_bytes = bytes + 4096;
  return c.pinned;
}
int DBuf::ensure(size_t nfloats) {
  if (n >= nfloats && p) return CG_OK;
  if (p) { cudaStreamSynchronize(ctx().stream); cudaFree(p); p = nullptr; n = 0; }
  ctx().alloc_gen++;
  if (nfloats == 0) nfloats = 1;
  if (cudaMalloc(&p, sizeof(float) * nfloats) != cudaSuccess) { p = nullptr; return set_err(CG_ERR_CUDA, "cudaMalloc(%zu floats) failed", nfloats); }
  n = nfloats;
  return CG_OK;
}
void DBuf::release() { if (p) cudaFree(p); p = nullptr; n = 0; }

// per-call temporary device buffers for the op-level entry points (unit-parity surface; not the hot path)
struct Tmp {
  std::vector<void*> v; bool ok = true;
  float* dev(size_t n) { void* p = nullptr; if (cudaMalloc(&p, sizeof(float) * (n ? n : 1)) != cudaSuccess) { ok = false; return nullptr; } v.push_back(p); return (float*)p; }
  float* up(const float* h, size_t n) {
    float* d = dev(n); if (!d) return nullptr;
    if (cudaMemcpyAsync(d, h, sizeof(float) * n, cudaMemcpyHostToDevice, ctx().stream) != cudaSuccess) ok = false;
    return d;
  }
  int down(float* h, const float* d, size_t n) {
    if (cudaMemcpyAsync(h, d, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx().stream) != cudaSuccess) return set_err(CG_ERR_CUDA, "D2H failed");
    return CG_OK;
  }
  int finish() {
    cudaError_t e = cudaStreamSynchronize(ctx().stream);
    if (e != cudaSuccess) return set_err(CG_ERR_CUDA, "stream sync: %s", cudaGetErrorString(e));
    return ok ? CG_OK : set_err(CG_ERR_CUDA, "temporary device allocation/copy failed");
  }
  ~Tmp() { cudaStreamSynchronize(ctx().stream); for (void* p : v) cudaFree(p); }
};
}  // namespace cg
namespace cg {
// sum over ranks, in place, on the current stream (graph-capturable); no-op for one rank
int dist_allreduce_sum_f32(float* buf, long n) {
  if (ctx().world <= 1) return CG_OK;
#ifdef CG_WITH_NCCL
  CG_NCCL_API();
  ncclResult_t r = N.AllReduce(buf, buf, (size_t)n, ncclFloat, ncclSum, (ncclComm_t)ctx().nccl, ctx().stream);
  if (r != ncclSuccess) return set_err(CG_ERR_NCCL, "ncclAllReduce: %s", N.GetErrorString(r));
  return CG_OK;
#else
  return set_err(CG_ERR_NCCL, "library built without NCCL");
#endif
}
int dist_allreduce_sum_f64(double* buf, long n) {
  if (ctx().world <= 1) return CG_OK;
#ifdef CG_WITH_NCCL
  CG_NCCL_API();
  ncclResult_t r = N.AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)ctx().nccl, ctx().stream);
  if (r != ncclSuccess) return set_err(CG_ERR_NCCL, "ncclAllReduce: %s", N.GetErrorString(r));
  return CG_OK;
#else
  return set_err(CG_ERR_NCCL, "library built without NCCL");
#endif
}
}  // namespace cg

using namespace cg;

extern "C" {

int cg_init(int device) {
  Ctx& c = ctx();
  if (c.inited) { if (c.device == device) return CG_OK; return set_err(CG_ERR_STATE, "already initialised on device %d", c.device); }
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    return set_err(CG_ERR_NODEVICE, "no CUDA device: libcatgen has no CPU fallback (B200 / sm_100a required)");
  if (device < 0 || device >= n) return set_err(CG_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
  CG_CUDA(cudaSetDevice(device));
  cudaDeviceProp p; CG_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) return set_err(CG_ERR_NODEVICE, "device %d is sm_%d%d; this library contains sm_100a code only", device, p.major, p.minor);
  c.sm_count = p.multiProcessorCount;
  CG_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  { const char* e = getenv("CATGEN_LANES"); c.lanes_on = !(e && e[0] == '0'); }
  { const char* e = getenv("CATGEN_SIDE"); c.side_on = !(e && e[0] == '0'); }
  c.device = device; c.inited = true; c.launches = 0; c.trace_launches = getenv("CATGEN_LAUNCH_TRACE") != nullptr;
  // the page-locked scratch the loss read-back uses: allocated here, not on the first cg_train_step that asks for a loss (cudaMallocHost took
  // 1-35 ms there -- profiles/r02_bench_hiccup.txt -- and fell inside whatever region that step belonged to)
  if (!pinned(4096)) return set_err(CG_ERR_CUDA, "cudaMallocHost failed");
  return CG_OK;
}
void cg_shutdown(void) {
  Ctx& c = ctx();
  if (!c.inited) return;
  cudaStreamSynchronize(c.stream);
#ifdef CG_WITH_NCCL
  if (c.nccl && nccl_api().ok) { nccl_api().CommDestroy((ncclComm_t)c.nccl); c.nccl = nullptr; }
#endif
  if (c.ws) cudaFree(c.ws); if (c.ws2) cudaFree(c.ws2); if (c.ws3) cudaFree(c.ws3); if (c.ws4) cudaFree(c.ws4); if (c.pinned) cudaFreeHost(c.pinned);
  c.ws = c.ws2 = c.ws3 = c.ws4 = c.pinned = nullptr; c.ws_bytes = c.ws2_bytes = c.ws3_bytes = c.ws4_bytes = c.pinned_bytes = 0;
  for (auto& L : c.lanes) {
    if (L.ws) cudaFree(L.ws); if (L.ws3) cudaFree(L.ws3); if (L.ws4) cudaFree(L.ws4);
    if (L.stream) cudaStreamDestroy(L.stream); if (L.done) cudaEventDestroy(L.done);
    L = Ctx::Lane();
  }
  if (c.fork_ev) { cudaEventDestroy(c.fork_ev); c.fork_ev = nullptr; }
  for (auto& W : c.side) {
    if (W.ws) cudaFree(W.ws); if (W.ws3) cudaFree(W.ws3);
    if (W.stream) cudaStreamDestroy(W.stream); if (W.fork) cudaEventDestroy(W.fork); if (W.done) cudaEventDestroy(W.done);
    W = Ctx::Side();
  }
  c.in_side = false;
  c.lane = -1;
  cudaStreamDestroy(c.stream); c.stream = nullptr; c.inited = false; c.device = -1; c.world = 1; c.rank = 0;
}
const char* cg_last_error(void) { return ctx().err; }
const char* cg_version(void) { return "catgen-b200 0.1 (sm_100a)"; }
int cg_sync(void) { CG_REQUIRE_INIT(); CG_CUDA(cudaStreamSynchronize(ctx().stream)); return CG_OK; }
int64_t cg_launch_count(void) { return ctx().launches; }
void cg_reset_launch_count(void) { ctx().launches = 0; }
// ---- instrumentation: replaces the reference's sys.clock() epoch timing (adversarial.lua:34,278-280)
int cg_timer_start(void) {
  CG_REQUIRE_INIT(); Ctx& c = ctx();
  if (!c.t0) { CG_CUDA(cudaEventCreate(&c.t0)); CG_CUDA(cudaEventCreate(&c.t1)); }
  CG_CUDA(cudaEventRecord(c.t0, c.stream)); return CG_OK;
}
int cg_timer_stop(float* ms) {
  CG_REQUIRE_INIT(); CG_ARG(ms); Ctx& c = ctx();
  if (!c.t0) return set_err(CG_ERR_STATE, "cg_timer_start was not called");
  CG_CUDA(cudaEventRecord(c.t1, c.stream)); CG_CUDA(cudaEventSynchronize(c.t1)); CG_CUDA(cudaEventElapsedTime(ms, c.t0, c.t1)); return CG_OK;
}
int cg_profile_enable(int on) {
  CG_REQUIRE_INIT(); Ctx& c = ctx(); cudaStreamSynchronize(c.stream);
  for (auto& r : c.prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  c.prof.clear(); c.prof_on = on != 0; return CG_OK;
}
int cg_profile_report(char* out, int cap) {
  CG_REQUIRE_INIT(); CG_ARG(out && cap > 64); Ctx& c = ctx();
  CG_CUDA(cudaStreamSynchronize(c.stream));
  struct Agg { const char* name; long n; double ms, flops, bytes, f1, b1; };
  std::vector<Agg> ag;
  for (auto& r : c.prof) {
    float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
    // one row per (kernel, algorithmic work per launch): launches of the same kernel on different layer shapes stay apart (per-shape TFLOP/s, GB/s)
    size_t i = 0; for (; i < ag.size(); ++i) if (!strcmp(ag[i].name, r.name) && ag[i].f1 == r.flops && ag[i].b1 == r.bytes) break;
    if (i == ag.size()) ag.push_back({r.name, 0, 0, 0, 0, r.flops, r.bytes});
    ag[i].n++; ag[i].ms += ms; ag[i].flops += r.flops; ag[i].bytes += r.bytes;
  }
  int o = snprintf(out, cap, "[");
  for (size_t i = 0; i < ag.size() && o < cap - 200; ++i)
    o += snprintf(out + o, cap - o, "%s{\"kernel\":\"%s\",\"launches\":%ld,\"ms\":%.6f,\"flops\":%.6e,\"bytes\":%.6e}", i ? "," : "", ag[i].name, ag[i].n, ag[i].ms, ag[i].flops, ag[i].bytes);
  snprintf(out + o, cap - o, "]");
  return CG_OK;
}
// one row per launch since cg_profile_enable(1): [kernel, lane (-1 = main stream), start_us, end_us] relative to the first launch.  With
// cg_set_concurrency(1) the rows of different lanes overlap: this is the step's timeline (what runs beside what, where the streams idle).
int cg_profile_timeline(char* out, int cap) {
  CG_REQUIRE_INIT(); CG_ARG(out && cap > 64); Ctx& c = ctx();
  CG_CUDA(cudaDeviceSynchronize());
  int o = snprintf(out, cap, "[");
  for (size_t i = 0; i < c.prof.size() && o < cap - 160; ++i) {
    float t0 = 0, t1 = 0; cudaEventElapsedTime(&t0, c.prof[0].a, c.prof[i].a); cudaEventElapsedTime(&t1, c.prof[0].a, c.prof[i].b);
    o += snprintf(out + o, cap - o, "%s[\"%s\",%d,%.1f,%.1f]", i ? "," : "", c.prof[i].name, c.prof[i].lane, 1e3 * t0, 1e3 * t1);
  }
  snprintf(out + o, cap - o, "]");
  return CG_OK;
}
int cg_set_graph_mode(int on) { ctx().graph_mode = on ? 1 : 0; return CG_OK; }
int cg_get_graph_mode(void) { return ctx().graph_mode; }
int cg_set_conv_engine(int e) { if (e != 0 && e != 1) return set_err(CG_ERR_ARG, "engine must be 0 or 1"); ctx().conv_engine = e; return CG_OK; }
int cg_get_conv_engine(void) { return ctx().conv_engine; }
int cg_set_precision(int m) { if (m != 0 && m != 1) return set_err(CG_ERR_ARG, "precision must be 0 or 1"); ctx().precision = m; return CG_OK; }
int cg_get_precision(void) { return ctx().precision; }
int cg_set_concurrency(int on) { CG_REQUIRE_INIT(); if (ctx().lane != -1 || ctx().in_side) return set_err(CG_ERR_STATE, "cg_set_concurrency inside a lane"); cudaStreamSynchronize(ctx().stream); ctx().lanes_on = ctx().side_on = on ? 1 : 0; return CG_OK; }
int cg_set_dead_grad_elim(int on) { ctx().dead_grad_elim = on ? 1 : 0; return CG_OK; }

// ------------------------------------------------------------------ models
int cg_model_create(cg_model** out, int kind, int C, int nz, uint64_t seed) {
  CG_REQUIRE_INIT(); CG_ARG(out); CG_ARG(C == 1 || C == 3); CG_ARG(nz > 0);
  cg_model* m = new (std::nothrow) cg_model();
  if (!m) return set_err(CG_ERR_STATE, "out of host memory");
  m->kind = kind; m->C = C; m->nz = nz; m->seed = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  int s = model_build(m);
  if (s != CG_OK) { delete m; return s; }
  CG_CUDA(cudaStreamSynchronize(ctx().stream));
  *out = m; return CG_OK;
}
int cg_model_free(cg_model* m) {
  if (!m) return CG_OK;
  cudaStreamSynchronize(ctx().stream);
  for (auto& L : m->layers) { cg::conv_tc_unregister_wslices(L.Wp); cg::conv_tc_unregister_wslices(L.Wd); }
  if (m->wq) cudaFree(m->wq); if (m->jobs_dev) cudaFree(m->jobs_dev);
  if (m->P) cudaFree(m->P); if (m->G) cudaFree(m->G); if (m->packed) cudaFree(m->packed); if (m->run) cudaFree(m->run);
  if (m->masks) cudaFree(m->masks); if (m->amax) cudaFree(m->amax); if (m->mq) cudaFree(m->mq); if (m->rng_dev) cudaFree(m->rng_dev);
  for (auto& b : m->fw) b.release(); for (auto& b : m->bw) b.release(); m->gwp.release();
  delete m; return CG_OK;
}
int cg_model_nparams(const cg_model* m, int64_t* n) { CG_ARG(m && n); *n = m->np; return CG_OK; }
int cg_model_get_params(cg_model* m, float* host) {
  CG_REQUIRE_INIT(); CG_ARG(m && host);
  CG_CUDA(cudaMemcpyAsync(host, m->P, sizeof(float) * m->np, cudaMemcpyDeviceToHost, ctx().stream)); CG_CUDA(cudaStreamSynchronize(ctx().stream)); return CG_OK;
}
int cg_model_set_params(cg_model* m, const float* host) {
  CG_REQUIRE_INIT(); CG_ARG(m && host);
  CG_CUDA(cudaMemcpyAsync(m->P, host, sizeof(float) * m->np, cudaMemcpyHostToDevice, ctx().stream)); CG_CUDA(cudaStreamSynchronize(ctx().stream));
  m->dirty = true; m->dirty32 = true; return CG_OK;
}
int cg_model_get_grads(cg_model* m, float* host) {
  CG_REQUIRE_INIT(); CG_ARG(m && host);
  CG_CUDA(cudaMemcpyAsync(host, m->G, sizeof(float) * m->np, cudaMemcpyDeviceToHost, ctx().stream)); CG_CUDA(cudaStreamSynchronize(ctx().stream)); return CG_OK;
}
int cg_model_zero_grads(cg_model* m) { CG_REQUIRE_INIT(); CG_ARG(m); CG_CUDA(cudaMemsetAsync(m->G, 0, sizeof(float) * m->np, ctx().stream)); return CG_OK; }
int cg_model_bn_running_len(const cg_model* m, int64_t* n) { CG_ARG(m && n); *n = m->nrun; return CG_OK; }
int cg_model_get_bn_running(cg_model* m, float* host) {
  CG_REQUIRE_INIT(); CG_ARG(m && host); if (!m->nrun) return CG_OK;
  CG_CUDA(cudaMemcpyAsync(host, m->run, sizeof(float) * m->nrun, cudaMemcpyDeviceToHost, ctx().stream)); CG_CUDA(cudaStreamSynchronize(ctx().stream)); return CG_OK;
}
int cg_model_set_bn_running(cg_model* m, const float* host) {
  CG_REQUIRE_INIT(); CG_ARG(m && host); if (!m->nrun) return CG_OK;
  CG_CUDA(cudaMemcpyAsync(m->run, host, sizeof(float) * m->nrun, cudaMemcpyHostToDevice, ctx().stream)); CG_CUDA(cudaStreamSynchronize(ctx().stream)); return CG_OK;
}
int cg_model_set_mode(cg_model* m, int training) { CG_ARG(m); m->training = training ? 1 : 0; return CG_OK; }

int cg_G_forward(cg_model* g, const float* z, int B, float* out) {
  CG_REQUIRE_INIT(); CG_ARG(g && z && out && B > 0); CG_ARG(g->kind == CG_G32UP || g->kind == CG_G32UPC);
  size_t nz = (size_t)B * g->nz, no = (size_t)B * g->C * 1024;
  float* st = (float*)workspace2(sizeof(float) * (nz + no)); if (!st) return CG_ERR_CUDA;
  CG_CUDA(cudaMemcpyAsync(st, z, sizeof(float) * nz, cudaMemcpyHostToDevice, ctx().stream));
  CG_TRY(G_forward_dev(g, st, B, st + nz));
  CG_CUDA(cudaMemcpyAsync(out, st + nz, sizeof(float) * no, cudaMemcpyDeviceToHost, ctx().stream));
  CG_CUDA(cudaStreamSynchronize(ctx().stream)); return CG_OK;
}
int cg_G_backward(cg_model* g, const float* gout, float* gz) {
  CG_REQUIRE_INIT(); CG_ARG(g && gout); CG_ARG(g->kind == CG_G32UP || g->kind == CG_G32UPC);
  int B = g->B; size_t no = (size_t)B * g->C * 1024, nz = (size_t)B * g->nz;
  float* st = (float*)workspace2(sizeof(float) * (no + nz)); if (!st) return CG_ERR_CUDA;
  CG_CUDA(cudaMemcpyAsync(st, gout, sizeof(float) * no, cudaMemcpyHostToDevice, ctx().stream));
  CG_TRY(G_backward_dev(g, st, st + no));
  if (gz) CG_CUDA(cudaMemcpyAsync(gz, st + no, sizeof(float) * nz, cudaMemcpyDeviceToHost, ctx().stream));
  CG_CUDA(cudaStreamSynchronize(ctx().stream)); return CG_OK;
}
int cg_D_forward(cg_model* d, const float* x, int B, float* out_sig, float* out_pre) {
  CG_REQUIRE_INIT(); CG_ARG(d && x && B > 0); CG_ARG(d->kind == CG_D32_ST3);
  size_t nx = (size_t)B * d->C * 1024;
  float* st = (float*)workspace2(sizeof(float) * (nx + 2 * (size_t)B)); if (!st) return CG_ERR_CUDA;
  CG_CUDA(cudaMemcpyAsync(st, x, sizeof(float) * nx, cudaMemcpyHostToDevice, ctx().stream));
  CG_TRY(D_forward_dev(d, st, B, st + nx, st + nx + B));
  if (out_sig) CG_CUDA(cudaMemcpyAsync(out_sig, st + nx, sizeof(float) * B, cudaMemcpyDeviceToHost, ctx().stream));
  if (out_pre) CG_CUDA(cudaMemcpyAsync(out_pre, st + nx + B, sizeof(float) * B, cudaMemcpyDeviceToHost, ctx().stream));
  CG_CUDA(cudaStreamSynchronize(ctx().stream)); return CG_OK;
}
int cg_D_backward(cg_model* d, const float* gout, float* gx) {
  CG_REQUIRE_INIT(); CG_ARG(d && gout); CG_ARG(d->kind == CG_D32_ST3);
  int B = d->B; size_t nx = (size_t)B * d->C * 1024;
  float* st = (float*)workspace2(sizeof(float) * (nx + (size_t)B)); if (!st) return CG_ERR_CUDA;
  CG_CUDA(cudaMemcpyAsync(st, gout, sizeof(float) * B, cudaMemcpyHostToDevice, ctx().stream));
  CG_TRY(D_backward_dev(d, st, st + B));
  if (gx) CG_CUDA(cudaMemcpyAsync(gx, st + B, sizeof(float) * nx, cudaMemcpyDeviceToHost, ctx().stream));
  CG_CUDA(cudaStreamSynchronize(ctx().stream)); return CG_OK;
}
int cg_V_forward(cg_model* v, const float* x, int B, float* out) {
  CG_REQUIRE_INIT(); CG_ARG(v && x && out && B > 0); CG_ARG(v->kind == CG_V32);
  size_t nx = (size_t)B * v->C * 1024;
  float* st = (float*)workspace2(sizeof(float) * (nx + 2 * (size_t)B)); if (!st) return CG_ERR_CUDA;
  CG_CUDA(cudaMemcpyAsync(st, x, sizeof(float) * nx, cudaMemcpyHostToDevice, ctx().stream));
  CG_TRY(V_forward_dev(v, st, B, st + nx));
  CG_CUDA(cudaMemcpyAsync(out, st + nx, sizeof(float) * 2 * B, cudaMemcpyDeviceToHost, ctx().stream));
  CG_CUDA(cudaStreamSynchronize(ctx().stream)); return CG_OK;
}
int cg_D_mask_floats(int B, int64_t* n) { CG_ARG(n && B > 0); *n = D_mask_floats(B); return CG_OK; }
int cg_D_get_masks(cg_model* d, float* host) {
  CG_REQUIRE_INIT(); CG_ARG(d && host); if (!d->masks || !d->masks_B) return set_err(CG_ERR_STATE, "no forward has run yet");
  CG_CUDA(cudaMemcpyAsync(host, d->masks, sizeof(float) * D_mask_floats(d->masks_B), cudaMemcpyDeviceToHost, ctx().stream));
  CG_CUDA(cudaStreamSynchronize(ctx().stream)); return CG_OK;
}
int cg_D_set_masks(cg_model* d, const float* host, int B, int count) {
  CG_REQUIRE_INIT(); CG_ARG(d && host && B > 0 && count > 0);
  size_t n = (size_t)D_mask_floats(B) * count;
  if (d->mq) { cudaStreamSynchronize(ctx().stream); cudaFree(d->mq); d->mq = nullptr; }
  CG_CUDA(cudaMalloc(&d->mq, sizeof(float) * n));
  CG_CUDA(cudaMemcpyAsync(d->mq, host, sizeof(float) * n, cudaMemcpyHostToDevice, ctx().stream)); CG_CUDA(cudaStreamSynchronize(ctx().stream));
  d->mq_count = count; d->mq_next = 0; d->mq_B = B; return CG_OK;
}

// ------------------------------------------------------------------ criterion / optimiser
int cg_bce(const float* p, const float* t, int n, float* loss, float* g) {
  CG_REQUIRE_INIT(); CG_ARG(p && t && n > 0);
  Tmp T; float* dp = T.up(p, n); float* dt = T.up(t, n); float* dg = T.dev(n); float* dl = T.dev(1);
  if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(bce(dp, dt, n, dl, dg));
  if (loss) CG_TRY(T.down(loss, dl, 1)); if (g) CG_TRY(T.down(g, dg, n));
  return T.finish();
}
int cg_penalty_clamp(cg_model* m, float l1, float l2sign, float l2, float clampv, float* loss_add) {
  CG_REQUIRE_INIT(); CG_ARG(m);
  Tmp T; float* dl = T.dev(1); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(penalty_clamp(m->G, m->P, m->np, l1, l2sign, l2, clampv, dl));
  if (loss_add) CG_TRY(T.down(loss_add, dl, 1));
  return T.finish();
}
int cg_trainer_create(cg_trainer** out, cg_model* G, cg_model* D) {
  CG_REQUIRE_INIT(); CG_ARG(out && G && D); CG_ARG((G->kind == CG_G32UP || G->kind == CG_G32UPC) && D->kind == CG_D32_ST3 && G->C == D->C);
  cg_trainer* t = new (std::nothrow) cg_trainer(); if (!t) return set_err(CG_ERR_STATE, "out of host memory");
  t->G = G; t->D = D;
  CG_CUDA(cudaMalloc(&t->mD, sizeof(float) * D->np)); CG_CUDA(cudaMalloc(&t->vD, sizeof(float) * D->np));
  CG_CUDA(cudaMalloc(&t->mG, sizeof(float) * G->np)); CG_CUDA(cudaMalloc(&t->vG, sizeof(float) * G->np));
  CG_CUDA(cudaMemsetAsync(t->mD, 0, sizeof(float) * D->np, ctx().stream)); CG_CUDA(cudaMemsetAsync(t->vD, 0, sizeof(float) * D->np, ctx().stream));
  CG_CUDA(cudaMemsetAsync(t->mG, 0, sizeof(float) * G->np, ctx().stream)); CG_CUDA(cudaMemsetAsync(t->vG, 0, sizeof(float) * G->np, ctx().stream));
  CG_CUDA(cudaMalloc(&t->t_dev, 2 * sizeof(int))); CG_CUDA(cudaMemsetAsync(t->t_dev, 0, 2 * sizeof(int), ctx().stream));
  *out = t; return CG_OK;
}
int cg_trainer_free(cg_trainer* t) {
  if (!t) return CG_OK;
  cudaStreamSynchronize(ctx().stream);
  cudaFree(t->mD); cudaFree(t->vD); cudaFree(t->mG); cudaFree(t->vG); cudaFree(t->t_dev);
  t->inputs.release(); t->targets.release(); t->samples.release(); t->dout.release(); t->df.release(); t->gimg.release(); t->scal.release();
  t->stage.release(); t->gin.release();
  for (auto& e : t->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
  delete t; return CG_OK;
}
int cg_adam_step(cg_trainer* t, int which, const cg_step_cfg* cfg) {
  CG_REQUIRE_INIT(); CG_ARG(t && cfg && (which == 0 || which == 1));
  if (which == 0) { CG_TRY(adam(t->D->P, t->D->G, t->mD, t->vD, t->D->np, t->t_dev + 0, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps)); t->D->dirty = true; }
  else { CG_TRY(adam(t->G->P, t->G->G, t->mG, t->vG, t->G->np, t->t_dev + 1, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps)); t->G->dirty = true; }
  return CG_OK;
}

int cg_dist_allreduce_grads(cg_model* m) {
  CG_REQUIRE_INIT(); CG_ARG(m);
  if (ctx().world <= 1) return CG_OK;
#ifdef CG_WITH_NCCL
  // sum over ranks then scale by 1/world: BCE is a mean over the LOCAL batch
  CG_NCCL_API();
  CG_TRY(cg::dist_allreduce_sum_f32(m->G, m->np));
  return scale_inplace(m->G, 1.f / ctx().world, m->np);
#else
  return set_err(CG_ERR_NCCL, "library built without NCCL");
#endif
}

// one adversarial.train loop body on device-resident inputs.  scal: [lossD(d_iters), pen(d_iters), lossG(g_iters), pen(g_iters)]
struct LaneGuard { int status; bool in; explicit LaneGuard(int b) : status(cg::lane_enter(b)), in(status == CG_OK) {} ~LaneGuard() { if (in) cg::lane_exit(); } };

static int train_step_core(cg_trainer* t, const cg_step_cfg* c, const float* real, const float* zD, const float* zG, float* lossD, float* lossG) {
  cg_model *G = t->G, *D = t->D;
  int B = c->B, hB = B / 2, C = G->C, nz = G->nz; size_t img = (size_t)C * 1024;
  CG_ARG(B >= 4 && B % 2 == 0);   // adversarial.lua:65-68 skips batches smaller than 4
  CG_TRY(t->inputs.ensure(B * img)); CG_TRY(t->targets.ensure(2 * (size_t)B)); CG_TRY(t->samples.ensure(B * img));
  CG_TRY(t->dout.ensure(B)); CG_TRY(t->df.ensure(B)); CG_TRY(t->gimg.ensure(B * img));
  int ns = 2 * (c->d_iters + c->g_iters); CG_TRY(t->scal.ensure(ns));
  float* tgtD = t->targets.p; float* tgtG = t->targets.p + B;
  CG_TRY(fill(tgtD, 1.f, hB)); CG_TRY(fill(tgtD + hB, 0.f, B - hB)); CG_TRY(fill(tgtG, 1.f, B));   // Y_NOT_GENERATOR=1, Y_GENERATOR=0 (train.lua:70-71)
  int si = 0;
  bool g_ahead = false;
  for (int k = 0; k < c->d_iters; ++k) {
    // (1.1) real half, (1.2) fake half from a separate G forward on B/2 noise vectors (adversarial.lua:223-238)
    CG_CUDA(cudaMemcpyAsync(t->inputs.p, real + (size_t)k * hB * img, sizeof(float) * hB * img, cudaMemcpyDeviceToDevice, ctx().stream));
    CG_TRY(G_forward_dev(G, zD + (size_t)k * hB * nz, hB, t->inputs.p + hB * img));
    // fevalG's generator forward depends on nothing fevalD changes (G's parameters move only in fevalG; its BN running
    // statistics are updated in issue order): start it now on its own lane, beside D's forward/backward/Adam below.
    if (k == c->d_iters - 1 && c->g_iters > 0 && ctx().lanes_on && !(ctx().sync_bn && ctx().world > 1)) {   // sync-BN: G's forward holds collectives, keep ONE issue order on the communicator
      CG_TRY(lanes_fork(4, 1));
      int st;
      { LaneGuard lane(4); st = lane.status; if (st == CG_OK) st = G_forward_dev(G, zG, B, t->samples.p); }
      CG_TRY(st);
      g_ahead = true;
    }
    // fevalD (adversarial.lua:72-167)
    CG_CUDA(cudaMemsetAsync(D->G, 0, sizeof(float) * D->np, ctx().stream));
    CG_TRY(D_forward_dev(D, t->inputs.p, B, t->dout.p, nullptr));
    CG_TRY(bce(t->dout.p, tgtD, B, t->scal.p + si, t->df.p));
    CG_TRY(D_backward_dev(D, t->df.p, nullptr));
    CG_TRY(dist_allreduce_sum_f32(D->G, D->np));   // BCE is a mean over the LOCAL batch: sum over ranks, 1/world folded into the pass below
    CG_TRY(penalty_clamp_adam(D->G, D->P, t->mD, t->vD, D->np, 1.f / ctx().world, c->D_L1, c->D_L1, c->D_L2, c->D_clamp, t->scal.p + si + 1,
                              t->t_dev + 0, c->lr, c->beta1, c->beta2, c->eps)); D->dirty = true;   // :92-112 then optim.adam :245
    si += 2;
  }
  for (int k = 0; k < c->g_iters; ++k) {
    // fevalG_on_D (adversarial.lua:171-215)
    CG_CUDA(cudaMemsetAsync(G->G, 0, sizeof(float) * G->np, ctx().stream));
    if (k == 0 && g_ahead) CG_TRY(lanes_join(4, 1));
    else CG_TRY(G_forward_dev(G, zG + (size_t)k * B * nz, B, t->samples.p));
    CG_TRY(D_forward_dev(D, t->samples.p, B, nullptr, nullptr));   // t->dout keeps the last D-phase outputs for d_out
    float* dsig = D->hsig;
    CG_TRY(bce(dsig, tgtG, B, t->scal.p + si, t->df.p));
    // The reference's MODEL_D:backward here also accumulates D's parameter gradients (adversarial.lua:193), which nothing reads:
    // the next fevalD zeroes them (:78).  By default only the input-gradient path runs; cg_set_dead_grad_elim(0) restores the
    // accumulation (then cg_model_get_grads(D) after a step returns what Torch's gradParameters would hold).
    D->skip_param_grads = ctx().dead_grad_elim;
    int bst = D_backward_dev(D, t->df.p, t->gimg.p);
    D->skip_param_grads = 0;
    CG_TRY(bst);
    CG_TRY(G_backward_dev(G, t->gimg.p, nullptr));
    CG_TRY(dist_allreduce_sum_f32(G->G, G->np));
    CG_TRY(penalty_clamp_adam(G->G, G->P, t->mG, t->vG, G->np, 1.f / ctx().world, c->G_L1, c->G_L2, c->G_L2, c->G_clamp, t->scal.p + si + 1,   // sign term uses G_L2 (adversarial.lua:206)
                              t->t_dev + 1, c->lr, c->beta1, c->beta2, c->eps)); G->dirty = true;   // :201-212 then optim.adam :262
    si += 2;
  }
  return CG_OK;
}
// losses of the step just enqueued: D2H + sync, kept OUTSIDE the captured graph
static int read_losses(cg_trainer* t, const cg_step_cfg* c, float* lossD, float* lossG) {
  if (!lossD && !lossG) return CG_OK;
  int ns = 2 * (c->d_iters + c->g_iters);
  float* h = (float*)pinned(sizeof(float) * ns); if (!h) return set_err(CG_ERR_CUDA, "pinned alloc failed");
  CG_CUDA(cudaMemcpyAsync(h, t->scal.p, sizeof(float) * ns, cudaMemcpyDeviceToHost, ctx().stream));
  CG_CUDA(cudaStreamSynchronize(ctx().stream));
  for (int k = 0; k < c->d_iters; ++k) if (lossD) lossD[k] = h[2 * k] + h[2 * k + 1];
  for (int k = 0; k < c->g_iters; ++k) if (lossG) lossG[k] = h[2 * (c->d_iters + k)] + h[2 * (c->d_iters + k) + 1];
  return CG_OK;
}
// The step as the caller sees it: eager for the first calls of a configuration (every buffer reaches its final size, function
// attributes are set), then captured once into a CUDA graph and replayed.  A step is ~1200 small launches and the per-launch
// cost dominated it (profiles/r01_bench_tc_engine.json); everything a replay needs -- Adam's step count, the dropout RNG
// offset -- lives in device memory.  Never captured: profiling runs and forwards that consume masks queued by cg_D_set_masks.  A failed capture disables graphs for that configuration and the step runs eagerly.
static int train_step_run(cg_trainer* t, const cg_step_cfg* c, const float* real, const float* zD, const float* zG, float* lossD, float* lossG) {
  Ctx& X = ctx();
  // Multi-rank steps are captured too (the two NCCL all-reduces become graph nodes; NCCL >= 2.9 supports stream capture): every
  // rank captures at the same call number, so the collectives inside the capture line up.  Keeping N > 1 eager made data
  // parallel look ~25% worse than N = 1 for a reason unrelated to communication.
  bool eligible = X.graph_mode && !X.prof_on && !(t->D->mq && t->D->mq_next < t->D->mq_count);
  if (!eligible) { CG_TRY(train_step_core(t, c, real, zD, zG, nullptr, nullptr)); return read_losses(t, c, lossD, lossG); }
  cg_trainer::StepGraph* sg = nullptr;
  // the key holds everything the recorded launch sequence depends on besides buffer addresses (those: alloc_gen below)
  const int mode_key = X.lanes_on * 2 + X.side_on + 4 * t->G->training + 8 * t->D->training + 16 * X.sync_bn + 32 * X.world + 4096 * X.precision;
  for (auto& e : t->graphs) if (!memcmp(&e.cfg, c, sizeof(cg_step_cfg)) && e.engine == X.conv_engine && e.elim == X.dead_grad_elim && e.lanes == mode_key) { sg = &e; break; }
  if (!sg) { t->graphs.emplace_back(); sg = &t->graphs.back(); sg->cfg = *c; sg->engine = X.conv_engine; sg->elim = X.dead_grad_elim; sg->lanes = mode_key; }
  // A graph bakes in raw device pointers (DBufs, workspaces, lane / side scratch).  Any reallocation since the capture --
  // a larger configuration, a larger eager forward -- makes them dangle: drop the graph, run one eager step so that every
  // buffer reaches its size again, and capture anew.
  static const bool trace = getenv("CATGEN_GRAPH_TRACE") != nullptr;
  if (sg->exec && sg->gen != X.alloc_gen) {
    if (trace) fprintf(stderr, "[catgen graph] B=%d: buffers were reallocated since the capture (generation %llu -> %llu): dropping the graph\n", c->B, (unsigned long long)sg->gen, (unsigned long long)X.alloc_gen);
    cudaGraphExecDestroy(sg->exec); sg->exec = nullptr; sg->warm = 1;
  }
  const int B = c->B, hB = B / 2; const size_t img = (size_t)t->G->C * 1024, nz = t->G->nz;
  const size_t nr = (size_t)c->d_iters * hB * img, nzd = (size_t)c->d_iters * hB * nz, nzg = (size_t)c->g_iters * B * nz;
  if (sg->failed || (!sg->exec && sg->warm < 2)) {
    sg->warm++;
    CG_TRY(train_step_core(t, c, real, zD, zG, nullptr, nullptr)); return read_losses(t, c, lossD, lossG);
  }
  CG_TRY(t->gin.ensure(nr + nzd + nzg));
  float* g0 = t->gin.p;
  // The graph is captured from -- and therefore always replayed from -- a state in which both networks' packed operands are fresh:
  // whatever ran since the last update (an eager forward, cg_model_set_params, a replay that ended with an Adam step) is settled by an
  // eager repack here; the host-side flags are put back to what an eager step leaves behind after every replay (sg->end_dirty_*).
  // (G runs at batch B/2 and B inside a step, D only at B: asking D for fp32 fallback operands because of G's half batch cost one k_repack_model
  //  launch per replay at small batches -- tools/launch_diff.py)
  { const int need32_D = !conv_tc_all_shapes_taken(c->B), need32_G = need32_D || !conv_tc_all_shapes_taken(c->B / 2); CG_TRY(model_repack(t->G, need32_G)); CG_TRY(model_repack(t->D, need32_D)); }
  CG_CUDA(cudaMemcpyAsync(g0, real, sizeof(float) * nr, cudaMemcpyDeviceToDevice, X.stream));
  CG_CUDA(cudaMemcpyAsync(g0 + nr, zD, sizeof(float) * nzd, cudaMemcpyDeviceToDevice, X.stream));
  CG_CUDA(cudaMemcpyAsync(g0 + nr + nzd, zG, sizeof(float) * nzg, cudaMemcpyDeviceToDevice, X.stream));
  if (!sg->exec) {
    int64_t l0 = X.launches;
    cudaGraph_t graph = nullptr;
    const uint64_t gen0 = X.alloc_gen;
    if (trace) fprintf(stderr, "[catgen graph] B=%d d=%d g=%d: capturing (generation %llu)\n", c->B, c->d_iters, c->g_iters, (unsigned long long)gen0);
    if (cudaStreamBeginCapture(X.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); sg->failed = true; }
    else {
      int st = train_step_core(t, c, g0, g0 + nr, g0 + nr + nzd, nullptr, nullptr);
      cudaError_t e = cudaStreamEndCapture(X.stream, &graph);
      if (st != CG_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&sg->exec, graph, 0) != cudaSuccess) { cudaGetLastError(); sg->exec = nullptr; sg->failed = true; }
      if (graph) cudaGraphDestroy(graph);
      if (!sg->failed && X.alloc_gen != gen0) {             // a buffer grew during capture (illegal in capture anyway): never replay this
        cudaGraphExecDestroy(sg->exec); sg->exec = nullptr; sg->warm = 1;
        X.launches = l0;
        CG_TRY(train_step_core(t, c, real, zD, zG, nullptr, nullptr)); return read_losses(t, c, lossD, lossG);
      }
      sg->gen = X.alloc_gen;
      sg->end_dirty_G = t->G->dirty; sg->end_dirty_D = t->D->dirty; sg->end_dirty32_G = t->G->dirty32; sg->end_dirty32_D = t->D->dirty32;
      sg->launches = X.launches - l0; X.launches = l0;      // captured, not executed yet
    }
    if (sg->failed) { CG_TRY(train_step_core(t, c, real, zD, zG, nullptr, nullptr)); return read_losses(t, c, lossD, lossG); }
  }
  CG_CUDA(cudaGraphLaunch(sg->exec, X.stream));
  // the replay ran the Adam updates on the device: leave the flags as the eager step would (a network updated after its last forward
  // of the step has stale packed operands for whatever runs next)
  t->G->dirty = sg->end_dirty_G; t->D->dirty = sg->end_dirty_D;
  t->G->dirty32 = sg->end_dirty32_G; t->D->dirty32 = sg->end_dirty32_D;   // the captured repacks refresh the fp32 fallback operands exactly when the capture step's did
  X.launches += sg->launches;
  return read_losses(t, c, lossD, lossG);
}
int cg_train_step_dev(cg_trainer* t, const cg_step_cfg* cfg, const float* real_dev, const float* zD_dev, const float* zG_dev, float* lossD, float* lossG) {
  CG_REQUIRE_INIT(); CG_ARG(t && cfg && real_dev && zD_dev && zG_dev);
  return train_step_run(t, cfg, real_dev, zD_dev, zG_dev, lossD, lossG);
}
int cg_train_step(cg_trainer* t, const cg_step_cfg* cfg, const float* real, const float* zD, const float* zG, float* lossD, float* lossG, float* d_out) {
  CG_REQUIRE_INIT(); CG_ARG(t && cfg && real && zD && zG);
  int B = cfg->B, hB = B / 2; size_t img = (size_t)t->G->C * 1024, nz = t->G->nz;
  size_t nr = (size_t)cfg->d_iters * hB * img, nzd = (size_t)cfg->d_iters * hB * nz, nzg = (size_t)cfg->g_iters * B * nz;
  CG_TRY(t->stage.ensure(nr + nzd + nzg));
  float* s = t->stage.p;
  CG_CUDA(cudaMemcpyAsync(s, real, sizeof(float) * nr, cudaMemcpyHostToDevice, ctx().stream));
  CG_CUDA(cudaMemcpyAsync(s + nr, zD, sizeof(float) * nzd, cudaMemcpyHostToDevice, ctx().stream));
  CG_CUDA(cudaMemcpyAsync(s + nr + nzd, zG, sizeof(float) * nzg, cudaMemcpyHostToDevice, ctx().stream));
  float ld[16], lg[16]; CG_ARG(cfg->d_iters <= 16 && cfg->g_iters <= 16);
  // D's outputs of the last D update must be captured before the G phase overwrites dout
  CG_TRY(train_step_run(t, cfg, s, s + nr, s + nr + nzd, ld, lg));
  if (lossD) for (int k = 0; k < cfg->d_iters; ++k) lossD[k] = ld[k];
  if (lossG) for (int k = 0; k < cfg->g_iters; ++k) lossG[k] = lg[k];
  if (d_out) { CG_CUDA(cudaMemcpyAsync(d_out, t->dout.p, sizeof(float) * B, cudaMemcpyDeviceToHost, ctx().stream)); CG_CUDA(cudaStreamSynchronize(ctx().stream)); }
  return CG_OK;
}
void* cg_dev_alloc(int64_t bytes) { if (!ctx().inited) return nullptr; void* p = nullptr; if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) return nullptr; return p; }
int cg_dev_free(void* p) { if (p) { cudaStreamSynchronize(ctx().stream); cudaFree(p); } return CG_OK; }
void* cg_host_alloc(int64_t bytes) { if (!ctx().inited) return nullptr; void* p = nullptr; if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
int cg_host_free(void* p) { if (p) { cudaStreamSynchronize(ctx().stream); cudaFreeHost(p); } return CG_OK; }
int cg_dev_upload(void* dst, const void* src, int64_t bytes) { CG_REQUIRE_INIT(); CG_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, ctx().stream)); CG_CUDA(cudaStreamSynchronize(ctx().stream)); return CG_OK; }
int cg_dev_download(void* dst, const void* src, int64_t bytes) { CG_REQUIRE_INIT(); CG_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx().stream)); CG_CUDA(cudaStreamSynchronize(ctx().stream)); return CG_OK; }
int cg_uniform_dev(float* dst, int64_t n, float lo, float hi, uint64_t seed, uint64_t offset) { CG_REQUIRE_INIT(); return uniform(dst, n, lo, hi, seed, offset); }

// ------------------------------------------------------------------ data parallel
int cg_dist_unique_id(char id_out[128]) {
#ifdef CG_WITH_NCCL
  CG_NCCL_API();
  ncclUniqueId id; ncclResult_t r = N.GetUniqueId(&id);
  if (r != ncclSuccess) return set_err(CG_ERR_NCCL, "ncclGetUniqueId: %s", N.GetErrorString(r));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(id_out, &id, 128); return CG_OK;
#else
  (void)id_out; return set_err(CG_ERR_NCCL, "library built without NCCL");
#endif
}
int cg_dist_init(int rank, int world, const char id[128]) {
  CG_REQUIRE_INIT(); CG_ARG(world >= 1 && rank >= 0 && rank < world);
  if (world == 1) { ctx().rank = 0; ctx().world = 1; return CG_OK; }
#ifdef CG_WITH_NCCL
  CG_NCCL_API();
  ncclUniqueId uid; memcpy(&uid, id, 128);
  ncclComm_t comm; ncclResult_t r = N.CommInitRank(&comm, world, uid, rank);
  if (r != ncclSuccess) return set_err(CG_ERR_NCCL, "ncclCommInitRank: %s", N.GetErrorString(r));
  ctx().nccl = comm; ctx().rank = rank; ctx().world = world; return CG_OK;
#else
  (void)id; return set_err(CG_ERR_NCCL, "library built without NCCL");
#endif
}
int cg_dist_world(void) { return ctx().world; }
int cg_dist_set_sync_bn(int on) { ctx().sync_bn = on ? 1 : 0; return CG_OK; }
int cg_dist_get_sync_bn(void) { return ctx().sync_bn; }

// ------------------------------------------------------------------ op level (Torch NCHW host tensors)
static int conv_op_args(int N, int Ci, int H, int W, int Co, int k) {
  if (N <= 0 || Ci <= 0 || H <= 0 || W <= 0 || Co <= 0) return set_err(CG_ERR_ARG, "non-positive conv dimension");
  if (k <= 0 || k % 2 == 0) return set_err(CG_ERR_ARG, "kW/kH has to be odd");   // layers/SpatialConvolutionUpsample.lua:5-7
  return CG_OK;
}
int cg_conv2d_fprop(const float* x, const float* W, const float* b, float* y, int N, int Ci, int H, int Wd, int Co, int k) {
  CG_REQUIRE_INIT(); CG_ARG(x && W && y); CG_TRY(conv_op_args(N, Ci, H, Wd, Co, k));
  Tmp T; size_t nx = (size_t)N * Ci * H * Wd, ny = (size_t)N * Co * H * Wd, nW = (size_t)Co * Ci * k * k;
  float *dx = T.up(x, nx), *dW = T.up(W, nW), *db = b ? T.up(b, Co) : nullptr, *xh = T.dev(nx), *Wp = T.dev(nW), *yh = T.dev(ny), *dy = T.dev(ny);
  if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  ConvSpec s; s.Ci = Ci; s.Co = Co; s.k = k;
  CG_TRY(nchw_to_nhwc(dx, xh, N, Ci, H * Wd)); CG_TRY(pack_fprop(dW, Wp, s));
  CG_TRY(conv_fwd(xh, Wp, db, yh, N, H, Wd, Ci, Co, k));
  CG_TRY(nhwc_to_nchw(yh, dy, N, Co, H * Wd)); CG_TRY(T.down(y, dy, ny));
  return T.finish();
}
int cg_conv2d_dgrad(const float* gy, const float* W, float* gx, int N, int Ci, int H, int Wd, int Co, int k) {
  CG_REQUIRE_INIT(); CG_ARG(gy && W && gx); CG_TRY(conv_op_args(N, Ci, H, Wd, Co, k));
  Tmp T; size_t nx = (size_t)N * Ci * H * Wd, ny = (size_t)N * Co * H * Wd, nW = (size_t)Co * Ci * k * k;
  float *dg = T.up(gy, ny), *dW = T.up(W, nW), *gh = T.dev(ny), *Wdp = T.dev(nW), *xh = T.dev(nx), *dx = T.dev(nx);
  if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  ConvSpec s; s.Ci = Ci; s.Co = Co; s.k = k;
  CG_TRY(nchw_to_nhwc(dg, gh, N, Co, H * Wd)); CG_TRY(pack_dgrad(dW, Wdp, s));
  CG_TRY(conv_dgrad(gh, Wdp, xh, N, H, Wd, Ci, Co, k));
  CG_TRY(nhwc_to_nchw(xh, dx, N, Ci, H * Wd)); CG_TRY(T.down(gx, dx, nx));
  return T.finish();
}
int cg_conv2d_wgrad(const float* x, const float* gy, float* gW, float* gb, int N, int Ci, int H, int Wd, int Co, int k) {
  CG_REQUIRE_INIT(); CG_ARG(x && gy && gW); CG_TRY(conv_op_args(N, Ci, H, Wd, Co, k));
  Tmp T; size_t nx = (size_t)N * Ci * H * Wd, ny = (size_t)N * Co * H * Wd, nW = (size_t)Co * Ci * k * k;
  float *dx = T.up(x, nx), *dg = T.up(gy, ny), *xh = T.dev(nx), *gh = T.dev(ny), *gwp = T.dev(nW), *dgW = T.up(gW, nW), *dgb = gb ? T.up(gb, Co) : nullptr;
  if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  ConvSpec s; s.Ci = Ci; s.Co = Co; s.k = k;
  CG_TRY(nchw_to_nhwc(dx, xh, N, Ci, H * Wd)); CG_TRY(nchw_to_nhwc(dg, gh, N, Co, H * Wd));
  CG_TRY(conv_wgrad(xh, gh, gwp, N, H, Wd, Ci, Co, k)); CG_TRY(unpack_wgrad_acc(gwp, dgW, s));   // accumulates, like accGradParameters
  if (dgb) CG_TRY(colsum_acc(gh, dgb, (long)N * H * Wd, Co));
  CG_TRY(T.down(gW, dgW, nW)); if (gb) CG_TRY(T.down(gb, dgb, Co));
  return T.finish();
}
int cg_conv_upsample_fwd(const float* x, const float* W, const float* b, float* y, int N, int Ci, int H, int Wd, int nOut, int k, int f) {
  CG_ARG(f >= 1 && nOut > 0);
  // parent conv to nOut*f*f planes; the :view to [N,nOut,H*f,W*f] does not move memory (SpatialConvolutionUpsample.lua:21)
  return cg_conv2d_fprop(x, W, b, y, N, Ci, H, Wd, nOut * f * f, k);
}
int cg_conv_upsample_bwd(const float* x, const float* gy, const float* W, float* gx, float* gW, float* gb, int N, int Ci, int H, int Wd, int nOut, int k, int f) {
  CG_ARG(f >= 1 && nOut > 0);
  // gradOutput is viewed back to [N,nOut*f*f,H,W] (SpatialConvolutionUpsample.lua:30-47): same memory
  if (gx) CG_TRY(cg_conv2d_dgrad(gy, W, gx, N, Ci, H, Wd, nOut * f * f, k));
  if (gW) CG_TRY(cg_conv2d_wgrad(x, gy, gW, gb, N, Ci, H, Wd, nOut * f * f, k));
  return CG_OK;
}
int cg_linear_fwd(const float* x, const float* W, const float* b, float* y, int N, int in, int out) {
  CG_REQUIRE_INIT(); CG_ARG(x && W && y && N > 0 && in > 0 && out > 0);
  Tmp T; float *dx = T.up(x, (size_t)N * in), *dW = T.up(W, (size_t)in * out), *db = b ? T.up(b, out) : nullptr, *Wp = T.dev((size_t)in * out), *dy = T.dev((size_t)N * out);
  if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  ConvSpec s; s.Ci = in; s.Co = out; s.k = 1;
  CG_TRY(pack_fprop(dW, Wp, s)); CG_TRY(conv_fwd(dx, Wp, db, dy, N, 1, 1, in, out, 1)); CG_TRY(T.down(y, dy, (size_t)N * out));
  return T.finish();
}
int cg_linear_bwd(const float* x, const float* gy, const float* W, float* gx, float* gW, float* gb, int N, int in, int out) {
  CG_REQUIRE_INIT(); CG_ARG(x && gy && W && N > 0 && in > 0 && out > 0);
  Tmp T; size_t nW = (size_t)in * out;
  float *dx = T.up(x, (size_t)N * in), *dg = T.up(gy, (size_t)N * out), *dW = T.up(W, nW), *Wdp = T.dev(nW), *dgx = T.dev((size_t)N * in), *gwp = T.dev(nW);
  float *dgW = gW ? T.up(gW, nW) : nullptr, *dgb = gb ? T.up(gb, out) : nullptr;
  if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  ConvSpec s; s.Ci = in; s.Co = out; s.k = 1;
  if (gx) { CG_TRY(pack_dgrad(dW, Wdp, s)); CG_TRY(conv_dgrad(dg, Wdp, dgx, N, 1, 1, in, out, 1)); CG_TRY(T.down(gx, dgx, (size_t)N * in)); }
  if (gW) { CG_TRY(conv_wgrad(dx, dg, gwp, N, 1, 1, in, out, 1)); CG_TRY(unpack_wgrad_acc(gwp, dgW, s)); CG_TRY(T.down(gW, dgW, nW)); }
  if (gb) { CG_TRY(colsum_acc(dg, dgb, N, out)); CG_TRY(T.down(gb, dgb, out)); }
  return T.finish();
}
int cg_bn2d_fwd(const float* x, const float* gamma, const float* beta, float* y, float* save_mean, float* save_invstd, float* run_mean, float* run_var, int N, int C, int HW) {
  CG_REQUIRE_INIT(); CG_ARG(x && gamma && beta && y && N > 0 && C > 0 && HW > 0);
  Tmp T; size_t n = (size_t)N * C * HW;
  float *dx = T.up(x, n), *xh = T.dev(n), *yh = T.dev(n), *dy = T.dev(n), *dg = T.up(gamma, C), *db = T.up(beta, C), *dm = T.dev(C), *di = T.dev(C);
  float *drm = run_mean ? T.up(run_mean, C) : nullptr, *drv = run_var ? T.up(run_var, C) : nullptr;
  if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(nchw_to_nhwc(dx, xh, N, C, HW));
  CG_TRY(bn_fwd_train(xh, dg, db, yh, dm, di, drm, drv, (long)N * HW, C, 1e-5f, 0.1f));
  CG_TRY(nhwc_to_nchw(yh, dy, N, C, HW)); CG_TRY(T.down(y, dy, n));
  if (save_mean) CG_TRY(T.down(save_mean, dm, C)); if (save_invstd) CG_TRY(T.down(save_invstd, di, C));
  if (run_mean) CG_TRY(T.down(run_mean, drm, C)); if (run_var) CG_TRY(T.down(run_var, drv, C));
  return T.finish();
}
int cg_bn2d_bwd(const float* x, const float* gy, const float* gamma, const float* save_mean, const float* save_invstd, float* gx, float* ggamma, float* gbeta, int N, int C, int HW) {
  CG_REQUIRE_INIT(); CG_ARG(x && gy && gamma && save_mean && save_invstd && N > 0 && C > 0 && HW > 0);
  Tmp T; size_t n = (size_t)N * C * HW;
  float *dx = T.up(x, n), *dgy = T.up(gy, n), *xh = T.dev(n), *gh = T.dev(n), *gxh = T.dev(n), *dgx = T.dev(n);
  float *dg = T.up(gamma, C), *dm = T.up(save_mean, C), *di = T.up(save_invstd, C);
  float *dgg = ggamma ? T.up(ggamma, C) : nullptr, *dgb = gbeta ? T.up(gbeta, C) : nullptr;
  if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(nchw_to_nhwc(dx, xh, N, C, HW)); CG_TRY(nchw_to_nhwc(dgy, gh, N, C, HW));
  CG_TRY(bn_bwd(xh, gh, dg, dm, di, gx ? gxh : nullptr, dgg, dgb, (long)N * HW, C));
  if (gx) { CG_TRY(nhwc_to_nchw(gxh, dgx, N, C, HW)); CG_TRY(T.down(gx, dgx, n)); }
  if (ggamma) CG_TRY(T.down(ggamma, dgg, C)); if (gbeta) CG_TRY(T.down(gbeta, dgb, C));
  return T.finish();
}
int cg_prelu_fwd(const float* x, float w, float* y, int64_t n) {
  CG_REQUIRE_INIT(); CG_ARG(x && y && n > 0);
  Tmp T; float *dx = T.up(x, n), *dw = T.up(&w, 1), *dy = T.dev(n); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(prelu_fwd(dx, dw, dy, n)); CG_TRY(T.down(y, dy, n)); return T.finish();
}
int cg_prelu_bwd(const float* x, const float* gy, float w, float* gx, float* gw, int64_t n) {
  CG_REQUIRE_INIT(); CG_ARG(x && gy && n > 0);
  Tmp T; float *dx = T.up(x, n), *dg = T.up(gy, n), *dw = T.up(&w, 1), *dgx = T.dev(n), *dgw = gw ? T.up(gw, 1) : nullptr; if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(prelu_bwd(dx, dg, dw, gx ? dgx : nullptr, dgw, n));
  if (gx) CG_TRY(T.down(gx, dgx, n)); if (gw) CG_TRY(T.down(gw, dgw, 1)); return T.finish();
}
int cg_leakyrelu_fwd(const float* x, float s, float* y, int64_t n) {
  CG_REQUIRE_INIT(); CG_ARG(x && y && n > 0);
  Tmp T; float *dx = T.up(x, n), *dy = T.dev(n); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(lrelu_fwd(dx, s, dy, n)); CG_TRY(T.down(y, dy, n)); return T.finish();
}
int cg_leakyrelu_bwd(const float* x, const float* gy, float s, float* gx, int64_t n) {
  CG_REQUIRE_INIT(); CG_ARG(x && gy && gx && n > 0);
  Tmp T; float *dx = T.up(x, n), *dg = T.up(gy, n), *dgx = T.dev(n); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(lrelu_bwd(dx, dg, s, dgx, n)); CG_TRY(T.down(gx, dgx, n)); return T.finish();
}
int cg_sigmoid_fwd(const float* x, float* y, int64_t n) {
  CG_REQUIRE_INIT(); CG_ARG(x && y && n > 0);
  Tmp T; float *dx = T.up(x, n), *dy = T.dev(n); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(sigmoid_fwd(dx, dy, n)); CG_TRY(T.down(y, dy, n)); return T.finish();
}
int cg_sigmoid_bwd(const float* y, const float* gy, float* gx, int64_t n) {
  CG_REQUIRE_INIT(); CG_ARG(y && gy && gx && n > 0);
  Tmp T; float *dy = T.up(y, n), *dg = T.up(gy, n), *dgx = T.dev(n); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(sigmoid_bwd(dy, dg, dgx, n)); CG_TRY(T.down(gx, dgx, n)); return T.finish();
}
// NCHW planes are independent for these ops: run the NHWC kernels with C = 1 and N = NC
int cg_upsample2x_fwd(const float* x, float* y, int NC, int H, int Wd) {
  CG_REQUIRE_INIT(); CG_ARG(x && y && NC > 0 && H > 0 && Wd > 0);
  Tmp T; size_t n = (size_t)NC * H * Wd; float *dx = T.up(x, n), *dy = T.dev(4 * n); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(upsample2x_fwd(dx, dy, NC, H, Wd, 1)); CG_TRY(T.down(y, dy, 4 * n)); return T.finish();
}
int cg_upsample2x_bwd(const float* gy, float* gx, int NC, int H, int Wd) {
  CG_REQUIRE_INIT(); CG_ARG(gy && gx && NC > 0 && H > 0 && Wd > 0);
  Tmp T; size_t n = (size_t)NC * H * Wd; float *dg = T.up(gy, 4 * n), *dx = T.dev(n); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(upsample2x_bwd(dg, dx, NC, H, Wd, 1)); CG_TRY(T.down(gx, dx, n)); return T.finish();
}
int cg_avgpool2_fwd(const float* x, float* y, int NC, int H, int Wd) {
  CG_REQUIRE_INIT(); CG_ARG(x && y && NC > 0 && H > 1 && Wd > 1);
  Tmp T; size_t n = (size_t)NC * H * Wd, no = (size_t)NC * (H / 2) * (Wd / 2); float *dx = T.up(x, n), *dy = T.dev(no); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(avgpool2_fwd(dx, dy, NC, H, Wd, 1)); CG_TRY(T.down(y, dy, no)); return T.finish();
}
int cg_avgpool2_bwd(const float* gy, float* gx, int NC, int H, int Wd) {
  CG_REQUIRE_INIT(); CG_ARG(gy && gx && NC > 0 && H > 1 && Wd > 1);
  Tmp T; size_t n = (size_t)NC * H * Wd, no = (size_t)NC * (H / 2) * (Wd / 2); float *dg = T.up(gy, no), *dx = T.dev(n); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(avgpool2_bwd(dg, dx, NC, H, Wd, 1)); CG_TRY(T.down(gx, dx, n)); return T.finish();
}
__global__ void k_u8_to_i32(const uint8_t* a, int32_t* b, long n, int dir) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) { if (dir == 0) b[i] = a[i]; else ((uint8_t*)a)[i] = (uint8_t)b[i]; }
}
int cg_maxpool2_fwd(const float* x, float* y, int32_t* idx, int NC, int H, int Wd) {
  CG_REQUIRE_INIT(); CG_ARG(x && y && NC > 0 && H > 1 && Wd > 1);
  Tmp T; size_t n = (size_t)NC * H * Wd, no = (size_t)NC * (H / 2) * (Wd / 2);
  float *dx = T.up(x, n), *dy = T.dev(no); uint8_t* di = (uint8_t*)T.dev(no / 4 + 4); int32_t* di32 = (int32_t*)T.dev(no); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(maxpool2_fwd(dx, dy, di, NC, H, Wd, 1)); CG_TRY(T.down(y, dy, no));
  if (idx) { CG_LAUNCH(k_u8_to_i32, grid1d(no, 256), 256, 0, di, di32, (long)no, 0); CG_TRY(T.down((float*)idx, (float*)di32, no)); }
  return T.finish();
}
int cg_maxpool2_bwd(const float* gy, const int32_t* idx, float* gx, int NC, int H, int Wd) {
  CG_REQUIRE_INIT(); CG_ARG(gy && idx && gx && NC > 0 && H > 1 && Wd > 1);
  Tmp T; size_t n = (size_t)NC * H * Wd, no = (size_t)NC * (H / 2) * (Wd / 2);
  float *dg = T.up(gy, no), *dx = T.dev(n); int32_t* di32 = (int32_t*)T.up((const float*)idx, no); uint8_t* di = (uint8_t*)T.dev(no / 4 + 4); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_LAUNCH(k_u8_to_i32, grid1d(no, 256), 256, 0, di, di32, (long)no, 1);
  CG_TRY(maxpool2_bwd(dg, di, dx, NC, H, Wd, 1)); CG_TRY(T.down(gx, dx, n)); return T.finish();
}
static int nth_of(int rot, int scl, int trn) { return (rot ? 1 : 0) + (scl ? 1 : 0) + (trn ? 2 : 0); }
int cg_affine_matrix_fwd(const float* theta, float* A, int B, int rot, int scl, int trn) {
  CG_REQUIRE_INIT(); CG_ARG(theta && A && B > 0 && nth_of(rot, scl, trn) > 0);
  Tmp T; float *dt = T.up(theta, (size_t)B * nth_of(rot, scl, trn)), *dA = T.dev((size_t)B * 6); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(affine_matrix_fwd(dt, dA, B, rot, scl, trn)); CG_TRY(T.down(A, dA, (size_t)B * 6)); return T.finish();
}
int cg_affine_matrix_bwd(const float* theta, const float* gA, float* gtheta, int B, int rot, int scl, int trn) {
  CG_REQUIRE_INIT(); CG_ARG(theta && gA && gtheta && B > 0 && nth_of(rot, scl, trn) > 0);
  int nth = nth_of(rot, scl, trn);
  Tmp T; float *dt = T.up(theta, (size_t)B * nth), *dg = T.up(gA, (size_t)B * 6), *dgt = T.dev((size_t)B * nth); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(affine_matrix_bwd(dt, dg, dgt, B, rot, scl, trn)); CG_TRY(T.down(gtheta, dgt, (size_t)B * nth)); return T.finish();
}
int cg_affine_grid_fwd(const float* A, float* grid, int B, int H, int Wd) {
  CG_REQUIRE_INIT(); CG_ARG(A && grid && B > 0 && H > 1 && Wd > 1);
  Tmp T; size_t n = (size_t)B * H * Wd * 2; float *dA = T.up(A, (size_t)B * 6), *dg = T.dev(n); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(affine_grid_fwd(dA, dg, B, H, Wd)); CG_TRY(T.down(grid, dg, n)); return T.finish();
}
int cg_affine_grid_bwd(const float* ggrid, float* gA, int B, int H, int Wd) {
  CG_REQUIRE_INIT(); CG_ARG(ggrid && gA && B > 0 && H > 1 && Wd > 1);
  Tmp T; size_t n = (size_t)B * H * Wd * 2; float *dg = T.up(ggrid, n), *dA = T.dev((size_t)B * 6); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(affine_grid_bwd(dg, dA, B, H, Wd)); CG_TRY(T.down(gA, dA, (size_t)B * 6)); return T.finish();
}
int cg_bilinear_fwd(const float* img, const float* grid, float* out, int B, int H, int Wd, int C) {
  CG_REQUIRE_INIT(); CG_ARG(img && grid && out && B > 0 && H > 0 && Wd > 0 && C > 0);
  Tmp T; size_t n = (size_t)B * H * Wd * C; float *di = T.up(img, n), *dg = T.up(grid, (size_t)B * H * Wd * 2), *dout = T.dev(n); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(bilinear_fwd(di, dg, dout, B, H, Wd, C)); CG_TRY(T.down(out, dout, n)); return T.finish();
}
int cg_bilinear_bwd(const float* img, const float* grid, const float* gout, float* gimg, float* ggrid, int B, int H, int Wd, int C) {
  CG_REQUIRE_INIT(); CG_ARG(img && grid && gout && gimg && ggrid && B > 0 && H > 0 && Wd > 0 && C > 0);
  Tmp T; size_t n = (size_t)B * H * Wd * C, ng = (size_t)B * H * Wd * 2;
  float *di = T.up(img, n), *dg = T.up(grid, ng), *dgo = T.up(gout, n), *dgi = T.dev(n), *dgg = T.dev(ng); if (!T.ok) return set_err(CG_ERR_CUDA, "staging failed");
  CG_TRY(bilinear_bwd(di, dg, dgo, dgi, dgg, B, H, Wd, C)); CG_TRY(T.down(gimg, dgi, n)); CG_TRY(T.down(ggrid, dgg, ng)); return T.finish();
}

}  // extern "C"
