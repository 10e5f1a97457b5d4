// common.cuh -- shared infrastructure of libcatgen (sm_100a only; no CPU fallback anywhere).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/catgen.h"

namespace cg {

struct Ctx {
  int device = -1;
  bool inited = false;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int conv_engine = 1;
  int dead_grad_elim = 1;       // cg_train_step: skip D's parameter gradients inside fevalG (zeroed unread by the next fevalD)
  int graph_mode = 1;           // replay the training step as a CUDA graph once warm (cg_set_graph_mode)
  int64_t launches = 0;
  uint64_t alloc_gen = 0;       // bumped whenever a device buffer a captured step graph may point into is (re)allocated
  char err[1024] = {0};
  // scratch (grown on demand, stream-ordered reuse)
  void* ws = nullptr; size_t ws_bytes = 0;
  void* ws2 = nullptr; size_t ws2_bytes = 0;
  void* ws3 = nullptr; size_t ws3_bytes = 0;
  void* ws4 = nullptr; size_t ws4_bytes = 0;
  void* pinned = nullptr; size_t pinned_bytes = 0;
  // lanes: independent branches of one model (D32_st3's four transformer branches) issue on their own stream with their own
  // scratch, forked from / joined to `stream` by events -- which CUDA-graph capture records as parallel paths.  lane_enter()
  // swaps `stream`, ws, ws3, ws4 with the lane's, so every launcher keeps using ctx().stream / workspace*() unchanged.
  struct Lane { cudaStream_t stream = nullptr; cudaEvent_t done = nullptr; void *ws = nullptr, *ws3 = nullptr, *ws4 = nullptr; size_t ws_bytes = 0, ws3_bytes = 0, ws4_bytes = 0; };
  static constexpr int kLanes = 5;   // 0-3: D's branches; 4: the generator forward of fevalG, issued ahead (capi.cu: train_step_core)
  Lane lanes[kLanes], saved;
  int lane = -1;                // -1: the main stream
  int lanes_on = 1;             // CATGEN_LANES=0 serialises the branches on the main stream
  cudaEvent_t fork_ev = nullptr;
  // side streams: one companion per stream above (index lane + 1).  A layer's weight-gradient chain is issued there while its
  // input-gradient conv and the element-wise ops behind it continue on the owning stream (conv_tc.cu: conv_bwd_tc).
  struct Side { cudaStream_t stream = nullptr; cudaEvent_t fork = nullptr, done = nullptr; bool pending = false; void *ws = nullptr, *ws3 = nullptr; size_t ws_bytes = 0, ws3_bytes = 0; };
  Side side[kLanes + 1], side_saved;
  bool in_side = false;
  int side_on = 1;              // CATGEN_SIDE=0 keeps the weight gradients on the owning stream
  // data parallel
  int rank = 0, world = 1;
  void* nccl = nullptr;
  int sync_bn = 0;              // data parallel: batch-norm statistics over the GLOBAL batch (all-reduced sums) instead of per rank
  // instrumentation (bench.py): per-launch CUDA events on `stream`, algorithmic flops/bytes per launch
  bool prof_on = false;
  double next_flops = 0, next_bytes = 0;
  double* next_stats_part = nullptr;   // one-shot: the next k_conv_ps launch may write per-CTA (sum, sum of squares) rows of its output there: [rows <= next_stats_cap][C][2]
  int next_stats_cap = 0, stats_rows = 0;   // stats_rows: set by that launch to the number of rows it wrote (0: request not honoured)
  float* next_pool2_out = nullptr;     // one-shot: the next k_conv_ps launch may write the 2 x 2 block sums of its output there instead of the output (conv_tc.cu)
  int pool2_done = 0;                  // set by that launch when it did
  int trace_launches = 0;              // CATGEN_LAUNCH_TRACE=1: every counted launch is named on stderr (diagnosis of launch-count differences)
  int precision = 0;                   // cg_set_precision: 1 = every FORWARD convolution of G and D runs with error-compensated operands (hi + lo fp16 pairs)
  int split_fwd = 0;                   // > 0 inside a forward executor while precision == 1: conv_ps_run packs [hi, lo, hi] x [hi, hi, lo]
  int fp32_operands_stale = 0;         // > 0 inside a model executor that skipped refreshing the fp32 fallback operands: a fallback must fail loudly
  unsigned int* next_amax = nullptr;   // one-shot: the next conv / split-K reduction launch also records max|output| there (atomicMax on float bits; the caller zeroes it)
  struct ProfRec { const char* name; cudaEvent_t a, b; double flops, bytes; int lane; };
  std::vector<ProfRec> prof;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
};
void prof_begin(const char* name);
void prof_end();
Ctx& ctx();

int set_err(int code, const char* fmt, ...);
void* workspace(size_t bytes);    // device scratch #1 (split-K partials, reductions)
void* workspace2(size_t bytes);   // device scratch #2: boundary staging of the host-pointer entry points ONLY
void* workspace3(size_t bytes);   // device scratch #3: tensor-core operand packing (never shared with staging: a grow
                                  // reallocates, and ws2 pointers are live across the whole forward/backward call)
void* workspace4(size_t bytes);   // device scratch #4: the packed gradient operand shared by dgrad and wgrad of one layer (outlives both runs' ws3 use)
void* pinned(size_t bytes);       // pinned host staging
int lanes_fork(int first = 0, int n = 4);   // lanes [first, first+n) may start after everything issued on the main stream so far
int lane_enter(int b);            // route launches + scratch to lane b (no-op when lanes are off)
int lane_exit();
int lanes_join(int first = 0, int n = 4);   // the main stream waits for those lanes (and their side streams)
int side_begin();                 // 1: launches now go to this stream's side stream (ordered after everything issued so far); 0: side streams off
int side_end();                   // back to the owning stream; the side work is pending
int side_wait();                  // the owning stream waits for its pending side work (before reusing what that work reads)
int side_wait_all();              // main stream: every side stream's pending work

#define CG_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return cg::set_err(CG_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)
#define CG_TRY(expr) do { int _s = (expr); if (_s != CG_OK) return _s; } while (0)
#define CG_REQUIRE_INIT() do { if (!cg::ctx().inited) return cg::set_err(CG_ERR_STATE, "cg_init has not been called"); } while (0)
#define CG_ARG(cond) do { if (!(cond)) return cg::set_err(CG_ERR_ARG, "%s:%d argument check failed: %s", __FILE__, __LINE__, #cond); } while (0)

// every kernel launch of this library goes through this macro so cg_launch_count() is exact
#define CG_LAUNCH(kernel, grid, block, smem, ...)                                             \
  do {                                                                                        \
    if (cg::ctx().prof_on) cg::prof_begin(#kernel);                                           \
    kernel<<<(grid), (block), (smem), cg::ctx().stream>>>(__VA_ARGS__);                       \
    if (cg::ctx().prof_on) cg::prof_end();                                                    \
    cg::ctx().next_flops = 0; cg::ctx().next_bytes = 0;                                       \
    cg::ctx().launches++;                                                                     \
    if (cg::ctx().trace_launches) fprintf(stderr, "[launch] %s\n", #kernel);                   \
    cudaError_t _e = cudaPeekAtLastError();                                                   \
    if (_e != cudaSuccess)                                                                    \
      return cg::set_err(CG_ERR_CUDA, "%s:%d launch %s -> %s", __FILE__, __LINE__, #kernel, cudaGetErrorString(_e)); \
  } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }
static inline int grid1d(long n, int block, int per_thread = 1) {
  long g = (n + (long)block * per_thread - 1) / ((long)block * per_thread);
  if (g < 1) g = 1;
  long cap = (long)ctx().sm_count * 16;   // grid-stride loops: a few waves of 148 SMs
  return (int)(g < cap ? g : cap);
}

// growable device buffer
struct DBuf {
  float* p = nullptr; size_t n = 0;
  int ensure(size_t nfloats);
  void release();
};

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox4x32-10 counter-based RNG (own implementation of the published algorithm, Salmon et al. 2011)
__host__ __device__ __forceinline__ void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__host__ __device__ __forceinline__ float u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }  // [0,1)

}  // namespace cg
