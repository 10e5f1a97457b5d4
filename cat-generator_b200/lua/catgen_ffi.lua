--[[ catgen_ffi.lua -- LuaJIT FFI binding of libcatgen.so (include/catgen.h).

STATUS: WRITTEN, NOT EXECUTED.  No Lua / LuaJIT / Torch7 exists in the build container or on the GPU box
(SURVEY.md section 0.1 #7), so nothing in this directory has ever run.  Every C function bound here is exercised
through the identical ctypes binding (catgen/lib.py) by tests/test_gpu_parity.py; the declarations below are a
hand copy of include/catgen.h and must be kept in sync with it.

Replaces: the THNN/THCUNN/cuDNN C functions that nn / cunn / cudnn reach through FFI from the modules that
/root/reference/models.lua instantiates (models.lua:138-228, :640-711, :814-906). ]]
local ffi = require("ffi")

ffi.cdef[[
typedef struct cg_model cg_model;
typedef struct cg_trainer cg_trainer;
typedef struct { int B, d_iters, g_iters; float D_L1, D_L2, G_L1, G_L2, D_clamp, G_clamp, lr, beta1, beta2, eps; } cg_step_cfg;
int cg_init(int device); void cg_shutdown(void); const char* cg_last_error(void); const char* cg_version(void); int cg_sync(void);
int cg_set_conv_engine(int engine); int cg_get_conv_engine(void); int cg_set_precision(int mode); int cg_get_precision(void);
int cg_set_dead_grad_elim(int on); int cg_set_concurrency(int on);
int cg_model_create(cg_model** out, int kind, int C, int nz, uint64_t seed); int cg_model_free(cg_model* m);
int cg_model_nparams(const cg_model* m, int64_t* n);
int cg_model_get_params(cg_model* m, float* host); int cg_model_set_params(cg_model* m, const float* host);
int cg_model_get_grads(cg_model* m, float* host); int cg_model_zero_grads(cg_model* m);
int cg_model_set_mode(cg_model* m, int training);
int cg_model_bn_running_len(const cg_model* m, int64_t* n); int cg_model_get_bn_running(cg_model* m, float* host); int cg_model_set_bn_running(cg_model* m, const float* host);
int cg_G_forward(cg_model* g, const float* z, int B, float* out); int cg_G_backward(cg_model* g, const float* gout, float* gz);
int cg_V_forward(cg_model* v, const float* x, int B, float* out);
int cg_D_forward(cg_model* d, const float* x, int B, float* out_sig, float* out_pre); int cg_D_backward(cg_model* d, const float* gout, float* gx);
int cg_bce(const float* p, const float* t, int n, float* loss, float* g);
int cg_penalty_clamp(cg_model* m, float l1, float l2sign, float l2, float clampv, float* loss_add);
int cg_trainer_create(cg_trainer** out, cg_model* G, cg_model* D); int cg_trainer_free(cg_trainer* t);
int cg_adam_step(cg_trainer* t, int which, const cg_step_cfg* cfg);
int cg_train_step(cg_trainer* t, const cg_step_cfg* cfg, const float* real, const float* zD, const float* zG, float* lossD, float* lossG, float* d_out);
int cg_dist_unique_id(char id_out[128]); int cg_dist_init(int rank, int world, const char id[128]);
int cg_dist_set_sync_bn(int on); int cg_dist_get_sync_bn(void); int cg_dist_world(void);
int cg_leakyrelu_fwd(const float* x, float slope, float* y, int64_t n); int cg_leakyrelu_bwd(const float* x, const float* gy, float slope, float* gx, int64_t n);
int cg_conv_upsample_fwd(const float* x, const float* W, const float* b, float* y, int N, int Ci, int H, int Wd, int nOut, int k, int f);
int cg_conv_upsample_bwd(const float* x, const float* gy, const float* W, float* gx, float* gW, float* gb, int N, int Ci, int H, int Wd, int nOut, int k, int f);
]]

local M = {}
M.G32UP, M.G32UPC, M.D32_ST3, M.V32 = 0, 1, 2, 3
M.lib = ffi.load(os.getenv("CATGEN_LIB") or "catgen")          -- libcatgen.so on the loader path
-- the reference's error convention is assert/error (layers/SpatialConvolutionUpsample.lua:5-7): non-zero status raises
function M.check(status) if status ~= 0 then error(ffi.string(M.lib.cg_last_error()), 2) end end
-- extra knobs come from the environment because lapp rejects unknown flags (train.lua:15-49)
local inited = false
function M.init(gpu)
   if not inited then M.check(M.lib.cg_init(gpu or tonumber(os.getenv("CATGEN_DEVICE") or "0"))); inited = true end
end
-- float* of a contiguous torch.FloatTensor (Torch7 exposes :data() through FFI)
function M.ptr(t) assert(t:isContiguous(), "catgen needs contiguous FloatTensors"); return t:data() end
return M
