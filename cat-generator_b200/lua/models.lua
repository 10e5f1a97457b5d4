--[[ models.lua -- drop-in for /root/reference/models.lua's create_G / create_D (models.lua:234-240, :268-277).
STATUS: WRITTEN, NOT EXECUTED (see catgen_ffi.lua).  The returned objects implement the part of the nn.Module
protocol train.lua, adversarial.lua and utils/nn_utils.lua use (train.lua:147-185,252-261; adversarial.lua:84-89,187-197;
utils/nn_utils.lua:428-462,620-680): forward, backward, getParameters, training, evaluate, zeroGradParameters, clone, cuda, float,
listModules, clearState, __tostring, `modules[1].gradInput` on D, and torch.save / torch.load through write / read.

Parameter ownership (the one deliberate difference from nn): the DEVICE copy is authoritative.  getParameters() returns host
FloatTensors that mirror it; `syncToHost()` refreshes them (adversarial.train does so at the end of every epoch, before train.lua's
saveAs), `syncFromHost()` pushes host edits back (done automatically by forward/backward when the host copy was handed out and
`self.hostDirty` was set by the caller).  Code that mutates PARAMETERS_D/G element-wise between steps must call syncFromHost(). ]]
local ffi = require("ffi")
local cg = require("catgen_ffi")
local models = {}

local Net = torch.class('catgen.Net')            -- a torch class so that torch.save / torch.load can carry it (train.lua:127-137,260)
local KIND_NAME = {[0] = "G32up", "G32up-c", "D32_st3", "V32"}
local function attach(self, kind, C, nz, seed)
   cg.init()
   local h = ffi.new("cg_model*[1]")
   cg.check(cg.lib.cg_model_create(h, kind, C, nz, seed or 1))
   local n = ffi.new("int64_t[1]"); cg.check(cg.lib.cg_model_nparams(h[0], n))
   self.h = ffi.gc(h[0], cg.lib.cg_model_free); self.kind, self.C, self.nz, self.nparams = kind, C, nz, tonumber(n[0])
   self.modules = {self}                       -- adversarial.lua:193 reads MODEL_D.modules[1].gradInput
   self.output = torch.FloatTensor(); self.gradInput = torch.FloatTensor()
   self.train = true
   return self
end
function Net:__init(kind, C, nz, seed) attach(self, kind, C, nz, seed) end
function Net:getParameters()                  -- train.lua:184-185: flat parameters and flat gradients (host mirrors)
   self.params = self.params or torch.FloatTensor(self.nparams)
   self.gradParams = self.gradParams or torch.FloatTensor(self.nparams)
   self.weight = self.params                   -- NN_UTILS.getNumberOfParameters sums listModules()[i].weight:nElement() (nn_utils.lua:453-462)
   self:syncToHost()
   return self.params, self.gradParams
end
function Net:syncToHost()
   if self.params then cg.check(cg.lib.cg_model_get_params(self.h, cg.ptr(self.params))) end
   if self.gradParams then cg.check(cg.lib.cg_model_get_grads(self.h, cg.ptr(self.gradParams))) end
end
function Net:syncFromHost() if self.params then cg.check(cg.lib.cg_model_set_params(self.h, cg.ptr(self.params))) end; self.hostDirty = false end
function Net:setParameters(p) cg.check(cg.lib.cg_model_set_params(self.h, cg.ptr(p:contiguous()))) end
function Net:zeroGradParameters() cg.check(cg.lib.cg_model_zero_grads(self.h)) end
function Net:training() self.train = true; cg.check(cg.lib.cg_model_set_mode(self.h, 1)); return self end
function Net:evaluate() self.train = false; cg.check(cg.lib.cg_model_set_mode(self.h, 0)); return self end
function Net:cuda() return self end             -- the network lives on the GPU already (NN_UTILS.activateCuda, train.lua:177)
function Net:float() return self end            -- train.lua:123,137,159: host-side view is float by construction
function Net:type() return 'torch.FloatTensor' end
function Net:listModules() return {self} end    -- nn_utils.lua:453-462, :630-643 (containsCopyLayers walks this list)
function Net:clearState() self.output:set(); self.gradInput:set(); return self end   -- nn_utils.lua:428-451 prepareNetworkForSave
function Net:clone()                            -- nn_utils.lua:630 activateCuda clones the network before wrapping it
   local c = catgen.Net(self.kind, self.C, self.nz, 1)
   local p = torch.FloatTensor(self.nparams); cg.check(cg.lib.cg_model_get_params(self.h, cg.ptr(p))); cg.check(cg.lib.cg_model_set_params(c.h, cg.ptr(p)))
   local n = ffi.new("int64_t[1]"); cg.check(cg.lib.cg_model_bn_running_len(self.h, n))
   if tonumber(n[0]) > 0 then
      local r = torch.FloatTensor(tonumber(n[0])); cg.check(cg.lib.cg_model_get_bn_running(self.h, cg.ptr(r))); cg.check(cg.lib.cg_model_set_bn_running(c.h, cg.ptr(r)))
   end
   if not self.train then c:evaluate() end
   return c
end
function Net:forward(input)
   if self.hostDirty then self:syncFromHost() end
   input = input:float():contiguous(); local B = input:size(1)
   if self.kind == cg.D32_ST3 then
      self.output:resize(B, 1)
      cg.check(cg.lib.cg_D_forward(self.h, cg.ptr(input), B, cg.ptr(self.output), nil))
   elseif self.kind == cg.V32 then              -- MODEL_V:forward in NN_UTILS.rateWithV (utils/nn_utils.lua:700): [B,2] SoftMax, evaluate() mode
      self.output:resize(B, 2)
      cg.check(cg.lib.cg_V_forward(self.h, cg.ptr(input), B, cg.ptr(self.output)))
   else
      self.output:resize(B, self.C, 32, 32)
      cg.check(cg.lib.cg_G_forward(self.h, cg.ptr(input), B, cg.ptr(self.output)))
   end
   return self.output                          -- reused on every call, like nn: callers clone what they keep
end
Net.updateOutput = Net.forward
function Net:backward(input, gradOutput)
   gradOutput = gradOutput:float():contiguous(); local B = input:size(1)
   if self.kind == cg.D32_ST3 then
      self.gradInput:resize(B, self.C, 32, 32)
      cg.check(cg.lib.cg_D_backward(self.h, cg.ptr(gradOutput), cg.ptr(self.gradInput)))
   else
      self.gradInput:resize(B, self.nz)
      cg.check(cg.lib.cg_G_backward(self.h, cg.ptr(gradOutput), cg.ptr(self.gradInput)))
   end
   return self.gradInput
end
function Net:__tostring() return string.format("catgen.Net(%s) [%d parameters, on the GPU]", KIND_NAME[self.kind], self.nparams) end
-- torch.save / torch.load (train.lua:260, :127-137): the FFI handle cannot be serialised; what is stored is what rebuilds it
function Net:write(file)
   local p = torch.FloatTensor(self.nparams); cg.check(cg.lib.cg_model_get_params(self.h, cg.ptr(p)))
   local n = ffi.new("int64_t[1]"); cg.check(cg.lib.cg_model_bn_running_len(self.h, n))
   local r = torch.FloatTensor(math.max(tonumber(n[0]), 1)):zero()
   if tonumber(n[0]) > 0 then cg.check(cg.lib.cg_model_get_bn_running(self.h, cg.ptr(r))) end
   file:writeObject({kind = self.kind, C = self.C, nz = self.nz, train = self.train, params = p, bn_running = r, nrun = tonumber(n[0])})
end
function Net:read(file)
   local t = file:readObject()
   attach(self, t.kind, t.C, t.nz, 1)
   cg.check(cg.lib.cg_model_set_params(self.h, cg.ptr(t.params)))
   if t.nrun > 0 then cg.check(cg.lib.cg_model_set_bn_running(self.h, cg.ptr(t.bn_running))) end
   if not t.train then self:evaluate() end
end

function models.create_G(dimensions, noiseDim)             -- models.lua:234-240
   assert(dimensions[2] == 32, "only the 32x32 generators are on the hot path (SURVEY.md section 2 row 15)")
   return catgen.Net(os.getenv("CATGEN_G") == "G32up" and cg.G32UP or cg.G32UPC, dimensions[1], noiseDim, OPT and OPT.seed or 1)
end
function models.create_D(dimensions, cuda)                 -- models.lua:268-277 -> create_D32_st3
   assert(dimensions[2] == 32, "only create_D32_st3 is on the hot path")
   assert(cuda, "libcatgen has no CPU path")
   return catgen.Net(cg.D32_ST3, dimensions[1], 100, (OPT and OPT.seed or 1) + 1)
end
function models.create_V(dimensions)                       -- models.lua:716-721 -> create_V32; forward / evaluate() only (train.lua:119-123)
   assert(dimensions[2] == 32, "only create_V32 is built")
   return catgen.Net(cg.V32, dimensions[1], 100, (OPT and OPT.seed or 1) + 2):evaluate()
end
return models
