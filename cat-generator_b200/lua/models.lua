--[[ models.lua -- drop-in for /root/reference/models.lua's create_G / create_D (models.lua:234-240, :268-277).
STATUS: WRITTEN, NOT EXECUTED (see catgen_ffi.lua).  The returned objects implement the part of the nn.Module
protocol train.lua and adversarial.lua use (train.lua:147-185; adversarial.lua:84-89,187-197): forward, backward,
getParameters, training, evaluate, zeroGradParameters, and `modules[1].gradInput` on D. ]]
local ffi = require("ffi")
local cg = require("catgen_ffi")
local models = {}

local Net = {}; Net.__index = Net
local function new_net(kind, C, nz, seed)
   cg.init()
   local h = ffi.new("cg_model*[1]")
   cg.check(cg.lib.cg_model_create(h, kind, C, nz, seed or 1))
   local n = ffi.new("int64_t[1]"); cg.check(cg.lib.cg_model_nparams(h[0], n))
   local self = setmetatable({h = ffi.gc(h[0], cg.lib.cg_model_free), kind = kind, C = C, nz = nz, nparams = tonumber(n[0])}, Net)
   self.modules = {self}                       -- adversarial.lua:193 reads MODEL_D.modules[1].gradInput
   return self
end
function Net:getParameters()                  -- train.lua:184-185: flat parameters and flat gradients
   self.params = self.params or torch.FloatTensor(self.nparams)
   self.gradParams = self.gradParams or torch.FloatTensor(self.nparams)
   cg.check(cg.lib.cg_model_get_params(self.h, cg.ptr(self.params)))
   cg.check(cg.lib.cg_model_get_grads(self.h, cg.ptr(self.gradParams)))
   return self.params, self.gradParams
end
function Net:setParameters(p) cg.check(cg.lib.cg_model_set_params(self.h, cg.ptr(p:contiguous()))) end
function Net:zeroGradParameters() cg.check(cg.lib.cg_model_zero_grads(self.h)) end
function Net:training() cg.check(cg.lib.cg_model_set_mode(self.h, 1)) end
function Net:evaluate() cg.check(cg.lib.cg_model_set_mode(self.h, 0)) end
function Net:forward(input)
   input = input:float():contiguous(); local B = input:size(1)
   if self.kind == cg.D32_ST3 then
      self.output = self.output or torch.FloatTensor(); self.output:resize(B, 1)
      cg.check(cg.lib.cg_D_forward(self.h, cg.ptr(input), B, cg.ptr(self.output), nil))
   else
      self.output = self.output or torch.FloatTensor(); self.output:resize(B, self.C, 32, 32)
      cg.check(cg.lib.cg_G_forward(self.h, cg.ptr(input), B, cg.ptr(self.output)))
   end
   return self.output                          -- reused on every call, like nn: callers clone what they keep
end
function Net:backward(input, gradOutput)
   gradOutput = gradOutput:float():contiguous(); local B = input:size(1)
   self.gradInput = self.gradInput or torch.FloatTensor()
   if self.kind == cg.D32_ST3 then
      self.gradInput:resize(B, self.C, 32, 32)
      cg.check(cg.lib.cg_D_backward(self.h, cg.ptr(gradOutput), cg.ptr(self.gradInput)))
   else
      self.gradInput:resize(B, self.nz)
      cg.check(cg.lib.cg_G_backward(self.h, cg.ptr(gradOutput), cg.ptr(self.gradInput)))
   end
   return self.gradInput
end
function Net:__tostring() return string.format("catgen %s [%d parameters]", ({[0]="G32up", "G32up-c", "D32_st3"})[self.kind], self.nparams) end

function models.create_G(dimensions, noiseDim)             -- models.lua:234-240
   assert(dimensions[2] == 32, "only the 32x32 generators are on the hot path (SURVEY.md section 2 row 15)")
   return new_net(os.getenv("CATGEN_G") == "G32up" and cg.G32UP or cg.G32UPC, dimensions[1], noiseDim)
end
function models.create_D(dimensions, cuda)                 -- models.lua:268-277 -> create_D32_st3
   assert(dimensions[2] == 32, "only create_D32_st3 is on the hot path")
   assert(cuda, "libcatgen has no CPU path")
   return new_net(cg.D32_ST3, dimensions[1], 100)
end
return models
