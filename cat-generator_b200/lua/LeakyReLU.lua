--[[ LeakyReLU.lua -- drop-in for /root/reference/LeakyReLU.lua (nn.LeakyReLU, slope 0.333).
STATUS: WRITTEN, NOT EXECUTED (see catgen_ffi.lua).  Same guard, constructor and protocol as the reference
(LeakyReLU.lua:2-31); updateGradInput gives gradOutput where input >= 0, INCLUDING input == 0 (:21-31). ]]
if nn.LeakyReLU then return end
local cg = require("catgen_ffi")
local M, base = torch.class('nn.LeakyReLU', 'nn.Module')   -- same class name: the reference's models.lua instantiates nn.LeakyReLU
function M:__init(slope)
   base.__init(self)
   self.negative_scale = slope or 0.333   -- field name kept: checkpoints of the reference carry it
end
function M:updateOutput(input)
   cg.init(); input = input:contiguous(); self.output:resizeAs(input)
   cg.check(cg.lib.cg_leakyrelu_fwd(cg.ptr(input), self.negative_scale, cg.ptr(self.output), input:nElement()))
   return self.output
end
function M:updateGradInput(input, gradOutput)
   input = input:contiguous(); gradOutput = gradOutput:contiguous(); self.gradInput:resizeAs(gradOutput)
   cg.check(cg.lib.cg_leakyrelu_bwd(cg.ptr(input), cg.ptr(gradOutput), self.negative_scale, cg.ptr(self.gradInput), input:nElement()))
   return self.gradInput
end
