--[[ layers/cudnnSpatialConvolutionUpsample.lua -- drop-in for /root/reference/layers/cudnnSpatialConvolutionUpsample.lua:1-58
(class cudnn.SpatialConvolutionUpsample, required by train.lua:106).  Same constructor as the reference
(nInputPlane, nOutputPlane, kW, kH, factor, groups): a cudnn.SpatialConvolution to nOutputPlane*factor^2 planes, stride 1, pad (k-1)/2,
followed by the contiguous view to (nOutputPlane, h*factor, w*factor); `groups` other than 1 is an error here (no model uses it).
STATUS: WRITTEN, NOT EXECUTED (see catgen_ffi.lua).  The arithmetic is cg_conv_upsample_fwd / _bwd, the functions the nn twin binds. ]]
require 'layers.SpatialConvolutionUpsample'
cudnn = cudnn or {}
local SCU, parent = torch.class('cudnn.SpatialConvolutionUpsample', 'nn.SpatialConvolutionUpsample')
function SCU:__init(nInputPlane, nOutputPlane, kW, kH, factor, groups)
   assert(groups == nil or groups == 1, 'groups ~= 1 is not supported')
   parent.__init(self, nInputPlane, nOutputPlane, kW, kH, factor)
   self.groups = 1
end
-- the reference also defines accUpdateGradParameters (SpatialConvolutionUpsample.lua:49-56): gradient step fused into the weights
function SCU:accUpdateGradParameters(input, gradOutput, scale)
   self.gradWeight:zero(); self.gradBias:zero()
   self:accGradParameters(input, gradOutput, 1)
   self.weight:add(-scale, self.gradWeight); self.bias:add(-scale, self.gradBias)
end
