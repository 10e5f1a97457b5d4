--[[ layers/SpatialConvolutionUpsample.lua -- drop-in for /root/reference/layers/SpatialConvolutionUpsample.lua and its
cudnn twin (layers/cudnnSpatialConvolutionUpsample.lua): conv to nOutputPlane*factor^2 planes, then a contiguous view
to (nOutputPlane, h*factor, w*factor) -- NOT a depth-to-space shuffle (SURVEY.md section 8 row A9).
STATUS: WRITTEN, NOT EXECUTED (see catgen_ffi.lua).  Unused by every model in models.lua; kept as module surface. ]]
local cg = require("catgen_ffi")
local SCU, parent = torch.class('nn.SpatialConvolutionUpsample', 'nn.Module')
function SCU:__init(nInputPlane, nOutputPlane, kW, kH, factor)
   parent.__init(self)
   factor = factor or 2
   assert(kW and kH and nInputPlane and nOutputPlane)
   assert(kW % 2 == 1, 'kW has to be odd'); assert(kH % 2 == 1, 'kH has to be odd'); assert(kW == kH, 'square kernels only')
   self.factor, self.kW, self.nInputPlaneU, self.nOutputPlaneU = factor, kW, nInputPlane, nOutputPlane
   local nOut = nOutputPlane * factor * factor
   local stdv = 1 / math.sqrt(kW * kH * nInputPlane)                       -- nn.SpatialConvolution:reset()
   self.weight = torch.FloatTensor(nOut, nInputPlane, kH, kW):uniform(-stdv, stdv); self.bias = torch.FloatTensor(nOut):uniform(-stdv, stdv)
   self.gradWeight = torch.FloatTensor(nOut, nInputPlane, kH, kW):zero(); self.gradBias = torch.FloatTensor(nOut):zero()
end
local function dims(input) if input:dim() == 4 then return input:size(1), input:size(3), input:size(4), true end return 1, input:size(2), input:size(3), false end
function SCU:updateOutput(input)
   cg.init(); input = input:contiguous(); local N, h, w, batched = dims(input); self.h, self.w = h, w
   if batched then self.output:resize(N, self.nOutputPlaneU, h * self.factor, w * self.factor) else self.output:resize(self.nOutputPlaneU, h * self.factor, w * self.factor) end
   cg.check(cg.lib.cg_conv_upsample_fwd(cg.ptr(input), cg.ptr(self.weight), cg.ptr(self.bias), cg.ptr(self.output), N, self.nInputPlaneU, h, w, self.nOutputPlaneU, self.kW, self.factor))
   return self.output
end
function SCU:updateGradInput(input, gradOutput)
   input = input:contiguous(); gradOutput = gradOutput:contiguous(); local N, h, w = dims(input); self.gradInput:resizeAs(input)
   cg.check(cg.lib.cg_conv_upsample_bwd(cg.ptr(input), cg.ptr(gradOutput), cg.ptr(self.weight), cg.ptr(self.gradInput), nil, nil, N, self.nInputPlaneU, h, w, self.nOutputPlaneU, self.kW, self.factor))
   return self.gradInput
end
function SCU:accGradParameters(input, gradOutput, scale)
   assert(scale == nil or scale == 1, "scale ~= 1 is not supported")
   input = input:contiguous(); gradOutput = gradOutput:contiguous(); local N, h, w = dims(input)
   cg.check(cg.lib.cg_conv_upsample_bwd(cg.ptr(input), cg.ptr(gradOutput), cg.ptr(self.weight), nil, cg.ptr(self.gradWeight), cg.ptr(self.gradBias), N, self.nInputPlaneU, h, w, self.nOutputPlaneU, self.kW, self.factor))
end
