--[[ utils/nn_utils.lua -- drop-in for the functions of /root/reference/utils/nn_utils.lua that train.lua, adversarial.lua and sample.lua
call on the hot path and next to it (F1): createNoiseInputs (:35-39), createImagesFromNoise (:45-69), createImages (:75-77),
sortImagesByPrediction (:89-117), switchToTrainingMode / switchToEvaluationMode (:334-349), prepareNetworkForSave (:428-451),
getNumberOfParameters (:453-462), activateCuda (:620-680), rateWithV (:686-711).  Uses the same globals (OPT, MODEL_G, MODEL_D, MODEL_V).
STATUS: WRITTEN, NOT EXECUTED (see catgen_ffi.lua).  The Python twin catgen/nn_utils.py IS executed (tests/test_gpu_parity.py
test_sampler_path_matches_oracle, tests/test_host_logic.py). ]]
local nn_utils = {}
function nn_utils.createNoiseInputs(N)
   local noiseInputs = torch.Tensor(N, OPT.noiseDim):float()
   noiseInputs:uniform(-1.0, 1.0)
   return noiseInputs
end
function nn_utils.createImagesFromNoise(noiseInputs, outputAsList)
   local images
   local N = noiseInputs:size(1)
   for i = 1, math.ceil(N / OPT.batchSize) do
      local batchStart, batchEnd = 1 + (i - 1) * OPT.batchSize, math.min(i * OPT.batchSize, N)
      local generated = MODEL_G:forward(noiseInputs[{{batchStart, batchEnd}}]):clone()     -- module output is reused: clone what is kept
      images = images or torch.Tensor(N, generated:size(2), generated:size(3), generated:size(4)):float()
      images[{{batchStart, batchEnd}, {}, {}, {}}] = generated
   end
   if not outputAsList then return images end
   local list = {}
   for i = 1, images:size(1) do list[#list + 1] = images[i]:float() end
   return list
end
function nn_utils.createImages(N, outputAsList) return nn_utils.createImagesFromNoise(nn_utils.createNoiseInputs(N), outputAsList) end
function nn_utils.sortImagesByPrediction(images, ascending, nbMaxOut)
   local predictions = torch.Tensor(images:size(1), 1)
   for i = 1, math.ceil(images:size(1) / OPT.batchSize) do
      local batchStart, batchEnd = 1 + (i - 1) * OPT.batchSize, math.min(i * OPT.batchSize, images:size(1))
      predictions[{{batchStart, batchEnd}, {1}}] = MODEL_D:forward(images[{{batchStart, batchEnd}, {}, {}, {}}]):clone()
   end
   local pairs_ = {}
   for i = 1, images:size(1) do pairs_[#pairs_ + 1] = {images[i], predictions[i][1]} end
   table.sort(pairs_, ascending and function(a, b) return a[2] < b[2] end or function(a, b) return a[2] > b[2] end)
   local resultImages, resultPredictions = {}, {}
   for i = 1, math.min(nbMaxOut, #pairs_) do resultImages[i] = pairs_[i][1]; resultPredictions[i] = pairs_[i][2] end
   return resultImages, resultPredictions
end
function nn_utils.switchToTrainingMode() if MODEL_AE then MODEL_AE:training() end; MODEL_G:training(); MODEL_D:training() end
function nn_utils.switchToEvaluationMode() if MODEL_AE then MODEL_AE:evaluate() end; MODEL_G:evaluate(); MODEL_D:evaluate() end
function nn_utils.prepareNetworkForSave(node, nogc)
   node:clearState()
   if node.syncToHost then node:syncToHost() end           -- catgen.Net: refresh the host mirrors train.lua's PARAMETERS_* point at
   if not nogc then collectgarbage() end
end
function nn_utils.getNumberOfParameters(net)
   local nparams, mods = 0, net:listModules()
   for i = 1, #mods do
      if mods[i].nparams then nparams = nparams + mods[i].nparams
      elseif mods[i].weight ~= nil then nparams = nparams + mods[i].weight:nElement() end
   end
   return nparams
end
function nn_utils.containsCopyLayers(net)
   local mods = net:listModules()
   for i = 1, #mods do if string.find(torch.type(mods[i]), "Copy") then return true end end
   return false
end
-- The reference clones the network and wraps it in Copy(Float->Cuda) / Copy(Cuda->Float) (:620-680).  A catgen.Net takes host float
-- tensors in and hands host float tensors out already -- the Copy pair is inside cg_G_forward / cg_D_forward -- so only the clone remains.
function nn_utils.activateCuda(net) return net:clone():cuda() end
function nn_utils.visualizeProgress(noiseInputs)           -- :119-332 needs the `display` rock; train.lua guards the call with --noplot
   if not DISP then return end
   nn_utils.switchToEvaluationMode()
   local images = nn_utils.createImagesFromNoise(noiseInputs, true)
   nn_utils.switchToTrainingMode()
   DISP.image(images, {win = OPT.window or 3, title = string.format("Generated images (epoch %d)", EPOCH or 0)})
end
function nn_utils.rateWithV(images)                        -- :686-711: 1 - mean P(fake) under the validator network V
   local imagesTensor, N
   if type(images) == 'table' then
      N = #images; imagesTensor = torch.Tensor(N, IMG_DIMENSIONS[1], IMG_DIMENSIONS[2], IMG_DIMENSIONS[3])
      for i = 1, N do imagesTensor[i] = images[i] end
   else N = images:size(1); imagesTensor = images end
   local predictions = MODEL_V:forward(imagesTensor)
   local sm = 0
   for i = 1, N do sm = sm + predictions[i][1] end
   return 1 - sm / N
end
return nn_utils
