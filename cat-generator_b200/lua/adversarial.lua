--[[ adversarial.lua -- drop-in for /root/reference/adversarial.lua's adversarial.train (:27-292).
STATUS: WRITTEN, NOT EXECUTED (see catgen_ffi.lua).  The per-sample Lua loops of the reference (:101-106, :225-238)
collapse into one cg_train_step call per step; the order of operations inside the step is the reference's
(csrc/capi.cu train_step_core cites the lines).  Uses the same globals train.lua sets (OPT, MODEL_G, MODEL_D,
CONFUSION, IMG_DIMENSIONS, EPOCH). ]]
local ffi = require("ffi")
local cg = require("catgen_ffi")
local adversarial = {accs = {}}
local trainer

function adversarial.train(trainData, maxAccuracyD, accsInterval)
   EPOCH = EPOCH or 1
   -- D_maxAcc <= 1 makes fevalD return false,false (adversarial.lua:157-166), which stock optim.adam cannot digest; refuse it loudly
   -- like the Python twin (catgen/adversarial.py) instead of ignoring the arguments
   if maxAccuracyD and maxAccuracyD <= 1 then error("catgen: D_maxAcc <= 1 (skipping D updates) is not supported") end
   local N_epoch = OPT.N_epoch; if N_epoch <= 0 then N_epoch = trainData:size() end
   local half = OPT.batchSize / 2
   if not trainer then
      local t = ffi.new("cg_trainer*[1]"); cg.check(cg.lib.cg_trainer_create(t, MODEL_G.h, MODEL_D.h)); trainer = ffi.gc(t[0], cg.lib.cg_trainer_free)
   end
   local time = sys.clock()
   print(string.format("<trainer> Epoch #%d [batchSize = %d]", EPOCH, OPT.batchSize))
   for t = 1, N_epoch, half do                                                      -- :51
      local thisB = math.min(OPT.batchSize, N_epoch - t + 1)                        -- :53
      thisB = thisB - thisB % 2                                                     -- fevalD needs B/2 real + B/2 fake (odd tails of trainData:size())
      if thisB < 4 then print(string.format("[INFO] skipping batch at t=%d, because its size is less than 4", t)); break end  -- :65-68
      local cfg = ffi.new("cg_step_cfg", {thisB, OPT.D_iterations, OPT.G_iterations, OPT.D_L1, OPT.D_L2, OPT.G_L1, OPT.G_L2,
                                          OPT.D_clamp, OPT.G_clamp, 1e-3, 0.9, 0.999, 1e-8})
      local C, H, W = IMG_DIMENSIONS[1], IMG_DIMENSIONS[2], IMG_DIMENSIONS[3]
      local real = torch.FloatTensor(OPT.D_iterations, math.floor(thisB / 2), C, H, W)
      for k = 1, OPT.D_iterations do for i = 1, math.floor(thisB / 2) do real[k][i]:copy(trainData[math.random(trainData:size())]) end end  -- :225-230
      local zD = torch.FloatTensor(OPT.D_iterations, math.floor(thisB / 2), OPT.noiseDim):uniform(-1, 1)   -- :233 via nn_utils.lua:35-39
      local zG = torch.FloatTensor(OPT.G_iterations, thisB, OPT.noiseDim):uniform(-1, 1)       -- :254
      local dout = torch.FloatTensor(thisB)
      cg.check(cg.lib.cg_train_step(trainer, cfg, cg.ptr(real), cg.ptr(zD), cg.ptr(zG), nil, nil, cg.ptr(dout)))
      for i = 1, thisB do                                                            -- :101-106
         local c = dout[i] > 0.5 and 2 or 1
         CONFUSION:add(c, (i <= math.floor(thisB / 2)) and 2 or 1)
      end
      xlua.progress(t + thisB, N_epoch)                                              -- :270
   end
   time = sys.clock() - time
   print(string.format("<trainer> time required for this epoch = %d s", time))                       -- :278-280
   print(string.format("<trainer> time to learn 1 sample = %f ms", 1000 * time / N_epoch))
   print("Confusion of D:"); print(CONFUSION)
   CONFUSION:updateValids()
   local tV = CONFUSION.totalValid; CONFUSION:zero()
   MODEL_G:syncToHost(); MODEL_D:syncToHost()      -- the single hand-off: PARAMETERS_* / GRAD_PARAMETERS_* (train.lua:184-185) mirror the device again
   return tV
end
return adversarial
