--[[ rocks/cutorch.lua -- stand-in for the `cutorch` rock so that an UNCHANGED train.lua runs on a Torch7 without CUDA rocks
(train.lua:102,109-110 call cutorch.setDevice / cutorch.manualSeed).  STATUS: WRITTEN, NOT EXECUTED (see catgen_ffi.lua).
Put this directory on package.path ahead of the real rocks: LUA_PATH="cat-generator_b200/lua/rocks/?.lua;cat-generator_b200/lua/?.lua;;" ]]
local cg = require("catgen_ffi")
cutorch = {}
function cutorch.setDevice(d) cg.init(d - 1) end                   -- train.lua:109 passes OPT.gpu + 1 (Torch devices are 1-based)
function cutorch.manualSeed(s) cutorch.seed = s end                 -- parameters are seeded in models.create_* from OPT.seed
function cutorch.synchronize() cg.check(cg.lib.cg_sync()) end
function cutorch.getDeviceCount() return tonumber(os.getenv("CATGEN_GPUS") or "1") end   -- --gpu admits 0..3 only (train.lua:57): more GPUs come from the environment
return cutorch
