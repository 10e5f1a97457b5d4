--[[ rocks/cunn.lua -- empty stand-in for the `cunn` rock (train.lua:101-107 requires it before building any model).  Every module of it
that /root/reference/models.lua instantiates lives inside libcatgen's G / D executors (csrc/model.cu), so nothing is needed here.
STATUS: WRITTEN, NOT EXECUTED (see catgen_ffi.lua). ]]
return {}
