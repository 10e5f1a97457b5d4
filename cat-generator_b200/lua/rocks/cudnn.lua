--[[ rocks/cudnn.lua -- stand-in for the `cudnn` rock: only the table, so that `cudnn.SpatialConvolutionUpsample`
(layers/cudnnSpatialConvolutionUpsample.lua) has a home.  STATUS: WRITTEN, NOT EXECUTED (see catgen_ffi.lua). ]]
cudnn = cudnn or {}
return cudnn
