// conv_tc.cu -- tcgen05 implicit-GEMM convolution engine (sm_100a).  NOT IMPLEMENTED YET in this commit:
// both entry points report CG_ERR_UNSUPPORTED so conv_ref.cu's fp32 CUDA-core kernels take every shape.
// The dispatcher treats UNSUPPORTED as "use the other CUDA engine", never as a CPU fallback.
#include "ops.cuh"
namespace cg {
int conv_fwd_tc(const float*, const float*, const float*, float*, int, int, int, int, int, int) { return CG_ERR_UNSUPPORTED; }
int conv_wgrad_tc(const float*, const float*, float*, int, int, int, int, int, int) { return CG_ERR_UNSUPPORTED; }
}  // namespace cg
