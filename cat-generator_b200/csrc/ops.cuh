// ops.cuh -- internal launchers.  All pointers are DEVICE pointers; activations are NHWC fp32
// ([N,H,W,C], "BHWD" in stn terms) unless a name says nchw.  Everything is enqueued on cg::ctx().stream.
#pragma once
#include "common.cuh"

namespace cg {

// ---- layout (boundary <-> internal)
int nchw_to_nhwc(const float* x, float* y, int N, int C, int HW);
int nhwc_to_nchw(const float* x, float* y, int N, int C, int HW);

// ---- pointwise
int prelu_fwd(const float* x, const float* w, float* y, long n);
int prelu_bwd(const float* x, const float* gy, const float* w, float* gx, float* gw_acc, long n);
int lrelu_fwd(const float* x, float s, float* y, long n);
int lrelu_bwd(const float* x, const float* gy, float s, float* gx, long n);
int sigmoid_fwd(const float* x, float* y, long n);
int sigmoid_bwd(const float* y, const float* gy, float* gx, long n);
int softmax_rows(const float* x, float* y, long rows, int C);   // nn.SoftMax over the last dimension
int add_inplace(float* a, const float* b, long n);
int fill(float* a, float v, long n);
int mask_channels(const float* x, const float* mask_nc, float* y, int N, int HW, int C);
int mask_elems(const float* x, const float* mask, float* y, long n);

// ---- NHWC spatial (H, W are the INPUT dims of the forward op)
int upsample2x_fwd(const float* x, float* y, int N, int H, int W, int C);
int upsample2x_bwd(const float* gy, float* gx, int N, int H, int W, int C);
int avgpool2_fwd(const float* x, float* y, int N, int H, int W, int C);
int avgpool2_bwd(const float* gy, float* gx, int N, int H, int W, int C);
int maxpool2_fwd(const float* x, float* y, uint8_t* idx, int N, int H, int W, int C);
int maxpool2_bwd(const float* gy, const uint8_t* idx, float* gx, int N, int H, int W, int C);

// ---- batch norm over rows of [M, C]
int bn_fwd_train(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* invstd,
                 float* run_mean, float* run_var, long M, int C, float eps, float mom);
// pre_part / pre_S: per-channel partial sums already produced elsewhere (the conv epilogue), [pre_S][C][2] doubles with room for one more row
int bn_fwd_train_pre(const double* pre_part, int pre_S, const float* x, const float* gamma, const float* beta, float* y, float* mean, float* invstd,
                     float* run_mean, float* run_var, long M, int C, float eps, float mom);
int bn_fwd_eval(const float* x, const float* gamma, const float* beta, float* y, const float* run_mean,
                const float* run_var, long M, int C, float eps);
int bn_bwd(const float* x, const float* gy, const float* gamma, const float* mean, const float* invstd,
           float* gx, float* ggamma_acc, float* gbeta_acc, long M, int C);
int colsum_acc(const float* x, float* out_acc, long M, int C);   // out[c] += sum_rows x[r,c]
// same, plus max|x| -> the power-of-two operand scale [s, 1/s]; *amax_zeroed must be 0 on entry and is 0 again on exit
int colsum_acc_absmax(const float* x, float* out_acc, long M, int C, unsigned int* amax_zeroed, float* scale2);

// ---- spatial transformer
int affine_matrix_fwd(const float* theta, float* A, int B, int rot, int scl, int trn);
int affine_matrix_bwd(const float* theta, const float* gA, float* gtheta, int B, int rot, int scl, int trn);
int affine_grid_fwd(const float* A, float* grid, int B, int H, int W);
int affine_grid_bwd(const float* ggrid, float* gA, int B, int H, int W);
int bilinear_fwd(const float* img, const float* grid, float* out, int B, int H, int W, int C);
int bilinear_bwd(const float* img, const float* grid, const float* gout, float* gimg, float* ggrid, int B, int H, int W, int C);

// ---- criterion / optimiser / rng
int bce(const float* p, const float* t, int n, float* loss_dev, float* g);
// g += sign(p)*l1sign + p*l2; clamp; *loss_add_dev = l1*|p|_1 + l2*|p|^2/2 (if non-null)
int penalty_clamp(float* g, const float* p, long n, float l1, float l1sign, float l2, float clampv, float* loss_add_dev);
int adam(float* x, const float* g, float* m, float* v, long n, int* t_dev, float lr, float b1, float b2, float eps);   // increments *t_dev, then steps
// The optimiser tail of one closure in ONE pass over the flat vectors: g <- clamp(g*gscale + sign(p)*l1sign + p*l2) ; *loss_add_dev as
// penalty_clamp ; then the optim.adam update with that g (t_dev incremented first).  Arithmetic per element is penalty_clamp's followed
// by adam's, so results are bit-identical to the two-pass form.  gscale: 1/world after a data-parallel sum-all-reduce, else 1.
int penalty_clamp_adam(float* g, float* p, float* m, float* v, long n, float gscale, float l1, float l1sign, float l2, float clampv, float* loss_add_dev,
                       int* t_dev, float lr, float b1, float b2, float eps);
// data parallel (capi.cu): in-place sum over ranks on the current stream; no-ops for one rank
int dist_allreduce_sum_f32(float* buf, long n);
int dist_allreduce_sum_f64(double* buf, long n);
int uniform(float* dst, long n, float lo, float hi, uint64_t seed, uint64_t offset);
// dst[i] = (u >= p_drop) ? keep_value : 0
int bernoulli_mask(float* dst, long n, float p_drop, float keep_value, uint64_t seed, const unsigned long long* offset_dev, uint64_t rel);
int rng_advance(unsigned long long* offset_dev, uint64_t by);
int scale_inplace(float* a, float s, long n);

// ---- parameter packing (Torch layout <-> kernel layout); see conv_ref.cu
struct ConvSpec {     // also describes nn.Linear as a 1x1 conv on a [N,1,1,in] tensor
  int Ci, Co, k;
  // Linear only: the Torch in/out feature index f = c*HW + s is permuted to the NHWC index s*C + c
  int in_hw = 1, out_hw = 1;
};
// index into the Torch-layout weight of (tap, input feature cip, output feature cop) in kernel (NHWC) order
__device__ __forceinline__ long torch_index(int k, int Ci, int Co, int in_hw, int out_hw, int ky, int kx, int cip, int cop) {
  if (in_hw == 1 && out_hw == 1) return (((long)cop * Ci + cip) * k + ky) * k + kx;          // conv, or plain Linear
  int Civ = Ci / in_hw, Cov = Co / out_hw;                                                      // Linear beside an nn.View
  int fi = (cip % Civ) * in_hw + cip / Civ;
  int fo = (cop % Cov) * out_hw + cop / Cov;
  return (long)fo * Ci + fi;
}
int pack_fprop(const float* W, float* Wp, const ConvSpec& s);    // Wp[(ky,kx,ci)][co]
int pack_dgrad(const float* W, float* Wd, const ConvSpec& s);    // Wd[(ky,kx,co)][ci], taps flipped
int pack_bias(const float* b, float* bp, const ConvSpec& s);     // permuted for Linear with out_hw > 1
int unpack_wgrad_acc(const float* gWp, float* gW_acc, const ConvSpec& s);   // gW (Torch layout) += gWp
int unpack_bias_acc(const float* gbp, float* gb_acc, const ConvSpec& s);

// ---- convolution engines (NHWC, stride 1, pad (k-1)/2).  bias may be null.
int conv_fwd(const float* x, const float* Wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co, int k);
int conv_dgrad(const float* gy, const float* Wd, float* gx, int N, int H, int W, int Ci, int Co, int k);
// gWp_out[(ky,kx,ci)][co] = sum_pixels x[p+tap,ci]*gy[p,co]  (overwritten, packed layout)
// gW_acc/done: optional direct accumulation into the Torch-layout gradient of a plain conv (see conv_ref.cu)
int conv_wgrad(const float* x, const float* gy, float* gWp_out, int N, int H, int W, int Ci, int Co, int k, float* gW_acc = nullptr, int* done = nullptr);
int parts_to_torch_acc(const float* part, int Z, long zstride, float* gW_acc, int Ci, int Co, int kk);
int splitk_reduce(const float* part, int S, long n, int ncol, const float* bias, float* out);   // out[i] = bias[i % ncol] + sum_z part[z][i], fixed z order
// both gradients of one layer (gWp_out overwritten, gx written)
int conv_backward(const float* x, const float* gy, const float* Wd, float* gWp_out, float* gx, int N, int H, int W, int Ci, int Co, int k, float* gW_acc = nullptr, int* done = nullptr,
                  const uint8_t* xq = nullptr, float* gb_acc = nullptr, int* bias_done = nullptr);   // gb_acc: bias gradient; *bias_done = 1 when the engine added it
// cached fp16 operand path of the tensor-core engine (conv_tc.cu): PReLU(BN(x)) -> 2x upsample -> packed operand in one pass
bool conv_tc_cached_ok(int H, int W, int Ci, int Co, int k);
size_t conv_tc_operand_bytes(int N, int H, int W, int Ci, int k);
int bn_prelu_up_pack(const float* x, const float* gamma, const float* beta, const float* mean, const float* invstd, const float* pw,
                     float* bn_out, uint8_t* xq, int N, int h, int w, int C, int up, int k);
int conv_fwd_tc_packed(const uint8_t* xq, const float* Wp, const float* bias, float* y, int N, int H, int W, int Ci, int Co, int k);
int conv_bwd_tc_gq(const float* x, const uint8_t* xq_prepacked, const uint8_t* gq, const float* scale2, const float* Wd, float* gx,
                   int N, int H, int W, int Ci, int Co, int k, float* gW_acc);


// ---- whole-model repack in ONE launch (conv_tc.cu): fp32 fprop/dgrad/bias operands of every layer plus, for layers the
// tensor-core engine takes, the fp16 weight slices of both directions.  The slices are registered under the fp32 operand's
// pointer, so the engine finds them instead of re-packing per call.
struct PackJob {
  const float *W, *b; float *Wp, *Wd, *bp; uint8_t *wqf, *wqd;   // wqf/wqd null: not a tensor-core layer
  ConvSpec s; int need_dgrad, CBf, CBd;
  int blk0, nblk;                                                 // this job's slice of the flat grid (blocks proportional to its weight count)
  int always32;                                                   // the fp32 operands Wp / Wd are refreshed on EVERY repack (the fused spatial transformer reads them: stn_fused.cu)
};
int repack_model(const PackJob* jobs_dev, int njobs, int total_blocks, int mode);   // mode: 1 fp32 operands, 2 bias + fp16 slices, 3 both
bool conv_tc_all_shapes_taken(int B);   // true when every conv / Linear call at batch B goes to the tensor-core engine (no fp32 operand is read)
bool conv_tc_wslice_plan(int Cin, int Cout, int k, int* CB, size_t* bytes);   // fp16 forward-conv slices for Cin -> Cout
void conv_tc_register_wslices(const float* key, const uint8_t* wq, int CB);
void conv_tc_unregister_wslices(const float* key);

}  // namespace cg
