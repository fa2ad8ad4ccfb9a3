// c_api.cu — extern "C" wrappers of the single kernels, PIT loss, overlap-add and the step tail
// (declared in include/ctn_b200.h).  The whole-model entry points live in model.cu.
#include "common.cuh"

namespace ctn {
int run_encoder_fwd(const float*, const float*, int, int, int, int, float*, cudaStream_t);
int run_encoder_bwd(const float*, const float*, const float*, const float*, int, int, int, int, float*, cudaStream_t);
int run_row_stats(const float*, const float*, int64_t, int, float*, cudaStream_t, int bf16 = 0);
int run_prep_normfold(const float*, const float*, const float*, int, int, int, int64_t, float*, float*, float*, int64_t,
                      int64_t, cudaStream_t);
int run_dwconv_fwd(const float*, const float*, NormStats, const float*, const float*, const float*, int, int, int, int,
                   int, int, float*, double*, const float*, cudaStream_t, int bf16 = 0, float* rs1_out = nullptr,
                   float* rs2_out = nullptr);
int run_dwconv_bwd(const float*, const float*, const float*, NormStats, const float*, const float*, const float*, int,
                   int, int, int, int, int, float*, float*, float*, float*, double*, float*, int, cudaStream_t);
int run_dwconv_bwd_gln_fused(const float*, const float*, const float*, NormStats, const float*, const double*, float*,
                             const float*, const float*, NormStats, const float*, const float*, const float*, int, int, int,
                             int, int, int, float*, float*, float*, float*, double*, float*, int, cudaStream_t);
int run_norm_bwd_reduce(const float*, const float*, const float*, NormStats, const float*, int, int, int, float*,
                        float*, double*, float*, int, cudaStream_t);
int run_norm_bwd_apply(float*, const float*, const float*, NormStats, const float*, const double*, int, int, int,
                       float*, cudaStream_t);
int run_decoder_fwd(const float*, const float*, const float*, int, int, int, int, int, int, int, float*, cudaStream_t);
int run_decoder_bwd(const float*, const float*, const float*, const float*, int, int, int, int, int, int, int, float*,
                    float*, float*, cudaStream_t);
int run_overlap_and_add(const float*, int64_t, int, int, int, float*, cudaStream_t);
int64_t pit_workspace_bytes(int, int);
int run_pit_forward(const float*, float*, const int64_t*, int, int, int, float*, float*, int64_t*, float*, float*, void*,
                    cudaStream_t);
int run_pit_backward(const float*, const float*, const int64_t*, const float*, const float*, int, int, int, float*,
                     cudaStream_t);
int run_reorder(const float*, const int64_t*, int, int, int64_t, float*, cudaStream_t);
int64_t sisnri_workspace_bytes(int, int);
int run_sisnri(const float*, const float*, const float*, const int64_t*, int, int, int, float*, float*, void*,
               cudaStream_t);
int run_clip_grad_norm(float*, int64_t, float, float*, void*, cudaStream_t);
int run_adam_step(float*, const float*, float*, float*, int64_t, float, float, float, float, float, int64_t*,
                  cudaStream_t);

int run_assemble_batch(const float*, const float*, const int64_t*, int, int, int, float*, float*, int64_t*, cudaStream_t);
int run_pack_valid(const float*, const int64_t*, const int64_t*, int, int, int, float*, cudaStream_t);

int run_bn_forward_stats(const float*, const float*, const float*, const float*, float*, float*, int64_t, int, int, double*,
                         float*, float*, float*, float*, float*, cudaStream_t);
int run_bn_bwd_finalize(const float*, const float*, const float*, const float*, const float*, int, int64_t, int, float*,
                        float*, float*, float*, cudaStream_t);
int run_bn_bwd_apply(float*, const float*, const float*, const float*, const float*, const float*, const float*, int64_t,
                     int, float*, cudaStream_t);

static NormStats make_stats(const double* gln_acc, const float* rowstat, int K, int Ch) {
  NormStats st;
  st.acc = gln_acc;
  st.row = rowstat;
  st.inv_count = 1.0 / ((double)K * (double)Ch);
  return st;
}
}  // namespace ctn

using namespace ctn;

extern "C" {

int64_t ctn_pit_workspace_bytes(int32_t B, int32_t C) { return pit_workspace_bytes(B, C); }

int64_t ctn_sisnri_workspace_bytes(int32_t B, int32_t C) { return sisnri_workspace_bytes(B, C); }

int32_t ctn_sisnri(const float* source, const float* reordered_est, const float* mixture, const int64_t* lengths, int32_t B,
                   int32_t C, int32_t T, float* sisnri, float* sisnr_est, void* ws, cudaStream_t stream) {
  CTN_REQUIRE(source && reordered_est && mixture && lengths && sisnri && ws, "sisnri: null pointer");
  return run_sisnri(source, reordered_est, mixture, lengths, B, C, T, sisnri, sisnr_est, ws, stream);
}

int32_t ctn_assemble_batch(const float* packed_mix, const float* packed_src, const int64_t* offsets, int32_t B, int32_t C,
                           int32_t T, float* padded_mixture, float* padded_source, int64_t* lengths,
                           cudaStream_t stream) {
  CTN_REQUIRE(packed_mix && offsets && padded_mixture && lengths, "assemble_batch: null pointer");
  return run_assemble_batch(packed_mix, packed_src, offsets, B, C, T, padded_mixture, padded_source, lengths, stream);
}

int32_t ctn_pack_valid(const float* inputs, const int64_t* lengths, const int64_t* out_offsets, int32_t B, int32_t C,
                       int32_t T, float* packed, cudaStream_t stream) {
  CTN_REQUIRE(inputs && lengths && out_offsets && packed, "pack_valid: null pointer");
  return run_pack_valid(inputs, lengths, out_offsets, B, C, T, packed, stream);
}

int32_t ctn_pit_forward(const float* source, float* est, const int64_t* lengths, int32_t B, int32_t C, int32_t T,
                        float* loss, float* max_snr, int64_t* idx, float* reorder, float* coef, void* pit_ws,
                        cudaStream_t stream) {
  CTN_REQUIRE(source && est && lengths && loss && max_snr && idx && coef && pit_ws, "pit_forward: null pointer");
  return run_pit_forward(source, est, lengths, B, C, T, loss, max_snr, idx, reorder, coef, pit_ws, stream);
}

int32_t ctn_pit_backward(const float* source, const float* est_masked, const int64_t* lengths, const float* coef,
                         const float* grad_loss, int32_t B, int32_t C, int32_t T, float* d_est, cudaStream_t stream) {
  CTN_REQUIRE(source && est_masked && lengths && coef && d_est, "pit_backward: null pointer");
  return run_pit_backward(source, est_masked, lengths, coef, grad_loss, B, C, T, d_est, stream);
}

int32_t ctn_reorder_source(const float* source, const int64_t* idx, int32_t B, int32_t C, int64_t inner, float* out,
                           cudaStream_t stream) {
  CTN_REQUIRE(source && idx && out, "reorder_source: null pointer");
  return run_reorder(source, idx, B, C, inner, out, stream);
}

int32_t ctn_overlap_and_add(const float* signal, int64_t outer, int32_t frames, int32_t frame_length,
                            int32_t frame_step, float* out, cudaStream_t stream) {
  CTN_REQUIRE(signal && out && outer >= 1 && frames >= 1 && frame_length >= 1, "overlap_and_add: bad arguments");
  return run_overlap_and_add(signal, outer, frames, frame_length, frame_step, out, stream);
}

int32_t ctn_clip_grad_norm(float* grads, int64_t n, float max_norm, float* norm_out, void* scratch,
                           cudaStream_t stream) {
  CTN_REQUIRE(grads && norm_out && scratch, "clip_grad_norm: null pointer");
  return run_clip_grad_norm(grads, n, max_norm, norm_out, scratch, stream);
}

int32_t ctn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int64_t* step_dev, cudaStream_t stream) {
  CTN_REQUIRE(params && grads && exp_avg && exp_avg_sq && step_dev && n > 0, "adam_step: bad arguments");
  return run_adam_step(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step_dev, stream);
}

int32_t ctn_encoder_fwd(const float* mix, const float* U, int32_t M, int32_t T, int32_t N, int32_t L, float* w,
                        cudaStream_t stream) {
  return run_encoder_fwd(mix, U, M, T, N, L, w, stream);
}

int32_t ctn_encoder_bwd(const float* mix, const float* w, const float* dw_a, const float* dw_b, int32_t M, int32_t T,
                        int32_t N, int32_t L, float* dU, cudaStream_t stream) {
  return run_encoder_bwd(mix, w, dw_a, dw_b, M, T, N, L, dU, stream);
}

int32_t ctn_row_stats(const float* x, const float* alpha, int64_t F, int32_t Ch, float* rowstat, cudaStream_t stream) {
  return run_row_stats(x, alpha, F, Ch, rowstat, stream);
}

int32_t ctn_conv1x1(const float* A, const float* W, int32_t w_is_kn, float* D, int64_t F, int32_t O, int32_t Kd,
                    int32_t K, const float* alpha_in, const float* c1, const float* c2, const double* gln_acc,
                    const float* rowstat, const float* res, double* stat_out, const float* alpha_out,
                    cudaStream_t stream) {
  GemmArgs a = {};
  a.A = A; a.W = W; a.w_is_kn = w_is_kn; a.D = D; a.F = F; a.O = O; a.Kd = Kd; a.K = K;
  a.alpha_in = alpha_in; a.c1 = c1; a.c2 = c2; a.st = make_stats(gln_acc, rowstat, K, Kd);
  a.res = res; a.stat_out = stat_out; a.alpha_out = alpha_out;
  CTN_REQUIRE((c1 == nullptr) == (c2 == nullptr), "conv1x1: c1 and c2 go together");
  CTN_REQUIRE(c1 == nullptr || gln_acc != nullptr || rowstat != nullptr, "conv1x1: norm fold needs statistics");
  return launch_gemm(a, stream);
}

int32_t ctn_conv1x1_planes(const float* A, const void* W_hi, const void* W_lo, int32_t mode, float* D, int64_t F, int32_t O,
                           int32_t Kd, int32_t K, cudaStream_t stream) {
  CTN_REQUIRE(A && W_hi && D && (W_lo || mode >= 2), "conv1x1_planes: null pointer");
  CTN_REQUIRE(mode >= 0 && mode <= 4, "conv1x1_planes: mode must be 0..4 (got %d)", mode);
  CTN_REQUIRE(Kd % 64 == 0 && O % 16 == 0, "conv1x1_planes: needs Kd %% 64 == 0 and O %% 16 == 0 (got Kd=%d O=%d)", Kd, O);
  GemmArgs a = {};
  a.A = A; a.D = D; a.F = F; a.O = O; a.Kd = Kd; a.K = K;
  a.W_hi = W_hi; a.W_lo = W_lo; a.tf32 = mode == 1 ? 1 : 0;
  a.half = mode >= 2 ? (mode == 4 ? 2 : 1) : 0;
  a.d_bf16 = mode == 3 ? 1 : 0;
  return launch_gemm(a, stream);
}

int32_t ctn_wgrad(const float* G, const float* Act, float* dW, int64_t F, int32_t O, int32_t I, int32_t K,
                  const float* alpha, const float* gamma, const float* beta, const double* gln_acc,
                  const float* rowstat, cudaStream_t stream) {
  WgradArgs a = {};
  a.G = G; a.Act = Act; a.dW = dW; a.F = F; a.O = O; a.I = I; a.K = K;
  a.alpha = alpha; a.gamma = gamma; a.beta = beta; a.st = make_stats(gln_acc, rowstat, K, I);
  CTN_REQUIRE(gamma == nullptr || (beta != nullptr && (gln_acc != nullptr || rowstat != nullptr)),
              "wgrad: norm prologue needs beta and statistics");
  return launch_wgrad(a, stream);
}

int32_t ctn_prep_normfold(const float* W, const float* gamma, const float* beta, int32_t O, int32_t I, float* Wg,
                          float* c1, float* c2, cudaStream_t stream) {
  return run_prep_normfold(W, gamma, beta, O, I, 1, 0, Wg, c1, c2, 0, 0, stream);
}

int32_t ctn_dwconv_fwd(const float* z1, const float* alpha1, const double* gln_acc1, const float* rowstat1,
                       const float* gamma1, const float* beta1, const float* Wd, int32_t M, int32_t K, int32_t H,
                       int32_t P, int32_t dilation, int32_t causal, float* z2, double* stat_out, const float* alpha2,
                       cudaStream_t stream) {
  return run_dwconv_fwd(z1, alpha1, make_stats(gln_acc1, rowstat1, K, H), gamma1, beta1, Wd, M, K, H, P, dilation,
                        causal, z2, stat_out, alpha2, stream);
}

int32_t ctn_dwconv_bwd(const float* dz2, const float* z1, const float* alpha1, const double* gln_acc1,
                       const float* rowstat1, const float* gamma1, const float* beta1, const float* Wd, int32_t M,
                       int32_t K, int32_t H, int32_t P, int32_t dilation, int32_t causal, float* dn1, float* dWd,
                       float* dgamma1, float* dbeta1, double* red1, cudaStream_t stream) {
  return run_dwconv_bwd(dz2, z1, alpha1, make_stats(gln_acc1, rowstat1, K, H), gamma1, beta1, Wd, M, K, H, P, dilation,
                        causal, dn1, dWd, dgamma1, dbeta1, red1, nullptr, 0, stream);
}

int32_t ctn_dwconv_bwd_gln_fused(const float* dn2, const float* z2, const float* alpha2, const double* gln_acc2,
                                 const float* gamma2, const double* red2, float* dalpha2, const float* z1,
                                 const float* alpha1, const double* gln_acc1, const float* gamma1, const float* beta1,
                                 const float* Wd, int32_t M, int32_t K, int32_t H, int32_t P, int32_t dilation,
                                 int32_t causal, float* dn1, float* dWd, float* dgamma1, float* dbeta1, double* red1,
                                 cudaStream_t stream) {
  CTN_REQUIRE(dn2 && z2 && alpha2 && gln_acc2 && gamma2 && red2 && dalpha2 && z1 && alpha1 && gln_acc1 && gamma1 && beta1 &&
              Wd && dn1 && dWd && dgamma1 && dbeta1, "dwconv_bwd_gln_fused: null pointer");
  return run_dwconv_bwd_gln_fused(dn2, z2, alpha2, make_stats(gln_acc2, nullptr, K, H), gamma2, red2, dalpha2, z1, alpha1,
                                  make_stats(gln_acc1, nullptr, K, H), gamma1, beta1, Wd, M, K, H, P, dilation, causal, dn1,
                                  dWd, dgamma1, dbeta1, red1, nullptr, 0, stream);
}

int32_t ctn_norm_bwd_reduce(const float* dn, const float* z, const float* alpha, const double* gln_acc,
                            const float* rowstat, const float* gamma, int32_t M, int32_t K, int32_t Ch, float* dgamma,
                            float* dbeta, double* red, cudaStream_t stream) {
  return run_norm_bwd_reduce(dn, z, alpha, make_stats(gln_acc, rowstat, K, Ch), gamma, M, K, Ch, dgamma, dbeta, red,
                             nullptr, 0, stream);
}

int32_t ctn_norm_bwd_apply(float* dn, const float* z, const float* alpha, const double* gln_acc, const float* rowstat,
                           const float* gamma, const double* red, int32_t M, int32_t K, int32_t Ch, float* dalpha,
                           cudaStream_t stream) {
  CTN_REQUIRE(rowstat != nullptr || (gln_acc != nullptr && red != nullptr), "norm_bwd_apply: gLN needs gln_acc and red");
  return run_norm_bwd_apply(dn, z, alpha, make_stats(gln_acc, rowstat, K, Ch), gamma, red, M, K, Ch, dalpha, stream);
}

int32_t ctn_batchnorm_stats(const float* z, const float* alpha, const float* weight, const float* bias, float* running_mean,
                            float* running_var, int64_t F, int32_t C, int32_t batch_stats, void* scratch, float* mean,
                            float* rstd, float* s, float* t, cudaStream_t stream) {
  CTN_REQUIRE(z && weight && bias && scratch && mean && rstd && s && t, "batchnorm_stats: null pointer");
  double* acc = reinterpret_cast<double*>(scratch);
  float* mode = reinterpret_cast<float*>(acc + 2 * (int64_t)C);
  if (batch_stats) CTN_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * (size_t)C, stream));
  return run_bn_forward_stats(z, alpha, weight, bias, running_mean, running_var, F, C, batch_stats ? 1 : 0, acc, mean, rstd,
                              s, t, mode, stream);
}

int32_t ctn_batchnorm_bwd(float* dn, const float* z, const float* alpha, const float* A, const float* Bsum,
                          const float* mean, const float* rstd, const float* s, int32_t batch_stats, int64_t F, int32_t C,
                          float* dweight, float* dbias, float* dalpha, void* scratch, cudaStream_t stream) {
  CTN_REQUIRE(dn && z && A && Bsum && mean && rstd && s && dweight && dbias && scratch, "batchnorm_bwd: null pointer");
  float* ca = reinterpret_cast<float*>(scratch);
  float* cq = ca + C;
  CTN_TRY(run_bn_bwd_finalize(A, Bsum, mean, rstd, nullptr, batch_stats ? 1 : 0, F, C, dweight, dbias, ca, cq, stream));
  return run_bn_bwd_apply(dn, z, alpha, s, mean, ca, cq, F, C, dalpha, stream);
}

int32_t ctn_decoder_fwd(const float* score, const float* w, const float* V, int32_t M, int32_t K, int32_t C, int32_t N,
                        int32_t L, int32_t T, int32_t softmax, float* est, cudaStream_t stream) {
  CTN_REQUIRE(T >= (K - 1) * (L / 2) + L, "decoder: T=%d shorter than the overlap-added length", T);
  return run_decoder_fwd(score, w, V, M, K, C, N, L, T, softmax, est, stream);
}

int32_t ctn_decoder_bwd(const float* d_est, const float* score, const float* w, const float* V, int32_t M, int32_t K,
                        int32_t C, int32_t N, int32_t L, int32_t T, int32_t softmax, float* d_score, float* d_w,
                        float* dV, cudaStream_t stream) {
  return run_decoder_bwd(d_est, score, w, V, M, K, C, N, L, T, softmax, d_score, d_w, dV, stream);
}

}  // extern "C"
