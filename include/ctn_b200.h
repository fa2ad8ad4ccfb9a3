/* ctn_b200.h — C ABI of the B200-native Conv-TasNet hot path (libctn_b200.so).
 *
 * The reference (OfekCohen1/Conv-TasNet) has no FFI: its boundary is the Python API of
 * src/conv_tasnet.py, src/pit_criterion.py and src/utils.py.  This header is the
 * C-ABI layer underneath the drop-in Python modules of conv_tasnet_b200/: plain pointers and
 * sizes, no torch types, caller-owned memory, explicit stream, no global mutable state other
 * than a thread-local error string.  Every entry point names the reference interface it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; fp32 unless stated
 *   - activations inside the library are channels-last: [M, K, Ch] (frame-major), K = (T-L)/(L/2)+1
 *   - return value: 0 = ok, nonzero = error (ctn_last_error() gives the message).  There is no
 *     CPU fallback: without a CUDA device every compute entry point returns an error.
 *   - parameters live in ONE flat fp32 buffer in reference state_dict order; ctn_param_layout()
 *     is the single source of truth for the offsets (each tensor 16-byte aligned).
 */
#ifndef CTN_B200_H
#define CTN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st* cudaStream_t;
#endif

/* Hyper-parameters of ConvTasNet.__init__ (src/conv_tasnet.py:14-30). */
typedef struct ctn_config {
  int32_t N, L, B, H, P, X, R, C;
  int32_t norm_type;      /* 0 = gLN, 1 = cLN, 2 = BatchNorm (chose_norm, src/conv_tasnet.py:298-309) */
  int32_t causal;         /* 0 / 1                          (src/conv_tasnet.py:182,264-269)        */
  int32_t mask_nonlinear; /* 0 = relu, 1 = softmax          (src/conv_tasnet.py:209-214)            */
} ctn_config;

enum { CTN_NORM_GLN = 0, CTN_NORM_CLN = 1, CTN_NORM_BN = 2, CTN_MASK_RELU = 0, CTN_MASK_SOFTMAX = 1 };

int32_t ctn_version(void);
const char* ctn_last_error(void);
/* number of CUDA kernels this library has launched in this process (monotonic; for bench.py's gpu_launches) */
int64_t ctn_launch_count(void);
/* debug (CTN_TIMING=1 in the environment, eager launches only): prints per-kernel time measured in place between the
 * launches recorded since the previous call, then clears the records; reset_only != 0 just clears.  No-op otherwise. */
int32_t ctn_timing_report(int32_t reset_only);

/* ---- parameter / workspace geometry (host only, no GPU needed) -------------------------- */
/* number of tensors in state_dict order (src/conv_tasnet.py state_dict(): 4 + 9*R*X + 2) */
int32_t ctn_param_tensors(const ctn_config* cfg);
/* total floats of the flat parameter (and gradient) buffer, including alignment padding */
int64_t ctn_param_floats(const ctn_config* cfg);
/* offsets[i], numels[i] for tensor i in state_dict order; returns 0 or error */
int32_t ctn_param_layout(const ctn_config* cfg, int64_t* offsets, int64_t* numels, int32_t n);
/* frames K for T samples (Encoder, src/conv_tasnet.py:113) */
int32_t ctn_num_frames(const ctn_config* cfg, int32_t T);
/* bytes of scratch the model calls need; `training` == 1 keeps the activation stash for backward (0 and 2: inference) */
int64_t ctn_workspace_bytes(const ctn_config* cfg, int32_t M, int32_t T, int32_t training);

/* ---- whole-path entry points ------------------------------------------------------------ */
/* ConvTasNet.forward (src/conv_tasnet.py:45-60): mixture [M,T] -> est [M,C,T] (right zero-padded).
 * training != 0 additionally leaves the stash in `workspace` for ctn_model_backward and runs the 1x1 convs with the
 * TF32x3 operand split (fp32-class forward: needed for gradient parity); training == 0 uses the bf16x3
 * split (outputs within 2e-5 of the training path, budget 1e-4).
 * training == 2: reduced-precision inference ("bf16 forward", BASELINE configs[2]; the reference has no such path —
 * src/utils.py:40, src/pit_criterion.py:72 — so the budget is the north star's 2e-2 against the fp32 result): the two
 * H-wide activations of every TemporalBlock (z1, z2) are STORED as bf16 and every 1x1 conv runs one bf16 x bf16 MMA per
 * k-step with fp32 accumulation (frame-major tcgen05 kernel); the B-wide residual stream, the normalisation statistics,
 * the encoder / decoder and all I/O stay fp32.  Needs N, B, H multiples of 64 and gLN or cLN. */
int32_t ctn_model_forward(const ctn_config* cfg, const float* params, const float* mixture,
                          int32_t M, int32_t T, float* est, void* workspace, int64_t workspace_bytes,
                          int32_t training, cudaStream_t stream);
/* The BatchNorm branch of chose_norm (src/conv_tasnet.py:306-309: nn.BatchNorm1d(H) after both PReLUs of every
 * TemporalBlock; the front norm stays cLN, :172).  norm_state holds the running statistics, per block
 * [running_mean1 H | running_var1 H | running_mean2 H | running_var2 H] (ctn_norm_state_floats floats; 0 unless BN).
 * batch_stats != 0 is nn.Module.train(): normalise with the batch statistics over (M, K) and update norm_state
 * (momentum 0.1, unbiased variance, eps 1e-5 — the nn.BatchNorm1d defaults); batch_stats == 0 is .eval(): normalise
 * with norm_state.  `training` keeps its meaning (stash for ctn_model_backward, which needs nothing extra: it
 * differentiates through the batch statistics when the forward used them).  BatchNorm weight / bias take the place
 * of gamma / beta in the parameter layout (same offsets, [H] each). */
int64_t ctn_norm_state_floats(const ctn_config* cfg);
int32_t ctn_model_forward_bn(const ctn_config* cfg, const float* params, float* norm_state, const float* mixture,
                             int32_t M, int32_t T, float* est, void* workspace, int64_t workspace_bytes,
                             int32_t training, int32_t batch_stats, cudaStream_t stream);
/* autograd of the above (replaces torch autograd over src/conv_tasnet.py:45-60):
 * d_est [M,C,T] -> grads (flat, same layout as params).  accumulate == 0 overwrites grads. */
int32_t ctn_model_backward(const ctn_config* cfg, const float* params, const float* mixture,
                           int32_t M, int32_t T, const float* d_est, float* grads, void* workspace,
                           int64_t workspace_bytes, int32_t accumulate, cudaStream_t stream);

/* The same backward cut into R+2 stages so a data-parallel caller can all-reduce each finished gradient
 * slice while the next stage computes (replaces nn.DataParallel's reduce_add_coalesced, src/train.py:84):
 *   stage 0      zero-init, decoder, mask conv        -> slice [Wm, V]
 *   stage 1..R   repeat R-stage (X blocks, last first) -> that repeat's slice
 *   stage R+1    bottleneck, first cLN, encoder        -> slice [U .. Wb]
 * Stages must run in ascending order on one stream.  ctn_grad_bucket gives each stage's flat slice. */
int32_t ctn_model_backward_stage(const ctn_config* cfg, const float* params, const float* mixture,
                                 int32_t M, int32_t T, const float* d_est, float* grads, void* workspace,
                                 int64_t workspace_bytes, int32_t accumulate, int32_t stage, cudaStream_t stream);
int32_t ctn_grad_bucket(const ctn_config* cfg, int32_t stage, int64_t* offset, int64_t* count);

/* cal_loss / cal_si_snr_with_pit / reorder_source (src/pit_criterion.py:12-99).
 *   source [B,C,T]; est [B,C,T] is masked IN PLACE beyond lengths[b] (pit_criterion.py:38);
 *   lengths int64 [B]; outputs: loss [1], max_snr [B], idx int64 [B] (index into the lexicographic
 *   permutation table), reorder [B,C,T] (may be NULL), coef [B,C,4] = (c_e, c_s, c_0, j) for backward.
 *   pit_ws: ctn_pit_workspace_bytes(B,C) bytes of scratch. */
int64_t ctn_pit_workspace_bytes(int32_t B, int32_t C);
int32_t ctn_pit_forward(const float* source, float* est, const int64_t* lengths, int32_t B, int32_t C,
                        int32_t T, float* loss, float* max_snr, int64_t* idx, float* reorder,
                        float* coef, void* pit_ws, cudaStream_t stream);
/* d loss / d est: d_est[b,i,t] = grad_loss * mask * (c_e*est + c_s*source[b,j] + c_0) */
int32_t ctn_pit_backward(const float* source, const float* est_masked, const int64_t* lengths,
                         const float* coef, const float* grad_loss, int32_t B, int32_t C, int32_t T,
                         float* d_est, cudaStream_t stream);
/* reorder_source (src/pit_criterion.py:80-99): out[b,c] = source[b, perms[idx[b]][c]] */
int32_t ctn_reorder_source(const float* source, const int64_t* idx, int32_t B, int32_t C, int64_t inner,
                           float* out, cudaStream_t stream);

/* cal_SISNRi / cal_SISNR (src/evaluate.py:94-130) on the padded evaluation batch, in one pass on the device:
 *   source [B,C,T], reordered_est [B,C,T] (cal_loss's reorder_estimate_source), mixture [B,T], lengths [B]
 *   -> sisnri [B] = mean_c( SI-SNR(source_c, est_c) - SI-SNR(source_c, mixture) ) over t < lengths[b]
 *   (the reference hard-codes C = 2, src/evaluate.py:103-110; this is the same average for any C).
 *   sisnr_est [B,C] (optional, may be NULL): cal_SISNR(source_c, est_c) itself (src/evaluate.py:114-130).
 *   ws: ctn_sisnri_workspace_bytes(B,C) bytes of scratch. */
int64_t ctn_sisnri_workspace_bytes(int32_t B, int32_t C);
int32_t ctn_sisnri(const float* source, const float* reordered_est, const float* mixture, const int64_t* lengths,
                   int32_t B, int32_t C, int32_t T, float* sisnri, float* sisnr_est, void* ws, cudaStream_t stream);

/* utils.overlap_and_add (src/utils.py:9-47): signal [outer, frames, frame_length] -> [outer, out_len] */
int32_t ctn_overlap_and_add(const float* signal, int64_t outer, int32_t frames, int32_t frame_length,
                            int32_t frame_step, float* out, cudaStream_t stream);

/* ---- batch assembly ---------------------------------------------------------- */
/* _collate_fn + pad_list + .cuda() (src/data.py:159-183,322-331; src/solver.py:184-187) on the device.
 * packed_mix: the B mixtures back to back (offsets[b] .. offsets[b+1] samples, offsets [B+1] on the device);
 * packed_src: the B sources back to back in the loader's [T_b, C] layout, or NULL (evaluation collate, :239-260).
 * -> padded_mixture [B,T] and padded_source [B,C,T] zero padded (pad_value 0), lengths [B] (T >= every length). */
int32_t ctn_assemble_batch(const float* packed_mix, const float* packed_src, const int64_t* offsets,
                           int32_t B, int32_t C, int32_t T, float* padded_mixture, float* padded_source,
                           int64_t* lengths, cudaStream_t stream);
/* utils.remove_pad (src/utils.py:50-67) without the per-item copies: item b of inputs [B,C,T] (C = 1 for [B,T]) is
 * written as a dense [C, lengths[b]] block starting at float out_offsets[b] * C of `packed` (out_offsets = exclusive
 * prefix sum of the lengths); the caller then does ONE device->host copy of `packed`. */
int32_t ctn_pack_valid(const float* inputs, const int64_t* lengths, const int64_t* out_offsets, int32_t B,
                       int32_t C, int32_t T, float* packed, cudaStream_t stream);

/* ---- data-parallel gradient exchange over NVLink peer memory (replaces nn.DataParallel's gather of the replicas'
 * gradients, src/train.py:83-85 / src/solver.py:194) ------------------------------------------------------------
 * Every rank of the node allocates its exchange buffer with ctn_peer_alloc (a CUDA allocation of its own, zero-filled),
 * exports it (64-byte CUDA-IPC handle, carried to the other processes by the caller, e.g. torch.distributed
 * all_gather_object), and maps every other rank's buffer with ctn_peer_open.  Layout of a buffer, by convention of the
 * caller: a 256-byte flag block (uint32, zero at start: barrier slots, epoch, block counter, error word [18]) followed
 * by the flat fp32 gradient buffer. */
int32_t ctn_peer_alloc(int64_t bytes, void** ptr);
int32_t ctn_peer_free(void* ptr);
int32_t ctn_peer_export(const void* ptr, uint8_t* handle64);
int32_t ctn_peer_open(const uint8_t* handle64, void** ptr);
int32_t ctn_peer_close(void* ptr);
/* ONE kernel: wait until every rank's gradients are complete, sum floats [offset, offset + count) of the `world` buffers
 * in rank order, multiply by `scale` and leave the result in EVERY rank's buffer (each rank reduces 1/world of the range
 * with peer loads and writes it to all ranks with peer stores), wait until every rank's share has landed.  bufs_host /
 * flags_host: HOST arrays of `world` device pointers (entry `rank` is this rank's own memory, the others are the
 * ctn_peer_open mappings).  Every rank must launch it once per exchange with the same offset / count / world (2..8);
 * offset and count multiples of 4.  Stream-ordered, no host synchronisation, capturable into a CUDA graph. */
int32_t ctn_peer_all_reduce(float* const* bufs_host, uint32_t* const* flags_host, int32_t rank, int32_t world,
                            int64_t offset, int64_t count, float scale, cudaStream_t stream);

/* ---- step tail (solver.py:192-196: clip_grad_norm_ + Adam) on the flat buffers ----------- */
/* total L2 norm -> norm_out[0]; grads *= min(1, max_norm/(norm+1e-6)) like torch clip_grad_norm_.
 * scratch: >= 8 * 1024 bytes */
int32_t ctn_clip_grad_norm(float* grads, int64_t n, float max_norm, float* norm_out, void* scratch,
                           cudaStream_t stream);
/* torch.optim.Adam (no amsgrad), step count read from the device counter step_dev[0] (incremented here) */
int32_t ctn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      float lr, float beta1, float beta2, float eps, float weight_decay,
                      int64_t* step_dev, cudaStream_t stream);

/* ---- single kernels, exported for the unit parity tests (tests/test_kernels_gpu.py) ------ */
/* Encoder (src/conv_tasnet.py:108-121): mix [M,T], U [N,L] -> w [M,K,N] */
int32_t ctn_encoder_fwd(const float* mix, const float* U, int32_t M, int32_t T, int32_t N, int32_t L,
                        float* w, cudaStream_t stream);
/* dU [N,L] += sum_f (dw_a + dw_b)[f,n] * [w>0] * frame[f,l]  (dw_b may be NULL) */
int32_t ctn_encoder_bwd(const float* mix, const float* w, const float* dw_a, const float* dw_b, int32_t M,
                        int32_t T, int32_t N, int32_t L, float* dU, cudaStream_t stream);
/* cLN statistics (src/conv_tasnet.py:332-333) of prelu(x, alpha) (alpha NULL = identity):
 * x [F,Ch] -> rowstat [F,2] = (mean, 1/sqrt(var+eps)) */
int32_t ctn_row_stats(const float* x, const float* alpha, int64_t F, int32_t Ch, float* rowstat,
                      cudaStream_t stream);
/* 1x1 conv as GEMM: D[F,O] = epi( pro(A[F,Kd]) . W ), W is [O,Kd] (w_is_kn == 0) or [Kd,O] (w_is_kn != 0).
 *   pro: prelu(alpha_in) if alpha_in != NULL
 *   epi: if c1 != NULL:  D = r_f*acc + c1[o] - mu_f*r_f*c2[o]   (norm folded), stats from
 *        gln_acc [M,2] doubles (sum, sumsq; count = K*Kd) when rowstat == NULL else rowstat [F,2]
 *        if res != NULL: D += res[F,O]
 *        if stat_out != NULL: stat_out[m] += (sum, sumsq) of prelu(acc, alpha_out) per sample m = f / K */
int32_t ctn_conv1x1(const float* A, const float* W, int32_t w_is_kn, float* D, int64_t F, int32_t O,
                    int32_t Kd, int32_t K, const float* alpha_in, const float* c1, const float* c2,
                    const double* gln_acc, const float* rowstat, const float* res, double* stat_out,
                    const float* alpha_out, cudaStream_t stream);
/* The same 1x1 convolution with the weight operand already split into the planes the tensor-core kernels read
 * ([O, Kd] row-major), as the whole-model path holds them (one split per step instead of one per call).  No prologue /
 * epilogue.  mode: 0 = bf16x3 (bf16 planes hi = bf16(W), lo = bf16(W - hi)); 1 = TF32x3 (fp32 planes hi = tf32(W),
 * lo = W - hi); 2..4 = the single-bf16 MMA of the reduced-precision inference path (W_hi bf16 only, W_lo unused):
 * 2 = A fp32 -> D fp32, 3 = A fp32 -> D stored as bf16, 4 = A stored as bf16 -> D fp32.
 * Replaces nn.Conv1d(kernel_size=1, bias=False) calls, src/conv_tasnet.py:223,262,191. */
int32_t ctn_conv1x1_planes(const float* A, const void* W_hi, const void* W_lo, int32_t mode, float* D, int64_t F,
                           int32_t O, int32_t Kd, int32_t K, cudaStream_t stream);

/* weight gradient: dW[O,I] += sum_f G[f,o] * act(f,i), act = Act or gamma*(prelu(Act,alpha)-mu)*r+beta when
 * gamma != NULL (stats as above, count = K*I) */
int32_t ctn_wgrad(const float* G, const float* Act, float* dW, int64_t F, int32_t O, int32_t I, int32_t K,
                  const float* alpha, const float* gamma, const float* beta, const double* gln_acc,
                  const float* rowstat, cudaStream_t stream);
/* norm fold constants for one 1x1 conv: Wg = W*gamma, c1 = W@beta, c2 = rowsum(Wg); W [O,I] */
int32_t ctn_prep_normfold(const float* W, const float* gamma, const float* beta, int32_t O, int32_t I,
                          float* Wg, float* c1, float* c2, cudaStream_t stream);
/* norm + dilated depthwise conv (+Chomp1d) (src/conv_tasnet.py:253-260,295):
 * z1 [M,K,H] -> z2 [M,K,H]; stat_out (nullable) gets the gLN sums of prelu(z2, alpha2) */
int32_t ctn_dwconv_fwd(const float* z1, const float* alpha1, const double* gln_acc1, const float* rowstat1,
                       const float* gamma1, const float* beta1, const float* Wd, int32_t M, int32_t K,
                       int32_t H, int32_t P, int32_t dilation, int32_t causal, float* z2, double* stat_out,
                       const float* alpha2, cudaStream_t stream);
/* backward of the above: dz2 -> dn1 [M,K,H]; dWd [H,P], dgamma1 [H], dbeta1 [H] accumulate;
 * red1 [M,2] doubles accumulate (sum gh, sum gh*yhat) for the gLN backward */
int32_t ctn_dwconv_bwd(const float* dz2, const float* z1, const float* alpha1, const double* gln_acc1,
                       const float* rowstat1, const float* gamma1, const float* beta1, const float* Wd,
                       int32_t M, int32_t K, int32_t H, int32_t P, int32_t dilation, int32_t causal,
                       float* dn1, float* dWd, float* dgamma1, float* dbeta1, double* red1,
                       cudaStream_t stream);
/* The same backward for gLN blocks with the backward of norm2 (+ its PReLU) applied on load (what ctn_norm_bwd_apply
 * would do in a separate pass): takes dn2 (gradient w.r.t. the NORMALISED depthwise output), z2, the gLN statistics of
 * prelu(z2), gamma2 and the per-sample sums red2 that ctn_norm_bwd_reduce left; dalpha2 [1] accumulates.
 * (autograd of src/conv_tasnet.py:253-260,295,358-360 in one kernel.) */
int32_t ctn_dwconv_bwd_gln_fused(const float* dn2, const float* z2, const float* alpha2, const double* gln_acc2,
                                 const float* gamma2, const double* red2, float* dalpha2, const float* z1,
                                 const float* alpha1, const double* gln_acc1, const float* gamma1,
                                 const float* beta1, const float* Wd, int32_t M, int32_t K, int32_t H, int32_t P,
                                 int32_t dilation, int32_t causal, float* dn1, float* dWd, float* dgamma1,
                                 float* dbeta1, double* red1, cudaStream_t stream);
/* norm backward, reduction pass: dgamma [Ch], dbeta [Ch], red [M,2] accumulate */
int32_t ctn_norm_bwd_reduce(const float* dn, const float* z, const float* alpha, const double* gln_acc,
                            const float* rowstat, const float* gamma, int32_t M, int32_t K, int32_t Ch,
                            float* dgamma, float* dbeta, double* red, cudaStream_t stream);
/* norm (+PReLU) backward, apply pass: dn -> dz IN PLACE; dalpha [1] accumulates (alpha NULL = no PReLU) */
int32_t ctn_norm_bwd_apply(float* dn, const float* z, const float* alpha, const double* gln_acc,
                           const float* rowstat, const float* gamma, const double* red, int32_t M,
                           int32_t K, int32_t Ch, float* dalpha, cudaStream_t stream);
/* BatchNorm branch (chose_norm fall-through, src/conv_tasnet.py:306-309), statistics: z [F,C] (F = M*K frames),
 * p = prelu(z, alpha) -> per-channel mean, rstd (eps 1e-5) and the affine map n = s*p + t (s = weight*rstd,
 * t = bias - mean*s).  batch_stats != 0: statistics of this batch over all F frames, running_mean / running_var (may be
 * NULL) updated with momentum 0.1 and the unbiased variance; batch_stats == 0: the running statistics.
 * scratch: >= 16*C + 16 bytes. */
int32_t ctn_batchnorm_stats(const float* z, const float* alpha, const float* weight, const float* bias,
                            float* running_mean, float* running_var, int64_t F, int32_t C, int32_t batch_stats,
                            void* scratch, float* mean, float* rstd, float* s, float* t, cudaStream_t stream);
/* ... backward: dn (gradient w.r.t. n) -> dz IN PLACE through the norm (and the batch statistics when batch_stats != 0)
 * and the PReLU; A = sum_f dn, Bsum = sum_f dn*p per channel (ctn_norm_bwd_reduce with gln_acc = rowstat = NULL returns
 * them as dbeta / dgamma); dweight, dbias [C] and dalpha [1] accumulate.  scratch: >= 8*C bytes. */
int32_t ctn_batchnorm_bwd(float* dn, const float* z, const float* alpha, const float* A, const float* Bsum,
                          const float* mean, const float* rstd, const float* s, int32_t batch_stats, int64_t F,
                          int32_t C, float* dweight, float* dbias, float* dalpha, void* scratch, cudaStream_t stream);
/* mask nonlinearity * w -> basis -> overlap-add -> pad (src/conv_tasnet.py:208-214,140-145,57-59):
 * score [M,K,C*N], w [M,K,N], V [L,N] -> est [M,C,T] */
int32_t ctn_decoder_fwd(const float* score, const float* w, const float* V, int32_t M, int32_t K, int32_t C,
                        int32_t N, int32_t L, int32_t T, int32_t softmax, float* est, cudaStream_t stream);
/* d_est [M,C,T] -> d_score [M,K,C*N], d_w [M,K,N] (overwritten), dV [L,N] (accumulates) */
int32_t ctn_decoder_bwd(const float* d_est, const float* score, const float* w, const float* V, int32_t M,
                        int32_t K, int32_t C, int32_t N, int32_t L, int32_t T, int32_t softmax,
                        float* d_score, float* d_w, float* dV, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CTN_B200_H */
