// batchnorm.cu — the BatchNorm1d branch of chose_norm (src/conv_tasnet.py:306-309: any norm_type other than gLN / cLN
// builds nn.BatchNorm1d(channel_size), statistics over (M, K) per channel, eps 1e-5, momentum 0.1, affine).
//
// BatchNorm is a per-channel affine map n = s[c] * p + t[c] (p = prelu(z)), with
//   training:   s = weight * rstd_batch,  t = bias - mean_batch * s   (biased batch variance; running statistics are
//               updated with the unbiased one, torch/nn/functional.py batch_norm)
//   evaluation: s = weight / sqrt(running_var + eps),  t = bias - running_mean * s
// so the existing consumers (depthwise stencil, norm-folded 1x1 conv, weight gradients) run unchanged with identity
// NormStats and (s, t) in place of (gamma, beta).  This file adds what is specific to BN: the per-channel batch
// statistics, the running-statistics update, and the backward through the batch statistics.
#include "common.cuh"

namespace ctn {
namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// per-channel (sum, sumsq) of prelu(z) over all F frames: acc[c], acc[C + c] += ...  (fp64 atomics, acc zeroed by the caller)
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ z, const float* __restrict__ alpha,
                                                       int64_t F, int C, int cgt, double* __restrict__ acc) {
  pdl_launch_dependents();
  pdl_wait();
  const float a = alpha != nullptr ? __ldg(alpha) : 1.f;
  const int rl = 256 / cgt;                 // frames covered per pass by this block
  const int cg = threadIdx.x % cgt, rlane = threadIdx.x / cgt;
  if (rlane >= rl) return;
  for (int c = cg * 4; c < C; c += cgt * 4) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    double ds[4] = {0.0, 0.0, 0.0, 0.0}, dq[4] = {0.0, 0.0, 0.0, 0.0};
    int cnt = 0;
    for (int64_t f = (int64_t)blockIdx.x * rl + rlane; f < F; f += (int64_t)gridDim.x * rl) {
      float4 v = ld4(z + f * C + c);
      v.x = prelu(v.x, a); v.y = prelu(v.y, a); v.z = prelu(v.z, a); v.w = prelu(v.w, a);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
      if (++cnt == 64) {  // bound the fp32 chain
        ds[0] += s.x; ds[1] += s.y; ds[2] += s.z; ds[3] += s.w;
        dq[0] += q.x; dq[1] += q.y; dq[2] += q.z; dq[3] += q.w;
        s = make_float4(0.f, 0.f, 0.f, 0.f); q = s; cnt = 0;
      }
    }
    ds[0] += s.x; ds[1] += s.y; ds[2] += s.z; ds[3] += s.w;
    dq[0] += q.x; dq[1] += q.y; dq[2] += q.z; dq[3] += q.w;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(acc + c + i, ds[i]);
      atomicAdd(acc + C + c + i, dq[i]);
    }
  }
}

// per channel: batch (or running) statistics -> (mean, rstd, s, t); running statistics update in training
__global__ void __launch_bounds__(256) bn_finalize_kernel(const double* __restrict__ acc, const float* __restrict__ weight,
                                                          const float* __restrict__ bias, float* __restrict__ run_mean,
                                                          float* __restrict__ run_var, int64_t F, int C, int use_batch,
                                                          float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                          float* __restrict__ s_out, float* __restrict__ t_out,
                                                          float* __restrict__ mode_out) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0) mode_out[0] = use_batch ? 1.f : 0.f;
  if (c >= C) return;
  float mean, rstd;
  if (use_batch) {
    const double n = (double)F;
    const double m = acc[c] / n;
    double var = acc[C + c] / n - m * m;
    var = var > 0.0 ? var : 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + 1e-5));
    if (run_mean != nullptr) {  // momentum 0.1, unbiased variance (torch.nn.BatchNorm1d defaults)
      const double unb = F > 1 ? var * n / (n - 1.0) : var;
      run_mean[c] = (float)(0.9 * (double)run_mean[c] + 0.1 * m);
      run_var[c] = (float)(0.9 * (double)run_var[c] + 0.1 * unb);
    }
  } else {
    mean = run_mean[c];
    rstd = 1.f / sqrtf(run_var[c] + 1e-5f);
  }
  const float s = weight[c] * rstd;
  mean_out[c] = mean;
  rstd_out[c] = rstd;
  s_out[c] = s;
  t_out[c] = fmaf(-mean, s, bias[c]);
}

// backward coefficients.  In: A[c] = sum_f dn, Bv[c] = sum_f dn * p (folded partial rows).  With xhat = (p - mean) rstd:
//   dbias += A,  dweight += rstd (Bv - mean A),
//   training:   dp = s (dn - A/F - xhat * dweight/F) = s (dn - ca - (p - mean) cq),  ca = A/F, cq = rstd * dweight / F
//   evaluation: dp = s dn  (the statistics are constants)
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ A, const float* __restrict__ Bv,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const float* __restrict__ mode, int mode_override,
                                                              int64_t F, int C, float* __restrict__ dweight,
                                                              float* __restrict__ dbias, float* __restrict__ ca,
                                                              float* __restrict__ cq) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float a = A[c], dw = rstd[c] * (Bv[c] - mean[c] * a);
  dbias[c] += a;
  dweight[c] += dw;
  const bool batch = mode_override >= 0 ? mode_override != 0 : mode[0] != 0.f;
  ca[c] = batch ? a / (float)F : 0.f;
  cq[c] = batch ? rstd[c] * dw / (float)F : 0.f;
}

// dz = s (dn - ca - (p - mean) cq) * prelu'(z) in place over dn; dalpha += sum dp * z * [z <= 0]
constexpr int BA_TF = 8;
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(float* __restrict__ dn, const float* __restrict__ z,
                                                           const float* __restrict__ alpha, const float* __restrict__ sv,
                                                           const float* __restrict__ mean, const float* __restrict__ ca,
                                                           const float* __restrict__ cq, int64_t F, int C,
                                                           float* __restrict__ dalpha) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[32];
  const bool hasp = alpha != nullptr;
  const float a = hasp ? __ldg(alpha) : 1.f;
  const int64_t f0 = (int64_t)blockIdx.x * BA_TF;
  const int nf = (int)(F - f0 < BA_TF ? F - f0 : BA_TF);
  double acc[1] = {0.0};
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    const float4 s = ld4(sv + c), mu = ld4(mean + c), va = ld4(ca + c), vq = ld4(cq + c);
    float sa = 0.f;
    for (int i = 0; i < nf; ++i) {
      const int64_t o = (f0 + i) * C + c;
      const float4 zz = ld4(z + o), d = ld4(dn + o);
      float4 p = zz;
      if (hasp) { p.x = prelu(p.x, a); p.y = prelu(p.y, a); p.z = prelu(p.z, a); p.w = prelu(p.w, a); }
      float4 da;
      da.x = s.x * (d.x - va.x - (p.x - mu.x) * vq.x);
      da.y = s.y * (d.y - va.y - (p.y - mu.y) * vq.y);
      da.z = s.z * (d.z - va.z - (p.z - mu.z) * vq.z);
      da.w = s.w * (d.w - va.w - (p.w - mu.w) * vq.w);
      if (hasp) {
        sa += (zz.x > 0.f ? 0.f : da.x * zz.x) + (zz.y > 0.f ? 0.f : da.y * zz.y) +
              (zz.z > 0.f ? 0.f : da.z * zz.z) + (zz.w > 0.f ? 0.f : da.w * zz.w);
        da.x *= zz.x > 0.f ? 1.f : a; da.y *= zz.y > 0.f ? 1.f : a;
        da.z *= zz.z > 0.f ? 1.f : a; da.w *= zz.w > 0.f ? 1.f : a;
      }
      st4(dn + o, da);
    }
    acc[0] += (double)sa;
  }
  if (hasp) {
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) atomicAdd(dalpha, (float)acc[0]);
  }
}

}  // namespace

int64_t bn_slot_bytes(int C) {  // one (block, norm) slot of the workspace: see BnSlot in model.cu
  const int64_t c4 = (C + 3) & ~3;
  return 2 * c4 * 8 + (8 * c4 + 4) * 4;
}

// forward: statistics of prelu(z) -> slot (mean, rstd, s, t); `acc` must be zero on entry when use_batch
int run_bn_forward_stats(const float* z, const float* alpha, const float* weight, const float* bias, float* run_mean,
                         float* run_var, int64_t F, int C, int use_batch, double* acc, float* mean, float* rstd, float* sv,
                         float* tv, float* mode, cudaStream_t s) {
  CTN_REQUIRE(C % 4 == 0, "batch norm: channels must be a multiple of 4 (got %d)", C);
  CTN_REQUIRE(use_batch || (run_mean != nullptr && run_var != nullptr), "batch norm: evaluation needs running statistics");
  if (use_batch) {
    int cgt = C / 4;
    if (cgt > 256) cgt = 256;
    while (256 % cgt != 0) --cgt;  // threads per frame row; the remaining column groups are looped
    const int rl = 256 / cgt;
    int grid = cdiv(F, rl * 8);
    grid = grid < 1 ? 1 : (grid > 296 ? 296 : grid);
    launch_kernel(bn_stats_kernel, grid, 256, 0, s, z, alpha, F, C, cgt, acc);
    CTN_TRY(check_launch("bn_stats_kernel"));
  }
  launch_kernel(bn_finalize_kernel, cdiv(C, 256), 256, 0, s, (const double*)acc, weight, bias, run_mean, run_var, F, C,
                use_batch, mean, rstd, sv, tv, mode);
  return check_launch("bn_finalize_kernel");
}

int run_bn_bwd_finalize(const float* A, const float* Bv, const float* mean, const float* rstd, const float* mode,
                        int mode_override, int64_t F, int C, float* dweight, float* dbias, float* ca, float* cq,
                        cudaStream_t s) {
  launch_kernel(bn_bwd_finalize_kernel, cdiv(C, 256), 256, 0, s, A, Bv, mean, rstd, mode, mode_override, F, C, dweight,
                dbias, ca, cq);
  return check_launch("bn_bwd_finalize_kernel");
}

int run_bn_bwd_apply(float* dn, const float* z, const float* alpha, const float* sv, const float* mean, const float* ca,
                     const float* cq, int64_t F, int C, float* dalpha, cudaStream_t s) {
  CTN_REQUIRE(C % 4 == 0, "batch norm: channels must be a multiple of 4 (got %d)", C);
  int threads = ((C / 4 + 31) / 32) * 32;
  threads = threads > 256 ? 256 : threads;
  launch_kernel(bn_bwd_apply_kernel, cdiv(F, BA_TF), threads, 0, s, dn, z, alpha, sv, mean, ca, cq, F, C, dalpha);
  return check_launch("bn_bwd_apply_kernel");
}

}  // namespace ctn
