// common.cuh — shared device helpers for the sm_100a Conv-TasNet kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ctn_b200.h"

#define CTN_EPS 1e-8f

namespace ctn {

// ---- host-side error plumbing -----------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> 0 / 1 (+ message)
// library-owned device scratch for the STANDALONE single-kernel entry points only (slot 0: weight planes, 1: partials);
// the whole-model path never allocates: it uses the caller's workspace
int lib_scratch(size_t bytes, void** out, int slot);

#define CTN_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::ctn::set_error(__VA_ARGS__);           \
      return 1;                                \
    }                                          \
  } while (0)

#define CTN_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::ctn::set_error("%s failed: %s", #call, cudaGetErrorString(e__));            \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

#define CTN_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != 0) return rc__;  \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with
// pdl_launch_dependents() (lets the NEXT kernel's CTAs be scheduled as soon as all of this kernel's CTAs are resident,
// e.g. on the SMs the last partial wave leaves idle) and pdl_wait() (blocks until the PREVIOUS kernel has completed and
// its writes are visible) before it touches any global data.  CTN_NO_PDL=1 disables the launch attribute (A/B).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();
// CTN_TIMING=1 (debug): every launch is followed by an event on its stream; ctn_timing_report() prints per-kernel totals
// measured in place (realistic cache state, unlike a profiler's cold-cache replays).  Off: one predictable branch.
void timing_note_stream(cudaStream_t s);

template <typename... KArgs, typename... Args>
static inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // eager launches gain ~10 % from the overlap; inside a captured graph the programmatic edges measured 0.5 % slower
  // than plain kernel-to-kernel edges (6.285 vs 6.25 ms per step), so captures record ordinary launches
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &cap);
  cfg.numAttrs = (pdl_enabled() && cap == cudaStreamCaptureStatusNone) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors are picked up by check_launch()
  timing_note_stream(stream);
}

// ---- normalisation statistics -----------------------------------------------------------
// gLN: per-sample (sum, sumsq) accumulated in fp64 by the producing kernel; consumers derive
//      mean / rstd on the fly (no finalize kernel).  cLN: per-frame (mean, rstd) floats.
// Both pointers null = identity (mu 0, r 1): the BatchNorm branch, whose per-channel scale/shift travel as the
// consumers' gamma / beta vectors (batchnorm.cu).
struct NormStats {
  const double* acc;  // [M][2] or nullptr
  const float* row;   // [F][2] or nullptr
  double inv_count;   // 1 / (K * Ch) for gLN
};

__device__ __forceinline__ void load_stats(const NormStats& s, int m, int64_t f, float& mu, float& r) {
  if (s.row != nullptr) {
    float2 v = reinterpret_cast<const float2*>(s.row)[f];
    mu = v.x;
    r = v.y;
  } else if (s.acc == nullptr) {
    mu = 0.f;
    r = 1.f;
  } else {
    double S = s.acc[2 * m], S2 = s.acc[2 * m + 1];
    double mean = S * s.inv_count;
    double var = S2 * s.inv_count - mean * mean;
    var = var > 0.0 ? var : 0.0;
    mu = (float)mean;
    r = (float)(1.0 / sqrt(var + 1e-8));
  }
}

__device__ __forceinline__ float prelu(float z, float a) { return z > 0.f ? z : a * z; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of up to NV doubles; result valid in thread 0.  scratch: NV * 32 doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) scratch[i * 32 + wid] = v[i];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double x = lane < nw ? scratch[i * 32 + lane] : 0.0;
      v[i] = warp_sum(x);
    }
  }
}

// one deferred fold of per-block partial rows (see reduce_partials_kernel); up to FOLD_MAX entries per launch
struct FoldEntry {
  const float* part;
  int nb, H, P;
  float* dW;
  float* dgamma;
  float* dbeta;
};
constexpr int FOLD_MAX = 48;
struct FoldBatch {
  FoldEntry e[FOLD_MAX];
};

// ---- kernel launchers implemented across the .cu files (host) ---------------------------
struct GemmArgs {
  const float* A;  // [F, Kd]
  const float* W;  // [O, Kd] (w_is_kn = 0) or [Kd, O] (w_is_kn = 1)
  float* D;        // [F, O]
  int64_t F;
  int O, Kd, K;  // K = frames per sample (row -> sample index)
  int w_is_kn;
  const float* alpha_in;   // prologue PReLU slope (device scalar) or nullptr
  const float* c1;         // norm-fold constants or nullptr
  const float* c2;
  NormStats st;            // stats for the norm fold
  const float* res;        // residual [F, O] or nullptr
  double* stat_out;        // [M][2] accumulates stats of prelu(acc, alpha_out) or nullptr
  const float* alpha_out;
  // bf16 hi/lo planes of the weight as an [O, Kd] row-major (K-major) matrix, for the tcgen05 path; when null the
  // standalone entry points split W into a library-owned scratch first
  const void* W_hi;
  const void* W_lo;
  int tf32;  // 1: planes are fp32 (tf32 hi + exact remainder), forward precision; 0: bf16 hi/lo planes
  // reduced-precision inference (frame-major kernel only; BASELINE configs[2] "bf16 forward"): one bf16 plane per operand.
  // half = 1: A is fp32 in memory, rounded to bf16 on the way in; half = 2: A is STORED as bf16 ([F, Kd] bf16, the pointer
  // travels in `A`); d_bf16 = 1: D is stored as bf16 ([F, O] bf16 behind the `D` pointer).  W_lo is not read.
  int half;
  int d_bf16;
  // optional fused norm-backward reduction on the OUTPUT (D = dn, the gradient w.r.t. a normalised activation):
  // with yhat = (prelu(nred_z, nred_alpha) - mu) * r (stats in `st`):  dgamma[o] += sum_f D*yhat, dbeta[o] += sum_f D,
  // red[m] += (sum D*gamma, sum D*gamma*yhat).  nred_part: scratch for the un-fused fallback.
  const float* nred_z;
  const float* nred_alpha;
  const float* nred_gamma;
  float* nred_dgamma;
  float* nred_dbeta;
  double* nred_red;
  float* nred_part;
};
int launch_gemm(const GemmArgs& a, cudaStream_t s);

struct WgradArgs {
  const float* G;    // [F, O]
  const float* Act;  // [F, I]
  float* dW;         // [O, I], accumulated with atomics
  int64_t F;
  int O, I, K;
  const float* alpha;  // PReLU slope or nullptr
  const float* gamma;  // nullptr = use Act as is
  const float* beta;
  NormStats st;
};
int launch_wgrad(const WgradArgs& a, cudaStream_t s);

}  // namespace ctn
