// optim.cu — the step tail of solver.py:192-196 on the flat parameter / gradient buffers:
// global-norm clipping (torch.nn.utils.clip_grad_norm_) and Adam (torch.optim.Adam, no amsgrad).
#include "common.cuh"

namespace ctn {
namespace {

constexpr int RB = 512;  // reduction blocks

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ part,
                                                    unsigned int* __restrict__ ticket, float max_norm,
                                                    float* __restrict__ norm_out, float* __restrict__ scale_out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double scratch[32];
  __shared__ bool is_last;
  double acc[1] = {0.0};
  const int64_t n4 = n / 4;
  float s = 0.f;
  int cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    if (++cnt == 64) { acc[0] += (double)s; s = 0.f; cnt = 0; }
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  acc[0] += (double)s;
  block_sum<1>(acc, scratch);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = acc[0];
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double t[1] = {0.0};
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t[0] += ((volatile double*)part)[i];
  block_sum<1>(t, scratch);
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(t[0]);
    norm_out[0] = norm;
    const float coef = max_norm / (norm + 1e-6f);
    scale_out[0] = coef < 1.f ? coef : 1.f;
    *ticket = 0u;
  }
}

__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ g, int64_t n, const float* __restrict__ scale) {
  pdl_launch_dependents();
  pdl_wait();
  const float sc = __ldg(scale);
  if (sc == 1.f) return;
  const int64_t n4 = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(g)[i];
    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    reinterpret_cast<float4*>(g)[i] = v;
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) g[i] *= sc;
}

__global__ void step_inc_kernel(int64_t* step) {
  pdl_launch_dependents();
  pdl_wait(); step[0] += 1; }

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                   float wd, const int64_t* __restrict__ step, int vec) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_bias[2];
  if (threadIdx.x == 0) {  // the fp64 pow()s once per block, not once per thread
    const double t = (double)step[0];
    s_bias[0] = (float)(1.0 - pow((double)b1, t));
    s_bias[1] = (float)sqrt(1.0 - pow((double)b2, t));
  }
  __syncthreads();
  const float bias1 = s_bias[0], bias2_sqrt = s_bias[1];
  const float step_size = lr / bias1;
  auto update = [&](float gi, float pi, float& mi, float& vi) {
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    mi = mi + (gi - mi) * (1.f - b1);  // exp_avg.lerp_(grad, 1 - beta1)
    vi = vi * b2 + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bias2_sqrt + eps;
    return pi - step_size * (mi / denom);
  };
  const int64_t n4 = vec ? n / 4 : 0;  // 16-byte aligned buffers: four parameters per thread and access
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    p4.x = update(g4.x, p4.x, m4.x, v4.x);
    p4.y = update(g4.y, p4.y, m4.y, v4.y);
    p4.z = update(g4.z, p4.z, m4.z, v4.z);
    p4.w = update(g4.w, p4.w, m4.w, v4.w);
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    reinterpret_cast<float4*>(p)[i] = p4;
  }
  if (blockIdx.x == 0) {
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      float mi = m[i], vi = v[i];
      p[i] = update(g[i], p[i], mi, vi);
      m[i] = mi;
      v[i] = vi;
    }
  }
}

}  // namespace

int run_clip_grad_norm(float* grads, int64_t n, float max_norm, float* norm_out, void* scratch, cudaStream_t s) {
  CTN_REQUIRE(n > 0, "clip_grad_norm: empty buffer");
  double* part = reinterpret_cast<double*>(scratch);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(part + RB);
  float* scale = reinterpret_cast<float*>(ticket + 2);
  CTN_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), s));
  launch_kernel(sumsq_kernel, RB, 256, 0, s, grads, n, part, ticket, max_norm, norm_out, scale);
  CTN_TRY(check_launch("sumsq_kernel"));
  launch_kernel(scale_kernel, 592, 256, 0, s, grads, n, scale);
  return check_launch("scale_kernel");
}

int run_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                  float wd, int64_t* step_dev, cudaStream_t s) {
  launch_kernel(step_inc_kernel, 1, 1, 0, s, step_dev);
  CTN_TRY(check_launch("step_inc_kernel"));
  const int vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  launch_kernel(adam_kernel, 1184, 256, 0, s, p, g, m, v, n, lr, b1, b2, eps, wd, step_dev, vec);
  return check_launch("adam_kernel");
}

}  // namespace ctn
