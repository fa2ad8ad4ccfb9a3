// sisnri.cu — batched SI-SNR improvement, the evaluation metric of src/evaluate.py:94-130, on the padded batch the
// evaluation loop already holds on the device (padded_source, the PIT-reordered estimate, padded_mixture, lengths):
// one streaming pass accumulates per (utterance, speaker) the eight masked moments
//   Sr, So, Sm, Srr, Soo, Smm, Sro, Srm        (r = reference, o = estimate, m = mixture; t < length only)
// in fp64, and the last block of an utterance turns them into SI-SNR(ref, est) - SI-SNR(ref, mix) averaged over the
// speakers.  With x~ = x - mean(x):  <r~,r~> = Srr - Sr^2/n,  <r~,o~> = Sro - Sr So/n,  <o~,o~> = Soo - So^2/n, and
// proj = a r~ with a = <r~,o~> / (<r~,r~> + eps):  |proj|^2 = a^2 <r~,r~>,  |noise|^2 = <o~,o~> - 2 a <r~,o~> + a^2 <r~,r~>.
// The reference does this per utterance on the host (remove_pad -> numpy): a device->host copy and C numpy passes each.
#include "common.cuh"

namespace ctn {
namespace {

constexpr int SI_THREADS = 256;
constexpr int SI_PER_THREAD = 8;
constexpr int SI_NMOM = 8;

__device__ __forceinline__ double sisnr_from_moments(double Srr, double Sr, double Soo, double So, double Sro, double n) {
  const double eps = 1e-8;
  const double Err = Srr - Sr * Sr / n, Eoo = Soo - So * So / n, Ero = Sro - Sr * So / n;
  const double a = Ero / (Err + eps);
  const double proj2 = a * a * Err;
  double noise2 = Eoo - 2.0 * a * Ero + a * a * Err;
  noise2 = noise2 > 0.0 ? noise2 : 0.0;
  const double ratio = proj2 / (noise2 + eps);
  return 10.0 * log(ratio + eps) / log(10.0);  // src/evaluate.py:128
}

// grid (chunks, C, B); mom [B][C][8] doubles, ticket [B] (zeroed by the launcher)
__global__ void __launch_bounds__(SI_THREADS) sisnri_kernel(const float* __restrict__ ref, const float* __restrict__ est,
                                                            const float* __restrict__ mix,
                                                            const int64_t* __restrict__ lengths, int C, int T,
                                                            double* __restrict__ mom, unsigned int* __restrict__ ticket,
                                                            float* __restrict__ out, float* __restrict__ sisnr_out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double scratch[SI_NMOM][SI_THREADS / 32];
  __shared__ bool is_last;
  const int b = blockIdx.z, c = blockIdx.y;
  int64_t len = lengths[b];
  len = len < 0 ? 0 : (len > T ? T : len);
  const float* r = ref + ((int64_t)b * C + c) * T;
  const float* o = est + ((int64_t)b * C + c) * T;
  const float* m = mix + (int64_t)b * T;
  double acc[SI_NMOM];
#pragma unroll
  for (int i = 0; i < SI_NMOM; ++i) acc[i] = 0.0;
  const int64_t t0 = (int64_t)blockIdx.x * SI_THREADS * SI_PER_THREAD;
#pragma unroll
  for (int it = 0; it < SI_PER_THREAD; ++it) {
    const int64_t t = t0 + (int64_t)it * SI_THREADS + threadIdx.x;
    if (t < len) {
      const double rv = r[t], ov = o[t], mv = m[t];
      acc[0] += rv; acc[1] += ov; acc[2] += mv;
      acc[3] += rv * rv; acc[4] += ov * ov; acc[5] += mv * mv;
      acc[6] += rv * ov; acc[7] += rv * mv;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < SI_NMOM; ++i) {
    const double v = warp_sum(acc[i]);
    if (lane == 0) scratch[i][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < SI_NMOM) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < SI_THREADS / 32; ++w) v += scratch[threadIdx.x][w];
    atomicAdd(mom + ((int64_t)b * C + c) * SI_NMOM + threadIdx.x, v);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int tk = atomicAdd(ticket + b, 1u);
    is_last = (tk == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();
  const double n = (double)len;
  double sum = 0.0;
  for (int cc = 0; cc < C; ++cc) {
    volatile double* M = mom + ((int64_t)b * C + cc) * SI_NMOM;
    const double s_est = sisnr_from_moments(M[3], M[0], M[4], M[1], M[6], n);
    const double s_mix = sisnr_from_moments(M[3], M[0], M[5], M[2], M[7], n);
    sum += s_est - s_mix;
    if (sisnr_out != nullptr) sisnr_out[(int64_t)b * C + cc] = (float)s_est;
  }
  out[b] = (float)(sum / C);
}

}  // namespace

int64_t sisnri_workspace_bytes(int B, int C) {
  return (int64_t)B * C * SI_NMOM * sizeof(double) + (int64_t)B * sizeof(unsigned int) + 64;
}

int run_sisnri(const float* ref, const float* est, const float* mix, const int64_t* lengths, int B, int C, int T,
               float* out, float* sisnr_out, void* ws, cudaStream_t s) {
  CTN_REQUIRE(B >= 1 && B <= 65535 && C >= 1 && C <= 65535 && T >= 1, "sisnri: bad shape B=%d C=%d T=%d", B, C, T);
  CTN_CUDA(cudaMemsetAsync(ws, 0, (size_t)sisnri_workspace_bytes(B, C), s));
  double* mom = reinterpret_cast<double*>(ws);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(mom + (int64_t)B * C * SI_NMOM);
  const int chunks = cdiv(T, SI_THREADS * SI_PER_THREAD);
  launch_kernel(sisnri_kernel, dim3(chunks, C, B), SI_THREADS, 0, s, ref, est, mix, lengths, C, T, mom, ticket, out, sisnr_out);
  return check_launch("sisnri_kernel");
}

}  // namespace ctn
