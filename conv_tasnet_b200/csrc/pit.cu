// pit.cu — utterance-level permutation-invariant SI-SNR (src/pit_criterion.py:12-99) as one streaming
// moments pass + a "last block" finalisation (pairwise SI-SNR, argmax over the C! lexicographic
// permutations, loss), the elementwise backward, and reorder_source.
#include "common.cuh"

namespace ctn {
namespace {

constexpr int PIT_MAXC = 4;
constexpr int PIT_NMOM = 4 * PIT_MAXC + PIT_MAXC * PIT_MAXC + PIT_MAXC;  // Se, See, Ss_all, Ss, Sss, Ses
constexpr int PIT_THREADS = 256;
constexpr int PIT_PER_THREAD = 2;

// moment slots for sample b (doubles):
//   [0,C)        Se[i]    masked sum of est
//   [C,2C)       See[i]   masked sum of est^2
//   [2C,3C)      Sa[j]    un-masked sum of source (target mean quirk, pit_criterion.py:42)
//   [3C,4C)      Ss[j]    masked sum of source
//   [4C,5C)      Sss[j]   masked sum of source^2
//   [5C,5C+C*C)  Ses[i][j]
struct PitWs {
  double* mom;            // [B][PIT_NMOM]
  unsigned int* ticket;   // [B] blocks finished per sample
  unsigned int* done;     // [1] samples finished
};

__device__ void unrank_perm(int idx, int C, int* perm) {
  // idx-th permutation of range(C) in lexicographic order (itertools.permutations, pit_criterion.py:67)
  int fact = 1;
  for (int i = 2; i < C; ++i) fact *= i;  // (C-1)!
  bool used[PIT_MAXC] = {false, false, false, false};
  for (int i = 0; i < C; ++i) {
    const int q = idx / fact;
    idx -= q * fact;
    int cnt = -1, c = 0;
    for (; c < C; ++c) {
      if (!used[c]) ++cnt;
      if (cnt == q) break;
    }
    used[c] = true;
    perm[i] = c;
    if (C - 1 - i > 0) fact /= (C - 1 - i);
  }
}

// templated on the number of speakers: every moment index is a compile-time constant (a run-time C put the 14..36
// fp64 accumulators in local memory), and the block reduces all moments with two barriers instead of two per moment
template <int C>
__global__ void __launch_bounds__(PIT_THREADS) pit_moments_kernel(const float* __restrict__ src, float* __restrict__ est,
                                                                  const int64_t* __restrict__ lengths, int B, int,
                                                                  int T, PitWs ws, float* __restrict__ loss,
                                                                  float* __restrict__ max_snr, int64_t* __restrict__ idx_out,
                                                                  float* __restrict__ coef) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int NM = 5 * C + C * C;
  __shared__ double scratch[NM][PIT_THREADS / 32];
  __shared__ bool is_last;
  const int b = blockIdx.y;
  int64_t len = lengths[b];
  len = len < 0 ? 0 : (len > T ? T : len);
  const int64_t t0 = (int64_t)blockIdx.x * PIT_THREADS * PIT_PER_THREAD;
  double mom[NM];
#pragma unroll
  for (int i = 0; i < NM; ++i) mom[i] = 0.0;
  for (int it = 0; it < PIT_PER_THREAD; ++it) {
    const int64_t t = t0 + (int64_t)it * PIT_THREADS + threadIdx.x;
    if (t >= T) break;
    const bool in = t < len;
    double e[PIT_MAXC], s[PIT_MAXC];
#pragma unroll
    for (int c = 0; c < PIT_MAXC; ++c) {
      e[c] = 0.0;
      s[c] = 0.0;
      if (c < C) {
        const int64_t o = ((int64_t)b * C + c) * T + t;
        const float sv = src[o];
        mom[2 * C + c] += (double)sv;
        if (in) {
          e[c] = (double)est[o];
          s[c] = (double)sv;
        } else {
          est[o] = 0.f;  // estimate_source *= mask, in place (pit_criterion.py:38)
        }
      }
    }
    if (in) {
#pragma unroll
      for (int i = 0; i < PIT_MAXC; ++i) {
        if (i < C) {
          mom[i] += e[i];
          mom[C + i] += e[i] * e[i];
          mom[3 * C + i] += s[i];
          mom[4 * C + i] += s[i] * s[i];
#pragma unroll
          for (int j = 0; j < PIT_MAXC; ++j)
            if (j < C) mom[5 * C + i * C + j] += e[i] * s[j];
        }
      }
    }
  }
  {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NM; ++i) {
      const double v = warp_sum(mom[i]);
      if (lane == 0) scratch[i][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < NM) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < PIT_THREADS / 32; ++w) v += scratch[threadIdx.x][w];
      atomicAdd(ws.mom + (int64_t)b * PIT_NMOM + threadIdx.x, v);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int tk = atomicAdd(ws.ticket + b, 1u);
    is_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();

  // ---- finalise sample b (one thread; C <= 4) ----
  volatile double* M = ws.mom + (int64_t)b * PIT_NMOM;
  const double n = (double)len;
  const double eps = 1e-8;
  double mue[PIT_MAXC], mus[PIT_MAXC], Eee[PIT_MAXC], Ess[PIT_MAXC];
  for (int c = 0; c < C; ++c) {
    mue[c] = M[c] / n;
    mus[c] = M[2 * C + c] / n;
    Eee[c] = M[C + c] - 2.0 * mue[c] * M[c] + n * mue[c] * mue[c];
    Ess[c] = M[4 * C + c] - 2.0 * mus[c] * M[3 * C + c] + n * mus[c] * mus[c];
  }
  double snr[PIT_MAXC][PIT_MAXC], dotm[PIT_MAXC][PIT_MAXC];
  for (int i = 0; i < C; ++i)
    for (int j = 0; j < C; ++j) {
      const double dot = M[5 * C + i * C + j] - mue[i] * M[3 * C + j] - mus[j] * M[i] + n * mue[i] * mus[j];
      const double E = Ess[j] + eps;
      const double a = dot / E;
      const double P = a * a * Ess[j];
      const double Q = Eee[i] - 2.0 * a * dot + a * a * Ess[j];
      dotm[i][j] = dot;
      // the reference works in fp32: round the ratio and the log like it does (pit_criterion.py:62-63)
      const float ratio = (float)(P / (Q + eps));
      snr[i][j] = (double)(10.f * log10f(ratio + 1e-8f));
    }
  int nperm = 1;
  for (int i = 2; i <= C; ++i) nperm *= i;
  int best = 0;
  float best_v = 0.f;
  for (int p = 0; p < nperm; ++p) {
    int perm[PIT_MAXC];
    unrank_perm(p, C, perm);
    float v = 0.f;
    for (int i = 0; i < C; ++i) v += (float)snr[i][perm[i]];  // fp32 accumulation in einsum order (i ascending)
    if (p == 0 || v > best_v) {  // strict > keeps the first maximum, like torch.argmax
      best_v = v;
      best = p;
    }
  }
  int perm[PIT_MAXC];
  unrank_perm(best, C, perm);
  max_snr[b] = best_v / (float)C;
  idx_out[b] = best;
  const double k00 = (10.0 / log(10.0)) * (-1.0 / ((double)B * (double)C));
  for (int i = 0; i < C; ++i) {
    const int j = perm[i];
    const double dot = dotm[i][j];
    const double E = Ess[j] + eps;
    const double a = dot / E;
    const double P = a * a * Ess[j];
    const double Q = Eee[i] - 2.0 * a * dot + a * a * Ess[j];
    const double rho = P / (Q + eps);
    const double qs = dot - a * Ess[j];
    const double k0 = k00 / (rho + eps);
    const double A_ = k0 * (-P * 2.0 / ((Q + eps) * (Q + eps)));
    const double B_ = k0 * (2.0 * a * Ess[j] / E / (Q + eps) + P * 2.0 * (qs / E) / ((Q + eps) * (Q + eps)));
    const double ce = A_, cs = B_ - A_ * a;
    const double sum_sbar = M[3 * C + j] - n * mus[j];
    const double c0 = -ce * mue[i] - cs * mus[j] - cs * sum_sbar / n;
    float* co = coef + ((int64_t)b * C + i) * 4;
    co[0] = (float)ce;
    co[1] = (float)cs;
    co[2] = (float)c0;
    co[3] = (float)j;
  }
  __threadfence();
  const unsigned int d = atomicAdd(ws.done, 1u);
  if (d == (unsigned int)B - 1) {
    __threadfence();
    float acc = 0.f;
    for (int bb = 0; bb < B; ++bb) acc += ((volatile float*)max_snr)[bb];  // fixed order => deterministic loss
    loss[0] = 0.f - acc / (float)B;
  }
}

__global__ void __launch_bounds__(256) pit_bwd_kernel(const float* __restrict__ src, const float* __restrict__ est,
                                                      const int64_t* __restrict__ lengths, const float* __restrict__ coef,
                                                      const float* __restrict__ grad_loss, int C, int T,
                                                      float* __restrict__ d_est) {
  pdl_launch_dependents();
  pdl_wait();
  const int bc = blockIdx.y, b = bc / C;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float* co = coef + (int64_t)bc * 4;
  const float gl = grad_loss != nullptr ? __ldg(grad_loss) : 1.f;
  const int j = (int)co[3];
  int64_t len = lengths[b];
  len = len < 0 ? 0 : (len > T ? T : len);
  float g = 0.f;
  if (t < len) g = gl * (co[0] * est[(int64_t)bc * T + t] + co[1] * src[((int64_t)b * C + j) * T + t] + co[2]);
  d_est[(int64_t)bc * T + t] = g;
}

__global__ void __launch_bounds__(256) reorder_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                                                      int C, int64_t inner, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int bc = blockIdx.y, b = bc / C, c = bc - b * C;
  int perm[PIT_MAXC];
  unrank_perm((int)idx[b], C, perm);
  const float* in = src + ((int64_t)b * C + perm[c]) * inner;  // the permutation, not its inverse (pit_criterion.py:92-98)
  float* o = out + (int64_t)bc * inner;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < inner; t += (int64_t)gridDim.x * blockDim.x)
    o[t] = in[t];
}

}  // namespace

int64_t pit_workspace_bytes(int B, int C) {
  (void)C;
  return (int64_t)B * PIT_NMOM * sizeof(double) + ((int64_t)B + 1) * sizeof(unsigned int) + 64;
}

int run_pit_forward(const float* source, float* est, const int64_t* lengths, int B, int C, int T, float* loss,
                    float* max_snr, int64_t* idx, float* reorder, float* coef, void* pit_ws, cudaStream_t s) {
  CTN_REQUIRE(C >= 1 && C <= PIT_MAXC, "PIT: C must be in [1,%d] (got %d)", PIT_MAXC, C);
  CTN_REQUIRE(B >= 1 && T >= 1 && B <= 65535 / C, "PIT: bad shape B=%d T=%d", B, T);
  CTN_CUDA(cudaMemsetAsync(pit_ws, 0, (size_t)pit_workspace_bytes(B, C), s));
  PitWs ws;
  ws.mom = reinterpret_cast<double*>(pit_ws);
  ws.ticket = reinterpret_cast<unsigned int*>(ws.mom + (int64_t)B * PIT_NMOM);
  ws.done = ws.ticket + B;
  const int chunks = cdiv(T, PIT_THREADS * PIT_PER_THREAD);
  auto kern = C == 1 ? pit_moments_kernel<1> : C == 2 ? pit_moments_kernel<2> : C == 3 ? pit_moments_kernel<3>
                                                                                       : pit_moments_kernel<4>;
  launch_kernel(kern, dim3(chunks, B), PIT_THREADS, 0, s, source, est, lengths, B, C, T, ws, loss, max_snr, idx, coef);
  CTN_TRY(check_launch("pit_moments_kernel"));
  if (reorder != nullptr) {
    int gx = cdiv(T, 256 * 8);
    launch_kernel(reorder_kernel, dim3(gx < 1 ? 1 : gx, B * C), 256, 0, s, est, idx, C, T, reorder);
    CTN_TRY(check_launch("reorder_kernel"));
  }
  return 0;
}

int run_pit_backward(const float* source, const float* est_masked, const int64_t* lengths, const float* coef,
                     const float* grad_loss, int B, int C, int T, float* d_est, cudaStream_t s) {
  CTN_REQUIRE(C >= 1 && C <= PIT_MAXC && B * C <= 65535, "PIT backward: bad shape");
  launch_kernel(pit_bwd_kernel, dim3(cdiv(T, 256), B * C), 256, 0, s, source, est_masked, lengths, coef, grad_loss, C, T, d_est);
  return check_launch("pit_bwd_kernel");
}

int run_reorder(const float* source, const int64_t* idx, int B, int C, int64_t inner, float* out, cudaStream_t s) {
  CTN_REQUIRE(C >= 1 && C <= PIT_MAXC && B * C <= 65535, "reorder_source: bad shape");
  int gx = cdiv(inner, 256 * 8);
  launch_kernel(reorder_kernel, dim3(gx < 1 ? 1 : gx, B * C), 256, 0, s, source, idx, C, inner, out);
  return check_launch("reorder_kernel");
}

}  // namespace ctn
