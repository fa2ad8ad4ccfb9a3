// codec.cu — the two ends of the network: Encoder (framing conv + ReLU, src/conv_tasnet.py:108-121) and Decoder (mask
// nonlinearity, mask * mixture_w, basis Linear(N, L), overlap_and_add, zero pad to T; src/conv_tasnet.py:131-146,57-59),
// forward and backward.  These are skinny GEMMs (L = 20 outputs / reduction steps) around HBM-bound tensors; they run on
// the CUDA cores.  All four kernels are templated on the frame length (LT = L for the paper's 20, else LT = 32 or 64 —
// the paper's L = 40 — with run-time guards) so the per-frame loops unroll without predicated-off instructions, keep several independent global
// loads in flight per thread, and pick their frame tile per launch so the grid fills whole waves of the SMs.
#include "common.cuh"

namespace ctn {
namespace {

constexpr int MAXC = 4;
constexpr int MAXL = 64;

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// frames per block in [lo, hi]: minimise (waves of the grid) x (rounds of `per_round` frames a block walks)
int pick_tile(int K, int M, int halo, int lo, int hi, int slots, int per_round) {
  int best = lo;
  int64_t best_cost = -1;
  for (int tk = lo; tk <= hi; ++tk) {
    const int64_t blocks = (int64_t)cdiv(K, tk) * M;
    const int64_t cost = ((blocks + slots - 1) / slots) * cdiv(tk + halo, per_round);
    if (best_cost < 0 || cost < best_cost || (cost == best_cost && tk > best)) {
      best = tk;
      best_cost = cost;
    }
  }
  return best;
}

template <typename Kern>
int slots_for(Kern kernel, int threads, size_t smem) {
  int bps = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, threads, smem) != cudaSuccess || bps < 1) bps = 1;
  return bps * sm_count();
}

int ensure_smem(const void* fn, size_t bytes) {
  if (bytes > 48 * 1024) {
    CTN_REQUIRE(bytes <= 227 * 1024, "kernel needs %zu bytes of shared memory (> 227 KB): N*L too large", bytes);
    CTN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// Encoder forward: w[m,k,n] = relu(sum_l U[n,l] * mix[m, k*S + l]).  Thread per basis n (its U row in registers),
// the tile's samples in shared memory (broadcast reads).
// ---------------------------------------------------------------------------------------
template <int LT>
__global__ void __launch_bounds__(256) encoder_fwd_kernel(const float* __restrict__ mix, const float* __restrict__ U,
                                                          int T, int K, int N, int L, int tk, float* __restrict__ w) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  float* xs = sm;  // [tk*S + L]
  const int S = L / 2, m = blockIdx.y, k0 = blockIdx.x * tk;
  const int nk = min(tk, K - k0);
  const int nx = (nk - 1) * S + L;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) xs[i] = mix[(int64_t)m * T + (int64_t)k0 * S + i];
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float u[LT];
#pragma unroll
    for (int l = 0; l < LT; ++l) u[l] = l < L ? __ldg(U + n * L + l) : 0.f;
    float* out = w + ((int64_t)m * K + k0) * N + n;
    int k = 0;
    for (; k + 1 < nk; k += 2) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int l = 0; l < LT; ++l) {
        if (l < L) {
          a0 = fmaf(u[l], xs[k * S + l], a0);
          a1 = fmaf(u[l], xs[(k + 1) * S + l], a1);
        }
      }
      out[(int64_t)k * N] = fmaxf(a0, 0.f);
      out[(int64_t)(k + 1) * N] = fmaxf(a1, 0.f);
    }
    if (k < nk) {
      float a0 = 0.f;
#pragma unroll
      for (int l = 0; l < LT; ++l)
        if (l < L) a0 = fmaf(u[l], xs[k * S + l], a0);
      out[(int64_t)k * N] = fmaxf(a0, 0.f);
    }
  }
}

// Encoder backward: dU[n,l] += sum_{f in tile} (dwa+dwb)[f,n] * [w[f,n] > 0] * mix[m, k*S+l].  Thread per n, the L
// accumulators in registers, 4 frames of independent loads in flight.
template <int LT>
__global__ void __launch_bounds__(256) encoder_bwd_kernel(const float* __restrict__ mix, const float* __restrict__ w,
                                                          const float* __restrict__ dwa, const float* __restrict__ dwb,
                                                          int T, int K, int N, int L, int tk, float* __restrict__ dU) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  float* xs = sm;  // [tk*S + L]
  const int S = L / 2, m = blockIdx.y, k0 = blockIdx.x * tk;
  const int nk = min(tk, K - k0);
  const int nx = (nk - 1) * S + L;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) xs[i] = mix[(int64_t)m * T + (int64_t)k0 * S + i];
  __syncthreads();
  for (int nb = 0; nb < N; nb += blockDim.x) {  // block-uniform loop (barriers inside)
    const int n = nb + threadIdx.x;
    const bool live = n < N;
    float acc[LT];
#pragma unroll
    for (int l = 0; l < LT; ++l) acc[l] = 0.f;
    const int64_t base = ((int64_t)m * K + k0) * N + n;
    for (int k = 0; live && k < nk; k += 4) {
      float g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        g[u] = 0.f;
        if (k + u < nk) {
          const int64_t idx = base + (int64_t)(k + u) * N;
          float gv = dwa[idx];
          if (dwb != nullptr) gv += dwb[idx];
          g[u] = w[idx] > 0.f ? gv : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k + u < nk) {
#pragma unroll
          for (int l = 0; l < LT; ++l)
            if (l < L) acc[l] = fmaf(g[u], xs[(k + u) * S + l], acc[l]);
        }
      }
    }
    // dU is [N][L]: a thread's L sums are 4L bytes apart from its neighbour's, so stage the block's [256][L] slab in
    // shared memory and add it with fully coalesced atomics (one 128-byte reduction per warp instruction)
    float* slab = xs + tk * S + L;
    const int n0 = n - threadIdx.x;
#pragma unroll
    for (int l = 0; l < LT; ++l)
      if (l < L) slab[threadIdx.x * L + l] = acc[l];
    __syncthreads();
    const int cnt = min((int)blockDim.x, N - n0) * L;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) atomicAdd(dU + n0 * L + i, slab[i]);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// Decoder forward.  grid (frame tiles, M); one warp per frame: lane owns the basis channels n = lane + 32 i, forms
// sw[c][i] = mask(score)[c][n] * w[n], accumulates its partial frame p[l] = sum_i sw[c][i] V[l][n] for all L outputs
// (V staged in shared memory in a lane-major float4 layout: conflict-free 16-byte loads), and the warp reduces the L
// partials with ONE 31-shuffle transpose-reduce (lane l ends up with output l) instead of L five-step butterflies.
// Then the block overlap-adds its span of output samples (ascending frame order, like index_add_) and zero pads.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float transpose_reduce32(float (&p)[32], int lane) {
  // returns sum over lanes of p[lane]; consumes p
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool up = (lane & w) != 0;
#pragma unroll
    for (int j = 0; j < w; ++j) {
      const float send = up ? p[j] : p[j + w];
      const float keep = up ? p[j + w] : p[j];
      p[j] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return p[0];
}

template <int LT, int NIT>  // NIT: basis channels per lane (8: N <= 256, 16: N <= 512)
__global__ void __launch_bounds__(256) decoder_fwd_kernel(const float* __restrict__ score, const float* __restrict__ w,
                                                          const float* __restrict__ V, int K, int C, int N, int L,
                                                          int T, int softmax, int tk, float* __restrict__ est) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float smf[];
  const int S = L / 2;
  const int halo = (L - 1) / S;  // frames before the tile that still reach into it
  const int ni = (N + 31) >> 5, ni4 = (ni + 3) >> 2;
  float4* Vt = reinterpret_cast<float4*>(smf);  // [L][ni4][32 lanes] float4 = V[l][lane + 32 (4 i4 + 0..3)]
  float* fr = smf + (size_t)L * ni4 * 128;      // [(tk + halo)][C][L]
  const int m = blockIdx.y, k0 = blockIdx.x * tk;
  const int kb = max(0, k0 - halo), ke = min(K, k0 + tk);
  for (int i = threadIdx.x; i < L * ni4 * 128; i += blockDim.x) {
    const int ii = i & 3, ln = (i >> 2) & 31, i4 = (i >> 7) % ni4, l = (i >> 7) / ni4;
    const int n = ln + 32 * (4 * i4 + ii);
    smf[i] = n < N ? V[l * N + n] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = kb + wid; k < ke; k += nw) {
    const int64_t f = (int64_t)m * K + k;
    float wv[NIT], sc[MAXC][NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int n = lane + 32 * i;
      const bool ok = i < ni && n < N;
      wv[i] = ok ? w[f * N + n] : 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) sc[c][i] = (ok && c < C) ? score[f * (int64_t)(C * N) + c * N + n] : -INFINITY;
    }
#pragma unroll
    for (int i = 0; i < NIT; ++i) {  // mask nonlinearity (src/conv_tasnet.py:209-212), times mixture_w
      const bool ok = i < ni && lane + 32 * i < N;
      if (softmax) {
        float mx = sc[0][i];
#pragma unroll
        for (int c = 1; c < MAXC; ++c) mx = fmaxf(mx, sc[c][i]);
        float e[MAXC], den = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          e[c] = (ok && c < C) ? expf(sc[c][i] - mx) : 0.f;
          den += e[c];
        }
#pragma unroll
        for (int c = 0; c < MAXC; ++c) sc[c][i] = ok ? e[c] / den * wv[i] : 0.f;
      } else {
#pragma unroll
        for (int c = 0; c < MAXC; ++c) sc[c][i] = ok ? fmaxf(sc[c][i], 0.f) * wv[i] : 0.f;
      }
    }
    float* out = fr + (size_t)(k - kb) * C * L;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
#pragma unroll
        for (int lb = 0; lb < LT; lb += 32) {  // 32 outputs per transpose-reduce (two rounds for 32 < L <= 64)
          if (lb < L) {
            float p[32];
#pragma unroll
            for (int l = 0; l < 32; ++l) p[l] = 0.f;
#pragma unroll
            for (int l = 0; l < 32; ++l) {
              if (lb + l < LT && lb + l < L) {
#pragma unroll
                for (int i4 = 0; i4 < NIT / 4; ++i4) {
                  if (i4 < ni4) {
                    const float4 v = Vt[((lb + l) * ni4 + i4) * 32 + lane];
                    p[l] = fmaf(sc[c][4 * i4 + 0], v.x, p[l]);
                    p[l] = fmaf(sc[c][4 * i4 + 1], v.y, p[l]);
                    p[l] = fmaf(sc[c][4 * i4 + 2], v.z, p[l]);
                    p[l] = fmaf(sc[c][4 * i4 + 3], v.w, p[l]);
                  }
                }
              }
            }
            const float tot = transpose_reduce32(p, lane);
            if (lb + lane < L) out[c * L + lb + lane] = tot;
          }
        }
      }
    }
  }
  __syncthreads();
  // output span of this tile: [k0*S, (k0+tk)*S), the last tile runs to T (tail + zero pad)
  const int t0 = k0 * S;
  const int t1 = (k0 + tk >= K) ? T : (k0 + tk) * S;
  for (int c = 0; c < C; ++c) {
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
      float acc = 0.f;
      const int khi = min(t / S, K - 1);
      int klo = (t - L + S) / S;  // ceil((t-L+1)/S) for t-L+1 >= 0
      if (t - L + 1 <= 0) klo = 0;
      for (int k = max(klo, kb); k <= khi; ++k) {
        const int l = t - k * S;
        if (l >= 0 && l < L) acc += fr[((k - kb) * C + c) * L + l];
      }
      est[((int64_t)m * C + c) * T + t] = acc;
    }
  }
}

// Decoder backward: thread per basis channel n (its V column and dV accumulators live in registers), loop over the
// tile's frames, two frames of independent loads in flight; the frame gradients df[k][c][l] = d_est[m,c,k*S+l] sit in
// shared memory (broadcast 16-byte reads when L % 4 == 0).
template <int LT>
__global__ void __launch_bounds__(256) decoder_bwd_kernel(const float* __restrict__ d_est, const float* __restrict__ score,
                                                          const float* __restrict__ w, const float* __restrict__ V, int K,
                                                          int C, int N, int L, int T, int softmax, int tk,
                                                          float* __restrict__ d_score, float* __restrict__ d_w,
                                                          float* __restrict__ dV) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float smf[];
  const int S = L / 2;
  float* df = smf;  // [tk][C][L]
  const int m = blockIdx.y, k0 = blockIdx.x * tk;
  const int nk = min(tk, K - k0);
  for (int i = threadIdx.x; i < nk * C * L; i += blockDim.x) {
    const int l = i % L, c = (i / L) % C, kk = i / (L * C);
    df[i] = d_est[((int64_t)m * C + c) * T + (int64_t)(k0 + kk) * S + l];
  }
  __syncthreads();
  constexpr bool VEC = (LT % 4 == 0) && LT < 32;  // LT == L exactly and 16-byte rows
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float vcol[LT], dvacc[LT];
#pragma unroll
    for (int l = 0; l < LT; ++l) {
      vcol[l] = l < L ? V[l * N + n] : 0.f;
      dvacc[l] = 0.f;
    }
    for (int kk0 = 0; kk0 < nk; kk0 += 2) {
      float wv[2], sc[2][MAXC];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const bool ok = kk0 + u < nk;
        const int64_t f = (int64_t)m * K + k0 + kk0 + u;
        wv[u] = ok ? w[f * N + n] : 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) sc[u][c] = (ok && c < C) ? score[f * (int64_t)(C * N) + c * N + n] : -INFINITY;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int kk = kk0 + u;
        if (kk >= nk) break;
        const int64_t f = (int64_t)m * K + k0 + kk;
        float mk[MAXC], dsw[MAXC];
        if (softmax) {
          float mx = sc[u][0];
#pragma unroll
          for (int c = 1; c < MAXC; ++c) mx = fmaxf(mx, sc[u][c]);
          float den = 0.f;
#pragma unroll
          for (int c = 0; c < MAXC; ++c) { mk[c] = c < C ? expf(sc[u][c] - mx) : 0.f; den += mk[c]; }
#pragma unroll
          for (int c = 0; c < MAXC; ++c) mk[c] /= den;
        } else {
#pragma unroll
          for (int c = 0; c < MAXC; ++c) mk[c] = fmaxf(sc[u][c], 0.f);
        }
        float dwv = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          dsw[c] = 0.f;
          if (c < C) {
            const float* dfc = df + (kk * C + c) * L;
            const float sw = mk[c] * wv[u];
            float acc = 0.f;
            if (VEC) {
#pragma unroll
              for (int l = 0; l < LT; l += 4) {
                const float4 d = *reinterpret_cast<const float4*>(dfc + l);
                acc = fmaf(d.x, vcol[l], acc); dvacc[l] = fmaf(d.x, sw, dvacc[l]);
                acc = fmaf(d.y, vcol[l + 1], acc); dvacc[l + 1] = fmaf(d.y, sw, dvacc[l + 1]);
                acc = fmaf(d.z, vcol[l + 2], acc); dvacc[l + 2] = fmaf(d.z, sw, dvacc[l + 2]);
                acc = fmaf(d.w, vcol[l + 3], acc); dvacc[l + 3] = fmaf(d.w, sw, dvacc[l + 3]);
              }
            } else {
#pragma unroll
              for (int l = 0; l < LT; ++l) {
                if (l < L) {
                  const float d = dfc[l];
                  acc = fmaf(d, vcol[l], acc);
                  dvacc[l] = fmaf(d, sw, dvacc[l]);
                }
              }
            }
            dsw[c] = acc;
            dwv = fmaf(acc, mk[c], dwv);
          }
        }
        d_w[f * N + n] = dwv;
        if (softmax) {
          float dot = 0.f;
#pragma unroll
          for (int c = 0; c < MAXC; ++c) dot = fmaf(dsw[c] * wv[u], mk[c], dot);
#pragma unroll
          for (int c = 0; c < MAXC; ++c)
            if (c < C) d_score[f * (int64_t)(C * N) + c * N + n] = mk[c] * (dsw[c] * wv[u] - dot);
        } else {
#pragma unroll
          for (int c = 0; c < MAXC; ++c)
            if (c < C) d_score[f * (int64_t)(C * N) + c * N + n] = sc[u][c] > 0.f ? dsw[c] * wv[u] : 0.f;
        }
      }
    }
#pragma unroll
    for (int l = 0; l < LT; ++l)
      if (l < L) atomicAdd(dV + l * N + n, dvacc[l]);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host launchers (C linkage wrappers live in c_api.cu)
// ---------------------------------------------------------------------------------------
int run_encoder_fwd(const float* mix, const float* U, int M, int T, int N, int L, float* w, cudaStream_t s) {
  CTN_REQUIRE(L >= 2 && T >= L, "encoder: need L >= 2 and T >= L (T=%d L=%d)", T, L);
  CTN_REQUIRE(L <= MAXL, "encoder: L <= %d supported (got %d)", MAXL, L);
  const int S = L / 2, K = (T - L) / S + 1;
  auto kern = L == 20 ? encoder_fwd_kernel<20> : L <= 32 ? encoder_fwd_kernel<32> : encoder_fwd_kernel<64>;
  const int tk = pick_tile(K, M, 0, 16, 64, slots_for(kern, 256, (size_t)(64 * S + L) * 4), 2);
  const size_t smem = (size_t)(tk * S + L) * sizeof(float);
  launch_kernel(kern, dim3(cdiv(K, tk), M), 256, smem, s, mix, U, T, K, N, L, tk, w);
  return check_launch("encoder_fwd_kernel");
}

int run_encoder_bwd(const float* mix, const float* w, const float* dwa, const float* dwb, int M, int T, int N, int L,
                    float* dU, cudaStream_t s) {
  const int S = L / 2, K = (T - L) / S + 1;
  CTN_REQUIRE(L <= MAXL, "encoder: L <= %d supported (got %d)", MAXL, L);
  auto kern = L == 20 ? encoder_bwd_kernel<20> : L <= 32 ? encoder_bwd_kernel<32> : encoder_bwd_kernel<64>;
  CTN_TRY(ensure_smem((const void*)kern, (size_t)(64 * S + L + 256 * L) * 4));
  const int tk = pick_tile(K, M, 0, 32, 64, slots_for(kern, 256, (size_t)(64 * S + L + 256 * L) * 4), 4);
  const size_t smem = (size_t)(tk * S + L + 256 * L) * sizeof(float);
  launch_kernel(kern, dim3(cdiv(K, tk), M), 256, smem, s, mix, w, dwa, dwb, T, K, N, L, tk, dU);
  return check_launch("encoder_bwd_kernel");
}

int run_decoder_fwd(const float* score, const float* w, const float* V, int M, int K, int C, int N, int L, int T,
                    int softmax, float* est, cudaStream_t s) {
  CTN_REQUIRE(C >= 1 && C <= MAXC, "decoder: C must be in [1,%d] (got %d)", MAXC, C);
  CTN_REQUIRE(L <= MAXL, "decoder: L <= %d supported (got %d)", MAXL, L);
  CTN_REQUIRE(N <= 512, "decoder: N <= 512 supported (got %d)", N);
  const int S = L / 2, halo = (L - 1) / S;
  const int ni4 = (((N + 31) >> 5) + 3) >> 2;
  void (*kern)(const float*, const float*, const float*, int, int, int, int, int, int, int, float*);
  if (N <= 256) kern = L == 20 ? decoder_fwd_kernel<20, 8> : L <= 32 ? decoder_fwd_kernel<32, 8> : decoder_fwd_kernel<64, 8>;
  else kern = L == 20 ? decoder_fwd_kernel<20, 16> : L <= 32 ? decoder_fwd_kernel<32, 16> : decoder_fwd_kernel<64, 16>;
  const size_t vbytes = (size_t)L * ni4 * 128 * 4;
  const size_t smem_hi = vbytes + (size_t)(32 + halo) * C * L * 4;
  CTN_TRY(ensure_smem((const void*)kern, smem_hi));
  const int tk = pick_tile(K, M, halo, 8, 32, slots_for(kern, 256, smem_hi), 8);
  const size_t smem = vbytes + (size_t)(tk + halo) * C * L * sizeof(float);
  launch_kernel(kern, dim3(cdiv(K, tk), M), 256, smem, s, score, w, V, K, C, N, L, T, softmax, tk, est);
  return check_launch("decoder_fwd_kernel");
}

int run_decoder_bwd(const float* d_est, const float* score, const float* w, const float* V, int M, int K, int C, int N,
                    int L, int T, int softmax, float* d_score, float* d_w, float* dV, cudaStream_t s) {
  CTN_REQUIRE(C >= 1 && C <= MAXC, "decoder: C must be in [1,%d] (got %d)", MAXC, C);
  CTN_REQUIRE(L <= MAXL, "decoder: L <= %d supported (got %d)", MAXL, L);
  auto kern = L == 20 ? decoder_bwd_kernel<20> : L <= 32 ? decoder_bwd_kernel<32> : decoder_bwd_kernel<64>;
  const size_t smem_hi = (size_t)(64 * C * L) * sizeof(float);
  CTN_TRY(ensure_smem((const void*)kern, smem_hi));
  const int tk = pick_tile(K, M, 0, 32, 64, slots_for(kern, 256, smem_hi), 2);
  const size_t smem = (size_t)(tk * C * L) * sizeof(float);
  launch_kernel(kern, dim3(cdiv(K, tk), M), 256, smem, s, d_est, score, w, V, K, C, N, L, T, softmax, tk, d_score, d_w, dV);
  return check_launch("decoder_bwd_kernel");
}

}  // namespace ctn
