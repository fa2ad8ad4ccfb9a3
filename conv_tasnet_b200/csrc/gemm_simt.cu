// gemm_simt.cu — fp32 CUDA-core GEMMs for the 1x1 convolutions (forward, dgrad) and their weight
// gradients.  These are the exact-fp32 path: every shape the model needs goes through here unless the
// tcgen05 path (gemm_tc.cu) accepts it.  Channels-last activations make every 1x1 conv a row-major
// [frames, Cin] x [Cout, Cin]^T product (src/conv_tasnet.py:174,191,223,262).
#include "common.cuh"

namespace ctn {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256, PAD = 4;

struct StatCache {
  float2 v[16];
  int m_lo;
};

__device__ __forceinline__ void stat_cache_init(StatCache& c, const NormStats& s, int m_lo, int m_hi) {
  if (threadIdx.x == 0) c.m_lo = m_lo;
  if (s.row == nullptr && (int)threadIdx.x <= m_hi - m_lo && threadIdx.x < 16) {
    float mu, r;
    load_stats(s, m_lo + threadIdx.x, 0, mu, r);
    c.v[threadIdx.x] = make_float2(mu, r);
  }
  __syncthreads();
}

__device__ __forceinline__ void get_stats(const StatCache& c, const NormStats& s, int m, int64_t f, float& mu,
                                          float& r) {
  if (s.row != nullptr) {
    float2 v = reinterpret_cast<const float2*>(s.row)[f];
    mu = v.x;
    r = v.y;
  } else if (m - c.m_lo < 16) {
    float2 v = c.v[m - c.m_lo];
    mu = v.x;
    r = v.y;
  } else {
    load_stats(s, m, f, mu, r);
  }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ------------------------------------------------------------------------------------------
// D[F,O] = epi( pro(A[F,Kd]) . W^T )
// ------------------------------------------------------------------------------------------
template <bool W_IS_KN>
__global__ void __launch_bounds__(NT, 2) gemm_kernel(GemmArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  __shared__ float rowsum[BM][2];
  __shared__ StatCache sc;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t f0 = (int64_t)blockIdx.x * BM;
  const int o0 = blockIdx.y * BN;
  const int Kd = a.Kd, O = a.O;
  const bool pro = a.alpha_in != nullptr;
  const float alpha_in = pro ? __ldg(a.alpha_in) : 1.f;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  const int lrow = tid >> 2, lk = (tid & 3) * 4;   // [row][k] operand tiles: 64 rows x 4 float4 per pass
  const int krow = tid >> 5, kcol = (tid & 31) * 4;  // [k][col] operand tiles: 8 k-rows x 32 float4 per pass

  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t f = f0 + lrow + i * 64;
      const int k = k0 + lk;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < a.F && k < Kd) {
        v = ld4(a.A + f * Kd + k);
        if (pro) {
          v.x = prelu(v.x, alpha_in); v.y = prelu(v.y, alpha_in);
          v.z = prelu(v.z, alpha_in); v.w = prelu(v.w, alpha_in);
        }
      }
      ra[i] = v;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!W_IS_KN) {
        const int o = o0 + lrow + i * 64;
        if (o < O && k < Kd) w = ld4(a.W + (int64_t)o * Kd + k);
      } else {
        const int kk = k0 + krow + i * 8, o = o0 + kcol;
        if (kk < Kd && o < O) w = ld4(a.W + (int64_t)kk * O + o);
      }
      rb[i] = w;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = lrow + i * 64;
      As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y;
      As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
      if (!W_IS_KN) {
        Bs[buf][lk + 0][r] = rb[i].x; Bs[buf][lk + 1][r] = rb[i].y;
        Bs[buf][lk + 2][r] = rb[i].z; Bs[buf][lk + 3][r] = rb[i].w;
      } else {
        *reinterpret_cast<float4*>(&Bs[buf][krow + i * 8][kcol]) = rb[i];
      }
    }
  };

  const int nk = (Kd + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  {
    const int64_t fl = f0 + BM - 1 < a.F ? f0 + BM - 1 : a.F - 1;
    stat_cache_init(sc, a.st, (int)(f0 / a.K), (int)(fl / a.K));  // contains __syncthreads
  }
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tile(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue ----
  const bool fold = a.c1 != nullptr;
  const bool stats = a.stat_out != nullptr;
  const float alpha_out = (stats && a.alpha_out) ? __ldg(a.alpha_out) : 1.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
    const int64_t f = f0 + r;
    const bool vf = f < a.F;
    float mu = 0.f, rs = 1.f;
    if (fold && vf) get_stats(sc, a.st, (int)(f / a.K), f, mu, rs);
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int o = o0 + jh * 64 + tx * 4;
      if (vf && o < O) {
        float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
        if (fold) {
          const float4 c1 = ld4(a.c1 + o), c2 = ld4(a.c2 + o);
          const float mr = mu * rs;
          v.x = rs * v.x + c1.x - mr * c2.x; v.y = rs * v.y + c1.y - mr * c2.y;
          v.z = rs * v.z + c1.z - mr * c2.z; v.w = rs * v.w + c1.w - mr * c2.w;
        }
        if (a.res != nullptr) {
          const float4 q = ld4(a.res + f * O + o);
          v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
        }
        *reinterpret_cast<float4*>(a.D + f * O + o) = v;
        if (stats) {
          const float p0 = prelu(v.x, alpha_out), p1 = prelu(v.y, alpha_out);
          const float p2 = prelu(v.z, alpha_out), p3 = prelu(v.w, alpha_out);
          s += (p0 + p1) + (p2 + p3);
          s2 += (p0 * p0 + p1 * p1) + (p2 * p2 + p3 * p3);
        }
      }
    }
    if (stats) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (tx == 0) {
        rowsum[r][0] = s;
        rowsum[r][1] = s2;
      }
    }
  }
  if (stats) {
    __syncthreads();
    if (tid < BM) {  // 4 full warps
      const int64_t f = f0 + tid;
      const bool vf = f < a.F;
      const int m = vf ? (int)(f / a.K) : -1;
      const float s = vf ? rowsum[tid][0] : 0.f, s2 = vf ? rowsum[tid][1] : 0.f;
      const int m0 = __shfl_sync(0xffffffffu, m, 0);
      const bool same = __all_sync(0xffffffffu, m == m0 || !vf);
      if (same) {
        const double ds = warp_sum((double)s), ds2 = warp_sum((double)s2);
        if ((tid & 31) == 0 && m0 >= 0) {
          atomicAdd(a.stat_out + 2 * m0, ds);
          atomicAdd(a.stat_out + 2 * m0 + 1, ds2);
        }
      } else if (vf) {
        atomicAdd(a.stat_out + 2 * m, (double)s);
        atomicAdd(a.stat_out + 2 * m + 1, (double)s2);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// dW[O,I] += sum_{f in chunk} G[f,o] * act(f,i)       (split over f, fp32 atomics)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 2) wgrad_kernel(WgradArgs a, int f_chunk) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  __shared__ StatCache sc;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int o0 = blockIdx.x * BM, i0 = blockIdx.y * BN;
  const int64_t fb = (int64_t)blockIdx.z * f_chunk;
  const int64_t fe = fb + f_chunk < a.F ? fb + f_chunk : a.F;
  if (fb >= fe) return;
  const int O = a.O, I = a.I;
  const bool norm = a.gamma != nullptr;
  const bool hasp = a.alpha != nullptr;
  const float alpha = hasp ? __ldg(a.alpha) : 1.f;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int krow = tid >> 5, kcol = (tid & 31) * 4;
  float4 gam = make_float4(1.f, 1.f, 1.f, 1.f), bet = make_float4(0.f, 0.f, 0.f, 0.f);
  if (norm && i0 + kcol < I) {
    gam = ld4(a.gamma + i0 + kcol);
    bet = ld4(a.beta + i0 + kcol);
  }
  stat_cache_init(sc, a.st, (int)(fb / a.K), (int)((fe - 1) / a.K));

  float4 ra[2], rb[2];
  auto load_tile = [&](int64_t fk) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t f = fk + krow + i * 8;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f), v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < fe) {
        if (o0 + kcol < O) g = ld4(a.G + f * O + o0 + kcol);
        if (i0 + kcol < I) {
          v = ld4(a.Act + f * I + i0 + kcol);
          if (hasp) {
            v.x = prelu(v.x, alpha); v.y = prelu(v.y, alpha);
            v.z = prelu(v.z, alpha); v.w = prelu(v.w, alpha);
          }
          if (norm) {
            float mu, r;
            get_stats(sc, a.st, (int)(f / a.K), f, mu, r);
            v.x = gam.x * (v.x - mu) * r + bet.x; v.y = gam.y * (v.y - mu) * r + bet.y;
            v.z = gam.z * (v.z - mu) * r + bet.z; v.w = gam.w * (v.w - mu) * r + bet.w;
          }
        }
      }
      ra[i] = g;
      rb[i] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      *reinterpret_cast<float4*>(&As[buf][krow + i * 8][kcol]) = ra[i];
      *reinterpret_cast<float4*>(&Bs[buf][krow + i * 8][kcol]) = rb[i];
    }
  };

  const int nk = (int)((fe - fb + BK - 1) / BK);
  load_tile(fb);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile(fb + (int64_t)(kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tile(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int o = o0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (o >= O) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ii = i0 + ((j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (ii < I) atomicAdd(a.dW + (int64_t)o * I + ii, acc[i][j]);
    }
  }
}

}  // namespace

int launch_gemm_simt(const GemmArgs& a, cudaStream_t s) {
  CTN_REQUIRE(a.Kd % 4 == 0 && a.O % 4 == 0, "conv1x1: channel counts must be multiples of 4 (got Kd=%d O=%d)", a.Kd, a.O);
  CTN_REQUIRE(a.F > 0 && a.K > 0, "conv1x1: empty input (F=%lld K=%d)", (long long)a.F, a.K);
  dim3 grid(cdiv(a.F, BM), cdiv(a.O, BN));
  if (a.w_is_kn)
    launch_kernel(gemm_kernel<true>, grid, NT, 0, s, a);
  else
    launch_kernel(gemm_kernel<false>, grid, NT, 0, s, a);
  return check_launch("gemm_kernel");
}

int launch_wgrad_simt(const WgradArgs& a, cudaStream_t s) {
  CTN_REQUIRE(a.I % 4 == 0 && a.O % 4 == 0, "wgrad: channel counts must be multiples of 4 (got O=%d I=%d)", a.O, a.I);
  CTN_REQUIRE(a.F > 0 && a.K > 0, "wgrad: empty input");
  const int tiles = cdiv(a.O, BM) * cdiv(a.I, BN);
  int splits = (2 * 148 + tiles - 1) / tiles;
  int f_chunk = (int)((a.F + splits - 1) / splits);
  f_chunk = ((f_chunk + BK - 1) / BK) * BK;
  if (f_chunk < 4 * BK) f_chunk = 4 * BK;
  splits = cdiv(a.F, f_chunk);
  dim3 grid(cdiv(a.O, BM), cdiv(a.I, BN), splits);
  launch_kernel(wgrad_kernel, grid, NT, 0, s, a, f_chunk);
  return check_launch("wgrad_kernel");
}

}  // namespace ctn
