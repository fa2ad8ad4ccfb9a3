// elementwise.cu — the bandwidth-bound kernels of the Conv-TasNet path: encoder framing, cLN row
// statistics, norm + dilated depthwise conv (+Chomp1d) forward/backward, gLN/cLN (+PReLU) backward,
// mask nonlinearity + decoder basis + overlap-add forward/backward, norm-fold weight prep.
// All activations are channels-last [M, K, Ch]; every global access is a 16-byte vector along channels.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace ctn {
namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// the same 4-channel accesses on a bf16 tensor (reduced-precision inference stores the H-wide activations as bf16)
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  const uint2 w = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                     __uint_as_float(w.y & 0xffff0000u));
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 w;
  w.x = *reinterpret_cast<const uint32_t*>(&a);
  w.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = w;
}
__device__ __forceinline__ float4 prelu4(float4 v, float a) {
  return make_float4(prelu(v.x, a), prelu(v.y, a), prelu(v.z, a), prelu(v.w, a));
}
__device__ __forceinline__ float dprelu(float z, float a) { return z > 0.f ? 1.f : a; }

// ---- mbarrier + bulk (TMA, non-tensor) copies: global rows -> shared memory without passing through registers ----
__device__ __forceinline__ uint32_t ew_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ew_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ew_smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void ew_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ew_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ew_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "EW_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra EW_DONE;\n\t"
      "bra EW_WAIT_LOOP;\n\t"
      "EW_DONE:\n\t"
      "}" ::"r"(ew_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// `bytes` (a multiple of 16) from 16-byte aligned global memory to 16-byte aligned shared memory; completion is counted on `bar`
__device__ __forceinline__ void ew_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   ew_smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(ew_smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------
// cLN statistics: one warp per frame, two-pass like torch.var (src/conv_tasnet.py:332-333)
// ---------------------------------------------------------------------------------------
// 16 bytes of a row: 4 fp32 or 8 bf16 channels -> floats
template <typename T> struct RowVec;
template <> struct RowVec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    v[4] = v[5] = v[6] = v[7] = 0.f;
  }
};
template <> struct RowVec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 w = *reinterpret_cast<const uint4*>(p);
    const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(u[i] << 16);
      v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
    }
  }
};

// One warp per RPW frames.  Rows of up to 512 channels are read ONCE — all of a lane's 16-byte loads of all RPW rows
// issued back to back (the first version's `for` loop of load + add serialised them: four memory latencies per pass, two
// passes, one row per warp: 2.6-4.3 TB/s) — and kept in registers for the second pass (the squared deviations, like
// torch.var); wider rows take the re-reading path.
template <typename T>
__global__ void __launch_bounds__(256) row_stats_kernel(const T* __restrict__ x, const float* __restrict__ alpha,
                                                        int64_t F, int Ch, float* __restrict__ rowstat) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VN = RowVec<T>::N;           // channels per 16-byte load
  constexpr int NVT = 512 / (32 * VN);       // loads per lane and row (fp32: 4, bf16: 2)
#ifndef CTN_RS_RPW_BF16
#define CTN_RS_RPW_BF16 1
#endif
  constexpr int RPW = VN == 8 ? CTN_RS_RPW_BF16 : 2;       // rows per warp (measured in the config-2 forward, F = 102k x 512: fp32 1 -> 43.5,
                                             // 2 -> 42.8 us; bf16 1 -> 44.7, 4 -> 59 us (89 registers, 2 blocks per SM))
  const int lane = threadIdx.x & 31;
  const int64_t f0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
  if (f0 >= F) return;
  const bool hasp = alpha != nullptr;
  const float a = hasp ? __ldg(alpha) : 1.f;
  if (Ch % VN == 0 && Ch <= 512) {
    float v[RPW][NVT][8];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
#pragma unroll
      for (int i = 0; i < NVT; ++i) {
        const int c = (lane + 32 * i) * VN;
        if (f0 + r < F && c < Ch) RowVec<T>::load(x + (f0 + r) * Ch + c, v[r][i]);
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      if (f0 + r >= F) break;  // warp-uniform
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NVT; ++i) {
        if ((lane + 32 * i) * VN < Ch) {
#pragma unroll
          for (int j = 0; j < VN; ++j) {
            if (hasp) v[r][i][j] = prelu(v[r][i][j], a);
            s += v[r][i][j];
          }
        }
      }
      const float mu = warp_sum(s) / (float)Ch;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NVT; ++i) {
        if ((lane + 32 * i) * VN < Ch) {
#pragma unroll
          for (int j = 0; j < VN; ++j) {
            const float d = v[r][i][j] - mu;
            q = fmaf(d, d, q);
          }
        }
      }
      const float var = warp_sum(q) / (float)Ch;
      if (lane == 0) {
        rowstat[2 * (f0 + r)] = mu;
        rowstat[2 * (f0 + r) + 1] = 1.f / sqrtf(var + CTN_EPS);
      }
    }
    return;
  }
  for (int r = 0; r < RPW && f0 + r < F; ++r) {
    const T* row = x + (f0 + r) * Ch;
    float s = 0.f;
    for (int c = lane * 4; c < Ch; c += 128) {
      float4 v = ld4(row + c);
      if (hasp) v = prelu4(v, a);
      s += (v.x + v.y) + (v.z + v.w);
    }
    const float mu = warp_sum(s) / (float)Ch;
    float q = 0.f;
    for (int c = lane * 4; c < Ch; c += 128) {
      float4 v = ld4(row + c);
      if (hasp) v = prelu4(v, a);
      const float d0 = v.x - mu, d1 = v.y - mu, d2 = v.z - mu, d3 = v.w - mu;
      q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    const float var = warp_sum(q) / (float)Ch;
    if (lane == 0) {
      rowstat[2 * (f0 + r)] = mu;
      rowstat[2 * (f0 + r) + 1] = 1.f / sqrtf(var + CTN_EPS);
    }
  }
}

// ---------------------------------------------------------------------------------------
// norm-fold constants: Wg = W*gamma, c1 = W@beta, c2 = rowsum(Wg).  One warp per output row.
// Batched over `nb` convs that sit at a constant stride in the flat parameter buffer.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_normfold_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int O, int I,
                                                            int64_t in_stride, float* __restrict__ Wg,
                                                            float* __restrict__ c1, float* __restrict__ c2,
                                                            int64_t wg_stride, int64_t c_stride) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= O) return;
  const int64_t b = blockIdx.y;
  const float* w = W + b * in_stride + (int64_t)o * I;
  const float* g = gamma + b * in_stride;
  const float* be = beta + b * in_stride;
  float* wg = Wg + b * wg_stride + (int64_t)o * I;
  double s1 = 0.0, s2 = 0.0;
  for (int i = lane; i < I; i += 32) {
    const float wv = w[i], v = wv * g[i];
    wg[i] = v;
    s1 += (double)wv * (double)be[i];
    s2 += (double)v;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    c1[b * c_stride + o] = (float)s1;
    c2[b * c_stride + o] = (float)s2;
  }
}

// ---------------------------------------------------------------------------------------
// norm1 apply + dilated depthwise conv (+Chomp1d) (+ gLN stats of prelu(z2))
// A block walks the frames k_j = r + j*d of one residue class r (mod dilation d): the P taps of output j are the
// window elements j + p - cshift of that sequence, so each input row is loaded and normalised ONCE and slides through
// a P-deep register window (instead of being re-read from L2 once per tap).
// grid (classes * segments, M), block = H/4 threads rounded up to a warp (<= 256, loops over channel groups)
// ---------------------------------------------------------------------------------------
#ifndef CTN_DW_TJ
#define CTN_DW_TJ 16
#endif
#ifndef CTN_DW_U
#define CTN_DW_U 4
#endif
constexpr int DW_TJ = CTN_DW_TJ;  // backward: outputs per block (per channel group); one partial row per block
constexpr int DW_U = CTN_DW_U;    // outputs per inner group: DW_U independent loads in flight
#ifndef CTN_DWF_TJ
#define CTN_DWF_TJ 16
#endif
#ifndef CTN_DWF_U
#define CTN_DWF_U 4
#endif
constexpr int DWF_TJ = CTN_DWF_TJ;  // forward (no partial rows: its tile can differ from the backward's); measured
                                    // (M = 3 x 4 s): eager in-place timing favours (32, 8) (16.4 vs 18.1 us) but the
                                    // graph-replayed step is fastest with (16, 4): 6.29 vs 6.36 ms (register footprint
                                    // under programmatic dependent launch); (8, 8) 25 us, (64, 8) 21 us
constexpr int MAXP = 8;
// the bulk-staged kernels issue one row copy per lane of warp 0 and keep per-row metadata in TJ + MAXP slots
static_assert(DW_TJ + MAXP - 1 <= 32 && DWF_TJ + MAXP - 1 <= 32, "depthwise tiles: at most 32 staged rows per tensor");

__host__ __device__ inline int dw_classes(int K, int dil) { return dil < K ? dil : K; }
__host__ __device__ inline int dw_blocks(int K, int dil, int tj = DW_TJ) {
  const int ncls = dw_classes(K, dil);
  const int per_class = (K + ncls - 1) / ncls;  // frames in the largest class
  return ncls * ((per_class + tj - 1) / tj);
}

#ifndef CTN_DWF_MINB
#define CTN_DWF_MINB 1
#endif
#ifndef CTN_DWB_MINB
#define CTN_DWB_MINB 1
#endif
template <int PT, typename T = float, int U_ = CTN_DWF_U>
__global__ void __launch_bounds__(256, CTN_DWF_MINB) dwconv_fwd_kernel(const T* __restrict__ z1, const float* __restrict__ alpha1,
                                                         NormStats st1, const float* __restrict__ gamma1,
                                                         const float* __restrict__ beta1, const float* __restrict__ Wd,
                                                         int K, int H, int P, int dil, int cshift,
                                                         T* __restrict__ z2, double* __restrict__ stat_out,
                                                         const float* __restrict__ alpha2) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[2 * 32];
  __shared__ float2 s_st;
  const int m = blockIdx.y;
  const int ncls = dw_classes(K, dil);
  const int r = blockIdx.x % ncls, j0 = (blockIdx.x / ncls) * DWF_TJ;
  const int nclass = (K - r + dil - 1) / dil;            // frames of this residue class
  const int nj = max(0, min(DWF_TJ, nclass - j0));        // outputs of this block
  if (st1.row == nullptr) {
    if (threadIdx.x == 0) {
      float mu, rr;
      load_stats(st1, m, 0, mu, rr);
      s_st = make_float2(mu, rr);
    }
    __syncthreads();
  }
  const float a1 = __ldg(alpha1);
  const bool do_stats = stat_out != nullptr;
  const float a2 = do_stats ? __ldg(alpha2) : 1.f;
  const int64_t base = (int64_t)m * K;
  constexpr int NP_ = PT ? PT : MAXP;
  const int PP = PT ? PT : P;
  double acc[2] = {0.0, 0.0};
  for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) {
    const float4 g = ld4(gamma1 + c), b = ld4(beta1 + c);
    float wd[4][NP_];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int p = 0; p < NP_; ++p) wd[j][p] = p < PP ? Wd[(c + j) * PP + p] : 0.f;
    // normalised input of sequence index idx (zero outside [0, K): the padding is applied after the norm)
    auto fetch = [&](int idx, float4& raw, float2& stv, bool& ok) {
      const int k = r + idx * dil;
      ok = idx >= 0 && k < K;
      raw = ok ? ld4(z1 + (base + k) * H + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      stv = (ok && st1.row != nullptr) ? reinterpret_cast<const float2*>(st1.row)[base + k] : s_st;
    };
    auto normalise = [&](const float4& raw, const float2& stv, bool ok) {
      if (!ok) return make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 v = prelu4(raw, a1);
      return make_float4(g.x * (v.x - stv.x) * stv.y + b.x, g.y * (v.y - stv.x) * stv.y + b.y,
                         g.z * (v.z - stv.x) * stv.y + b.z, g.w * (v.w - stv.x) * stv.y + b.w);
    };
    float4 w[NP_];  // w[p] = normalised input at sequence index j + p - cshift for the current output j
#pragma unroll
    for (int p = 0; p < NP_; ++p) {
      float4 raw; float2 stv; bool ok;
      if (p < PP) { fetch(j0 + p - cshift, raw, stv, ok); w[p] = normalise(raw, stv, ok); }
      else w[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s = 0.f, s2 = 0.f;
    for (int jj = 0; jj < nj; jj += U_) {
      // the U_ rows that enter the window after each of the next U_ outputs: independent loads, issued together
      float4 nraw[U_]; float2 nst[U_]; bool nok[U_];
#pragma unroll
      for (int u = 0; u < U_; ++u) fetch(j0 + jj + u + PP - cshift, nraw[u], nst[u], nok[u]);
#pragma unroll
      for (int u = 0; u < U_; ++u) {
        if (jj + u < nj) {
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int p = 0; p < NP_; ++p) {
            o.x = fmaf(wd[0][p], w[p].x, o.x); o.y = fmaf(wd[1][p], w[p].y, o.y);
            o.z = fmaf(wd[2][p], w[p].z, o.z); o.w = fmaf(wd[3][p], w[p].w, o.w);
          }
          if (sizeof(T) == 2) {  // bf16 storage: the statistics describe the values the consumers will read
            o.x = __bfloat162float(__float2bfloat16_rn(o.x)); o.y = __bfloat162float(__float2bfloat16_rn(o.y));
            o.z = __bfloat162float(__float2bfloat16_rn(o.z)); o.w = __bfloat162float(__float2bfloat16_rn(o.w));
          }
          st4(z2 + (base + r + (int64_t)(j0 + jj + u) * dil) * H + c, o);
          if (do_stats) {
            const float4 q = prelu4(o, a2);
            s += (q.x + q.y) + (q.z + q.w);
            s2 += (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w);
          }
        }
        // slide
#pragma unroll
        for (int p = 0; p + 1 < NP_; ++p) w[p] = w[p + 1];
        const float4 nv = normalise(nraw[u], nst[u], nok[u]);
#pragma unroll
        for (int p = 0; p < NP_; ++p)
          if (p == PP - 1) w[p] = nv;
      }
    }
    acc[0] += (double)s;
    acc[1] += (double)s2;
  }
  if (do_stats) {
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) {
      atomicAdd(stat_out + 2 * m, acc[0]);
      atomicAdd(stat_out + 2 * m + 1, acc[1]);
    }
  }
}


// The same forward with the block's input rows STAGED IN SHARED MEMORY by bulk copies (cp.async.bulk, one per 2 KB row,
// issued by the lanes of warp 0, completion on one mbarrier): every byte the block needs is requested at once and costs
// no registers while in flight.  The register-window kernel above keeps U rows per thread in flight and walks its tile in
// TJ / U + 1 dependent rounds of ~2 us (memory latency under load) — measured 2.3 TB/s where the 4-round, 8-loads-per-
// thread gln_bwd_apply reaches 4.6 TB/s; here the walk reads shared memory and the only exposed latency is one round.
// Tile row i holds sequence index j0 - cshift + i of residue class r; rows outside [0, K) are not copied (zero padding
// is applied after the norm, as in the reference's Conv1d padding of the normalised tensor).
template <int PT, typename T = float>
__global__ void __launch_bounds__(256) dwconv_fwd_bulk_kernel(const T* __restrict__ z1, const float* __restrict__ alpha1,
                                                              NormStats st1, const float* __restrict__ gamma1,
                                                              const float* __restrict__ beta1,
                                                              const float* __restrict__ Wd, int K, int H, int P, int dil,
                                                              int cshift, T* __restrict__ z2,
                                                              double* __restrict__ stat_out,
                                                              const float* __restrict__ alpha2,
                                                              float* __restrict__ rs1_out,
                                                              float* __restrict__ rs2_out) {
  pdl_launch_dependents();
  extern __shared__ __align__(128) uint8_t ew_dsm[];
  T* tile = reinterpret_cast<T*>(ew_dsm);  // [DWF_TJ + PP - 1][H]
  __shared__ uint64_t bar;
  __shared__ double red[2 * 32];
  __shared__ float2 s_rs[DWF_TJ + MAXP];  // (mean, rstd) of every tile row
  const int m = blockIdx.y;
  const int ncls = dw_classes(K, dil);
  const int r = blockIdx.x % ncls, j0 = (blockIdx.x / ncls) * DWF_TJ;
  const int nclass = (K - r + dil - 1) / dil;
  const int nj = max(0, min(DWF_TJ, nclass - j0));
  constexpr int NP_ = PT ? PT : MAXP;
  const int PP = PT ? PT : P;
  const int nrows = nj > 0 ? nj + PP - 1 : 0;
  const int idx0 = j0 - cshift;
  const int64_t base = (int64_t)m * K;
  if (threadIdx.x == 0) ew_mbar_init(&bar, 1);
  __syncthreads();
  pdl_wait();
  if (threadIdx.x < 32) {
    const int idx = idx0 + (int)threadIdx.x, k = r + idx * dil;
    const bool ok = (int)threadIdx.x < nrows && idx >= 0 && k < K;
    const unsigned mask = __ballot_sync(0xffffffffu, ok);
    if (threadIdx.x == 0) ew_mbar_expect_tx(&bar, (uint32_t)__popc(mask) * (uint32_t)(H * sizeof(T)));
    __syncwarp();
    // (dil = 1: one copy of the whole contiguous range instead of a copy per row measured the same, 13.7 vs 14.0 us)
    if (ok) ew_bulk_load(tile + (size_t)threadIdx.x * H, z1 + (base + k) * H, (uint32_t)(H * sizeof(T)), &bar);
  }
  // cLN (rs1_out / rs2_out given): the per-frame statistics are formed HERE from the staged rows — of prelu(z1) for every
  // input row the block holds (two passes over shared memory, like torch.var) and of prelu(z2) for its output rows, which
  // are kept in the shared-memory rows their inputs have left — instead of two row_stats passes over [F, H] per block
  // (21 % of the causal-cLN forward).  Rows the block owns as outputs are also written out for the later consumers.
  const bool own_stats = rs1_out != nullptr;
  if ((int)threadIdx.x < nrows) {
    const int idx = idx0 + (int)threadIdx.x, k = r + idx * dil;
    const bool ok = idx >= 0 && k < K;
    float mu = 0.f, rr = 0.f;  // rr = 0 marks a padding row
    if (ok && !own_stats) load_stats(st1, m, base + k, mu, rr);
    s_rs[threadIdx.x] = make_float2(mu, ok ? (own_stats ? 1.f : rr) : 0.f);
  }
  const float a1 = __ldg(alpha1);
  const bool do_stats = stat_out != nullptr;
  const float a2 = (do_stats || rs2_out != nullptr) ? __ldg(alpha2) : 1.f;
  __syncthreads();
  double acc[2] = {0.0, 0.0};
  bool waited = false;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  // (mean, rstd) over the H channels of prelu(row, slope): one warp per row, the summation order of row_stats_kernel
  auto row_stats_of = [&](const T* rowp, float slope, float& mu, float& rstd) {
    float sum = 0.f, q = 0.f;
    if (H <= 512) {  // the lane's (up to 16) values stay in registers between the two passes
      float4 v[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int c = lane * 4 + 128 * t;
        v[t] = c < H ? prelu4(ld4(rowp + c), slope) : make_float4(0.f, 0.f, 0.f, 0.f);
        sum += v[t].x; sum += v[t].y; sum += v[t].z; sum += v[t].w;
      }
      mu = warp_sum(sum) / (float)H;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (lane * 4 + 128 * t < H) {
          const float d0 = v[t].x - mu, d1 = v[t].y - mu, d2 = v[t].z - mu, d3 = v[t].w - mu;
          q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
        }
      }
    } else {
      for (int c = lane * 4; c < H; c += 128) {
        const float4 v = prelu4(ld4(rowp + c), slope);
        sum += v.x; sum += v.y; sum += v.z; sum += v.w;
      }
      mu = warp_sum(sum) / (float)H;
      for (int c = lane * 4; c < H; c += 128) {
        const float4 v = prelu4(ld4(rowp + c), slope);
        const float d0 = v.x - mu, d1 = v.y - mu, d2 = v.z - mu, d3 = v.w - mu;
        q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
      }
    }
    rstd = 1.f / sqrtf(warp_sum(q) / (float)H + CTN_EPS);
  };
  if (own_stats) {
    ew_mbar_wait(&bar, 0);
    waited = true;
    for (int i = wid; i < nrows; i += nwarp) {
      if (s_rs[i].y == 0.f) continue;  // padding row (warp-uniform)
      float mu, rstd;
      row_stats_of(tile + (size_t)i * H, a1, mu, rstd);
      if (lane == 0) {
        s_rs[i] = make_float2(mu, rstd);
        if (i >= cshift && i < cshift + nj)  // tile row i = output frame j0 + i - cshift: written once, by its owner
          reinterpret_cast<float2*>(rs1_out)[base + r + (int64_t)(idx0 + i) * dil] = make_float2(mu, rstd);
      }
    }
    __syncthreads();
  }
  for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) {
    const float4 g = ld4(gamma1 + c), b = ld4(beta1 + c);
    float wd[4][NP_];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int p = 0; p < NP_; ++p) wd[j][p] = p < PP ? Wd[(c + j) * PP + p] : 0.f;
    if (!waited) {  // the parameter loads above overlap the bulk copies
      ew_mbar_wait(&bar, 0);
      waited = true;
    }
    auto row = [&](int i) {  // normalised tile row i (zero for a padding row)
      const float2 stv = s_rs[i];
      if (stv.y == 0.f) return make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 v = prelu4(ld4(tile + (size_t)i * H + c), a1);
      return make_float4(g.x * (v.x - stv.x) * stv.y + b.x, g.y * (v.y - stv.x) * stv.y + b.y,
                         g.z * (v.z - stv.x) * stv.y + b.z, g.w * (v.w - stv.x) * stv.y + b.w);
    };
    float4 w[NP_];  // w[p] = normalised input at tile row jj + p for the current output jj
#pragma unroll
    for (int p = 0; p < NP_; ++p) w[p] = (p < PP - 1 && p < nrows) ? row(p) : make_float4(0.f, 0.f, 0.f, 0.f);
    float s = 0.f, s2 = 0.f;
#pragma unroll 4
    for (int jj = 0; jj < nj; ++jj) {
      const float4 nv = row(jj + PP - 1);
#pragma unroll
      for (int p = 0; p < NP_; ++p)
        if (p == PP - 1) w[p] = nv;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < NP_; ++p) {
        o.x = fmaf(wd[0][p], w[p].x, o.x); o.y = fmaf(wd[1][p], w[p].y, o.y);
        o.z = fmaf(wd[2][p], w[p].z, o.z); o.w = fmaf(wd[3][p], w[p].w, o.w);
      }
      if (sizeof(T) == 2) {  // bf16 storage: the statistics describe the values the consumers will read
        o.x = __bfloat162float(__float2bfloat16_rn(o.x)); o.y = __bfloat162float(__float2bfloat16_rn(o.y));
        o.z = __bfloat162float(__float2bfloat16_rn(o.z)); o.w = __bfloat162float(__float2bfloat16_rn(o.w));
      }
      st4(z2 + (base + r + (int64_t)(j0 + jj) * dil) * H + c, o);
      // (input row jj has left every window of this thread's channels: its shared-memory row now keeps output jj)
      if (rs2_out != nullptr) st4(tile + (size_t)jj * H + c, o);
      if (do_stats) {
        const float4 q = prelu4(o, a2);
        s += (q.x + q.y) + (q.z + q.w);
        s2 += (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w);
      }
#pragma unroll
      for (int p = 0; p + 1 < NP_; ++p) w[p] = w[p + 1];
    }
    acc[0] += (double)s;
    acc[1] += (double)s2;
  }
  if (!waited) ew_mbar_wait(&bar, 0);  // never leave with copies in flight
  if (rs2_out != nullptr) {
    __syncthreads();
    for (int jj = wid; jj < nj; jj += nwarp) {
      float mu, rstd;
      row_stats_of(tile + (size_t)jj * H, a2, mu, rstd);
      if (lane == 0) reinterpret_cast<float2*>(rs2_out)[base + r + (int64_t)(j0 + jj) * dil] = make_float2(mu, rstd);
    }
  }
  if (do_stats) {
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) {
      atomicAdd(stat_out + 2 * m, acc[0]);
      atomicAdd(stat_out + 2 * m + 1, acc[1]);
    }
  }
}

// backward: dn1[k] = sum_p Wd[p] * dz2[k - off_p];  dWd[p] += dz2[k - off_p] * n1[k];
// plus the per-channel / per-sample reductions the norm1 backward needs (dgamma1, dbeta1, red1).
// Same stride-d walk as the forward: the taps of dz2 for input frame k_j are the sequence elements j + cshift - p, a
// P-deep sliding register window; z1 is read once for the centre.  Per-channel sums leave the block as one row of
// `part` ([P+2][H]: taps, dgamma, dbeta) — no atomics; reduce_partials_kernel folds the rows.
template <int PT>
__global__ void __launch_bounds__(256, CTN_DWB_MINB) dwconv_bwd_kernel(const float* __restrict__ dz2, const float* __restrict__ z1,
                                                         const float* __restrict__ alpha1, NormStats st1,
                                                         const float* __restrict__ gamma1, const float* __restrict__ beta1,
                                                         const float* __restrict__ Wd, int K, int H, int P, int dil,
                                                         int cshift, float* __restrict__ dn1, float* __restrict__ part,
                                                         double* __restrict__ red1) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[2 * 32];
  __shared__ float2 s_st;
  const int m = blockIdx.y;
  const int ncls = dw_classes(K, dil);
  const int r = blockIdx.x % ncls, j0 = (blockIdx.x / ncls) * DW_TJ;
  const int nclass = (K - r + dil - 1) / dil;
  const int nj = max(0, min(DW_TJ, nclass - j0));
  if (st1.row == nullptr) {
    if (threadIdx.x == 0) {
      float mu, rr;
      load_stats(st1, m, 0, mu, rr);
      s_st = make_float2(mu, rr);
    }
    __syncthreads();
  }
  const float a1 = __ldg(alpha1);
  const int64_t base = (int64_t)m * K;
  constexpr int NP_ = PT ? PT : MAXP;
  const int PP = PT ? PT : P;
  float* prow = part + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * (int64_t)(PP + 2) * H;
  double acc[2] = {0.0, 0.0};
  for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) {
    const float4 g = ld4(gamma1 + c), b = ld4(beta1 + c);
    float wd[4][NP_];
    float4 dwd[NP_];
#pragma unroll
    for (int p = 0; p < NP_; ++p) {
#pragma unroll
      for (int j = 0; j < 4; ++j) wd[j][p] = p < PP ? Wd[(c + j) * PP + p] : 0.f;
      dwd[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    auto fetch_dz = [&](int idx) {  // dz2 at sequence index idx, zero outside
      const int k = r + idx * dil;
      return (idx >= 0 && k < K) ? ld4(dz2 + (base + k) * H + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // v[i] = dz2 at sequence index j + cshift - (PP-1) + i; tap p reads i = PP-1-p
    float4 v[NP_];
#pragma unroll
    for (int i = 0; i < NP_; ++i) v[i] = i < PP ? fetch_dz(j0 + cshift - (PP - 1) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
    float s = 0.f, s2 = 0.f;
    for (int jj = 0; jj < nj; jj += DW_U) {
      float4 nv[DW_U], zc[DW_U]; float2 stv[DW_U];
#pragma unroll
      for (int u = 0; u < DW_U; ++u) {
        nv[u] = fetch_dz(j0 + jj + u + cshift + 1);  // enters the window after output jj+u
        const int k = r + (j0 + jj + u) * dil;
        const bool ok = jj + u < nj;
        zc[u] = ok ? ld4(z1 + (base + k) * H + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        stv[u] = (ok && st1.row != nullptr) ? reinterpret_cast<const float2*>(st1.row)[base + k] : s_st;
      }
#pragma unroll
      for (int u = 0; u < DW_U; ++u) {
        if (jj + u < nj) {
          const float mu = stv[u].x, rr = stv[u].y;
          const float4 a = prelu4(zc[u], a1);
          const float4 yh = make_float4((a.x - mu) * rr, (a.y - mu) * rr, (a.z - mu) * rr, (a.w - mu) * rr);
          const float4 n1 = make_float4(g.x * yh.x + b.x, g.y * yh.y + b.y, g.z * yh.z + b.z, g.w * yh.w + b.w);
          float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int p = 0; p < NP_; ++p) {
            if (p < PP) {
              float4 t = v[0];
#pragma unroll
              for (int i = 1; i < NP_; ++i)
                if (i == PP - 1 - p) t = v[i];
              d.x = fmaf(wd[0][p], t.x, d.x); d.y = fmaf(wd[1][p], t.y, d.y);
              d.z = fmaf(wd[2][p], t.z, d.z); d.w = fmaf(wd[3][p], t.w, d.w);
              dwd[p].x = fmaf(t.x, n1.x, dwd[p].x); dwd[p].y = fmaf(t.y, n1.y, dwd[p].y);
              dwd[p].z = fmaf(t.z, n1.z, dwd[p].z); dwd[p].w = fmaf(t.w, n1.w, dwd[p].w);
            }
          }
          st4(dn1 + (base + r + (int64_t)(j0 + jj + u) * dil) * H + c, d);
          dg.x = fmaf(d.x, yh.x, dg.x); dg.y = fmaf(d.y, yh.y, dg.y);
          dg.z = fmaf(d.z, yh.z, dg.z); dg.w = fmaf(d.w, yh.w, dg.w);
          db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
          const float4 gh = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
          s += (gh.x + gh.y) + (gh.z + gh.w);
          s2 += (gh.x * yh.x + gh.y * yh.y) + (gh.z * yh.z + gh.w * yh.w);
        }
#pragma unroll
        for (int i = 0; i + 1 < NP_; ++i) v[i] = v[i + 1];
#pragma unroll
        for (int i = 0; i < NP_; ++i)
          if (i == PP - 1) v[i] = nv[u];
      }
    }
#pragma unroll
    for (int p = 0; p < NP_; ++p)
      if (p < PP) st4(prow + (int64_t)p * H + c, dwd[p]);
    st4(prow + (int64_t)PP * H + c, dg);
    st4(prow + (int64_t)(PP + 1) * H + c, db);
    acc[0] += (double)s;
    acc[1] += (double)s2;
  }
  if (red1 != nullptr) {
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) {
      atomicAdd(red1 + 2 * m, acc[0]);
      atomicAdd(red1 + 2 * m + 1, acc[1]);
    }
  }
}


// dwconv_bwd_kernel with both operand tiles staged in shared memory by bulk copies (see dwconv_fwd_bulk_kernel): the
// nj + P - 1 rows of dz2 the block's outputs tap and the nj rows of z1 at the outputs themselves.  Same grid, same
// partial rows as the register-window kernel.
template <int PT>
__global__ void __launch_bounds__(256) dwconv_bwd_bulk_kernel(const float* __restrict__ dz2, const float* __restrict__ z1,
                                                              const float* __restrict__ alpha1, NormStats st1,
                                                              const float* __restrict__ gamma1,
                                                              const float* __restrict__ beta1,
                                                              const float* __restrict__ Wd, int K, int H, int P, int dil,
                                                              int cshift, float* __restrict__ dn1,
                                                              float* __restrict__ part, double* __restrict__ red1) {
  pdl_launch_dependents();
  extern __shared__ __align__(128) uint8_t ew_dsm[];
  __shared__ uint64_t bar;
  __shared__ double red[2 * 32];
  __shared__ float2 s_rs[DW_TJ];          // (mean, rstd) of the z1 rows
  __shared__ int s_ok[DW_TJ + MAXP];      // dz2 tile row inside [0, K)
  const int m = blockIdx.y;
  const int ncls = dw_classes(K, dil);
  const int r = blockIdx.x % ncls, j0 = (blockIdx.x / ncls) * DW_TJ;
  const int nclass = (K - r + dil - 1) / dil;
  const int nj = max(0, min(DW_TJ, nclass - j0));
  constexpr int NP_ = PT ? PT : MAXP;
  const int PP = PT ? PT : P;
  const int nrows = nj > 0 ? nj + PP - 1 : 0;
  const int idx0 = j0 + cshift - (PP - 1);  // sequence index of dz2 tile row 0
  const int64_t base = (int64_t)m * K;
  float* dz_t = reinterpret_cast<float*>(ew_dsm);     // [DW_TJ + PP - 1][H]
  float* z_t = dz_t + (size_t)(DW_TJ + PP - 1) * H;   // [DW_TJ][H]
  if (threadIdx.x == 0) ew_mbar_init(&bar, 1);
  __syncthreads();
  pdl_wait();
  if (threadIdx.x < 32) {
    const int idx = idx0 + (int)threadIdx.x, k = r + idx * dil;
    const bool ok = (int)threadIdx.x < nrows && idx >= 0 && k < K;
    const unsigned mask = __ballot_sync(0xffffffffu, ok);
    if (threadIdx.x == 0) ew_mbar_expect_tx(&bar, (uint32_t)(__popc(mask) + nj) * (uint32_t)(H * 4));
    __syncwarp();
    if (ok) ew_bulk_load(dz_t + (size_t)threadIdx.x * H, dz2 + (base + k) * H, (uint32_t)(H * 4), &bar);
    if ((int)threadIdx.x < nj)
      ew_bulk_load(z_t + (size_t)threadIdx.x * H, z1 + (base + r + (int64_t)(j0 + (int)threadIdx.x) * dil) * H,
                   (uint32_t)(H * 4), &bar);
    if ((int)threadIdx.x < nrows) s_ok[threadIdx.x] = ok ? 1 : 0;
    if ((int)threadIdx.x < nj) {
      float mu, rr;
      load_stats(st1, m, base + r + (int64_t)(j0 + (int)threadIdx.x) * dil, mu, rr);
      s_rs[threadIdx.x] = make_float2(mu, rr);
    }
  }
  const float a1 = __ldg(alpha1);
  float* prow = part + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * (int64_t)(PP + 2) * H;
  __syncthreads();
  double acc[2] = {0.0, 0.0};
  bool waited = false;
  for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) {
    const float4 g = ld4(gamma1 + c), b = ld4(beta1 + c);
    float wd[4][NP_];
    float4 dwd[NP_];
#pragma unroll
    for (int p = 0; p < NP_; ++p) {
#pragma unroll
      for (int j = 0; j < 4; ++j) wd[j][p] = p < PP ? Wd[(c + j) * PP + p] : 0.f;
      dwd[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (!waited) {
      ew_mbar_wait(&bar, 0);
      waited = true;
    }
    auto dzrow = [&](int i) {
      return s_ok[i] ? ld4(dz_t + (size_t)i * H + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // v[i] = dz2 tile row jj + i for the current output jj; tap p reads i = PP-1-p
    float4 v[NP_];
#pragma unroll
    for (int i = 0; i < NP_; ++i) v[i] = (i < PP - 1 && i < nrows) ? dzrow(i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
    float s = 0.f, s2 = 0.f;
#pragma unroll 4
    for (int jj = 0; jj < nj; ++jj) {
      const float4 nv = dzrow(jj + PP - 1);
#pragma unroll
      for (int i = 0; i < NP_; ++i)
        if (i == PP - 1) v[i] = nv;
      const float2 stv = s_rs[jj];
      const float mu = stv.x, rr = stv.y;
      const float4 a = prelu4(ld4(z_t + (size_t)jj * H + c), a1);
      const float4 yh = make_float4((a.x - mu) * rr, (a.y - mu) * rr, (a.z - mu) * rr, (a.w - mu) * rr);
      const float4 n1 = make_float4(g.x * yh.x + b.x, g.y * yh.y + b.y, g.z * yh.z + b.z, g.w * yh.w + b.w);
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < NP_; ++p) {
        if (p < PP) {
          float4 t = v[0];
#pragma unroll
          for (int i = 1; i < NP_; ++i)
            if (i == PP - 1 - p) t = v[i];
          d.x = fmaf(wd[0][p], t.x, d.x); d.y = fmaf(wd[1][p], t.y, d.y);
          d.z = fmaf(wd[2][p], t.z, d.z); d.w = fmaf(wd[3][p], t.w, d.w);
          dwd[p].x = fmaf(t.x, n1.x, dwd[p].x); dwd[p].y = fmaf(t.y, n1.y, dwd[p].y);
          dwd[p].z = fmaf(t.z, n1.z, dwd[p].z); dwd[p].w = fmaf(t.w, n1.w, dwd[p].w);
        }
      }
      st4(dn1 + (base + r + (int64_t)(j0 + jj) * dil) * H + c, d);
      dg.x = fmaf(d.x, yh.x, dg.x); dg.y = fmaf(d.y, yh.y, dg.y);
      dg.z = fmaf(d.z, yh.z, dg.z); dg.w = fmaf(d.w, yh.w, dg.w);
      db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
      const float4 gh = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
      s += (gh.x + gh.y) + (gh.z + gh.w);
      s2 += (gh.x * yh.x + gh.y * yh.y) + (gh.z * yh.z + gh.w * yh.w);
#pragma unroll
      for (int i = 0; i + 1 < NP_; ++i) v[i] = v[i + 1];
    }
#pragma unroll
    for (int p = 0; p < NP_; ++p)
      if (p < PP) st4(prow + (int64_t)p * H + c, dwd[p]);
    st4(prow + (int64_t)PP * H + c, dg);
    st4(prow + (int64_t)(PP + 1) * H + c, db);
    acc[0] += (double)s;
    acc[1] += (double)s2;
  }
  if (!waited) ew_mbar_wait(&bar, 0);
  if (red1 != nullptr) {
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) {
      atomicAdd(red1 + 2 * m, acc[0]);
      atomicAdd(red1 + 2 * m + 1, acc[1]);
    }
  }
}


// dwconv_bwd_bulk_kernel with the gLN backward of norm2 (+ the PReLU in front of it) applied as the rows enter the
// sliding window: the tiles staged are dn2 and z2 (nj + P - 1 rows each) and z1 (nj rows); a thread forms
//   dz2 = r2 (dn2 gamma2 - m1 - yhat2 m2) prelu'(z2),   yhat2 = (prelu(z2) - mu2) r2,   (m1, m2) = red2 / (K H)
// for its own 4 channels from shared memory, so the separate gln_bwd_apply pass over [F, H] (read dn2, read z2, write dz2,
// one launch per block) disappears and nothing waits on global memory after the one bulk round.  dalpha2 is summed over
// the rows the block owns as outputs.  gLN only (per-sample scalars).
template <int PT>
__global__ void __launch_bounds__(256) dwconv_bwd_gln_bulk_kernel(
    const float* __restrict__ dn2, const float* __restrict__ z2, const float* __restrict__ alpha2, NormStats st2,
    const float* __restrict__ gamma2, const double* __restrict__ red2, float* __restrict__ dalpha2,
    const float* __restrict__ z1, const float* __restrict__ alpha1, NormStats st1, const float* __restrict__ gamma1,
    const float* __restrict__ beta1, const float* __restrict__ Wd, int K, int H, int P, int dil, int cshift,
    float* __restrict__ dn1, float* __restrict__ part, double* __restrict__ red1) {
  pdl_launch_dependents();
  extern __shared__ __align__(128) uint8_t ew_dsm[];
  __shared__ uint64_t bar;
  __shared__ double red[3 * 32];
  __shared__ float s_c[6];
  __shared__ int s_ok[DW_TJ + MAXP];  // dn2 / z2 tile row inside [0, K)
  const int m = blockIdx.y;
  const int ncls = dw_classes(K, dil);
  const int r = blockIdx.x % ncls, j0 = (blockIdx.x / ncls) * DW_TJ;
  const int nclass = (K - r + dil - 1) / dil;
  const int nj = max(0, min(DW_TJ, nclass - j0));
  constexpr int NP_ = PT ? PT : MAXP;
  const int PP = PT ? PT : P;
  const int nrows = nj > 0 ? nj + PP - 1 : 0;
  const int idx0 = j0 + cshift - (PP - 1);  // sequence index of tile row 0
  const int own0 = PP - 1 - cshift;         // tile rows [own0, own0 + nj) are the block's own output frames
  const int64_t base = (int64_t)m * K;
  float* dn_t = reinterpret_cast<float*>(ew_dsm);      // [DW_TJ + PP - 1][H]
  float* z2_t = dn_t + (size_t)(DW_TJ + PP - 1) * H;   // [DW_TJ + PP - 1][H]
  float* z1_t = z2_t + (size_t)(DW_TJ + PP - 1) * H;   // [DW_TJ][H]
  if (threadIdx.x == 0) ew_mbar_init(&bar, 1);
  __syncthreads();
  pdl_wait();
  if (threadIdx.x < 32) {
    const int idx = idx0 + (int)threadIdx.x, k = r + idx * dil;
    const bool ok = (int)threadIdx.x < nrows && idx >= 0 && k < K;
    const unsigned mask = __ballot_sync(0xffffffffu, ok);
    if (threadIdx.x == 0) ew_mbar_expect_tx(&bar, (uint32_t)(2 * __popc(mask) + nj) * (uint32_t)(H * 4));
    __syncwarp();
    if (ok) {
      ew_bulk_load(dn_t + (size_t)threadIdx.x * H, dn2 + (base + k) * H, (uint32_t)(H * 4), &bar);
      ew_bulk_load(z2_t + (size_t)threadIdx.x * H, z2 + (base + k) * H, (uint32_t)(H * 4), &bar);
    }
    if ((int)threadIdx.x < nj)
      ew_bulk_load(z1_t + (size_t)threadIdx.x * H, z1 + (base + r + (int64_t)(j0 + (int)threadIdx.x) * dil) * H,
                   (uint32_t)(H * 4), &bar);
    if ((int)threadIdx.x < nrows) s_ok[threadIdx.x] = ok ? 1 : 0;
    if (threadIdx.x == 0) {
      float mu, rr;
      load_stats(st1, m, 0, mu, rr);
      s_c[0] = mu; s_c[1] = rr;
      load_stats(st2, m, 0, mu, rr);
      s_c[2] = mu; s_c[3] = rr;
      const double cnt = (double)K * (double)H;
      s_c[4] = (float)(red2[2 * m] / cnt);
      s_c[5] = (float)(red2[2 * m + 1] / cnt);
    }
  }
  const float a1 = __ldg(alpha1), a2 = __ldg(alpha2);
  float* prow = part + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * (int64_t)(PP + 2) * H;
  __syncthreads();
  const float mu1 = s_c[0], r1 = s_c[1], mu2 = s_c[2], r2 = s_c[3], m1 = s_c[4], m2 = s_c[5];
  double acc[3] = {0.0, 0.0, 0.0};
  bool waited = false;
  for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) {
    const float4 g = ld4(gamma1 + c), b = ld4(beta1 + c), g2 = ld4(gamma2 + c);
    float wd[4][NP_];
    float4 dwd[NP_];
#pragma unroll
    for (int p = 0; p < NP_; ++p) {
#pragma unroll
      for (int j = 0; j < 4; ++j) wd[j][p] = p < PP ? Wd[(c + j) * PP + p] : 0.f;
      dwd[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (!waited) {
      ew_mbar_wait(&bar, 0);
      waited = true;
    }
    float sa = 0.f;
    auto dzrow = [&](int i) {  // dz2 of tile row i, formed from the staged dn2 / z2 rows (zero outside [0, K))
      if (!s_ok[i]) return make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 d = ld4(dn_t + (size_t)i * H + c), zz = ld4(z2_t + (size_t)i * H + c);
      const float4 v = prelu4(zz, a2);
      float4 da;
      da.x = r2 * (d.x * g2.x - m1 - (v.x - mu2) * r2 * m2);
      da.y = r2 * (d.y * g2.y - m1 - (v.y - mu2) * r2 * m2);
      da.z = r2 * (d.z * g2.z - m1 - (v.z - mu2) * r2 * m2);
      da.w = r2 * (d.w * g2.w - m1 - (v.w - mu2) * r2 * m2);
      if (i >= own0 && i < own0 + nj)  // the block owns this frame as an output: count its PReLU-slope gradient once
        sa += (zz.x > 0.f ? 0.f : da.x * zz.x) + (zz.y > 0.f ? 0.f : da.y * zz.y) +
              (zz.z > 0.f ? 0.f : da.z * zz.z) + (zz.w > 0.f ? 0.f : da.w * zz.w);
      return make_float4(da.x * dprelu(zz.x, a2), da.y * dprelu(zz.y, a2), da.z * dprelu(zz.z, a2),
                         da.w * dprelu(zz.w, a2));
    };
    float4 v[NP_];  // v[i] = dz2 tile row jj + i for the current output jj; tap p reads i = PP-1-p
#pragma unroll
    for (int i = 0; i < NP_; ++i) v[i] = (i < PP - 1 && i < nrows) ? dzrow(i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
    float s = 0.f, s2 = 0.f;
#pragma unroll 4
    for (int jj = 0; jj < nj; ++jj) {
      const float4 nv = dzrow(jj + PP - 1);
#pragma unroll
      for (int i = 0; i < NP_; ++i)
        if (i == PP - 1) v[i] = nv;
      const float4 a = prelu4(ld4(z1_t + (size_t)jj * H + c), a1);
      const float4 yh = make_float4((a.x - mu1) * r1, (a.y - mu1) * r1, (a.z - mu1) * r1, (a.w - mu1) * r1);
      const float4 n1 = make_float4(g.x * yh.x + b.x, g.y * yh.y + b.y, g.z * yh.z + b.z, g.w * yh.w + b.w);
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < NP_; ++p) {
        if (p < PP) {
          float4 t = v[0];
#pragma unroll
          for (int i = 1; i < NP_; ++i)
            if (i == PP - 1 - p) t = v[i];
          d.x = fmaf(wd[0][p], t.x, d.x); d.y = fmaf(wd[1][p], t.y, d.y);
          d.z = fmaf(wd[2][p], t.z, d.z); d.w = fmaf(wd[3][p], t.w, d.w);
          dwd[p].x = fmaf(t.x, n1.x, dwd[p].x); dwd[p].y = fmaf(t.y, n1.y, dwd[p].y);
          dwd[p].z = fmaf(t.z, n1.z, dwd[p].z); dwd[p].w = fmaf(t.w, n1.w, dwd[p].w);
        }
      }
      st4(dn1 + (base + r + (int64_t)(j0 + jj) * dil) * H + c, d);
      dg.x = fmaf(d.x, yh.x, dg.x); dg.y = fmaf(d.y, yh.y, dg.y);
      dg.z = fmaf(d.z, yh.z, dg.z); dg.w = fmaf(d.w, yh.w, dg.w);
      db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
      const float4 gh = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
      s += (gh.x + gh.y) + (gh.z + gh.w);
      s2 += (gh.x * yh.x + gh.y * yh.y) + (gh.z * yh.z + gh.w * yh.w);
#pragma unroll
      for (int i = 0; i + 1 < NP_; ++i) v[i] = v[i + 1];
    }
#pragma unroll
    for (int p = 0; p < NP_; ++p)
      if (p < PP) st4(prow + (int64_t)p * H + c, dwd[p]);
    st4(prow + (int64_t)PP * H + c, dg);
    st4(prow + (int64_t)(PP + 1) * H + c, db);
    acc[0] += (double)s;
    acc[1] += (double)s2;
    acc[2] += (double)sa;
  }
  if (!waited) ew_mbar_wait(&bar, 0);
  block_sum<3>(acc, red);
  if (threadIdx.x == 0) {
    if (red1 != nullptr) {
      atomicAdd(red1 + 2 * m, acc[0]);
      atomicAdd(red1 + 2 * m + 1, acc[1]);
    }
    atomicAdd(dalpha2, (float)acc[2]);
  }
}

// fold `nb` partial rows of [Q][H] floats: q < P -> dW[c*P + q], q == P -> dgamma[c], q == P+1 -> dbeta[c]
// grid (ceil(Q*H / 256), splits); each block sums a slice of the rows and adds it atomically (few atomics per output)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ part, int nb, int H, int P,
                                                              float* __restrict__ dW, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = (P + 2) * H;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int per = (nb + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * per, b1 = min(nb, b0 + per);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int b = b0;
  for (; b + 3 < b1; b += 4) {
    s0 += part[(int64_t)b * n + i];
    s1 += part[(int64_t)(b + 1) * n + i];
    s2 += part[(int64_t)(b + 2) * n + i];
    s3 += part[(int64_t)(b + 3) * n + i];
  }
  for (; b < b1; ++b) s0 += part[(int64_t)b * n + i];
  const float v = (s0 + s1) + (s2 + s3);
  const int q = i / H, c = i - q * H;
  if (q < P) atomicAdd(dW + c * P + q, v);
  else if (q == P) atomicAdd(dgamma + c, v);
  else atomicAdd(dbeta + c, v);
}

// ---------------------------------------------------------------------------------------
// norm backward, reduction pass: per-channel sums (dn*yhat, dn) leave the block as a [2][Ch] row of `part`
// (folded by reduce_partials_kernel with P = 0); red[m] += (sum dn*gamma, sum dn*gamma*yhat)
// ---------------------------------------------------------------------------------------
#ifndef CTN_NR_TK
#define CTN_NR_TK 16
#endif
constexpr int NR_TK = CTN_NR_TK;
#ifndef CTN_NR_U
#define CTN_NR_U 4
#endif
constexpr int NR_U = CTN_NR_U;  // frames (2 x NR_U independent 16-byte loads) in flight per thread
__global__ void __launch_bounds__(256) norm_bwd_reduce_kernel(const float* __restrict__ dn, const float* __restrict__ z,
                                                              const float* __restrict__ alpha, NormStats st,
                                                              const float* __restrict__ gamma, int K, int Ch,
                                                              float* __restrict__ part, double* __restrict__ redout) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[2 * 32];
  __shared__ float2 s_st;
  const int m = blockIdx.y, k0 = blockIdx.x * NR_TK;
  const int nk = min(NR_TK, K - k0);
  if (st.row == nullptr) {
    if (threadIdx.x == 0) {
      float mu, r;
      load_stats(st, m, 0, mu, r);
      s_st = make_float2(mu, r);
    }
    __syncthreads();
  }
  const bool hasp = alpha != nullptr;
  const float a = hasp ? __ldg(alpha) : 1.f;
  const int64_t base = (int64_t)m * K;
  float* prow = part + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * (int64_t)2 * Ch;
  double acc[2] = {0.0, 0.0};
  for (int c = threadIdx.x * 4; c < Ch; c += blockDim.x * 4) {
    const float4 g = ld4(gamma + c);
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
    float s = 0.f, s2 = 0.f;
    for (int kk = 0; kk < nk; kk += NR_U) {
      float4 zv[NR_U], dv[NR_U];
      float2 stv[NR_U];
#pragma unroll
      for (int u = 0; u < NR_U; ++u) {
        const bool vk = kk + u < nk;
        const int64_t f = base + k0 + kk + u;
        zv[u] = vk ? ld4(z + f * Ch + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        dv[u] = vk ? ld4(dn + f * Ch + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        stv[u] = (vk && st.row != nullptr) ? reinterpret_cast<const float2*>(st.row)[f] : s_st;
      }
#pragma unroll
      for (int u = 0; u < NR_U; ++u) {
        if (kk + u >= nk) break;
        const float mu = stv[u].x, r = stv[u].y;
        const float4 v = hasp ? prelu4(zv[u], a) : zv[u];
        const float4 d = dv[u];
        const float4 yh = make_float4((v.x - mu) * r, (v.y - mu) * r, (v.z - mu) * r, (v.w - mu) * r);
        dg.x = fmaf(d.x, yh.x, dg.x); dg.y = fmaf(d.y, yh.y, dg.y);
        dg.z = fmaf(d.z, yh.z, dg.z); dg.w = fmaf(d.w, yh.w, dg.w);
        db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
        const float4 gh = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
        s += (gh.x + gh.y) + (gh.z + gh.w);
        s2 += (gh.x * yh.x + gh.y * yh.y) + (gh.z * yh.z + gh.w * yh.w);
      }
    }
    st4(prow + c, dg);
    st4(prow + Ch + c, db);
    acc[0] += (double)s;
    acc[1] += (double)s2;
  }
  if (redout != nullptr) {
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) {
      atomicAdd(redout + 2 * m, acc[0]);
      atomicAdd(redout + 2 * m + 1, acc[1]);
    }
  }
}

// apply pass, gLN: dz = r*(dn*gamma - m1 - yhat*m2) * prelu'(z); dalpha += sum da * z * [z<=0]
#ifndef CTN_GA_TK
#define CTN_GA_TK 16
#endif
constexpr int GA_TK = CTN_GA_TK;
#ifndef CTN_GA_U
#define CTN_GA_U 4
#endif
constexpr int GA_U = CTN_GA_U;
#ifdef CTN_GA_MINB
#define CTN_GA_BOUNDS __launch_bounds__(256, CTN_GA_MINB)
#else
#define CTN_GA_BOUNDS __launch_bounds__(256)
#endif
__global__ void CTN_GA_BOUNDS gln_bwd_apply_kernel(float* __restrict__ dn, const float* __restrict__ z,
                                                            const float* __restrict__ alpha, NormStats st,
                                                            const float* __restrict__ gamma, const double* __restrict__ redin,
                                                            int K, int Ch, float* __restrict__ dalpha) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[32];
  __shared__ float4 s_st;
  const int m = blockIdx.y, k0 = blockIdx.x * GA_TK;
  const int nk = min(GA_TK, K - k0);
  if (threadIdx.x == 0) {
    float mu, r;
    load_stats(st, m, 0, mu, r);
    const double cnt = (double)K * (double)Ch;
    s_st = make_float4(mu, r, (float)(redin[2 * m] / cnt), (float)(redin[2 * m + 1] / cnt));
  }
  __syncthreads();
  const float mu = s_st.x, r = s_st.y, m1 = s_st.z, m2 = s_st.w;
  const bool hasp = alpha != nullptr;
  const float a = hasp ? __ldg(alpha) : 1.f;
  const int64_t base = (int64_t)m * K;
  double acc[1] = {0.0};
  for (int c = threadIdx.x * 4; c < Ch; c += blockDim.x * 4) {
    const float4 g = ld4(gamma + c);
    float s = 0.f;
    for (int kk = 0; kk < nk; kk += GA_U) {  // GA_U frames (2 x GA_U independent 16-byte loads) in flight per thread
      float4 zv[GA_U], dv[GA_U];
#pragma unroll
      for (int u = 0; u < GA_U; ++u) {
        const bool vk = kk + u < nk;
        const int64_t f = base + k0 + kk + u;
        zv[u] = vk ? ld4(z + f * Ch + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        dv[u] = vk ? ld4(dn + f * Ch + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < GA_U; ++u) {
        if (kk + u >= nk) break;
        const int64_t f = base + k0 + kk + u;
        const float4 zz = zv[u], d = dv[u];
        const float4 v = hasp ? prelu4(zz, a) : zz;
        float4 da;
        da.x = r * (d.x * g.x - m1 - (v.x - mu) * r * m2);
        da.y = r * (d.y * g.y - m1 - (v.y - mu) * r * m2);
        da.z = r * (d.z * g.z - m1 - (v.z - mu) * r * m2);
        da.w = r * (d.w * g.w - m1 - (v.w - mu) * r * m2);
        if (hasp) {
          s += (zz.x > 0.f ? 0.f : da.x * zz.x) + (zz.y > 0.f ? 0.f : da.y * zz.y) +
               (zz.z > 0.f ? 0.f : da.z * zz.z) + (zz.w > 0.f ? 0.f : da.w * zz.w);
          da.x *= dprelu(zz.x, a); da.y *= dprelu(zz.y, a); da.z *= dprelu(zz.z, a); da.w *= dprelu(zz.w, a);
        }
        st4(dn + f * Ch + c, da);
      }
    }
    acc[0] += (double)s;
  }
  if (hasp) {
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) atomicAdd(dalpha, (float)acc[0]);
  }
}

// apply pass, cLN: one warp per frame (means over channels inside the warp)
__global__ void __launch_bounds__(256) cln_bwd_apply_kernel(float* __restrict__ dn, const float* __restrict__ z,
                                                            const float* __restrict__ alpha, const float* __restrict__ rowstat,
                                                            const float* __restrict__ gamma, int64_t F, int Ch,
                                                            float* __restrict__ dalpha) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[32];
  const int lane = threadIdx.x & 31;
  const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool hasp = alpha != nullptr;
  const float a = hasp ? __ldg(alpha) : 1.f;
  float sa = 0.f;
  if (f < F) {
    const float mu = rowstat[2 * f], r = rowstat[2 * f + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane * 4; c < Ch; c += 128) {
      const float4 g = ld4(gamma + c), d = ld4(dn + f * Ch + c);
      float4 v = ld4(z + f * Ch + c);
      if (hasp) v = prelu4(v, a);
      const float4 gh = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
      s1 += (gh.x + gh.y) + (gh.z + gh.w);
      s2 += gh.x * (v.x - mu) * r + gh.y * (v.y - mu) * r + gh.z * (v.z - mu) * r + gh.w * (v.w - mu) * r;
    }
    const float m1 = warp_sum(s1) / (float)Ch, m2 = warp_sum(s2) / (float)Ch;
    for (int c = lane * 4; c < Ch; c += 128) {
      const float4 g = ld4(gamma + c), d = ld4(dn + f * Ch + c);
      const float4 zz = ld4(z + f * Ch + c);
      const float4 v = hasp ? prelu4(zz, a) : zz;
      float4 da;
      da.x = r * (d.x * g.x - m1 - (v.x - mu) * r * m2);
      da.y = r * (d.y * g.y - m1 - (v.y - mu) * r * m2);
      da.z = r * (d.z * g.z - m1 - (v.z - mu) * r * m2);
      da.w = r * (d.w * g.w - m1 - (v.w - mu) * r * m2);
      if (hasp) {
        sa += (zz.x > 0.f ? 0.f : da.x * zz.x) + (zz.y > 0.f ? 0.f : da.y * zz.y) +
              (zz.z > 0.f ? 0.f : da.z * zz.z) + (zz.w > 0.f ? 0.f : da.w * zz.w);
        da.x *= dprelu(zz.x, a); da.y *= dprelu(zz.y, a); da.z *= dprelu(zz.z, a); da.w *= dprelu(zz.w, a);
      }
      st4(dn + f * Ch + c, da);
    }
  }
  if (hasp) {
    double acc[1] = {(double)sa};
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) atomicAdd(dalpha, (float)acc[0]);
  }
}

// utils.overlap_and_add as a standalone op: out[o, t] = sum_k sig[o, k, t - k*step], ascending k
__global__ void __launch_bounds__(256) ola_kernel(const float* __restrict__ sig, int frames, int flen, int step,
                                                  int64_t out_len, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t o = blockIdx.y;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= out_len) return;
  int64_t khi = t / step;
  if (khi > frames - 1) khi = frames - 1;
  int64_t klo = t - flen + 1 <= 0 ? 0 : (t - flen + step) / step;
  float acc = 0.f;
  for (int64_t k = klo; k <= khi; ++k) acc += sig[(o * frames + k) * flen + (t - k * step)];
  out[o * out_len + t] = acc;
}

// which depthwise kernels stage their tiles by bulk copies (default: both; measured on B200, paper config, M = 3 x 4 s:
// forward 17.3 -> 14.1 us, backward 24.1 -> 19.8 us in place, graph-replayed step 6.00 -> 5.75 ms).  CTN_DW_BULK (A/B):
// bit 0 = forward, bit 1 = backward; 0 = the register-window kernels.
static int dw_bulk_mask() {
  static const int v = getenv("CTN_DW_BULK") ? atoi(getenv("CTN_DW_BULK")) : 3;
  return v;
}

static int block_for_channels(int Ch) {
  int t = ((Ch / 4 + 31) / 32) * 32;
  return t > 256 ? 256 : (t < 32 ? 32 : t);
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host launchers (C linkage wrappers live in c_api.cu)
// ---------------------------------------------------------------------------------------
int run_row_stats(const float* x, const float* alpha, int64_t F, int Ch, float* rowstat, cudaStream_t s, int bf16) {
  CTN_REQUIRE(Ch % 4 == 0, "row_stats: channels must be a multiple of 4 (got %d)", Ch);
  if (bf16)
    launch_kernel(row_stats_kernel<__nv_bfloat16>, cdiv(F, 8 * CTN_RS_RPW_BF16), 256, 0, s, reinterpret_cast<const __nv_bfloat16*>(x), alpha, F, Ch, rowstat);
  else
    launch_kernel(row_stats_kernel<float>, cdiv(F, 8 * 2), 256, 0, s, x, alpha, F, Ch, rowstat);
  return check_launch("row_stats_kernel");
}

int run_prep_normfold(const float* W, const float* gamma, const float* beta, int O, int I, int nb, int64_t in_stride,
                      float* Wg, float* c1, float* c2, int64_t wg_stride, int64_t c_stride, cudaStream_t s) {
  launch_kernel(prep_normfold_kernel, dim3(cdiv(O, 8), nb), 256, 0, s, W, gamma, beta, O, I, in_stride, Wg, c1, c2, wg_stride, c_stride);
  return check_launch("prep_normfold_kernel");
}

// true when run_dwconv_fwd can form the cLN per-frame statistics itself (the bulk-staged kernel): the caller then skips
// its two row_stats launches per block and passes the statistics buffers as rs1_out / rs2_out
// Measured (B200, causal cLN, 32 x 4 s forward): fp32 storage 13.16 -> 12.14 ms with the statistics formed in the
// depthwise kernel; bf16 storage 9.49 -> 9.63 ms (the two row_stats passes it replaces read half the bytes, the extra
// work in the kernel is the same) — so the default fuses for fp32-stored activations only.  CTN_ROWSTAT_FUSION=0 / 1
// forces it off / on for both (A/B).
bool dwconv_fwd_fuses_rowstats(int H, int P, int bf16) {
  static const int mode = getenv("CTN_ROWSTAT_FUSION") == nullptr ? -1 : (getenv("CTN_ROWSTAT_FUSION")[0] == '0' ? 0 : 1);
  if (mode == 0 || (mode < 0 && bf16)) return false;
  const size_t esz = bf16 ? 2 : 4;
  return (dw_bulk_mask() & 1) && ((size_t)H * esz) % 16 == 0 && (size_t)(DWF_TJ + P - 1) * H * esz <= 160 * 1024 &&
         P >= 1 && P <= MAXP;
}

int run_dwconv_fwd(const float* z1, const float* alpha1, NormStats st1, const float* gamma1, const float* beta1,
                   const float* Wd, int M, int K, int H, int P, int dil, int causal, float* z2, double* stat_out,
                   const float* alpha2, cudaStream_t s, int bf16, float* rs1_out, float* rs2_out) {
  CTN_REQUIRE((rs1_out == nullptr && rs2_out == nullptr) || (dwconv_fwd_fuses_rowstats(H, P, bf16) && alpha2 != nullptr),
              "dwconv_fwd: in-kernel row statistics need the bulk-staged kernel (and alpha2)");
  CTN_REQUIRE(H % 4 == 0, "dwconv: H must be a multiple of 4 (got %d)", H);
  CTN_REQUIRE(P >= 1 && P <= MAXP, "dwconv: kernel size P must be in [1,%d] (got %d)", MAXP, P);
  CTN_REQUIRE(causal || (P % 2 == 1), "dwconv: non-causal needs odd P (reference output length changes otherwise)");
  const int cshift = causal ? P - 1 : (P - 1) / 2;
  const dim3 grid(dw_blocks(K, dil, DWF_TJ), M);
  // bulk-staged variant: the block's rows come to shared memory by cp.async.bulk (rows must be multiples of 16 bytes)
  const size_t esz = bf16 ? 2 : 4;
  const size_t bulk_smem = (size_t)(DWF_TJ + P - 1) * H * esz;
  if ((dw_bulk_mask() & 1) && ((size_t)H * esz) % 16 == 0 && bulk_smem <= 160 * 1024) {
#define CTN_DWF_BULK(PT, TY, zi, zo)                                                                                   \
  do {                                                                                                                 \
    if (bulk_smem > 48 * 1024)                                                                                         \
      CTN_CUDA(cudaFuncSetAttribute(dwconv_fwd_bulk_kernel<PT, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                    (int)bulk_smem));                                                                  \
    launch_kernel(dwconv_fwd_bulk_kernel<PT, TY>, grid, block_for_channels(H), bulk_smem, s, zi, alpha1, st1, gamma1,  \
                  beta1, Wd, K, H, P, dil, cshift, zo, stat_out, alpha2, rs1_out, rs2_out);                            \
  } while (0)
    if (bf16) {
      const __nv_bfloat16* zi = reinterpret_cast<const __nv_bfloat16*>(z1);
      __nv_bfloat16* zo = reinterpret_cast<__nv_bfloat16*>(z2);
      if (P == 3) CTN_DWF_BULK(3, __nv_bfloat16, zi, zo);
      else CTN_DWF_BULK(0, __nv_bfloat16, zi, zo);
      return check_launch("dwconv_fwd_bulk_kernel<bf16>");
    }
    if (P == 3) CTN_DWF_BULK(3, float, z1, z2);
    else CTN_DWF_BULK(0, float, z1, z2);
#undef CTN_DWF_BULK
    return check_launch("dwconv_fwd_bulk_kernel");
  }
  if (bf16) {  // reduced-precision inference: z1 and z2 are stored as bf16 (the pointers are reinterpreted)
    const __nv_bfloat16* zi = reinterpret_cast<const __nv_bfloat16*>(z1);
    __nv_bfloat16* zo = reinterpret_cast<__nv_bfloat16*>(z2);
    if (P == 3)
      launch_kernel(dwconv_fwd_kernel<3, __nv_bfloat16, 8>, grid, block_for_channels(H), 0, s, zi, alpha1, st1, gamma1, beta1, Wd, K, H, P, dil, cshift, zo, stat_out, alpha2);
    else
      launch_kernel(dwconv_fwd_kernel<0, __nv_bfloat16, 8>, grid, block_for_channels(H), 0, s, zi, alpha1, st1, gamma1, beta1, Wd, K, H, P, dil, cshift, zo, stat_out, alpha2);
    return check_launch("dwconv_fwd_kernel<bf16>");
  }
  if (P == 3)
    launch_kernel(dwconv_fwd_kernel<3>, grid, block_for_channels(H), 0, s, z1, alpha1, st1, gamma1, beta1, Wd, K, H, P, dil, cshift, z2, stat_out, alpha2);
  else
    launch_kernel(dwconv_fwd_kernel<0>, grid, block_for_channels(H), 0, s, z1, alpha1, st1, gamma1, beta1, Wd, K, H, P, dil, cshift, z2, stat_out, alpha2);
  return check_launch("dwconv_fwd_kernel");
}

// the same fold for a batch of independent (partial rows -> gradient slots) entries in ONE launch: blockIdx.z = entry
__global__ void __launch_bounds__(256) reduce_partials_batch_kernel(FoldBatch fb) {
  pdl_launch_dependents();
  pdl_wait();
  const FoldEntry& e = fb.e[blockIdx.z];
  const int n = (e.P + 2) * e.H;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int per = (e.nb + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * per, b1 = min(e.nb, b0 + per);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int b = b0;
  for (; b + 3 < b1; b += 4) {
    s0 += e.part[(int64_t)b * n + i];
    s1 += e.part[(int64_t)(b + 1) * n + i];
    s2 += e.part[(int64_t)(b + 2) * n + i];
    s3 += e.part[(int64_t)(b + 3) * n + i];
  }
  for (; b < b1; ++b) s0 += e.part[(int64_t)b * n + i];
  const float v = (s0 + s1) + (s2 + s3);
  const int q = i / e.H, c = i - q * e.H;
  if (q < e.P) atomicAdd(e.dW + c * e.P + q, v);
  else if (q == e.P) atomicAdd(e.dgamma + c, v);
  else atomicAdd(e.dbeta + c, v);
}

static int fold_partials(const float* part, int nb, int H, int P, float* dW, float* dgamma, float* dbeta,
                         cudaStream_t s) {
  int splits = nb / 32;
  splits = splits < 1 ? 1 : (splits > 16 ? 16 : splits);
  launch_kernel(reduce_partials_kernel, dim3(cdiv((int64_t)(P + 2) * H, 256), splits), 256, 0, s, part, nb, H, P, dW, dgamma, dbeta);
  return check_launch("reduce_partials_kernel");
}

int64_t dwconv_bwd_partial_floats(int M, int K, int H, int P, int dil) {
  return (int64_t)dw_blocks(K, dil) * M * (P + 2) * H;
}
int64_t norm_bwd_partial_floats(int M, int K, int Ch) { return (int64_t)cdiv(K, NR_TK) * M * 2 * Ch; }

int run_dwconv_bwd(const float* dz2, const float* z1, const float* alpha1, NormStats st1, const float* gamma1,
                   const float* beta1, const float* Wd, int M, int K, int H, int P, int dil, int causal, float* dn1,
                   float* dWd, float* dgamma1, float* dbeta1, double* red1, float* part, int defer_fold,
                   cudaStream_t s) {
  CTN_REQUIRE(H % 4 == 0 && P >= 1 && P <= MAXP, "dwconv_bwd: bad H/P (%d/%d)", H, P);
  if (part == nullptr) {  // standalone call: library-owned scratch
    void* scr = nullptr;
    CTN_TRY(lib_scratch((size_t)dwconv_bwd_partial_floats(M, K, H, P, dil) * 4, &scr, 1));
    part = reinterpret_cast<float*>(scr);
  }
  const int cshift = causal ? P - 1 : (P - 1) / 2;
  const dim3 grid(dw_blocks(K, dil), M);
  const size_t bulk_smem = (size_t)(2 * DW_TJ + P - 1) * H * 4;
  if ((dw_bulk_mask() & 2) && bulk_smem <= 160 * 1024) {
    if (P == 3) {
      if (bulk_smem > 48 * 1024)
        CTN_CUDA(cudaFuncSetAttribute(dwconv_bwd_bulk_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));
      launch_kernel(dwconv_bwd_bulk_kernel<3>, grid, block_for_channels(H), bulk_smem, s, dz2, z1, alpha1, st1, gamma1, beta1, Wd, K, H, P, dil, cshift, dn1, part, red1);
    } else {
      if (bulk_smem > 48 * 1024)
        CTN_CUDA(cudaFuncSetAttribute(dwconv_bwd_bulk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));
      launch_kernel(dwconv_bwd_bulk_kernel<0>, grid, block_for_channels(H), bulk_smem, s, dz2, z1, alpha1, st1, gamma1, beta1, Wd, K, H, P, dil, cshift, dn1, part, red1);
    }
    CTN_TRY(check_launch("dwconv_bwd_bulk_kernel"));
    if (defer_fold) return 0;
    return fold_partials(part, grid.x * grid.y, H, P, dWd, dgamma1, dbeta1, s);
  }
  if (P == 3)
    launch_kernel(dwconv_bwd_kernel<3>, grid, block_for_channels(H), 0, s, dz2, z1, alpha1, st1, gamma1, beta1, Wd, K, H, P, dil, cshift, dn1, part, red1);
  else
    launch_kernel(dwconv_bwd_kernel<0>, grid, block_for_channels(H), 0, s, dz2, z1, alpha1, st1, gamma1, beta1, Wd, K, H, P, dil, cshift, dn1, part, red1);
  CTN_TRY(check_launch("dwconv_bwd_kernel"));
  if (defer_fold) return 0;  // the caller folds a whole stage's partial rows in one launch (run_fold_batch)
  return fold_partials(part, grid.x * grid.y, H, P, dWd, dgamma1, dbeta1, s);
}

// dwconv backward with the gLN backward of norm2 fused on load (dwconv_bwd_gln_bulk_kernel); both norms gLN
int run_dwconv_bwd_gln_fused(const float* dn2, const float* z2, const float* alpha2, NormStats st2, const float* gamma2,
                             const double* red2, float* dalpha2, const float* z1, const float* alpha1, NormStats st1,
                             const float* gamma1, const float* beta1, const float* Wd, int M, int K, int H, int P, int dil,
                             int causal, float* dn1, float* dWd, float* dgamma1, float* dbeta1, double* red1, float* part,
                             int defer_fold, cudaStream_t s) {
  CTN_REQUIRE(H % 4 == 0 && P >= 1 && P <= MAXP, "dwconv_bwd: bad H/P (%d/%d)", H, P);
  CTN_REQUIRE(st1.row == nullptr && st2.row == nullptr && st1.acc != nullptr && st2.acc != nullptr && red2 != nullptr,
              "dwconv_bwd_gln_fused: both norms must be gLN (per-sample statistics)");
  if (part == nullptr) {
    void* scr = nullptr;
    CTN_TRY(lib_scratch((size_t)dwconv_bwd_partial_floats(M, K, H, P, dil) * 4, &scr, 1));
    part = reinterpret_cast<float*>(scr);
  }
  const int cshift = causal ? P - 1 : (P - 1) / 2;
  const dim3 grid(dw_blocks(K, dil), M);
  const size_t bulk_smem = (size_t)(3 * DW_TJ + 2 * (P - 1)) * H * 4;  // dn2, z2 (TJ + P - 1 rows each) and z1 (TJ rows)
  CTN_REQUIRE(bulk_smem <= 200 * 1024, "dwconv_bwd_gln_fused: H = %d needs %zu bytes of shared memory", H, bulk_smem);
  if (P == 3) {
    if (bulk_smem > 48 * 1024)
      CTN_CUDA(cudaFuncSetAttribute(dwconv_bwd_gln_bulk_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));
    launch_kernel(dwconv_bwd_gln_bulk_kernel<3>, grid, block_for_channels(H), bulk_smem, s, dn2, z2, alpha2, st2, gamma2, red2, dalpha2, z1, alpha1, st1, gamma1, beta1, Wd, K, H, P, dil, cshift, dn1, part, red1);
  } else {
    if (bulk_smem > 48 * 1024)
      CTN_CUDA(cudaFuncSetAttribute(dwconv_bwd_gln_bulk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));
    launch_kernel(dwconv_bwd_gln_bulk_kernel<0>, grid, block_for_channels(H), bulk_smem, s, dn2, z2, alpha2, st2, gamma2, red2, dalpha2, z1, alpha1, st1, gamma1, beta1, Wd, K, H, P, dil, cshift, dn1, part, red1);
  }
  CTN_TRY(check_launch("dwconv_bwd_gln_bulk_kernel"));
  if (defer_fold) return 0;
  return fold_partials(part, grid.x * grid.y, H, P, dWd, dgamma1, dbeta1, s);
}

int run_norm_bwd_reduce(const float* dn, const float* z, const float* alpha, NormStats st, const float* gamma, int M,
                        int K, int Ch, float* dgamma, float* dbeta, double* red, float* part, int defer_fold,
                        cudaStream_t s) {
  CTN_REQUIRE(Ch % 4 == 0, "norm_bwd: channels must be a multiple of 4 (got %d)", Ch);
  if (part == nullptr) {
    void* scr = nullptr;
    CTN_TRY(lib_scratch((size_t)norm_bwd_partial_floats(M, K, Ch) * 4, &scr, 1));
    part = reinterpret_cast<float*>(scr);
  }
  const dim3 grid(cdiv(K, NR_TK), M);
  // (a bulk-copy staged variant of this kernel — two 32 KB copies per block, 3 blocks per SM — measured slower: 14.8 vs
  // 13.9 us in place, step 5.86 vs 5.80 ms; with 8 loads per thread in flight and 5 blocks per SM the register path wins)
  launch_kernel(norm_bwd_reduce_kernel, grid, block_for_channels(Ch), 0, s, dn, z, alpha, st, gamma, K, Ch, part, red);
  CTN_TRY(check_launch("norm_bwd_reduce_kernel"));
  if (defer_fold) return 0;
  return fold_partials(part, grid.x * grid.y, Ch, 0, nullptr, dgamma, dbeta, s);
}

int run_norm_bwd_apply(float* dn, const float* z, const float* alpha, NormStats st, const float* gamma,
                       const double* red, int M, int K, int Ch, float* dalpha, cudaStream_t s) {
  CTN_REQUIRE(Ch % 4 == 0, "norm_bwd: channels must be a multiple of 4 (got %d)", Ch);
  if (st.row != nullptr) {
    launch_kernel(cln_bwd_apply_kernel, cdiv((int64_t)M * K, 8), 256, 0, s, dn, z, alpha, st.row, gamma, (int64_t)M * K, Ch, dalpha);
    return check_launch("cln_bwd_apply_kernel");
  }
  launch_kernel(gln_bwd_apply_kernel, dim3(cdiv(K, GA_TK), M), block_for_channels(Ch), 0, s, dn, z, alpha, st, gamma, red, K, Ch, dalpha);
  return check_launch("gln_bwd_apply_kernel");
}

int dwconv_bwd_blocks(int M, int K, int dil) { return dw_blocks(K, dil) * M; }
int norm_bwd_blocks(int M, int K) { return cdiv(K, NR_TK) * M; }

int run_fold_batch(const FoldBatch& fb, int n_entries, cudaStream_t s) {
  CTN_REQUIRE(n_entries >= 1 && n_entries <= FOLD_MAX, "fold batch: %d entries", n_entries);
  int maxn = 0, maxnb = 0;
  for (int i = 0; i < n_entries; ++i) {
    const int n = (fb.e[i].P + 2) * fb.e[i].H;
    maxn = n > maxn ? n : maxn;
    maxnb = fb.e[i].nb > maxnb ? fb.e[i].nb : maxnb;
  }
  int splits = maxnb / 32;
  splits = splits < 1 ? 1 : (splits > 16 ? 16 : splits);
  launch_kernel(reduce_partials_batch_kernel, dim3(cdiv(maxn, 256), splits, n_entries), 256, 0, s, fb);
  return check_launch("reduce_partials_batch_kernel");
}

int run_overlap_and_add(const float* sig, int64_t outer, int frames, int flen, int step, float* out, cudaStream_t s) {
  CTN_REQUIRE(step >= 1 && step <= flen, "overlap_and_add: frame_step must be in [1, frame_length]");
  CTN_REQUIRE(outer <= 65535, "overlap_and_add: too many outer rows (%lld)", (long long)outer);
  const int64_t out_len = (int64_t)(frames - 1) * step + flen;
  launch_kernel(ola_kernel, dim3(cdiv(out_len, 256), (unsigned)outer), 256, 0, s, sig, frames, flen, step, out_len, out);
  return check_launch("ola_kernel");
}

}  // namespace ctn
