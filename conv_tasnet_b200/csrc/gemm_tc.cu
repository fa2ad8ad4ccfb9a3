// gemm_tc.cu — the 1x1 convolutions on Blackwell tensor cores (tcgen05.mma, accumulators in TMEM).
//
// fp32 parity on tensor cores by operand splitting: every fp32 operand x is split as x = hi + lo and the product is
// accumulated in fp32 (TMEM) as lo*hi + hi*lo + hi*hi (3 MMAs).
//   * forward 1x1 convs: TF32 split (hi = tf32(x), lo = x - hi exact; kind::tf32) -> 2.7e-7 max-rel-err end to end on
//     the paper config, better than the reference's own fp32 (5.5e-7).  This matters beyond the 1e-4 output budget:
//     forward rounding decides which pre-activations sit on the other side of a PReLU kink, and the gradient noise
//     grows like sqrt(forward error) (see DESIGN.md "gradient parity").
//   * data / weight gradients: bf16 split (hi = bf16(x), lo = bf16(x - hi); kind::f16) -> 5e-6 per GEMM, smooth
//     error, half the tensor and shared-memory cost.
//   Plain single-pass TF32 gives 1.4e-3 end to end and fails the 1e-4 budget.
//
// forward / dgrad kernel (tc_gemm_kernel):   D[f, o] = epi( sum_c pro(A[f, c]) * W[o, c] )
//   MMA M = 128 output channels (weights, K-major, pre-split bf16 hi/lo planes, loaded by TMA with 128B swizzle)
//   MMA N = NF frames (16..256, chosen at launch to fill whole waves of 148 SMs)
//   activations are fp32 in HBM: 8 converter warps load them (coalesced 16 B), apply the prologue (PReLU), split to
//   bf16 hi/lo and write the 128B-swizzled K-major UMMA layout; the same warps run the epilogue from TMEM
//   (tcgen05.ld 32x32b): lane = output channel, column = frame, so every store is a 128 B line.
//   epilogues: norm fold (r*acc + c1 - mu*r*c2), residual add, gLN statistics of prelu(out) (fp64 atomics).
// weight-gradient kernel (tc_wgrad_kernel):  dW[o, i] += sum_f G[f, o] * act(f, i)     (split over f)
//   both operands are MN-major (the reduction index f is the slow one): converter warps write the MN-major
//   128B-swizzled layout, optionally applying gamma*(prelu(z)-mu)*r+beta on the fly.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ctn {
#ifdef CTN_TRACE
__device__ long long g_trace[512][64];
#define TR(slot) do { g_trace[((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) & 511][slot] = clock64(); } while (0)
#else
#define TR(slot) do { } while (0)
#endif
namespace {

#ifndef CTN_CONV_THREADS
#define CTN_CONV_THREADS 256
#endif
constexpr int CONV_THREADS = CTN_CONV_THREADS;  // converter / epilogue threads (8 or 16 warps)
#ifndef CTN_CONV_THREADS_BF16
#define CTN_CONV_THREADS_BF16 512
#endif
constexpr int CONV_THREADS_BF16 = CTN_CONV_THREADS_BF16;
#ifndef CTN_TC_MAX_GROUPS
#define CTN_TC_MAX_GROUPS (CTN_CONV_THREADS > 256 ? 2 : 4)
#endif
#ifndef CTN_TF32_NF_MAX
#define CTN_TF32_NF_MAX 160
#endif
constexpr int TF32_NF_MAX = CTN_TF32_NF_MAX;  // frames per forward tile: (groups + 1) * NF <= 512 TMEM columns
constexpr int TC_MAX_GROUPS = CTN_TC_MAX_GROUPS;  // main TF32 accumulators (register budget of the epilogue)
constexpr int BM = 128;               // MMA M (output channels)
constexpr int BK = 64;                // bf16 elements per 128-byte swizzle row
constexpr int STAGES = 2;             // operand stages (hi/lo planes of W and A)
constexpr int W_PLANE_BYTES = BM * BK * 2;  // 16 KB
constexpr int RAW_STAGES = 4;          // ring of raw fp32 activation tiles (TMA destination)

struct TcGemmArgs {
  const float* A;  // [F, Kd] fp32
  float* D;        // [F, O]
  int64_t F;
  int O, Kd, K, NF, stages, raw_stages, groups;
  const float* alpha_in;
  const float* c1;
  const float* c2;
  NormStats st;
  const float* res;
  double* stat_out;
  const float* alpha_out;
  const float* nred_z;
  const float* nred_alpha;
  const float* nred_gamma;
  float* nred_dgamma;
  float* nred_dbeta;
  double* nred_red;
};

// ------------------------------------------------------------------------------------------------
// forward / dgrad GEMM
// ------------------------------------------------------------------------------------------------
// smem map (dynamic, 1024-byte aligned): OST operand stages of [W_hi 16K | W_lo 16K | A_hi NF*128 | A_lo NF*128], a ring
// of RST raw fp32 activation tiles (TMA destination), then barriers, the TMEM base address, and per-column epilogue
// metadata float2 (r, mu*r) + int sample index.
// Warp roles: 0 = TMA producer of the weight planes, 3 = TMA producer of the raw activation tiles, 1 = MMA issuer,
// 2 = TMEM allocator, 4..11 = converters (smem raw fp32 -> prologue -> hi/lo split -> swizzled operand planes; no global
// loads, so nothing is in flight when they fence) and afterwards the epilogue.
template <bool TF32, bool FOLD, bool RES, bool STATS, bool NRED = false>
__global__ void __launch_bounds__(TF32 ? 128 + CONV_THREADS : 128 + CONV_THREADS_BF16, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
               const __grid_constant__ CUtensorMap map_a, TcGemmArgs a) {
  // converter / epilogue threads: the bf16 (gradient) flavour runs 16 warps (measured: its longer K loop gains from the
  // extra converters and its one-accumulator epilogue fits the register budget), the TF32 flavour 8 (150 registers)
  constexpr int CT = TF32 ? CONV_THREADS : CONV_THREADS_BF16;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int NF = a.NF, NST = a.stages, RST = a.raw_stages;
  const int a_plane = NF * 128;
  const int stage_bytes = 2 * W_PLANE_BYTES + 2 * a_plane;
  const int raw_bytes = a_plane * (TF32 ? 1 : 2);  // one 32-float box per row (TF32) or two (bf16: 64 K-elements)
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t raw_base = smem_base + NST * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NST * stage_bytes + RST * raw_bytes);
  uint64_t* empty = full + STAGES;
  uint64_t* raw_full = empty + STAGES;
  uint64_t* raw_empty = raw_full + RAW_STAGES;
  uint64_t* tmem_full = raw_empty + RAW_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float2* s_col = reinterpret_cast<float2*>(tmem_ptr + 2);  // [256] (r, mu*r)
  int* s_m = reinterpret_cast<int*>(s_col + 256);           // [256] sample index

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TR(0);
  const int64_t f0 = (int64_t)blockIdx.x * NF;
  const int o0 = blockIdx.y * BM;
  constexpr int KB = TF32 ? 32 : 64;  // elements per 128-byte swizzle row = K extent of one stage
  const int nkb = a.Kd / KB;

  // Set-up, spread over the control warps so that the first TMA loads leave before the block-wide barrier: each producer
  // thread initialises the barriers its own loads signal and issues its first ring of loads at once (after
  // griddepcontrol.wait: both operands may have been written by the kernel right before this one).
  const int pre_w = nkb < NST ? nkb : NST, pre_r = nkb < RST ? nkb : RST;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(full + s, 1 + CT);  // weight TMA producer + every converter thread
      mbar_init(empty + s, 1);
    }
    fence_barrier_init();
    pdl_wait();  // (the planes may come from the kernel right before this one: BatchNorm fold, stand-alone calls)
    for (int kb = 0; kb < pre_w; ++kb) {
      uint8_t* st = smem + kb * stage_bytes;
      mbar_expect_tx(full + kb, 2 * W_PLANE_BYTES);
      tma_load_2d(st, &map_hi, full + kb, kb * KB, o0);
      tma_load_2d(st + W_PLANE_BYTES, &map_lo, full + kb, kb * KB, o0);
    }
  } else if (warp == 3 && lane == 0) {
    for (int r = 0; r < RST; ++r) {
      mbar_init(raw_full + r, 1);
      mbar_init(raw_empty + r, CT);
    }
    fence_barrier_init();
    pdl_wait();
    for (int kb = 0; kb < pre_r; ++kb) {
      uint8_t* dst = smem + NST * stage_bytes + kb * raw_bytes;
      mbar_expect_tx(raw_full + kb, raw_bytes);
      tma_load_2d(dst, &map_a, raw_full + kb, kb * KB, (int)f0);
      if (!TF32) tma_load_2d(dst + a_plane, &map_a, raw_full + kb, kb * KB + 32, (int)f0);
    }
  } else if (warp == 1 && lane == 0) {
    if (smem_base & 1023u) __trap();  // SWIZZLE_128B operands need a 1024-byte aligned base
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc<512>(tmem_ptr);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // dependents may be scheduled only now that this CTA owns its tensor memory: a dependent CTA that became resident
  // earlier could allocate first and then sit in griddepcontrol.wait on a grid whose CTA is blocked in tcgen05.alloc
  pdl_launch_dependents();
  pdl_wait();  // everything above overlapped the previous kernel's tail; global data is touched only below
  if (threadIdx.x == 0) TR(1);
  // TMEM accumulators (NF fp32 columns each).  The tensor core truncates (round-toward-zero) when it adds into its
  // fp32 accumulator, a coherent bias that grows with the length of the accumulation chain (measured: 2.7e-6 over a
  // 96-MMA chain vs 4e-7 for fp32 FFMA).  The TF32 (forward) flavour therefore keeps the small correction products
  // lo*hi + hi*lo in their own accumulator (index NG) and spreads the hi*hi chain round-robin over NG accumulators;
  // the epilogue sums them in fp32 round-to-nearest.  The bf16 (gradient) flavour uses a single accumulator.
  const int NG = TF32 ? a.groups : 1;

  if (warp == 0) {
    // ===== TMA producer: weight hi/lo planes into the operand stages =====
    if (lane == 0) {
      for (int kb = pre_w; kb < nkb; ++kb) {
        const int s = kb % NST, ph = (kb / NST) & 1;
        mbar_wait(empty + s, ph ^ 1);
        uint8_t* st = smem + s * stage_bytes;
        mbar_expect_tx(full + s, 2 * W_PLANE_BYTES);
        tma_load_2d(st, &map_hi, full + s, kb * KB, o0);
        tma_load_2d(st + W_PLANE_BYTES, &map_lo, full + s, kb * KB, o0);
      }
    }
  } else if (warp == 3) {
    // ===== TMA producer: raw fp32 activation tiles [NF frames x 32 floats], rows past F are zero-filled =====
    if (lane == 0) {
      for (int kb = pre_r; kb < nkb; ++kb) {
        const int r = kb % RST, ph = (kb / RST) & 1;
        mbar_wait(raw_empty + r, ph ^ 1);
        uint8_t* dst = smem + NST * stage_bytes + r * raw_bytes;
        mbar_expect_tx(raw_full + r, raw_bytes);
        tma_load_2d(dst, &map_a, raw_full + r, kb * KB, (int)f0);
        if (!TF32) tma_load_2d(dst + a_plane, &map_a, raw_full + r, kb * KB + 32, (int)f0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, NF, 0, 0, TF32 ? 2u : 1u);
      const uint32_t dhi = desc_hi_sw128(1024);
      const uint32_t t_corr = tmem_base + (uint32_t)(NG * NF);
      int cstep = 0, gsel = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % NST, ph = (kb / NST) & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        if (kb < 16) TR(8 + kb);
        const uint32_t sb = smem_base + s * stage_bytes;
        uint32_t w_hi = desc_lo(sb, 16), w_lo = desc_lo(sb + W_PLANE_BYTES, 16);
        uint32_t a_hi = desc_lo(sb + 2 * W_PLANE_BYTES, 16), a_lo = desc_lo(sb + 2 * W_PLANE_BYTES + a_plane, 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // one MMA consumes 32 bytes of K per row (16 bf16 or 8 tf32): +2 in the descriptor
          if (TF32) {
            umma_lo<true>(t_corr, w_lo, a_hi, dhi, idesc, cstep != 0);                         // lo*hi -> correction
            umma_lo<true>(t_corr, w_hi, a_lo, dhi, idesc, 1);                                  // hi*lo -> correction
            umma_lo<true>(tmem_base + (uint32_t)(gsel * NF), w_hi, a_hi, dhi, idesc, cstep >= NG);  // hi*hi -> main[gsel]
            gsel = gsel + 1 == NG ? 0 : gsel + 1;
          } else {
            umma_lo<false>(tmem_base, w_lo, a_hi, dhi, idesc, cstep != 0);
            umma_lo<false>(tmem_base, w_hi, a_lo, dhi, idesc, 1);
            umma_lo<false>(tmem_base, w_hi, a_hi, dhi, idesc, 1);
          }
          ++cstep;
          w_hi += 2; w_lo += 2; a_hi += 2; a_lo += 2;
        }
        umma_commit(empty + s);  // frees the stage when these MMAs have read it
      }
      umma_commit(tmem_full);
      TR(2);
    }
  } else if (warp >= 4) {
    // ===== converters: raw smem tile -> prologue -> hi/lo split -> operand planes; then the epilogue =====
    const int t = threadIdx.x - 128;  // 0..255
    const bool pro = a.alpha_in != nullptr;
    const float alpha_in = pro ? __ldg(a.alpha_in) : 1.f;
    const int nvalid = (int)(a.F - f0 < NF ? a.F - f0 : NF);  // frames of this tile that exist
    const int nchunks = NF * 8;                 // 16-byte operand chunks per plane
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % NST, ph = (kb / NST) & 1;
      const int r = kb % RST, rph = (kb / RST) & 1;
      const uint32_t raw = raw_base + r * raw_bytes;
      const uint32_t dst = smem_base + s * stage_bytes + 2 * W_PLANE_BYTES;
      mbar_wait(raw_full + r, rph);       // TMA has landed the raw tile
      if (t == 0 && kb < 16) TR(24 + kb);
      mbar_wait(empty + s, ph ^ 1);       // the MMAs that read this operand stage have finished
      if (TF32) {
        // same element size in and out: the swizzled position of a 16-byte chunk is identical in the raw tile and
        // in the operand planes, so the conversion is position-agnostic
        // batches of 5 chunks per thread: all shared loads first (ILP), then convert + store
        for (int q0 = t; q0 < nchunks; q0 += 5 * CT) {
          float4 x[5];
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const int q = q0 + u * CT;
            if (q < nchunks) x[u] = lds128f(raw + q * 16);
          }
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const int q = q0 + u * CT;
            if (q < nchunks) {
              float4 y = x[u];
              if (pro) { y.x = prelu(y.x, alpha_in); y.y = prelu(y.y, alpha_in); y.z = prelu(y.z, alpha_in); y.w = prelu(y.w, alpha_in); }
              uint4 hi, lo;
              split4_tf32(y, hi, lo);
              sts128(dst + q * 16, hi);
              sts128(dst + a_plane + q * 16, lo);
            }
          }
        }
      } else {
        // bf16 chunk oc of row `row` = K elements [8 oc, 8 oc + 8) = raw box (oc >> 2), raw chunks 2 (oc & 3), +1
        for (int q0 = t; q0 < nchunks; q0 += 5 * CT) {
          float4 x0[5], x1[5];
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const int q = q0 + u * CT;
            if (q < nchunks) {
              const int row = q >> 3, oc = q & 7, sw = row & 7;
              const uint32_t rb = raw + (oc >> 2) * a_plane + row * 128;
              const int c0 = (oc & 3) * 2;
              x0[u] = lds128f(rb + ((c0 ^ sw) << 4));
              x1[u] = lds128f(rb + (((c0 + 1) ^ sw) << 4));
            }
          }
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const int q = q0 + u * CT;
            if (q < nchunks) {
              const int row = q >> 3, oc = q & 7, sw = row & 7;
              float x[8] = {x0[u].x, x0[u].y, x0[u].z, x0[u].w, x1[u].x, x1[u].y, x1[u].z, x1[u].w};
              if (pro) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = prelu(x[i], alpha_in);
              }
              uint4 hi, lo;
              split8(x, hi, lo);
              const uint32_t off = row * 128 + ((oc ^ sw) << 4);
              sts128(dst + off, hi);
              sts128(dst + a_plane + off, lo);
            }
          }
        }
      }
      fence_proxy_async();  // make the generic-proxy writes visible to the tensor core (async proxy)
      mbar_arrive(full + s);
      mbar_arrive(raw_empty + r);
      if (t == 0 && kb < 16) TR(40 + kb);
    }

    // ---- epilogue: TMEM -> registers -> global; lane = output channel, column = frame ----
    if (t < NF) {  // per-column (frame) metadata (sample index, norm-fold scalars), off the mainloop's critical path
      int m = -1;
      float mu = 0.f, r = 1.f;
      if (t < nvalid) {
        m = (int)((uint32_t)(f0 + t) / (uint32_t)a.K);
        if (FOLD || NRED) load_stats(a.st, m, f0 + t, mu, r);
      }
      s_m[t] = m;
      s_col[t] = make_float2(r, mu * r);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory");  // column metadata written by the converter threads
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (t == 0) TR(3);
    const int q = warp & 3, part = (warp - 4) >> 2;  // TMEM lane quarter; column range among the warps sharing it
    constexpr int CPARTS = CT / 128;
    const int o = o0 + q * 32 + lane;
    const int O = a.O;
    const float c1 = FOLD ? __ldg(a.c1 + o) : 0.f, c2 = FOLD ? __ldg(a.c2 + o) : 0.f;
    const float alpha_out = (STATS && a.alpha_out) ? __ldg(a.alpha_out) : 1.f;
    const int nch = NF >> 3;  // 8-column chunks of the tile
    const int jb = 8 * ((part * nch) / CPARTS);
    const int je = min(8 * (((part + 1) * nch) / CPARTS), nvalid);
    float* dptr = a.D + f0 * O + o;
    const float* rptr = RES ? a.res + f0 * O + o : nullptr;
    const uint32_t col_s = smem_u32(s_col);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float s1 = 0.f, s2 = 0.f;
    int cur_m = (STATS || NRED) && jb < je ? s_m[jb] : -1;
    const float* zptr = NRED ? a.nred_z + f0 * O + o : nullptr;
    const float nr_alpha = (NRED && a.nred_alpha) ? __ldg(a.nred_alpha) : 1.f;
    const bool nr_prelu = NRED && a.nred_alpha != nullptr;
    const float nr_gamma = NRED ? __ldg(a.nred_gamma + o) : 0.f;
    float nr_dg = 0.f, nr_db = 0.f;
    constexpr int NACC = TF32 ? TC_MAX_GROUPS + 1 : 1;  // TMEM tiles summed per output (NG main + 1 correction)
    uint32_t rawA[NACC][8], rawB[NACC][8];
    auto issue = [&](uint32_t (&dst)[NACC][8], int j) {
#pragma unroll
      for (int gidx = 0; gidx < NACC; ++gidx)
        if (gidx <= (TF32 ? NG : 0)) tmem_ld8_issue(taddr + (uint32_t)(gidx * NF + j), dst[gidx]);
    };
    auto process = [&](uint32_t (&cur)[NACC][8], int j) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __uint_as_float(cur[0][i]);
#pragma unroll
      for (int gidx = 1; gidx < NACC; ++gidx)
        if (gidx <= (TF32 ? NG : 0)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += __uint_as_float(cur[gidx][i]);
        }
      const int nj = min(8, je - j);
      const bool uniform = !(STATS || NRED) || (s_m[j] == cur_m && s_m[j + nj - 1] == cur_m);
      if (nj == 8 && uniform) {  // fast path: whole chunk valid, one sample
        float resv[8];
        if (RES) {
#pragma unroll
          for (int i = 0; i < 8; ++i) resv[i] = __ldg(rptr + (j + i) * O);
        }
        if (NRED) {
#pragma unroll
          for (int i = 0; i < 8; ++i) resv[i] = __ldg(zptr + (j + i) * O);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float val = acc[i];
          if (FOLD) {
            const float2 cm = lds64(col_s + (j + i) * 8);
            val = fmaf(cm.x, val, fmaf(-cm.y, c2, c1));
          }
          if (RES) val += resv[i];
          dptr[(j + i) * O] = val;
          if (STATS) {
            const float p = prelu(val, alpha_out);
            s1 += p;
            s2 = fmaf(p, p, s2);
          }
          if (NRED) {
            const float2 cm = lds64(col_s + (j + i) * 8);
            const float av = nr_prelu ? prelu(resv[i], nr_alpha) : resv[i];
            const float yh = fmaf(av, cm.x, -cm.y);
            nr_dg = fmaf(val, yh, nr_dg);
            nr_db += val;
            const float gh = val * nr_gamma;
            s1 += gh;
            s2 = fmaf(gh, yh, s2);
          }
        }
      } else {  // chunk crosses the end of the tile or a sample boundary (at most a couple per tile)
        for (int i = 0; i < nj; ++i) {
          float val = acc[i];
          if (FOLD) {
            const float2 cm = lds64(col_s + (j + i) * 8);
            val = fmaf(cm.x, val, fmaf(-cm.y, c2, c1));
          }
          if (RES) val += __ldg(rptr + (j + i) * O);
          dptr[(j + i) * O] = val;
          if (STATS || NRED) {
            const int m = s_m[j + i];
            if (m != cur_m) {  // warp-uniform
              const double d1 = warp_sum((double)s1), d2 = warp_sum((double)s2);
              double* sacc = NRED ? a.nred_red : a.stat_out;
              if (lane == 0 && sacc != nullptr) {
                atomicAdd(sacc + 2 * cur_m, d1);
                atomicAdd(sacc + 2 * cur_m + 1, d2);
              }
              cur_m = m;
              s1 = s2 = 0.f;
            }
          }
          if (STATS) {
            const float p = prelu(val, alpha_out);
            s1 += p;
            s2 = fmaf(p, p, s2);
          }
          if (NRED) {
            const float2 cm = lds64(col_s + (j + i) * 8);
            const float zv = __ldg(zptr + (j + i) * O);
            const float av = nr_prelu ? prelu(zv, nr_alpha) : zv;
            const float yh = fmaf(av, cm.x, -cm.y);
            nr_dg = fmaf(val, yh, nr_dg);
            nr_db += val;
            const float gh = val * nr_gamma;
            s1 += gh;
            s2 = fmaf(gh, yh, s2);
          }
        }
      }
    };
    if (CT <= 256 || TC_MAX_GROUPS <= 2 || !TF32) {
      // software pipeline: the TMEM loads of the next 8 columns are in flight while this chunk is processed
      if (jb < je) issue(rawA, jb);
      if (t == 0) TR(6);
      for (int j = jb; j < je; j += 16) {
        tmem_ld_wait();
        if (t == 0 && j == jb) TR(7);
        if (j + 8 < je) issue(rawB, j + 8);
        process(rawA, j);
        if (j + 8 < je) {
          tmem_ld_wait();
          if (j + 16 < je) issue(rawA, j + 16);
          process(rawB, j + 8);
        }
      }
    } else {  // 16 epilogue warps: four per TMEM lane quarter hide the load latency between them (and registers are scarce)
      for (int j = jb; j < je; j += 8) {
        issue(rawA, j);
        tmem_ld_wait();
        process(rawA, j);
      }
    }
    if ((STATS || NRED) && cur_m >= 0) {
      const double d1 = warp_sum((double)s1), d2 = warp_sum((double)s2);
      double* sacc = NRED ? a.nred_red : a.stat_out;
      if (lane == 0 && sacc != nullptr) {
        atomicAdd(sacc + 2 * cur_m, d1);
        atomicAdd(sacc + 2 * cur_m + 1, d2);
      }
    }
    if (NRED) {  // per-channel partial sums of this tile's frames
      atomicAdd(a.nred_dgamma + o, nr_dg);
      atomicAdd(a.nred_dbeta + o, nr_db);
    }
  }
  if (threadIdx.x == 128) TR(4);
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
  if (threadIdx.x == 0) TR(5);
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
constexpr int WK = 32;       // f rows per stage
constexpr int WSTAGES = 4;

struct TcWgradArgs {
  const float* G;    // [F, O]
  const float* Act;  // [F, I]
  float* dW;         // [O, I]
  int64_t F;
  int O, I, K;
  int f_chunk;
  const float* alpha;
  const float* gamma;
  const float* beta;
  NormStats st;
};

constexpr int WG_STAT_CACHE = 16;  // samples whose (mean, rstd) a CTA caches; more fall back to load_stats per row
#ifndef CTN_WGT
#define CTN_WGT 256
#endif
#ifndef CTN_WNG
#define CTN_WNG 2
#endif
constexpr int WGT = CTN_WGT;               // threads of one converter group
constexpr int WNG = CTN_WNG;               // converter groups, taking k-blocks round robin
constexpr int W_THREADS = 128 + WNG * WGT; // 4 control warps + the converter / epilogue warps
template <int NI>
__global__ void __launch_bounds__(W_THREADS, 1) tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_dw, TcWgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  constexpr int A_PLANE = WK * BM * 2;   // 8 KB   (2 groups of 64 o-columns x 32 rows x 128 B)
  constexpr int B_PLANE = WK * NI * 2;   // 16 KB for NI = 256
  constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  constexpr int GROUP = WK * 128;        // bytes of one 64-column group (LBO)
  uint8_t* tail = smem + WSTAGES * STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + WSTAGES;
  uint64_t* tmem_full = empty + WSTAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float2* s_wst = reinterpret_cast<float2*>(tail + 128);  // [WG_STAT_CACHE]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TR(0);
  const int o0 = blockIdx.x * BM, i0 = blockIdx.y * NI;
  const int64_t fb = (int64_t)blockIdx.z * a.f_chunk;
  const int64_t fe = fb + a.f_chunk < a.F ? fb + a.f_chunk : a.F;
  const int nkb = fb < fe ? (int)((fe - fb + WK - 1) / WK) : 0;

  if (warp == 1 && lane == 0) {
    if (smem_base & 1023u) __trap();
    for (int s = 0; s < WSTAGES; ++s) {
      mbar_init(full + s, WGT);  // one converter group per k-block
      mbar_init(empty + s, 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc<256>(tmem_ptr);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // dependents may be scheduled only now that this CTA owns its tensor memory: a dependent CTA that became resident
  // earlier could allocate first and then sit in griddepcontrol.wait on a grid whose CTA is blocked in tcgen05.alloc
  pdl_launch_dependents();
  pdl_wait();  // everything above overlapped the previous kernel's tail; global data is touched only below
  if (threadIdx.x == 0) TR(1);

  if (nkb > 0) {
    if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc(BM, NI, 1, 1);
        for (int kb = 0; kb < nkb; ++kb) {
          const int s = kb % WSTAGES, ph = (kb / WSTAGES) & 1;
          mbar_wait(full + s, ph);
          tc_fence_after();
          if (kb < 16) TR(8 + kb);
          const uint32_t sb = smem_base + s * STAGE;
          const uint32_t g_hi = sb, g_lo = sb + A_PLANE, x_hi = sb + 2 * A_PLANE, x_lo = x_hi + B_PLANE;
#pragma unroll
          for (int k = 0; k < WK / 16; ++k) {
            const uint32_t ko = k * 16 * 128;  // 16 k-rows of 128 bytes
            const uint64_t dgh = make_desc(g_hi + ko, GROUP, 1024), dgl = make_desc(g_lo + ko, GROUP, 1024);
            const uint64_t dxh = make_desc(x_hi + ko, GROUP, 1024), dxl = make_desc(x_lo + ko, GROUP, 1024);
            umma_bf16(tmem_base, dgl, dxh, idesc, (kb | k) != 0);
            umma_bf16(tmem_base, dgh, dxl, idesc, 1);
            umma_bf16(tmem_base, dgh, dxh, idesc, 1);
          }
          umma_commit(empty + s);
        }
        umma_commit(tmem_full);
        TR(2);
      }
    } else if (warp >= 4) {
      const int t = threadIdx.x - 128;
      const int grp = t / WGT, tg = t % WGT;  // two groups of WGT threads alternate k-blocks (see tc_gemm_kernel)
      const bool norm = a.gamma != nullptr, hasp = a.alpha != nullptr;
      const float alpha = hasp ? __ldg(a.alpha) : 1.f;
      // G tile: 32 rows x 16 chunks (8 floats each): GIT chunks per thread (rows tg/16 + GR*it)
      constexpr int GR = WGT / 16, GIT = WK / GR;
      const int gc = tg & 15, gr = tg >> 4;
      // Act tile: 32 rows x NI/8 chunks
      constexpr int XC = NI / 8;        // chunks per row
      constexpr int XR = WGT / XC;      // rows covered per pass
      constexpr int XIT = WK / XR;
      const int xc = tg % XC, xr = tg / XC;
      float gam[8], bet[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        gam[i] = norm ? __ldg(a.gamma + i0 + xc * 8 + i) : 1.f;
        bet[i] = norm ? __ldg(a.beta + i0 + xc * 8 + i) : 0.f;
      }
      // per-sample (gLN / identity) statistics of the samples this f-chunk touches, derived ONCE per CTA: load_stats is
      // fp64 (mean, rsqrt of the variance) and used to run per row per thread (+11 us per launch, measured in place)
      const bool per_sample = norm && a.st.row == nullptr;
      const int m_lo = (int)((uint32_t)fb / (uint32_t)a.K);
      if (per_sample) {
        if (t < WG_STAT_CACHE && (int64_t)(m_lo + t) * a.K < fe) {
          float mu, r;
          load_stats(a.st, m_lo + t, 0, mu, r);
          s_wst[t] = make_float2(mu, r);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(WNG * WGT) : "memory");
      }
      for (int kb = grp; kb < nkb; kb += WNG) {
        const int s = kb % WSTAGES, ph = (kb / WSTAGES) & 1;
        const int64_t fk = fb + (int64_t)kb * WK;
        float4 gv[GIT][2], xv[XIT][2];
#pragma unroll
        for (int it = 0; it < GIT; ++it) {
          const int64_t f = fk + gr + it * GR;
          const bool ok = f < fe;
          const float4* src = reinterpret_cast<const float4*>(a.G + f * a.O + o0 + gc * 8);
          gv[it][0] = ok ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
          gv[it][1] = ok ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int it = 0; it < XIT; ++it) {
          const int64_t f = fk + xr + it * XR;
          const bool ok = f < fe;
          const float4* src = reinterpret_cast<const float4*>(a.Act + f * a.I + i0 + xc * 8);
          xv[it][0] = ok ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
          xv[it][1] = ok ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(empty + s, ph ^ 1);
        const uint32_t st = smem_base + s * STAGE;
#pragma unroll
        for (int it = 0; it < GIT; ++it) {
          const int k = gr + it * GR;
          const float x[8] = {gv[it][0].x, gv[it][0].y, gv[it][0].z, gv[it][0].w,
                              gv[it][1].x, gv[it][1].y, gv[it][1].z, gv[it][1].w};
          uint4 hi, lo;
          split8(x, hi, lo);
          const int off = (gc >> 3) * GROUP + k * 128 + (((gc & 7) ^ (k & 7)) << 4);
          sts128(st + off, hi);
          sts128(st + A_PLANE + off, lo);
        }
#pragma unroll
        for (int it = 0; it < XIT; ++it) {
          const int k = xr + it * XR;
          const int64_t f = fk + k;
          float x[8] = {xv[it][0].x, xv[it][0].y, xv[it][0].z, xv[it][0].w,
                        xv[it][1].x, xv[it][1].y, xv[it][1].z, xv[it][1].w};
          if (f < fe) {
            if (hasp) {
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = prelu(x[i], alpha);
            }
            if (norm) {
              float mu, r;
              const int m = (int)((uint32_t)f / (uint32_t)a.K);
              if (per_sample && m - m_lo < WG_STAT_CACHE) {
                const float2 v = s_wst[m - m_lo];
                mu = v.x;
                r = v.y;
              } else {
                load_stats(a.st, m, f, mu, r);
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = gam[i] * (x[i] - mu) * r + bet[i];
            }
          }
          uint4 hi, lo;
          split8(x, hi, lo);
          const int off = (xc >> 3) * GROUP + k * 128 + (((xc & 7) ^ (k & 7)) << 4);
          sts128(st + 2 * A_PLANE + off, hi);
          sts128(st + 2 * A_PLANE + B_PLANE + off, lo);
        }
        fence_proxy_async();
        mbar_arrive(full + s);
      }
      // ---- epilogue: add the partial tile to dW with TMA reductions ----
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      if (t == 0) TR(3);
      // TMEM gives lane = output row o, registers = columns i.  The tile is staged in shared memory (the operand stages
      // are free: every MMA has completed) as NI/32 boxes of [128 rows x 128 bytes] in the 128-byte-swizzle layout
      // (conflict-free 16-byte stores down a column), and one thread hands each box to the TMA engine as a bulk
      // reduce-add into dW (cp.reduce.async.bulk.tensor): the L2 does the additions at line granularity while the CTA
      // retires, instead of 32k red.global.add.v4 issued and waited for by the warps (measured: the split-K atomics were
      // ~10k of the kernel's ~37k cycles).  CTN_WGRAD_RED=1 (compile time) keeps the red.v4 path for A/B.
      constexpr int EW = (WNG * WGT / 32) >= 16 ? 16 : 8;  // epilogue warps (the first EW converter warps)
      constexpr int CSPLIT = EW / 4;  // warps sharing a TMEM lane quarter: each takes NI / CSPLIT columns
      constexpr int QC = NI / CSPLIT;
      const int q = warp & 3, part = (warp - 4) >> 2;
#ifndef CTN_WGRAD_RED
      if (warp - 4 < EW) {
        const int jb = part * QC, je = jb + QC;
        const int row = q * 32 + lane;
        for (int j = jb; j < je; j += 8) {
          float acc[8];
          tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)j, acc);
          const uint32_t box = smem_base + (uint32_t)((j >> 5) * (BM * 128) + row * 128);
          const int c = (j & 31) >> 2;  // 16-byte chunk inside the 128-byte box row
          sts128(box + (uint32_t)(((c ^ (row & 7)) << 4)), make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]),
                                                                       __float_as_uint(acc[2]), __float_as_uint(acc[3])));
          sts128(box + (uint32_t)((((c + 1) ^ (row & 7)) << 4)), make_uint4(__float_as_uint(acc[4]), __float_as_uint(acc[5]),
                                                                             __float_as_uint(acc[6]), __float_as_uint(acc[7])));
        }
        fence_proxy_async();  // the generic-proxy stores above must be visible to the TMA engine (async proxy)
        asm volatile("bar.sync 2, %0;" ::"n"(EW * 32) : "memory");
        if (t == 0) TR(6);
        if (warp == 4 && lane == 0) {
#pragma unroll 1
          for (int b = 0; b < NI / 32; ++b)
            tma_reduce_add_2d(&map_dw, smem + b * (BM * 128), i0 + 32 * b, o0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory may go once the engine has read it
          TR(4);
        }
      }
#else
      if (warp - 4 < EW) {
      const int jb = part * QC, je = jb + QC;
      constexpr int PITCH = NI + 4;  // floats; 16-byte row alignment, conflict-free 16-byte stores down a column
      const uint32_t srow = smem_base + (uint32_t)(((q * 32 + lane) * PITCH + jb) * 4);
      for (int j = jb; j < je; j += 8) {
        float acc[8];
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)j, acc);
        sts128(srow + (uint32_t)((j - jb) * 4), make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]),
                                                           __float_as_uint(acc[2]), __float_as_uint(acc[3])));
        sts128(srow + (uint32_t)((j - jb) * 4 + 16), make_uint4(__float_as_uint(acc[4]), __float_as_uint(acc[5]),
                                                                __float_as_uint(acc[6]), __float_as_uint(acc[7])));
      }
      __syncwarp();
      constexpr int LPR = QC / 4;    // lanes per row (float4 each) of the quadrant
      constexpr int RPI = 32 / LPR;  // rows per warp instruction
      const int rsub = lane / LPR, col = jb + (lane % LPR) * 4;
#pragma unroll 4
      for (int r = 0; r < 32; r += RPI) {
        const int row = q * 32 + r + rsub;
        const float4 v = lds128f(smem_base + (uint32_t)((row * PITCH + col) * 4));
        red_add_v4(a.dW + (int64_t)(o0 + row) * a.I + i0 + col, v.x, v.y, v.z, v.w);
      }
      }
#endif
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<256>(tmem_base);
  if (threadIdx.x == 0) TR(5);
}

// ------------------------------------------------------------------------------------------------
// weight planes: fp32 [R, C] -> bf16 hi / lo planes, optionally transposed to [C, R]; batched over blockIdx.z
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ src, int R, int C, int64_t src_stride,
                                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                           int64_t dst_stride, int transpose) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float tile[32][33];
  const float* s = src + (int64_t)blockIdx.z * src_stride;
  __nv_bfloat16* h = hi + (int64_t)blockIdx.z * dst_stride;
  __nv_bfloat16* l = lo + (int64_t)blockIdx.z * dst_stride;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < C) ? s[(int64_t)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    float v;
    int64_t idx;
    bool ok;
    if (!transpose) {
      const int r = r0 + i, c = c0 + tx;
      v = tile[i][tx]; idx = (int64_t)r * C + c; ok = r < R && c < C;
    } else {
      const int c = c0 + i, r = r0 + tx;  // dst is [C, R]
      v = tile[tx][i]; idx = (int64_t)c * R + r; ok = r < R && c < C;
    }
    if (ok) {
      const __nv_bfloat16 hv = __float2bfloat16_rn(v);
      h[idx] = hv;
      l[idx] = __float2bfloat16_rn(v - __bfloat162float(hv));
    }
  }
}

// the TF32 flavour: hi = tf32(x) and the exact remainder, both stored as fp32 [R, C] planes (no transpose needed:
// only the forward weights use it)
__global__ void __launch_bounds__(256) split_planes_tf32_kernel(const float* __restrict__ src, int64_t n, int64_t src_stride,
                                                                float* __restrict__ hi, float* __restrict__ lo,
                                                                int64_t dst_stride) {
  pdl_launch_dependents();
  pdl_wait();
  const float* s = src + (int64_t)blockIdx.y * src_stride;
  float* h = hi + (int64_t)blockIdx.y * dst_stride;
  float* l = lo + (int64_t)blockIdx.y * dst_stride;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = s[i];
    uint32_t hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
    h[i] = __uint_as_float(hb);
    l[i] = v - __uint_as_float(hb);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// weight plane [rows, cols] row-major (bf16, or fp32 for the TF32 flavour) -> tensor map with a
// [128 rows x 128 bytes] box, 128-byte swizzle
static int make_plane_map(CUtensorMap* map, const void* plane, int rows, int cols, bool tf32) {
  EncodeTiledFn enc = get_encode();
  CTN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * (tf32 ? 4 : 2)};
  const cuuint32_t box[2] = {(cuuint32_t)(tf32 ? 32 : 64), BM};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(plane), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CTN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)r);
  return 0;
}

static size_t tc_gemm_tail_bytes() { return (2 * STAGES + 2 * RAW_STAGES + 1) * 8 + 8 + 256 * 4 * 3; }
static int tc_gemm_raw_stages(int NF, bool tf32) {
  const size_t stage = 2 * W_PLANE_BYTES + 2 * (size_t)NF * 128, raw = (size_t)NF * 128 * (tf32 ? 1 : 2);
  const size_t budget = 227 * 1024 - tc_gemm_tail_bytes();
  if (budget < STAGES * stage) return 0;
  int rst = (int)((budget - STAGES * stage) / raw);
  return rst > RAW_STAGES ? RAW_STAGES : rst;
}
static size_t tc_gemm_smem(int NF, bool tf32, int raw_stages) {
  return (size_t)STAGES * (2 * W_PLANE_BYTES + 2 * NF * 128) + (size_t)raw_stages * NF * 128 * (tf32 ? 1 : 2) +
         tc_gemm_tail_bytes();
}

static int pick_nf(int64_t F, int o_tiles, bool tf32) {
  // frames per tile (multiple of 16, <= 256) minimising (waves * tile time) on 148 SMs; the TF32 flavour needs room
  // for >= 2 main accumulators + 1 correction accumulator in the 512 TMEM columns
  int best = 128;
  double best_cost = 1e30;
  for (int nf = 64; nf <= (tf32 ? TF32_NF_MAX : 256); nf += 16) {
    if (tc_gemm_raw_stages(nf, tf32) < 2) continue;  // operand stages + raw ring must fit in 227 KB
    const int64_t tiles = (F + nf - 1) / nf * o_tiles;
    const int64_t waves = (tiles + 147) / 148;
    const double cost = (double)waves * (nf + 40);  // + fixed per-tile overhead (prologue/epilogue)
    if (cost < best_cost) {
      best_cost = cost;
      best = nf;
    }
  }
  return best;
}

// fp32 activation matrix [F, Kd] row-major -> tensor map with a [NF rows x 32 floats] box, 128-byte swizzle, zero fill
static int make_act_map(CUtensorMap* map, const float* A, int64_t F, int Kd, int NF) {
  EncodeTiledFn enc = get_encode();
  CTN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  const cuuint64_t dims[2] = {(cuuint64_t)Kd, (cuuint64_t)F};
  const cuuint64_t strides[1] = {(cuuint64_t)Kd * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)NF};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CTN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (activations) failed with code %d", (int)r);
  return 0;
}

}  // namespace

int launch_gemm_simt(const GemmArgs& a, cudaStream_t s);

bool tc_gemm_eligible(const GemmArgs& a) {
  return a.W_hi != nullptr && a.W_lo != nullptr && a.Kd % BK == 0 && a.O % BM == 0 && a.F >= 16;
}

int launch_gemm_tc(const GemmArgs& g, cudaStream_t s) {
  CUtensorMap mh, ml;
  const bool tf32 = g.tf32 != 0;
  CTN_TRY(make_plane_map(&mh, g.W_hi, g.O, g.Kd, tf32));
  CTN_TRY(make_plane_map(&ml, g.W_lo, g.O, g.Kd, tf32));
  TcGemmArgs a;
  a.A = g.A; a.D = g.D; a.F = g.F; a.O = g.O; a.Kd = g.Kd; a.K = g.K;
  a.NF = pick_nf(g.F, g.O / BM, tf32);
  a.groups = 512 / a.NF - 1;
  if (a.groups > TC_MAX_GROUPS) a.groups = TC_MAX_GROUPS;
  a.alpha_in = g.alpha_in; a.c1 = g.c1; a.c2 = g.c2; a.st = g.st; a.res = g.res;
  a.stat_out = g.stat_out; a.alpha_out = g.alpha_out;
  a.nred_z = g.nred_z; a.nred_alpha = g.nred_alpha; a.nred_gamma = g.nred_gamma;
  a.nred_dgamma = g.nred_dgamma; a.nred_dbeta = g.nred_dbeta; a.nred_red = g.nred_red;
  a.stages = STAGES;
  a.raw_stages = tc_gemm_raw_stages(a.NF, tf32);
  const size_t smem = tc_gemm_smem(a.NF, tf32, a.raw_stages);
  CUtensorMap ma;
  CTN_TRY(make_act_map(&ma, g.A, g.F, g.Kd, a.NF));
  CTN_REQUIRE(a.raw_stages >= 2 && smem <= 227 * 1024, "tc_gemm: shared memory %zu too large", smem);
  dim3 grid(cdiv(g.F, a.NF), g.O / BM);
  const bool fold = g.c1 != nullptr, res = g.res != nullptr, stats = g.stat_out != nullptr;
  const int threads = 128 + (tf32 ? CONV_THREADS : CONV_THREADS_BF16);
#define CTN_TC_LAUNCH(...)                                                                                       \
  do {                                                                                                           \
    static unsigned long long attr_mask = 0;  /* the attribute is per device */                                  \
    int dev__ = 0;                                                                                               \
    CTN_CUDA(cudaGetDevice(&dev__));                                                                             \
    if (!((attr_mask >> (dev__ & 63)) & 1ull)) {                                                                 \
      CTN_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                    227 * 1024));                                                                \
      attr_mask |= 1ull << (dev__ & 63);                                                                         \
    }                                                                                                            \
    launch_kernel(tc_gemm_kernel<__VA_ARGS__>, grid, threads, smem, s, mh, ml, ma, a);                                   \
  } while (0)
  if (tf32) {  // forward 1x1 convs
    if (!fold && !res && !stats) CTN_TC_LAUNCH(true, false, false, false);
    else if (!fold && !res && stats) CTN_TC_LAUNCH(true, false, false, true);
    else if (fold && !res && !stats) CTN_TC_LAUNCH(true, true, false, false);
    else if (fold && res && !stats) CTN_TC_LAUNCH(true, true, true, false);
    else return launch_gemm_simt(g, s);  // combinations the model never issues
  } else {     // data gradients
    if (g.nred_z != nullptr && !fold && !res && !stats) CTN_TC_LAUNCH(false, false, false, false, true);
    else if (!fold && !res && !stats) CTN_TC_LAUNCH(false, false, false, false);
    else if (!fold && !res && stats) CTN_TC_LAUNCH(false, false, false, true);   // inference-precision forward convs
    else if (fold && !res && !stats) CTN_TC_LAUNCH(false, true, false, false);
    else if (fold && res && !stats) CTN_TC_LAUNCH(false, true, true, false);
    else if (!fold && res && !stats) CTN_TC_LAUNCH(false, false, true, false);
    else return launch_gemm_simt(g, s);
  }
#undef CTN_TC_LAUNCH
  return check_launch(g.tf32 ? (g.O > g.Kd ? "tc_gemm_kernel<tf32> up" : "tc_gemm_kernel<tf32> down")
                             : (g.O > g.Kd ? "tc_gemm_kernel<bf16> up" : "tc_gemm_kernel<bf16> down"));
}

bool tc_wgrad_eligible(const WgradArgs& a) {
  return a.O % BM == 0 && (a.I % 256 == 0 || a.I == 128) && a.F >= 64;
}

int launch_wgrad_tc(const WgradArgs& w, cudaStream_t s) {
  TcWgradArgs a;
  a.G = w.G; a.Act = w.Act; a.dW = w.dW; a.F = w.F; a.O = w.O; a.I = w.I; a.K = w.K;
  a.alpha = w.alpha; a.gamma = w.gamma; a.beta = w.beta; a.st = w.st;
  const int ni = w.I % 256 == 0 ? 256 : 128;
  const int tiles = (w.O / BM) * (w.I / ni);
  int splits = (148 + tiles - 1) / tiles;
  int f_chunk = (int)((w.F + splits - 1) / splits);
  f_chunk = ((f_chunk + WK - 1) / WK) * WK;
  splits = cdiv(w.F, f_chunk);
  a.f_chunk = f_chunk;
  dim3 grid(w.O / BM, w.I / ni, splits);
  static unsigned long long attr_mask = 0;  // the attribute is per device
  int dev = 0;
  CTN_CUDA(cudaGetDevice(&dev));
  if (!((attr_mask >> (dev & 63)) & 1ull)) {
    CTN_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CTN_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_mask |= 1ull << (dev & 63);
  }
  CUtensorMap md;  // dW [O, I] fp32, boxes of [128 rows x 32 floats], 128-byte swizzle: the target of the TMA reduce-adds
  {
    EncodeTiledFn enc = get_encode();
    CTN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    const cuuint64_t dims[2] = {(cuuint64_t)w.I, (cuuint64_t)w.O};
    const cuuint64_t strides[1] = {(cuuint64_t)w.I * 4};
    const cuuint32_t box[2] = {32, BM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&md, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, w.dW, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CTN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (dW) failed with code %d", (int)r);
  }
  if (ni == 256) {
    const size_t smem = (size_t)WSTAGES * (2 * WK * BM * 2 + 2 * WK * 256 * 2) + 128 + 8 * WG_STAT_CACHE;
    launch_kernel(tc_wgrad_kernel<256>, grid, W_THREADS, smem, s, md, a);
  } else {
    const size_t smem = (size_t)WSTAGES * (2 * WK * BM * 2 + 2 * WK * 128 * 2) + 128 + 8 * WG_STAT_CACHE;
    launch_kernel(tc_wgrad_kernel<128>, grid, W_THREADS, smem, s, md, a);
  }
  return check_launch(w.gamma != nullptr ? "tc_wgrad_kernel (norm prologue)" : "tc_wgrad_kernel (plain)");
}

int run_split_planes_tf32(const float* src, int64_t n, int nb, int64_t src_stride, void* hi, void* lo,
                          int64_t dst_stride, cudaStream_t s) {
  int gx = cdiv(n, 256 * 4);
  launch_kernel(split_planes_tf32_kernel, dim3(gx < 1 ? 1 : gx, nb), 256, 0, s, src, n, src_stride, reinterpret_cast<float*>(hi), reinterpret_cast<float*>(lo), dst_stride);
  return check_launch("split_planes_tf32_kernel");
}

int run_split_planes(const float* src, int R, int C, int nb, int64_t src_stride, void* hi, void* lo, int64_t dst_stride,
                     int transpose, cudaStream_t s) {
  dim3 grid(cdiv(C, 32), cdiv(R, 32), nb);
  launch_kernel(split_planes_kernel, grid, 256, 0, s, src, R, C, src_stride, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), dst_stride, transpose);
  return check_launch("split_planes_kernel");
}

}  // namespace ctn

#ifdef CTN_TRACE
extern "C" int ctn_debug_read_trace(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, ctn::g_trace, sizeof(long long) * 64 * n);
}
#endif
