// gemm_ts.cu — the 1x1 convolutions (forward and data gradients) as FRAME-MAJOR tcgen05 GEMMs with the activation
// operand in tensor memory.
//
//   D[f, o] = epi( sum_c pro(A[f, c]) * W[o, c] )            A [F, Kd] fp32, W [O, Kd] (pre-split hi/lo planes)
//
// MMA M = 128 frames (TMEM lanes), MMA N = a run of output channels, K = input channels.  The activations are the A
// operand and live in TENSOR MEMORY (tcgen05.mma "ts" form, layout measured with scratch/ts_probe.cu): raw fp32 tiles
// arrive by TMA, converter threads (thread = frame row) read their row conflict-free from the 128B-swizzled tile,
// apply the prologue (PReLU), split hi/lo and write the operand with tcgen05.st — the split planes never touch shared
// memory, which was the bottleneck of the channel-major kernel in gemm_tc.cu (raw write + split read/write + three
// operand reads per k-step).  The weights are the B operand: hi/lo planes by TMA, 128-byte swizzle, K-major.
//
// Work decomposition: the (frame tile x output channel) space is flattened into 16-channel units and cut into one
// contiguous share per CTA (persistent, one CTA per SM): with F = 9,597 frames the paper configuration has only 75
// frame tiles, and whole (tile x 256 channel) blocks would leave a 150-block grid two waves deep on 148 SMs.  A share is
// processed as segments (frame tile, first channel, n <= nmax channels); every segment streams the tile's K extent.
// Thread-block clusters: the kernel is bound by L2 -> SM delivery (measured: ~6,300 B/clk chip-wide, the TMA latency
// climbs to 4k cycles), and most of those bytes are weights that every frame tile re-reads.  A cluster of `cl` CTAs
// works on `cl` consecutive frame tiles x the SAME channel segments; each CTA loads 1/cl of a weight stage and
// TMA-multicasts it to all of them, and a stage is recycled when every CTA's MMAs have committed (multicast commit).
// Pipeline per CTA, all on mbarriers: TMA(W) -> w ring | TMA(A raw) -> raw ring -> converters -> A ring in TMEM |
// one thread issues the MMAs | accumulators in TMEM (double buffered when they fit) -> 4 epilogue warps.
// The epilogue reads lane = frame, so per-frame quantities (norm fold scalars, sample index) are per-thread; the tile
// is transposed through a swizzled shared-memory buffer so that global stores (and the residual loads) are whole
// 128-byte rows.
//
// Reduced-precision inference (HALF > 0, the "bf16 forward" of BASELINE configs[2]; no reference path, 2e-2 budget):
// ONE bf16 plane per operand and one MMA per k-step; the activations are either fp32 in memory and rounded by the
// converters (HALF = 1) or STORED as bf16 (HALF = 2: half the bytes, the converter only applies the PReLU), and the output
// can be stored as bf16 (DBF) — the two H-wide tensors of a block (z1, z2) then cost 2 bytes per element end to end.
// Precision otherwise: as gemm_tc.cu — TF32x3 for the training forward (the correction products lo*hi + hi*lo in their own
// accumulator), bf16x3 for data gradients and inference.
#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ctn {
#ifdef CTN_TS_TRACE
__device__ long long g_ts_trace[160][256];
#define TST(slot) do { if (blockIdx.x < 160 && (slot) < 256) g_ts_trace[blockIdx.x][slot] = clock64(); } while (0)
#else
#define TST(slot) do { } while (0)
#endif
namespace {

constexpr int TS_THREADS = 640;   // 4 control warps, 8 converter warps (two groups of 4), 8 epilogue warps (two per lane quarter)
constexpr int TS_EPI_WARPS = 8;
constexpr int TS_MT = 128;        // frames per tile (MMA M)
constexpr int TS_ACOLS = 64;      // TMEM columns of one A stage: hi plane 32 | lo plane 32
constexpr int TS_MAXST = 6;       // ring depth limit (barrier slots)
constexpr int TS_RAW_BOX = TS_MT * 128;  // bytes of one raw activation box: 128 frames x 32 floats
constexpr int TS_EPI_BYTES = TS_EPI_WARPS * 4096;  // transposition buffers of the epilogue warps

struct TsGemmArgs {
  float* D;
  const float* res;
  int64_t F;
  int O, Kd, K;
  int ngroups;    // groups of `cl` consecutive frame tiles
  int cl;         // cluster size (1, 2 or 4): CTAs sharing the weight stream by TMA multicast
  int nmax;       // widest segment (multiple of 16): weight stage = 2 planes of nmax rows
  int nd;         // accumulator buffers (1 or 2)
  int w_stages, raw_stages, a_stages;
  const float* alpha_in;
  const float* c1;
  const float* c2;
  NormStats st;
  double* stat_out;
  const float* alpha_out;
};

struct Seg {
  int ft, c0, n;
};
// the CTA's share of the flattened (frame tile, 16-channel unit) space, cut into segments of <= nmax channels
struct SegIter {
  int u, u_end, uo, nmax16, cl, rank;
  __device__ __forceinline__ SegIter(const TsGemmArgs& a, int cta_rank) {
    uo = a.O >> 4;
    cl = a.cl;
    rank = cta_rank;
    const uint32_t total = (uint32_t)a.ngroups * (uint32_t)uo, ncl = gridDim.x / (uint32_t)cl, cid = blockIdx.x / (uint32_t)cl;
    const uint32_t base = total / ncl, extra = total - base * ncl;  // the first `extra` clusters take one more unit
    u = (int)(cid * base + min(cid, extra));
    u_end = u + (int)base + (cid < extra ? 1 : 0);
    nmax16 = a.nmax >> 4;
  }
  __device__ __forceinline__ bool next(Seg& s) {
    if (u >= u_end) return false;
    const int grp = u / uo, cu = u - grp * uo;
    const int rem = min(u_end - u, uo - cu);
    const int nseg = (rem + nmax16 - 1) / nmax16;
    const int n16 = (rem + nseg - 1) / nseg;
    s.ft = grp * cl + rank;
    s.c0 = cu << 4;
    s.n = n16 << 4;
    u += n16;
    return true;
  }
};

template <bool TF32, bool FOLD, bool RES, bool STATS, int HALF = 0, bool DBF = false>
__global__ void __launch_bounds__(TS_THREADS, 1)
ts_gemm_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
               const __grid_constant__ CUtensorMap map_a, TsGemmArgs a) {
  static_assert(!(TF32 && HALF), "the single-plane mode is a bf16 flavour");
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int KB = TF32 ? 32 : 64;                        // K elements per k-block (128 bytes of a weight row)
  // raw activation stage: 32 fp32 per row (TF32), 64 fp32 = two boxes (bf16 from fp32), or 64 bf16 = one box (HALF = 2)
  constexpr int RAW_STAGE = (TF32 || HALF == 2) ? TS_RAW_BOX : 2 * TS_RAW_BOX;
  constexpr int NACC = TF32 ? 2 : 1;                        // accumulators per output (main, correction)
  constexpr int NPL = HALF ? 1 : 2;                         // operand planes (hi, lo)
  const int WST = a.w_stages, RST = a.raw_stages, AST = a.a_stages, ND = a.nd;
  const int w_plane = a.nmax * 128, w_stage = NPL * w_plane;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t raw_base = smem_base + WST * w_stage;
  const uint32_t epi_base = raw_base + RST * RAW_STAGE;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + WST * w_stage + RST * RAW_STAGE + TS_EPI_BYTES);
  uint64_t* w_empty = w_full + TS_MAXST;
  uint64_t* raw_full = w_empty + TS_MAXST;
  uint64_t* raw_empty = raw_full + TS_MAXST;
  uint64_t* a_full = raw_empty + TS_MAXST;
  uint64_t* a_empty = a_full + TS_MAXST;
  uint64_t* d_full = a_empty + TS_MAXST;
  uint64_t* d_empty = d_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = a.Kd / KB;
  if (threadIdx.x == 0) TST(0);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_hi);
    prefetch_tensormap(&map_lo);
  } else if (warp == 3 && lane == 0) {
    prefetch_tensormap(&map_a);
  }
  if (warp == 1 && lane == 0) {
    if (smem_base & 1023u) __trap();  // SWIZZLE_128B operands need a 1024-byte aligned base
    for (int s = 0; s < TS_MAXST; ++s) {
      mbar_init(w_full + s, 1);
      mbar_init(w_empty + s, a.cl);   // one commit per CTA of the cluster
      mbar_init(raw_full + s, 1);
      mbar_init(raw_empty + s, 128);  // the converter group that read the tile
      mbar_init(a_full + s, 128);     // the converter group that wrote the stage
      mbar_init(a_empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(d_full + s, 1);
      mbar_init(d_empty + s, 32 * TS_EPI_WARPS);  // the epilogue threads
    }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc<512>(tmem_ptr);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int CL = a.cl;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
  if (CL > 1) cluster_sync();  // every CTA's barriers are initialised before a peer multicasts into it
  // this CTA holds its tensor memory now: a dependent kernel's CTAs may be scheduled (they wait for the dealloc below)
  pdl_launch_dependents();
  pdl_wait();  // global data (weight planes, activations) may come from the kernel right before this one
  if (threadIdx.x == 0) TST(1);

  const int dstride = NACC * a.nmax;                        // TMEM columns of one accumulator buffer
  const uint32_t a_col0 = 512u - (uint32_t)(AST * TS_ACOLS);  // the A ring sits at the top of the 512 columns

  if (warp == 0) {
    // ===== TMA producer: weight hi/lo planes.  One box per plane and CTA: this CTA's slice (nmax / cl rows) of the stage,
    // multicast to every CTA of the cluster (rows past the segment are not read by the MMAs, rows past O are zero-filled).
    // Measured: a bulk tensor copy costs ~80 issue cycles whatever its size, so 16-row boxes (18 per stage) made the
    // producer the bottleneck of the whole pipeline (1,500 cycles per k-block). =====
    if (lane == 0) {
      int it = 0;
      SegIter si(a, rank);
      Seg s;
      const int slice = a.nmax / CL;
      while (si.next(s)) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int ws = it % WST, ph = (it / WST) & 1;
          mbar_wait(w_empty + ws, ph ^ 1);
          if (it < 40) TST(48 + it);
          uint8_t* st = smem + ws * w_stage + rank * slice * 128;
          mbar_expect_tx(w_full + ws, w_stage);  // all the slices land here
          if (CL > 1) {
            tma_load_2d_mc(st, &map_hi, w_full + ws, kb * KB, s.c0 + rank * slice, cmask);
            if (NPL == 2) tma_load_2d_mc(st + w_plane, &map_lo, w_full + ws, kb * KB, s.c0 + rank * slice, cmask);
          } else {
            tma_load_2d(st, &map_hi, w_full + ws, kb * KB, s.c0);
            if (NPL == 2) tma_load_2d(st + w_plane, &map_lo, w_full + ws, kb * KB, s.c0);
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===== TMA producer: raw fp32 activation tiles [128 frames x 32 floats]; rows past F are zero-filled =====
    if (lane == 0) {
      int it = 0;
      SegIter si(a, rank);
      Seg s;
      while (si.next(s)) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int r = it % RST, ph = (it / RST) & 1;
          mbar_wait(raw_empty + r, ph ^ 1);
          uint8_t* dst = smem + WST * w_stage + r * RAW_STAGE;
          mbar_expect_tx(raw_full + r, RAW_STAGE);
          tma_load_2d(dst, &map_a, raw_full + r, kb * KB, s.ft * TS_MT);
          if (!TF32 && HALF != 2) tma_load_2d(dst + TS_RAW_BOX, &map_a, raw_full + r, kb * KB + 32, s.ft * TS_MT);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t dhi = desc_hi_sw128(1024);
      int it = 0, sidx = 0;
      SegIter si(a, rank);
      Seg s;
      while (si.next(s)) {
        const int db = sidx % ND, dph = (sidx / ND) & 1;
        mbar_wait(d_empty + db, dph ^ 1);  // the epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t idesc = make_idesc(TS_MT, s.n, 0, 0, TF32 ? 2u : 1u);
        const uint32_t d_main = tmem_base + (uint32_t)(db * dstride);
        const uint32_t d_corr = d_main + (uint32_t)a.nmax;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int ws = it % WST, wph = (it / WST) & 1;
          const int as = it % AST, aph = (it / AST) & 1;
          mbar_wait(w_full + ws, wph);
          if (it < 40) TST(8 + it);
          mbar_wait(a_full + as, aph);
          if (it < 40) TST(136 + it);
          tc_fence_after();
          const uint32_t sb = smem_base + ws * w_stage;
          uint32_t w_hi = desc_lo(sb, 16), w_lo = desc_lo(sb + w_plane, 16);
          uint32_t a_hi = tmem_base + a_col0 + (uint32_t)(as * TS_ACOLS), a_lo = a_hi + 32;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // one MMA consumes 32 bytes of K per weight row = 8 TMEM columns of A
            const uint32_t first = (kb | k) != 0;
            if (TF32) {
              umma_ts<true>(d_corr, a_hi, w_lo, dhi, idesc, first);  // hi*lo -> correction
              umma_ts<true>(d_corr, a_lo, w_hi, dhi, idesc, 1);      // lo*hi -> correction
              umma_ts<true>(d_main, a_hi, w_hi, dhi, idesc, first);  // hi*hi -> main
            } else if (HALF) {
              umma_ts<false>(d_main, a_hi, w_hi, dhi, idesc, first);
            } else {
              umma_ts<false>(d_main, a_hi, w_lo, dhi, idesc, first);
              umma_ts<false>(d_main, a_lo, w_hi, dhi, idesc, 1);
              umma_ts<false>(d_main, a_hi, w_hi, dhi, idesc, 1);
            }
            w_hi += 2; w_lo += 2; a_hi += 8; a_lo += 8;
          }
          if (CL > 1) umma_commit_mc(w_empty + ws, cmask);  // the weight stage is shared: tell every CTA of the cluster
          else umma_commit(w_empty + ws);                     // both rings are free once these MMAs have read them
          umma_commit(a_empty + as);
        }
        umma_commit(d_full + db);
        ++sidx;
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===== converters: own row of the raw tile -> prologue -> hi/lo split -> tcgen05.st into the A ring =====
    const int q = warp & 3, grp = (warp - 4) >> 2;  // TMEM lane quarter; the two groups take k-blocks alternately
    const int row = q * 32 + lane;
    const bool pro = a.alpha_in != nullptr;
    const float alpha_in = pro ? __ldg(a.alpha_in) : 1.f;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + a_col0;
    const int sw = row & 7;
    int it = 0;
    SegIter si(a, rank);
    Seg s;
    while (si.next(s)) {
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        if ((it & 1) != grp) continue;
        const int r = it % RST, rph = (it / RST) & 1;
        const int as = it % AST, aph = (it / AST) & 1;
        const uint32_t raw = raw_base + r * RAW_STAGE + row * 128;
        const uint32_t dst = lane_addr + (uint32_t)(as * TS_ACOLS);
        mbar_wait(raw_full + r, rph);
        if (row == 0 && it < 40) TST(88 + it);
        if (TF32) {
          float4 x[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) x[c] = lds128f(raw + ((c ^ sw) << 4));
          mbar_wait(a_empty + as, aph ^ 1);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float4 y = x[4 * h + c];
              if (pro) { y.x = prelu(y.x, alpha_in); y.y = prelu(y.y, alpha_in); y.z = prelu(y.z, alpha_in); y.w = prelu(y.w, alpha_in); }
              uint4 hv, lv;
              split4_tf32(y, hv, lv);
              hi[4 * c] = hv.x; hi[4 * c + 1] = hv.y; hi[4 * c + 2] = hv.z; hi[4 * c + 3] = hv.w;
              lo[4 * c] = lv.x; lo[4 * c + 1] = lv.y; lo[4 * c + 2] = lv.z; lo[4 * c + 3] = lv.w;
            }
            tmem_st16(dst + 16 * h, hi);
            tmem_st16(dst + 32 + 16 * h, lo);
          }
        } else if (HALF == 2) {
          // the activations are stored as bf16: the row's 128 bytes ARE the operand (64 elements = 32 packed columns);
          // only the PReLU prologue touches them (unpack, prelu, round back)
          uint4 x[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 t = lds128f(raw + ((c ^ sw) << 4));
            x[c] = make_uint4(__float_as_uint(t.x), __float_as_uint(t.y), __float_as_uint(t.z), __float_as_uint(t.w));
          }
          mbar_wait(a_empty + as, aph ^ 1);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t hi[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t w4[4] = {x[4 * h + c].x, x[4 * h + c].y, x[4 * h + c].z, x[4 * h + c].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint32_t w = w4[i];
                if (pro) {
                  const float e0 = prelu(__uint_as_float(w << 16), alpha_in), e1 = prelu(__uint_as_float(w & 0xffff0000u), alpha_in);
                  w = cvt_bf16x2(e0, e1);
                }
                hi[4 * c + i] = w;
              }
            }
            tmem_st16(dst + 16 * h, hi);
          }
        } else {
          float4 x[2][8];  // the two 32-float boxes of the k-block
#pragma unroll
          for (int c = 0; c < 8; ++c) x[0][c] = lds128f(raw + ((c ^ sw) << 4));
          mbar_wait(a_empty + as, aph ^ 1);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // 32 floats -> 16 packed hi columns + 16 packed lo columns
            if (h == 0) {
#pragma unroll
              for (int c = 0; c < 8; ++c) x[1][c] = lds128f(raw + TS_RAW_BOX + ((c ^ sw) << 4));
            }
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 y0 = x[h][2 * c], y1 = x[h][2 * c + 1];
              float v[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
              if (pro) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = prelu(v[i], alpha_in);
              }
              uint4 hv, lv;
              split8(v, hv, lv);  // element 2i in the low half of word i: the order the tensor core expects
              hi[4 * c] = hv.x; hi[4 * c + 1] = hv.y; hi[4 * c + 2] = hv.z; hi[4 * c + 3] = hv.w;
              lo[4 * c] = lv.x; lo[4 * c + 1] = lv.y; lo[4 * c + 2] = lv.z; lo[4 * c + 3] = lv.w;
            }
            tmem_st16(dst + 16 * h, hi);
            if (!HALF) tmem_st16(dst + 32 + 16 * h, lo);  // (HALF = 1: the activations are rounded to one bf16 plane)
          }
        }
        mbar_arrive(raw_empty + r);  // the tile's values have been consumed by the conversions above
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(a_full + as);
        if (row == 0 && it < 40) TST(176 + it);
      }
    }
  } else if (warp >= 12) {
    // ===== epilogue: TMEM (lane = frame) -> registers -> transposed through shared memory -> whole-row global stores;
    // two warps per TMEM lane quarter take alternate 32-column chunks =====
    const int q = warp & 3, hsel = (warp - 12) >> 2;
    const int row = q * 32 + lane;
    const int O = a.O;
    const uint32_t tbuf = epi_base + (warp - 12) * 4096;  // [32 rows][8 chunks of 16 B], chunk index XOR (row & 7)
    const float alpha_out = (STATS && a.alpha_out) ? __ldg(a.alpha_out) : 1.f;
    int sidx = 0;
    SegIter si(a, rank);
    Seg s;
    while (si.next(s)) {
      const int db = sidx % ND, dph = (sidx / ND) & 1;
      const int64_t f = (int64_t)s.ft * TS_MT + row;
      const bool valid = f < a.F;
      int m = -1;
      float mu = 0.f, r = 1.f;
      if (valid) {
        m = (int)((uint32_t)f / (uint32_t)a.K);
        if (FOLD) load_stats(a.st, m, f, mu, r);
      }
      const float mur = mu * r;
      float s1 = 0.f, s2 = 0.f;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(db * dstride);
      if (RES) {
        // The residual tile [128 frames x n channels] is pulled into L2 while the MMAs of this segment run: the coalesced
        // phase below reads it in dependent rounds of 4 row-loads per 32-column chunk (8 rounds per 256-column tile), and
        // from HBM each round cost ~2k cycles — the whole epilogue of the contracting conv was residual-load latency
        // (8 x 60 s forward: 436 us per launch against a 241 us HBM bound, with the MMAs waiting on the single
        // accumulator buffer).
        const int et = (warp - 12) * 32 + lane, lpr = (s.n + 31) >> 5;  // 128-byte lines per row
        for (int l = et; l < TS_MT * lpr; l += 32 * TS_EPI_WARPS) {
          const int rr = l / lpr, cc = l - rr * lpr;
          const int64_t ff = (int64_t)s.ft * TS_MT + rr;
          if (ff < a.F) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + ff * O + s.c0 + cc * 32));
        }
      }
      mbar_wait(d_full + db, dph);
      tc_fence_after();
      if (row == 0 && sidx < 8) TST(216 + sidx);
      bool released = false;
      for (int j = 32 * hsel; j < s.n; j += 64) {
        const bool full = j + 32 <= s.n;  // else 16 columns: the tail of a segment whose width is an odd multiple of 16
        uint32_t rm[2][16], rc[2][16];
        tmem_ld16_issue(taddr + (uint32_t)j, rm[0]);
        if (TF32) tmem_ld16_issue(taddr + (uint32_t)(a.nmax + j), rc[0]);
        if (full) {
          tmem_ld16_issue(taddr + (uint32_t)(j + 16), rm[1]);
          if (TF32) tmem_ld16_issue(taddr + (uint32_t)(a.nmax + j + 16), rc[1]);
        }
        tmem_ld_wait();
        if (j + 64 >= s.n) {  // this warp's last read of the accumulator buffer: hand it back to the MMA issuer
          tc_fence_before();
          mbar_arrive(d_empty + db);
          released = true;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && !full) break;
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            v[i] = TF32 ? __uint_as_float(rm[h][i]) + __uint_as_float(rc[h][i]) : __uint_as_float(rm[h][i]);
          if (FOLD) {
            const float4* c1 = reinterpret_cast<const float4*>(a.c1 + s.c0 + j + 16 * h);
            const float4* c2 = reinterpret_cast<const float4*>(a.c2 + s.c0 + j + 16 * h);
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // warp-uniform addresses: broadcast loads
              const float4 k1 = __ldg(c1 + i), k2 = __ldg(c2 + i);
              v[4 * i] = fmaf(r, v[4 * i], fmaf(-mur, k2.x, k1.x));
              v[4 * i + 1] = fmaf(r, v[4 * i + 1], fmaf(-mur, k2.y, k1.y));
              v[4 * i + 2] = fmaf(r, v[4 * i + 2], fmaf(-mur, k2.z, k1.z));
              v[4 * i + 3] = fmaf(r, v[4 * i + 3], fmaf(-mur, k2.w, k1.w));
            }
          }
          if (DBF) {  // the consumers (and the statistics) see the stored, rounded values
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
          }
          if (STATS && valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float p = prelu(v[i], alpha_out);
              s1 += p;
              s2 = fmaf(p, p, s2);
            }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
            sts128(tbuf + lane * 128 + (((4 * h + c) ^ (lane & 7)) << 4),
                   make_uint4(__float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]), __float_as_uint(v[4 * c + 2]),
                              __float_as_uint(v[4 * c + 3])));
        }
        __syncwarp();
        // coalesced phase: all shared loads (and residual loads) first, then the stores
        const int64_t fq = (int64_t)s.ft * TS_MT + q * 32;
        if (full) {  // 8 lanes per row, 4 rows per instruction
          const int c4 = lane & 7, rsub = lane >> 3;
#pragma unroll
          for (int t0 = 0; t0 < 8; t0 += 4) {
            float4 val[4], rv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int rr = 4 * (t0 + t) + rsub;
              val[t] = lds128f(tbuf + rr * 128 + ((c4 ^ (rr & 7)) << 4));
              if (RES) {
                rv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (fq + rr < a.F) rv[t] = __ldg(reinterpret_cast<const float4*>(a.res + (fq + rr) * O + s.c0 + j + c4 * 4));
              }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int rr = 4 * (t0 + t) + rsub;
              if (RES) { val[t].x += rv[t].x; val[t].y += rv[t].y; val[t].z += rv[t].z; val[t].w += rv[t].w; }
              if (fq + rr < a.F) {
                if (DBF) stg_bf16x4(reinterpret_cast<__nv_bfloat16*>(a.D) + (fq + rr) * O + s.c0 + j + c4 * 4, val[t]);
                else stg128(a.D + (fq + rr) * O + s.c0 + j + c4 * 4, val[t]);
              }
            }
          }
        } else {     // 4 lanes per row, 8 rows per instruction
          const int c4 = lane & 3, rsub = lane >> 2;
          float4 val[4], rv[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int rr = 8 * t + rsub;
            val[t] = lds128f(tbuf + rr * 128 + ((c4 ^ (rr & 7)) << 4));
            if (RES) {
              rv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (fq + rr < a.F) rv[t] = __ldg(reinterpret_cast<const float4*>(a.res + (fq + rr) * O + s.c0 + j + c4 * 4));
            }
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int rr = 8 * t + rsub;
            if (RES) { val[t].x += rv[t].x; val[t].y += rv[t].y; val[t].z += rv[t].z; val[t].w += rv[t].w; }
            if (fq + rr < a.F) {
              if (DBF) stg_bf16x4(reinterpret_cast<__nv_bfloat16*>(a.D) + (fq + rr) * O + s.c0 + j + c4 * 4, val[t]);
              else stg128(a.D + (fq + rr) * O + s.c0 + j + c4 * 4, val[t]);
            }
          }
        }
        __syncwarp();
      }
      if (!released) {  // a segment narrower than this warp's first chunk
        tc_fence_before();
        mbar_arrive(d_empty + db);
      }
      if (STATS) {  // per-sample sums: the rows of a warp span at most a few samples
        const int m_lo = __shfl_sync(0xffffffffu, m, 0);
        int m_hi = m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m_hi = max(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, o));
        if (m_lo >= 0) {
          for (int mm = m_lo; mm <= m_hi; ++mm) {
            const double d1 = warp_sum(m == mm ? (double)s1 : 0.0), d2 = warp_sum(m == mm ? (double)s2 : 0.0);
            if (lane == 0) {
              atomicAdd(a.stat_out + 2 * mm, d1);
              atomicAdd(a.stat_out + 2 * mm + 1, d2);
            }
          }
        }
      }
      if (row == 0 && sidx < 8) TST(224 + sidx);
      ++sidx;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync();  // no peer multicasts into / commits onto this CTA's shared memory after it has exited
  if (warp == 2) tmem_dealloc<512>(tmem_base);
  if (threadIdx.x == 0) TST(2);
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

// 2-D row-major tensor [rows, cols] -> tensor map with a [box_rows x 128 bytes] box, 128-byte swizzle, zero fill
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int cols, bool f32, int box_rows) {
  EncodeTiledFn enc = get_encode();
  CTN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * (f32 ? 4 : 2)};
  const cuuint32_t box[2] = {(cuuint32_t)(f32 ? 32 : 64), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CTN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)r);
  return 0;
}

static int sm_count() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (n[dev] == 0) {
    if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e != nullptr && e[0] != 0) ? atoi(e) : dflt;
}

}  // namespace

bool ts_gemm_eligible(const GemmArgs& a) {
  const int kb = a.tf32 ? 32 : 64;
  if (a.half && a.tf32) return false;
  return a.W_hi != nullptr && (a.W_lo != nullptr || a.half) && a.Kd % kb == 0 && a.O % 16 == 0 && a.F >= 1 &&
         a.F < ((int64_t)1 << 31) - TS_MT && a.nred_z == nullptr;
}

// launch with a cluster dimension (and programmatic dependent launch on eager streams, see launch_kernel in common.cuh)
template <typename... KArgs, typename... Args>
static void launch_clustered(void (*kernel)(KArgs...), int grid, int cl, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &cap);
  if (pdl_enabled() && cap == cudaStreamCaptureStatusNone) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cl > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cl;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors are picked up by check_launch()
  timing_note_stream(stream);
}

// clusters of `cl` CTAs with `smem` bytes each that the device can hold at once (every flavour has the same block size
// and register bound, so one instantiation answers for all); cached per (device, cl, shared-memory size)
static int max_clusters(int cl, size_t smem) {
  if (cl == 1) return sm_count();
  struct Entry { int dev, cl; size_t smem; int n; };
  static Entry cache[32];
  static int ncache = 0;
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < ncache; ++i)
    if (cache[i].dev == dev && cache[i].cl == cl && cache[i].smem == smem) return cache[i].n;
  auto kernel = ts_gemm_kernel<false, false, false, false>;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sm_count() / cl * cl);
  cfg.blockDim = dim3(TS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = 0;
  }
  if (n > sm_count() / cl) n = sm_count() / cl;
  if (ncache < 32) cache[ncache++] = Entry{dev, cl, smem, n};
  return n;
}

int launch_gemm_ts(const GemmArgs& g, cudaStream_t s) {
  const bool tf32 = g.tf32 != 0;
  const bool fold = g.c1 != nullptr, res = g.res != nullptr, stats = g.stat_out != nullptr;
  TsGemmArgs a;
  a.D = g.D; a.res = g.res; a.F = g.F; a.O = g.O; a.Kd = g.Kd; a.K = g.K;
  a.alpha_in = g.alpha_in; a.c1 = g.c1; a.c2 = g.c2; a.st = g.st; a.stat_out = g.stat_out; a.alpha_out = g.alpha_out;
  const int ntiles = (int)((g.F + TS_MT - 1) / TS_MT);
  // cluster size: frame tiles that share one weight stream (CTN_TS_CL overrides; 1 when there are too few tiles)
  static const int cl_env = env_int("CTN_TS_CL", 0);
  int cl = cl_env > 0 ? cl_env : 2;
  if (cl != 1 && cl != 2 && cl != 4) cl = 2;
  while (cl > 1 && ntiles < 2 * cl) cl >>= 1;
  const int nacc = tf32 ? 2 : 1;
  const int npl = g.half ? 1 : 2;
  const size_t raw_stage = (tf32 || g.half == 2) ? TS_RAW_BOX : 2 * TS_RAW_BOX;
  const size_t fixed = TS_EPI_BYTES + 1024, budget = 227 * 1024;
  int nclusters = 0;
  size_t smem = 0;
  for (;; cl >>= 1) {  // (falls back to smaller clusters if the device cannot hold them)
    a.cl = cl;
    a.ngroups = (ntiles + cl - 1) / cl;
    const int64_t units = (int64_t)a.ngroups * (g.O / 16);
    static const int grid_env = env_int("CTN_TS_GRID", 0);
    int want = (grid_env > 0 ? grid_env : sm_count()) / cl;  // clusters: one CTA per SM, >= 64 output channels per cluster
    if (want > (units + 3) / 4) want = (int)((units + 3) / 4);
    if (want < 1) want = 1;
    // segment width: the share of a cluster in as few equal segments as the accumulator columns allow
    static const int ncap_env = env_int("CTN_TS_NCAP", 0);
    // (measured on B200: 256 against 160 — a contracting conv is then ONE segment per frame tile, so its activations are
    // converted once instead of twice, at the price of a single-buffered accumulator — 8 x 60 s forward 40.8 -> 38.7 ms,
    // bf16 32 x 4 s forward 9.50 -> 9.26 ms, 16 x 4 s training 25.75 -> 25.40 ms, headline unchanged)
    const int ncap = ncap_env > 0 ? ncap_env : 256;
    int share = (int)((units + want - 1) / want) * 16;
    if (share > g.O) share = g.O;
    const int nseg = (share + ncap - 1) / ncap;
    const int gran = 8 * cl < 16 ? 16 : 8 * cl;  // a CTA's slice of a weight stage is whole 8-row swizzle atoms
    a.nmax = ((share + nseg - 1) / nseg + gran - 1) / gran * gran;
    if (a.nmax > 256) a.nmax = 256;
    a.nd = (2 * nacc * a.nmax + 2 * TS_ACOLS <= 512) ? 2 : 1;
    CTN_REQUIRE(a.nd * nacc * a.nmax + 2 * TS_ACOLS <= 512, "ts_gemm: segment width %d does not fit tensor memory", a.nmax);
    a.a_stages = (512 - a.nd * nacc * a.nmax) / TS_ACOLS;
    if (a.a_stages > TS_MAXST) a.a_stages = TS_MAXST;
    // shared memory: weight ring + raw activation ring + transposition buffers + barriers
    const size_t w_stage = npl * (size_t)a.nmax * 128;
    // Ring depths of the rings the CONVERTERS wait on (raw tiles, TMEM A stages) must be EVEN: the two converter groups
    // take k-blocks alternately, so with an odd depth a stage alternates between the groups and each group only observes
    // every other phase of the stage's barrier — always with the same parity.  mbarrier.try_wait.parity then cannot tell
    // "my phase completed" from "the phase two before it completed": a group that runs two uses ahead reads a tile
    // that has not landed (or overwrites an operand the MMAs have not read) and arrives in the wrong phase.  Seen with a
    // 3-deep raw ring and bf16-stored activations (256-cycle k-blocks: wrong tiles, then hangs / launch failures at
    // F = 102k, scratch/run21.sh); latent for a 3-deep A ring.  With even depths a stage belongs to one group, which sees
    // every phase.  (The weight ring and the accumulators are waited on by single threads / all epilogue warps.)
    static const int rst4 = env_int("CTN_TS_RST4", 0);  // experiment: 4-deep raw ring for the one-box flavours
    a.raw_stages = (rst4 && raw_stage == TS_RAW_BOX) ? 4 : 2;
    a.a_stages &= ~1;
    CTN_REQUIRE(a.a_stages >= 2, "ts_gemm: no room for two operand stages in tensor memory (nmax %d)", a.nmax);
    a.w_stages = (int)((budget - fixed - a.raw_stages * raw_stage) / w_stage);
    if (a.w_stages > TS_MAXST) a.w_stages = TS_MAXST;
    CTN_REQUIRE(a.w_stages >= 2, "ts_gemm: shared memory budget exceeded (nmax %d)", a.nmax);
    {  // debug overrides of the ring depths
      static const int wst = env_int("CTN_TS_WST", 0), rst = env_int("CTN_TS_RST", 0), ast = env_int("CTN_TS_AST", 0);
      if (wst >= 2 && wst <= a.w_stages) a.w_stages = wst;
      if (rst >= 2 && rst <= 4) a.raw_stages = rst & ~1;
      if (ast >= 2 && ast <= a.a_stages) a.a_stages = ast & ~1;
    }
    smem = a.w_stages * w_stage + a.raw_stages * raw_stage + fixed;
    nclusters = want;
    if (cl == 1) break;
    const int fit = max_clusters(cl, smem);
    if (fit >= 1) {
      if (nclusters > fit) nclusters = fit;
      break;
    }
  }
  const int grid = nclusters * cl;
  CUtensorMap mh, ml, ma;
  CTN_TRY(make_map(&mh, g.W_hi, g.O, g.Kd, tf32, a.nmax / cl));
  CTN_TRY(make_map(&ml, g.half ? g.W_hi : g.W_lo, g.O, g.Kd, tf32, a.nmax / cl));  // (unused in the single-plane mode)
  CTN_TRY(make_map(&ma, g.A, g.F, g.Kd, g.half != 2, TS_MT));  // fp32 rows of 32 floats, or bf16 rows of 64 elements
#define CTN_TS_LAUNCH(...)                                                                                       \
  do {                                                                                                           \
    static unsigned long long attr_mask = 0; /* the attributes are per device */                                 \
    int dev__ = 0;                                                                                               \
    CTN_CUDA(cudaGetDevice(&dev__));                                                                             \
    if (!((attr_mask >> (dev__ & 63)) & 1ull)) {                                                                 \
      CTN_CUDA(cudaFuncSetAttribute(ts_gemm_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                    227 * 1024));                                                                \
      attr_mask |= 1ull << (dev__ & 63);                                                                         \
    }                                                                                                            \
    launch_clustered(ts_gemm_kernel<__VA_ARGS__>, grid, cl, smem, s, mh, ml, ma, a);                             \
  } while (0)
  if (g.half) {  // reduced-precision inference: the combinations the model's forward issues
    const bool dbf = g.d_bf16 != 0;
    if (g.half == 1 && dbf && !fold && !res && stats) CTN_TS_LAUNCH(false, false, false, true, 1, true);
    else if (g.half == 1 && dbf && !fold && !res && !stats) CTN_TS_LAUNCH(false, false, false, false, 1, true);
    else if (g.half == 1 && !dbf && fold && !res && !stats) CTN_TS_LAUNCH(false, true, false, false, 1, false);
    else if (g.half == 1 && !dbf && !fold && !res && !stats) CTN_TS_LAUNCH(false, false, false, false, 1, false);
    else if (g.half == 2 && !dbf && fold && res && !stats) CTN_TS_LAUNCH(false, true, true, false, 2, false);
    else if (g.half == 2 && !dbf && !fold && !res && !stats) CTN_TS_LAUNCH(false, false, false, false, 2, false);
    else return -1;
  } else if (tf32) {
    if (!fold && !res && !stats) CTN_TS_LAUNCH(true, false, false, false);
    else if (!fold && !res && stats) CTN_TS_LAUNCH(true, false, false, true);
    else if (fold && !res && !stats) CTN_TS_LAUNCH(true, true, false, false);
    else if (fold && res && !stats) CTN_TS_LAUNCH(true, true, true, false);
    else return -1;  // a combination the model never issues: the caller falls back
  } else {
    if (!fold && !res && !stats) CTN_TS_LAUNCH(false, false, false, false);
    else if (!fold && !res && stats) CTN_TS_LAUNCH(false, false, false, true);
    else if (fold && !res && !stats) CTN_TS_LAUNCH(false, true, false, false);
    else if (fold && res && !stats) CTN_TS_LAUNCH(false, true, true, false);
    else if (!fold && res && !stats) CTN_TS_LAUNCH(false, false, true, false);
    else return -1;
  }
#undef CTN_TS_LAUNCH
  if (g.half) return check_launch(g.O > g.Kd ? "ts_gemm_kernel<bf16 x1> up" : "ts_gemm_kernel<bf16 x1> down");
  return check_launch(g.tf32 ? (g.O > g.Kd ? "ts_gemm_kernel<tf32> up" : "ts_gemm_kernel<tf32> down")
                             : (g.O > g.Kd ? "ts_gemm_kernel<bf16> up" : "ts_gemm_kernel<bf16> down"));
}

}  // namespace ctn

#ifdef CTN_TS_TRACE
extern "C" int ctn_debug_read_ts_trace(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, ctn::g_ts_trace, sizeof(long long) * 256 * n);
}
#endif
