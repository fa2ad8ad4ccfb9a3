// model.cu — parameter/workspace geometry and the whole-path orchestration:
// ConvTasNet.forward (src/conv_tasnet.py:45-60) and its hand-written backward, as a fixed sequence
// of kernel launches on one stream (graph-capturable: no allocation, no host sync, no host-side
// data-dependent control flow).
#include <stdarg.h>

#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ctn {

// ---- error plumbing ------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CTN_NO_PDL");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}
static unsigned long long g_launches = 0;  // kernels launched through this library (bench.py's gpu_launches)

// ---- CTN_TIMING=1: in-place per-kernel timing (debug) ------------------------------------------
struct TimingRec {
  const char* name;
  cudaEvent_t ev;
};
static bool timing_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CTN_TIMING");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
static TimingRec* g_trec = nullptr;
static int g_ntrec = 0, g_trec_cap = 0;
static thread_local cudaStream_t g_tstream = nullptr;
void timing_note_stream(cudaStream_t s) { g_tstream = s; }
static void timing_mark(const char* what) {
  if (g_ntrec == g_trec_cap) {
    g_trec_cap = g_trec_cap ? 2 * g_trec_cap : 4096;
    g_trec = (TimingRec*)realloc(g_trec, sizeof(TimingRec) * g_trec_cap);
  }
  cudaEvent_t ev;
  cudaEventCreate(&ev);
  cudaEventRecord(ev, g_tstream);
  g_trec[g_ntrec++] = TimingRec{what, ev};
}

int check_launch(const char* what) {
  __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED);
  if (timing_enabled()) timing_mark(what);
  static const int debug_launch = getenv("CTN_DEBUG_LAUNCH") != nullptr && getenv("CTN_DEBUG_LAUNCH")[0] == '1';
  if (debug_launch) {  // debug: name every launch and wait for it (finds the kernel that faults or never ends)
    fprintf(stderr, "[ctn] %s ...", what);
    fflush(stderr);
    const cudaError_t e = cudaDeviceSynchronize();
    fprintf(stderr, " %s\n", cudaGetErrorString(e));
    fflush(stderr);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// ---- launchers defined in the other translation units ---------------------------------------
int launch_gemm_simt(const GemmArgs& a, cudaStream_t s);
int launch_wgrad_simt(const WgradArgs& a, cudaStream_t s);
int run_encoder_fwd(const float*, const float*, int, int, int, int, float*, cudaStream_t);
int run_encoder_bwd(const float*, const float*, const float*, const float*, int, int, int, int, float*, cudaStream_t);
int run_row_stats(const float*, const float*, int64_t, int, float*, cudaStream_t, int bf16 = 0);
bool dwconv_fwd_fuses_rowstats(int H, int P, int bf16);
int run_prep_normfold(const float*, const float*, const float*, int, int, int, int64_t, float*, float*, float*, int64_t,
                      int64_t, cudaStream_t);
int run_dwconv_fwd(const float*, const float*, NormStats, const float*, const float*, const float*, int, int, int, int,
                   int, int, float*, double*, const float*, cudaStream_t, int bf16 = 0, float* rs1_out = nullptr,
                   float* rs2_out = nullptr);
int run_dwconv_bwd(const float*, const float*, const float*, NormStats, const float*, const float*, const float*, int,
                   int, int, int, int, int, float*, float*, float*, float*, double*, float*, int, cudaStream_t);
int run_norm_bwd_reduce(const float*, const float*, const float*, NormStats, const float*, int, int, int, float*,
                        float*, double*, float*, int, cudaStream_t);
int run_dwconv_bwd_gln_fused(const float*, const float*, const float*, NormStats, const float*, const double*, float*,
                             const float*, const float*, NormStats, const float*, const float*, const float*, int, int, int,
                             int, int, int, float*, float*, float*, float*, double*, float*, int, cudaStream_t);
int64_t dwconv_bwd_partial_floats(int M, int K, int H, int P, int dil);
int64_t norm_bwd_partial_floats(int M, int K, int Ch);
int dwconv_bwd_blocks(int M, int K, int dil);
int norm_bwd_blocks(int M, int K);
int run_fold_batch(const FoldBatch& fb, int n_entries, cudaStream_t s);
int run_norm_bwd_apply(float*, const float*, const float*, NormStats, const float*, const double*, int, int, int,
                       float*, cudaStream_t);
int run_decoder_fwd(const float*, const float*, const float*, int, int, int, int, int, int, int, float*, cudaStream_t);
int run_decoder_bwd(const float*, const float*, const float*, const float*, int, int, int, int, int, int, int, float*,
                    float*, float*, cudaStream_t);

int run_bn_forward_stats(const float*, const float*, const float*, const float*, float*, float*, int64_t, int, int, double*,
                         float*, float*, float*, float*, float*, cudaStream_t);
int run_bn_bwd_finalize(const float*, const float*, const float*, const float*, const float*, int, int64_t, int, float*,
                        float*, float*, float*, cudaStream_t);
int run_bn_bwd_apply(float*, const float*, const float*, const float*, const float*, const float*, const float*, int64_t,
                     int, float*, cudaStream_t);

bool tc_gemm_eligible(const GemmArgs& a);
int launch_gemm_tc(const GemmArgs& a, cudaStream_t s);
bool ts_gemm_eligible(const GemmArgs& a);
int launch_gemm_ts(const GemmArgs& a, cudaStream_t s);  // -1: epilogue combination not built (fall back)
bool tc_wgrad_eligible(const WgradArgs& a);
int launch_wgrad_tc(const WgradArgs& a, cudaStream_t s);
int run_split_planes(const float*, int, int, int, int64_t, void*, void*, int64_t, int, cudaStream_t);
int run_split_planes_tf32(const float*, int64_t, int, int64_t, void*, void*, int64_t, cudaStream_t);

// CTN_FORCE_SIMT=1 routes every GEMM to the fp32 CUDA-core kernels (A/B debugging only; the default is tcgen05)
static bool force_simt() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CTN_FORCE_SIMT");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// Library-owned device scratch of the STAND-ALONE single-kernel entry points (weight planes, partial rows); the
// whole-model path never comes here (it works in the caller's workspace).  One buffer per (device, slot, calling
// thread) — two host threads driving one device never share a buffer — grown under a mutex; growing synchronises the
// device first (cudaFree of the old buffer would do so anyway: an earlier launch of this thread may still read it).
int lib_scratch(size_t bytes, void** out, int slot) {
  struct Slot { void* buf; size_t cap; };
  static thread_local Slot slots[16][2] = {};
  static std::mutex mu;
  int dev = 0;
  CTN_CUDA(cudaGetDevice(&dev));
  CTN_REQUIRE(dev < 16 && slot >= 0 && slot < 2, "lib_scratch: bad device/slot");
  Slot& sl = slots[dev][slot];
  if (sl.cap < bytes) {
    std::lock_guard<std::mutex> lock(mu);
    if (sl.buf) {
      CTN_CUDA(cudaDeviceSynchronize());
      CTN_CUDA(cudaFree(sl.buf));
      sl.buf = nullptr;
      sl.cap = 0;
    }
    CTN_CUDA(cudaMalloc(&sl.buf, bytes));
    sl.cap = bytes;
  }
  *out = sl.buf;
  return 0;
}

static bool env_flag(const char* name) {
  const char* e = getenv(name);
  return e != nullptr && e[0] == '1';
}
int launch_gemm(const GemmArgs& a0, cudaStream_t s) {
  // frame-major kernel (gemm_ts.cu): any O % 16 == 0; channel-major kernel (gemm_tc.cu): O % 128 == 0.
  // CTN_GEMM_SS=1 (debug) never uses the frame-major kernel.
  static const bool use_ss = env_flag("CTN_GEMM_SS");
  if (a0.half) {  // reduced-precision inference exists in the frame-major kernel only (bf16-stored operands / outputs)
    CTN_REQUIRE(!force_simt() && ts_gemm_eligible(a0), "reduced-precision 1x1 conv: needs pre-split bf16 weight planes, "
                "Kd %% 64 == 0 and O %% 16 == 0 (got Kd=%d O=%d)", a0.Kd, a0.O);
    const int rc = launch_gemm_ts(a0, s);
    CTN_REQUIRE(rc >= 0, "reduced-precision 1x1 conv: this prologue / epilogue combination is not built");
    return rc;
  }
  const bool off_grid = use_ss ? (a0.Kd % 64 != 0 || a0.O % 128 != 0 || a0.F < 16) : (a0.Kd % 64 != 0 || a0.O % 16 != 0);
  if (force_simt() || off_grid) return launch_gemm_simt(a0, s);
  static const bool simt_fwd = env_flag("CTN_SIMT_FWD"), simt_bwd = env_flag("CTN_SIMT_BWD");  // A/B debugging only
  if ((simt_fwd && !a0.w_is_kn) || (simt_bwd && a0.w_is_kn)) return launch_gemm_simt(a0, s);
  GemmArgs a = a0;
  if (a.W_hi == nullptr) {  // standalone call: split (and transpose if needed) the fp32 weight first
    void* scr = nullptr;
    a.tf32 = a.w_is_kn ? 0 : 1;  // [O,Kd] weights = a forward conv (tf32 split); [Kd,O] = a data gradient (bf16 split)
    const size_t plane = (size_t)a.O * a.Kd * (a.tf32 ? 4 : 2);
    CTN_TRY(lib_scratch(2 * plane, &scr, 0));
    a.W_hi = scr;
    a.W_lo = reinterpret_cast<char*>(scr) + plane;
    if (a.w_is_kn)  // W is [Kd, O]: transpose to [O, Kd]
      CTN_TRY(run_split_planes(a.W, a.Kd, a.O, 1, 0, const_cast<void*>(a.W_hi), const_cast<void*>(a.W_lo), 0, 1, s));
    else
      CTN_TRY(run_split_planes_tf32(a.W, (int64_t)a.O * a.Kd, 1, 0, const_cast<void*>(a.W_hi), const_cast<void*>(a.W_lo),
                                    0, s));
  }
  // Which tcgen05 kernel: the frame-major one (gemm_ts.cu, activations in tensor memory) or the channel-major one
  // (gemm_tc.cu).  Measured on B200 (profiles/r2_summary.md): for the bf16x3 flavours (data gradients, inference) the
  // frame-major kernel is 1.3-1.5x faster once every SM has several frame tiles (F = 51k: 65 / 76 us against 99 / 101 us)
  // and level at the headline's F = 9,597 (27 / 31 against 29.5 / 29.5 us: it wins the K = 256 "up" shape only); the TF32x3
  // training forward needs two accumulators per output, cannot double-buffer them in 512 TMEM columns, and stays on
  // the channel-major kernel.  Shapes off the 128-channel grid (O % 128 != 0) only exist in the frame-major kernel.
  // CTN_TS_MASK (debug) overrides: 1 = tf32 without fold, 2 = tf32 with fold, 4 = bf16.
  static const int ts_mask = getenv("CTN_TS_MASK") ? atoi(getenv("CTN_TS_MASK")) : -1;
  const int flavour = a.tf32 ? (a.c1 != nullptr ? 2 : 1) : 4;
  bool want_ts;
  if (use_ss) want_ts = false;
  else if (ts_mask >= 0) want_ts = (ts_mask & flavour) != 0;
  else want_ts = !a.tf32 && (a.F >= 16384 || a.O > a.Kd);
  if (a.O % 128 != 0 || a.F < 16) want_ts = !use_ss;
  if (want_ts && ts_gemm_eligible(a)) {
    const int rc = launch_gemm_ts(a, s);
    if (rc >= 0) return rc;
  }
  if (!tc_gemm_eligible(a)) return launch_gemm_simt(a0, s);
  return launch_gemm_tc(a, s);
}

// GEMM whose output is the gradient w.r.t. a normalised activation, plus the norm-backward reduction over that output:
// fused into the tcgen05 epilogue when the shape is on the MMA grid, otherwise GEMM followed by norm_bwd_reduce
static bool nred_fusion_enabled() {
  return env_flag("CTN_NRED_FUSION") && !force_simt() && !env_flag("CTN_SIMT_BWD");
}
static bool nred_fused_for(const GemmArgs& a) {
  return nred_fusion_enabled() && a.W_hi != nullptr && a.Kd % 64 == 0 && a.O % 128 == 0 && a.F >= 16;
}
static int launch_gemm_nred(const GemmArgs& a, int M, int defer_fold, cudaStream_t s) {
  // measured on B200 (M=3 x 4 s): the fused epilogue exposes the z2 loads of a GEMM whose epilogue is not overlapped
  // with anything yet, 7.70 ms/step vs 7.52 ms un-fused -> opt-in until the GEMM is persistent (CTN_NRED_FUSION=1).
  // defer_fold = 0 (BatchNorm: the per-channel sums are needed right away) always takes the un-fused route.
  const bool fused = defer_fold && nred_fused_for(a);
  if (fused) return launch_gemm(a, s);
  GemmArgs b = a;
  b.nred_z = nullptr;
  CTN_TRY(launch_gemm(b, s));
  return run_norm_bwd_reduce(a.D, a.nred_z, a.nred_alpha, a.st, a.nred_gamma, M, a.K, a.O, a.nred_dgamma, a.nred_dbeta,
                             a.nred_red, a.nred_part, defer_fold, s);
}
int launch_wgrad(const WgradArgs& a, cudaStream_t s) {
  static const bool simt_bwd = env_flag("CTN_SIMT_BWD") || env_flag("CTN_SIMT_WGRAD");
  if (force_simt() || simt_bwd || !tc_wgrad_eligible(a)) return launch_wgrad_simt(a, s);
  return launch_wgrad_tc(a, s);
}

// ---- flat parameter layout (reference state_dict order) -------------------------
static inline int64_t al4(int64_t n) { return (n + 3) & ~(int64_t)3; }

struct ParamLayout {
  int64_t U, g0, b0, Wb, blk0, blk_stride, Wm, V, total;
  int64_t W1, a1, g1, b1, Wd, a2, g2, b2, W2;  // offsets inside one TemporalBlock
};

static ParamLayout make_layout(const ctn_config& c) {
  ParamLayout L;
  int64_t o = 0;
  L.U = o; o += al4((int64_t)c.N * c.L);
  L.g0 = o; o += al4(c.N);
  L.b0 = o; o += al4(c.N);
  L.Wb = o; o += al4((int64_t)c.B * c.N);
  L.blk0 = o;
  int64_t q = 0;
  L.W1 = q; q += al4((int64_t)c.H * c.B);
  L.a1 = q; q += 4;
  L.g1 = q; q += al4(c.H);
  L.b1 = q; q += al4(c.H);
  L.Wd = q; q += al4((int64_t)c.H * c.P);
  L.a2 = q; q += 4;
  L.g2 = q; q += al4(c.H);
  L.b2 = q; q += al4(c.H);
  L.W2 = q; q += al4((int64_t)c.B * c.H);
  L.blk_stride = q;
  o += q * c.R * c.X;
  L.Wm = o; o += al4((int64_t)c.C * c.N * c.B);
  L.V = o; o += al4((int64_t)c.L * c.N);
  L.total = o;
  return L;
}

static int validate(const ctn_config* c) {
  CTN_REQUIRE(c != nullptr, "null config");
  CTN_REQUIRE(c->N > 0 && c->L >= 2 && c->B > 0 && c->H > 0 && c->P > 0 && c->X > 0 && c->R > 0 && c->C > 0,
              "config: all of N,L,B,H,P,X,R,C must be positive and L >= 2");
  CTN_REQUIRE(c->N % 4 == 0 && c->B % 4 == 0 && c->H % 4 == 0, "config: N, B, H must be multiples of 4 (got %d,%d,%d)",
              c->N, c->B, c->H);
  CTN_REQUIRE(c->C <= 4, "config: C <= 4 supported (got %d)", c->C);
  CTN_REQUIRE(c->P <= 8, "config: P <= 8 supported (got %d)", c->P);
  CTN_REQUIRE(c->X <= 20, "config: X <= 20 (dilation 2^x)");
  CTN_REQUIRE(c->L <= 64 && c->N <= 512, "config: L <= 64 and N <= 512 supported (got L=%d N=%d)", c->L, c->N);
  CTN_REQUIRE(c->norm_type == CTN_NORM_GLN || c->norm_type == CTN_NORM_CLN || c->norm_type == CTN_NORM_BN,
              "config: norm_type must be 0 (gLN), 1 (cLN) or 2 (BN)");
  CTN_REQUIRE(c->causal || (c->P % 2 == 1), "config: non-causal needs odd P");
  CTN_REQUIRE(c->mask_nonlinear == CTN_MASK_RELU || c->mask_nonlinear == CTN_MASK_SOFTMAX,
              "Unsupported mask non-linear function");
  return 0;
}

// ---- workspace plan --------------------------------------------------------------------------
struct Plan {
  int M, T, K, nblk, training;
  int64_t F;
  // byte offsets
  int64_t w, rowstat0, x, z1, z2, gacc, rs1, rs2, score, Wbg, c1b, c2b, W2g, c1, c2;
  int64_t g, dn2, dn1, d_score, d_w, dn0, red, part;
  int64_t bn_acc, bn_fwd, bn_ab, bn_coef, bn_mode;  // BatchNorm slots, structure-of-arrays over (block, norm)
  int64_t part_dw, part_nr;  // floats per block of the dwconv / norm-2 partial rows (each block keeps its own)
  int64_t pl_W1, pl_W2g, pl_Wbg, pl_Wm, pl_W1T, pl_W2T, pl_WbT, pl_WmT;  // bf16 hi planes; lo plane follows at +pl_lo
  int64_t pl_lo;
  int64_t x_stride, z_stride, rs_stride;  // bytes between consecutive blocks' buffers (0 when not stashed)
  int64_t total;
};

static inline int64_t al256(int64_t n) { return (n + 255) & ~(int64_t)255; }

static Plan make_plan(const ctn_config& c, int M, int T, int training) {
  Plan p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.T = T; p.training = training;
  const int S = c.L / 2;
  p.K = (T - c.L) / S + 1;
  p.F = (int64_t)M * p.K;
  p.nblk = c.R * c.X;
  const int64_t F = p.F;
  const bool cln = c.norm_type == CTN_NORM_CLN;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o += al256(bytes); return r; };
  p.w = take(F * c.N * 4);
  p.rowstat0 = take(F * 2 * 4);
  p.x_stride = al256(F * c.B * 4);
  p.x = o; o += p.x_stride * (training ? p.nblk + 1 : 2);
  p.z_stride = al256(F * c.H * 4);
  p.z1 = o; o += p.z_stride * (training ? p.nblk : 1);
  p.z2 = o; o += p.z_stride * (training ? p.nblk : 1);
  p.gacc = take((int64_t)p.nblk * 2 * M * 2 * 8);
  p.rs_stride = cln ? al256(F * 2 * 4) : 0;
  p.rs1 = o; o += p.rs_stride * (training ? p.nblk : 1);
  p.rs2 = o; o += p.rs_stride * (training ? p.nblk : 1);
  if (c.norm_type == CTN_NORM_BN) {  // per (block, norm): fp64 (sum, sumsq)[H]; (mean, rstd, s, t)[H]; (A, B)[H]; (ca, cq)[H]; mode
    const int64_t ns = (int64_t)p.nblk * 2;
    p.bn_acc = take(ns * 2 * c.H * 8);
    p.bn_fwd = take(ns * 4 * c.H * 4);
    p.bn_ab = take(ns * 2 * c.H * 4);
    p.bn_coef = take(ns * 2 * c.H * 4);
    p.bn_mode = take(ns * 4);
  }
  p.score = take(F * c.C * c.N * 4);
  p.Wbg = take((int64_t)c.B * c.N * 4);
  p.c1b = take(c.B * 4);
  p.c2b = take(c.B * 4);
  p.W2g = take((int64_t)p.nblk * c.B * c.H * 4);
  p.c1 = take((int64_t)p.nblk * c.B * 4);
  p.c2 = take((int64_t)p.nblk * c.B * 4);
  {  // bf16 hi/lo weight planes for the tcgen05 GEMMs: one region of hi planes, an identical region of lo planes
    const int64_t start = o;
    p.pl_W1 = take((int64_t)p.nblk * c.H * c.B * 4);   // forward planes are fp32 (tf32 split)
    p.pl_W2g = take((int64_t)p.nblk * c.B * c.H * 4);
    p.pl_Wbg = take((int64_t)c.B * c.N * 4);
    p.pl_Wm = take((int64_t)c.C * c.N * c.B * 4);
    if (training) {
      p.pl_W1T = take((int64_t)p.nblk * c.H * c.B * 2);
      p.pl_W2T = take((int64_t)p.nblk * c.B * c.H * 2);
      p.pl_WbT = take((int64_t)c.B * c.N * 2);
      p.pl_WmT = take((int64_t)c.C * c.N * c.B * 2);
    }
    p.pl_lo = o - start;
    o += p.pl_lo;
  }
  if (training) {
    p.g = take(2 * al256(F * c.B * 4));
    p.dn2 = take(F * c.H * 4);
    p.dn1 = take(F * c.H * 4);
    p.d_score = take(F * c.C * c.N * 4);
    p.d_w = take(F * c.N * 4);
    p.dn0 = take(F * c.N * 4);
    p.red = take((int64_t)(p.nblk * 2 + 1) * M * 2 * 8);
    {  // per-block partial rows (folded once per backward stage) + one shared region for the front cLN
      int64_t pf = 0;
      for (int x = 0; x < c.X; ++x) {
        const int64_t q = dwconv_bwd_partial_floats(M, p.K, c.H, c.P, 1 << x);
        pf = pf > q ? pf : q;
      }
      p.part_dw = (pf + 63) & ~(int64_t)63;
      p.part_nr = (norm_bwd_partial_floats(M, p.K, c.H) + 63) & ~(int64_t)63;
      const int64_t front = norm_bwd_partial_floats(M, p.K, c.N);
      p.part = take(((int64_t)p.nblk * (p.part_dw + p.part_nr) + front) * 4);
    }
  }
  p.total = o;
  return p;
}

struct Ctx {
  const ctn_config& c;
  ParamLayout L;
  Plan p;
  const float* params;
  char* ws;
  cudaStream_t s;
  int bf16_act = 0;           // reduced-precision inference: z1 / z2 stored as bf16, one bf16 plane per GEMM operand
  float* bn_state = nullptr;  // BN running statistics [nblk][rm1 H | rv1 H | rm2 H | rv2 H] (forward only)
  int bn_batch = 1;           // BN: 1 = batch statistics (+ running update), 0 = running statistics
  template <typename T>
  T* at(int64_t off) const { return reinterpret_cast<T*>(ws + off); }
  const float* blk(int b, int64_t off) const { return params + L.blk0 + (int64_t)b * L.blk_stride + off; }
  // block input b (b == nblk is the separator output)
  float* x(int b) const { return at<float>(p.x + p.x_stride * (p.training ? b : (b & 1))); }
  float* z1(int b) const { return at<float>(p.z1 + p.z_stride * (p.training ? b : 0)); }
  float* z2(int b) const { return at<float>(p.z2 + p.z_stride * (p.training ? b : 0)); }
  // BatchNorm slot views
  int64_t bn_slot(int b, int which) const { return (int64_t)b * 2 + which; }
  double* bn_acc(int b, int which) const { return at<double>(p.bn_acc) + bn_slot(b, which) * 2 * c.H; }
  float* bn_fwd(int b, int which, int field) const {  // field: 0 mean, 1 rstd, 2 s, 3 t
    return at<float>(p.bn_fwd) + (bn_slot(b, which) * 4 + field) * c.H;
  }
  float* bn_ab(int b, int which, int field) const { return at<float>(p.bn_ab) + (bn_slot(b, which) * 2 + field) * c.H; }
  float* bn_coef(int b, int which, int field) const { return at<float>(p.bn_coef) + (bn_slot(b, which) * 2 + field) * c.H; }
  float* bn_mode(int b, int which) const { return at<float>(p.bn_mode) + bn_slot(b, which); }
  // (gamma, beta) as the consumers see them: the parameters, or BatchNorm's per-channel (s, t)
  const float* gam(int b, int which) const {
    return c.norm_type == CTN_NORM_BN ? bn_fwd(b, which, 2) : blk(b, which ? L.g2 : L.g1);
  }
  const float* bet(int b, int which) const {
    return c.norm_type == CTN_NORM_BN ? bn_fwd(b, which, 3) : blk(b, which ? L.b2 : L.b1);
  }
  NormStats stats(int b, int which) const {
    NormStats st;
    if (c.norm_type == CTN_NORM_BN) {
      st.acc = nullptr; st.row = nullptr; st.inv_count = 0.0;  // identity: the normalisation lives in (s, t)
    } else if (c.norm_type == CTN_NORM_GLN) {
      st.acc = at<double>(p.gacc) + ((int64_t)b * 2 + which) * p.M * 2;
      st.row = nullptr;
      st.inv_count = 1.0 / ((double)p.K * (double)c.H);
    } else {
      st.acc = nullptr;
      st.row = at<float>((which ? p.rs2 : p.rs1) + p.rs_stride * (p.training ? b : 0));
      st.inv_count = 0.0;
    }
    return st;
  }
  double* stat_out(int b, int which) const {
    return c.norm_type == CTN_NORM_GLN ? at<double>(p.gacc) + ((int64_t)b * 2 + which) * p.M * 2 : nullptr;
  }
  double* red(int b, int which) const { return at<double>(p.red) + ((int64_t)b * 2 + which) * p.M * 2; }
};

static int check_io(const ctn_config* cfg, int M, int T, const void* ws, int64_t ws_bytes, int training) {
  CTN_TRY(validate(cfg));
  CTN_REQUIRE(training >= 0 && training <= 2, "mode must be 0 (inference), 1 (training) or 2 (bf16 inference)");
  if (training == 2)
    CTN_REQUIRE(cfg->N % 64 == 0 && cfg->B % 64 == 0 && cfg->H % 64 == 0 && cfg->norm_type != CTN_NORM_BN,
                "bf16 inference needs N, B, H multiples of 64 and gLN / cLN (got N=%d B=%d H=%d norm=%d)", cfg->N, cfg->B,
                cfg->H, cfg->norm_type);
  CTN_REQUIRE(M >= 1 && M <= 65535, "batch size M must be in [1, 65535] (got %d)", M);
  CTN_REQUIRE(T >= cfg->L, "input has %d samples, fewer than one frame (L=%d)", T, cfg->L);
  const int64_t need = ctn_workspace_bytes(cfg, M, T, training);
  CTN_REQUIRE(ws != nullptr && ws_bytes >= need, "workspace too small: need %lld bytes, got %lld", (long long)need,
              (long long)ws_bytes);
  CTN_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
  return 0;
}

// BatchNorm of block b, norm `which` (0: after the first PReLU, 1: after the depthwise conv's PReLU): batch statistics of
// prelu(z) (training) or the running statistics (evaluation) -> per-channel (mean, rstd, s, t) in the block's slot
static int bn_forward(const Ctx& X, int b, int which, const float* z) {
  const ctn_config& c = X.c;
  const ParamLayout& L = X.L;
  CTN_REQUIRE(X.bn_state != nullptr, "norm_type BN needs the running-statistics buffer: call ctn_model_forward_bn");
  float* rm = X.bn_state + ((int64_t)b * 4 + which * 2) * c.H;
  return run_bn_forward_stats(z, X.blk(b, which ? L.a2 : L.a1), X.blk(b, which ? L.g2 : L.g1),
                              X.blk(b, which ? L.b2 : L.b1), rm, rm + c.H, X.p.F, c.H, X.bn_batch, X.bn_acc(b, which),
                              X.bn_fwd(b, which, 0), X.bn_fwd(b, which, 1), X.bn_fwd(b, which, 2),
                              X.bn_fwd(b, which, 3), X.bn_mode(b, which), X.s);
}

static int model_forward(const Ctx& X, const float* mixture, float* est) {
  const ctn_config& c = X.c;
  const Plan& p = X.p;
  const ParamLayout& L = X.L;
  const int M = p.M, K = p.K, nblk = p.nblk;
  const int64_t F = p.F;
  cudaStream_t s = X.s;
  const bool gln = c.norm_type == CTN_NORM_GLN, cln = c.norm_type == CTN_NORM_CLN, bn = c.norm_type == CTN_NORM_BN;

  if (gln) CTN_CUDA(cudaMemsetAsync(X.at<char>(p.gacc), 0, (size_t)nblk * 2 * M * 2 * 8, s));
  if (bn) CTN_CUDA(cudaMemsetAsync(X.at<char>(p.bn_acc), 0, (size_t)nblk * 2 * 2 * c.H * 8, s));
  // norm-fold constants for the bottleneck and every block's pointwise conv (one launch each).  BatchNorm: the
  // per-channel scale of a block is only known once its statistics are, so its fold happens inside the block loop.
  CTN_TRY(run_prep_normfold(X.params + L.Wb, X.params + L.g0, X.params + L.b0, c.B, c.N, 1, 0, X.at<float>(p.Wbg),
                            X.at<float>(p.c1b), X.at<float>(p.c2b), 0, 0, s));
  if (!bn)
    CTN_TRY(run_prep_normfold(X.blk(0, L.W2), X.blk(0, L.g2), X.blk(0, L.b2), c.B, c.H, nblk, L.blk_stride,
                              X.at<float>(p.W2g), X.at<float>(p.c1), X.at<float>(p.c2), (int64_t)c.B * c.H, c.B, s));
  // split planes of every forward GEMM weight (batched over the blocks).  Training: TF32 split (fp32-class forward,
  // needed for gradient parity).  Inference: bf16 split — 1e-5-class outputs (budget 1e-4), half the
  // tensor-core and shared-memory cost.
  const int fwd_tf32 = p.training ? 1 : 0;
  const int64_t esz = fwd_tf32 ? 4 : 2;
  if (fwd_tf32) {
    CTN_TRY(run_split_planes_tf32(X.blk(0, L.W1), (int64_t)c.H * c.B, nblk, L.blk_stride, X.at<char>(p.pl_W1),
                                  X.at<char>(p.pl_W1 + p.pl_lo), (int64_t)c.H * c.B, s));
    if (!bn)
      CTN_TRY(run_split_planes_tf32(X.at<float>(p.W2g), (int64_t)c.B * c.H, nblk, (int64_t)c.B * c.H,
                                    X.at<char>(p.pl_W2g), X.at<char>(p.pl_W2g + p.pl_lo), (int64_t)c.B * c.H, s));
    CTN_TRY(run_split_planes_tf32(X.at<float>(p.Wbg), (int64_t)c.B * c.N, 1, 0, X.at<char>(p.pl_Wbg),
                                  X.at<char>(p.pl_Wbg + p.pl_lo), 0, s));
    CTN_TRY(run_split_planes_tf32(X.params + L.Wm, (int64_t)c.C * c.N * c.B, 1, 0, X.at<char>(p.pl_Wm),
                                  X.at<char>(p.pl_Wm + p.pl_lo), 0, s));
  } else {
    CTN_TRY(run_split_planes(X.blk(0, L.W1), c.H, c.B, nblk, L.blk_stride, X.at<char>(p.pl_W1),
                             X.at<char>(p.pl_W1 + p.pl_lo), (int64_t)c.H * c.B, 0, s));
    if (!bn)
      CTN_TRY(run_split_planes(X.at<float>(p.W2g), c.B, c.H, nblk, (int64_t)c.B * c.H, X.at<char>(p.pl_W2g),
                               X.at<char>(p.pl_W2g + p.pl_lo), (int64_t)c.B * c.H, 0, s));
    CTN_TRY(run_split_planes(X.at<float>(p.Wbg), c.B, c.N, 1, 0, X.at<char>(p.pl_Wbg), X.at<char>(p.pl_Wbg + p.pl_lo),
                             0, 0, s));
    CTN_TRY(run_split_planes(X.params + L.Wm, c.C * c.N, c.B, 1, 0, X.at<char>(p.pl_Wm), X.at<char>(p.pl_Wm + p.pl_lo),
                             0, 0, s));
  }
  // encoder + first cLN statistics + bottleneck (cLN folded into the GEMM epilogue)
  float* w = X.at<float>(p.w);
  CTN_TRY(run_encoder_fwd(mixture, X.params + L.U, M, p.T, c.N, c.L, w, s));
  CTN_TRY(run_row_stats(w, nullptr, F, c.N, X.at<float>(p.rowstat0), s));
  const int half = X.bf16_act ? 1 : 0;  // reduced-precision inference: one bf16 plane per operand (gemm_ts.cu)
  {
    GemmArgs a = {};
    a.A = w; a.W = X.at<float>(p.Wbg); a.D = X.x(0); a.F = F; a.O = c.B; a.Kd = c.N; a.K = K;
    a.c1 = X.at<float>(p.c1b); a.c2 = X.at<float>(p.c2b);
    a.st.row = X.at<float>(p.rowstat0);
    a.W_hi = X.at<char>(p.pl_Wbg); a.W_lo = X.at<char>(p.pl_Wbg + p.pl_lo); a.tf32 = fwd_tf32;
    a.half = half;
    CTN_TRY(launch_gemm(a, s));
  }
  for (int b = 0; b < nblk; ++b) {
    const int dil = 1 << (b % c.X);
    {  // z1 = x W1^T, gLN stats of prelu(z1)
      GemmArgs a = {};
      a.A = X.x(b); a.W = X.blk(b, L.W1); a.D = X.z1(b); a.F = F; a.O = c.H; a.Kd = c.B; a.K = K;
      a.stat_out = X.stat_out(b, 0); a.alpha_out = X.blk(b, L.a1);
      a.W_hi = X.at<char>(p.pl_W1) + (int64_t)b * c.H * c.B * esz;
      a.W_lo = X.at<char>(p.pl_W1 + p.pl_lo) + (int64_t)b * c.H * c.B * esz;
      a.tf32 = fwd_tf32;
      a.half = half; a.d_bf16 = half;  // z1 stored as bf16
      CTN_TRY(launch_gemm(a, s));
    }
    // cLN: the per-frame statistics of prelu(z1) and prelu(z2) come out of the depthwise kernel itself when it stages its
    // rows in shared memory (elementwise.cu); otherwise one row_stats pass over [F, H] each
    const bool fuse_rs = cln && dwconv_fwd_fuses_rowstats(c.H, c.P, half);
    if (cln && !fuse_rs)
      CTN_TRY(run_row_stats(X.z1(b), X.blk(b, L.a1), F, c.H, const_cast<float*>(X.stats(b, 0).row), s, half));
    if (bn) CTN_TRY(bn_forward(X, b, 0, X.z1(b)));
    CTN_TRY(run_dwconv_fwd(X.z1(b), X.blk(b, L.a1), X.stats(b, 0), X.gam(b, 0), X.bet(b, 0), X.blk(b, L.Wd), M, K,
                           c.H, c.P, dil, c.causal, X.z2(b), X.stat_out(b, 1), X.blk(b, L.a2), s, half,
                           fuse_rs ? const_cast<float*>(X.stats(b, 0).row) : nullptr,
                           fuse_rs ? const_cast<float*>(X.stats(b, 1).row) : nullptr));
    if (cln && !fuse_rs)
      CTN_TRY(run_row_stats(X.z2(b), X.blk(b, L.a2), F, c.H, const_cast<float*>(X.stats(b, 1).row), s, half));
    if (bn) {  // statistics of prelu(z2) -> (s, t) -> this block's folded pointwise weight and its operand planes
      CTN_TRY(bn_forward(X, b, 1, X.z2(b)));
      float* W2g = X.at<float>(p.W2g) + (int64_t)b * c.B * c.H;
      CTN_TRY(run_prep_normfold(X.blk(b, L.W2), X.gam(b, 1), X.bet(b, 1), c.B, c.H, 1, 0, W2g,
                                X.at<float>(p.c1) + (int64_t)b * c.B, X.at<float>(p.c2) + (int64_t)b * c.B, 0, 0, s));
      char* hi = X.at<char>(p.pl_W2g) + (int64_t)b * c.B * c.H * esz;
      if (fwd_tf32)
        CTN_TRY(run_split_planes_tf32(W2g, (int64_t)c.B * c.H, 1, 0, hi, hi + p.pl_lo, 0, s));
      else
        CTN_TRY(run_split_planes(W2g, c.B, c.H, 1, 0, hi, hi + p.pl_lo, 0, 0, s));
    }
    {  // out = x + norm2(prelu(z2)) W2^T with norm2 folded
      GemmArgs a = {};
      a.A = X.z2(b); a.W = X.at<float>(p.W2g) + (int64_t)b * c.B * c.H; a.D = X.x(b + 1);
      a.F = F; a.O = c.B; a.Kd = c.H; a.K = K;
      a.alpha_in = X.blk(b, L.a2);
      a.c1 = X.at<float>(p.c1) + (int64_t)b * c.B; a.c2 = X.at<float>(p.c2) + (int64_t)b * c.B;
      a.st = X.stats(b, 1);
      a.res = X.x(b);
      a.W_hi = X.at<char>(p.pl_W2g) + (int64_t)b * c.B * c.H * esz;
      a.W_lo = X.at<char>(p.pl_W2g + p.pl_lo) + (int64_t)b * c.B * c.H * esz;
      a.tf32 = fwd_tf32;
      a.half = half ? 2 : 0;  // z2 is stored as bf16
      CTN_TRY(launch_gemm(a, s));
    }
  }
  {  // mask conv
    GemmArgs a = {};
    a.A = X.x(nblk); a.W = X.params + L.Wm; a.D = X.at<float>(p.score); a.F = F; a.O = c.C * c.N; a.Kd = c.B; a.K = K;
    a.W_hi = X.at<char>(p.pl_Wm); a.W_lo = X.at<char>(p.pl_Wm + p.pl_lo); a.tf32 = fwd_tf32;
    a.half = half;
    CTN_TRY(launch_gemm(a, s));
  }
  return run_decoder_fwd(X.at<float>(p.score), w, X.params + L.V, M, K, c.C, c.N, c.L, p.T,
                         c.mask_nonlinear == CTN_MASK_SOFTMAX, est, s);
}

// stages: 0 = zero-init + decoder + mask conv; 1..R = repeat R-stage (its X blocks, last first); R+1 = bottleneck,
// first cLN, encoder.  Each stage completes one contiguous slice of the flat gradient (ctn_grad_bucket).
static int model_backward(const Ctx& X, const float* mixture, const float* d_est, float* grads, int accumulate,
                          int stage_lo, int stage_hi) {
  const ctn_config& c = X.c;
  const Plan& p = X.p;
  const ParamLayout& L = X.L;
  const int M = p.M, K = p.K, nblk = p.nblk;
  const int64_t F = p.F;
  cudaStream_t s = X.s;
  const bool gln = c.norm_type == CTN_NORM_GLN, bn = c.norm_type == CTN_NORM_BN;
  auto gblk = [&](int b, int64_t off) { return grads + L.blk0 + (int64_t)b * L.blk_stride + off; };
  // backward through one norm + the PReLU in front of it, in place over dn (gradient w.r.t. the normalised activation)
  auto norm_apply = [&](int b, int which, float* dn, const float* z) -> int {
    const float* alpha = X.blk(b, which ? L.a2 : L.a1);
    float* dalpha = gblk(b, which ? L.a2 : L.a1);
    if (!bn)
      return run_norm_bwd_apply(dn, z, alpha, X.stats(b, which), X.blk(b, which ? L.g2 : L.g1), X.red(b, which), M, K, c.H,
                                dalpha, s);
    // BatchNorm: the folded per-channel sums -> dweight, dbias and the coefficients of the apply pass
    CTN_TRY(run_bn_bwd_finalize(X.bn_ab(b, which, 0), X.bn_ab(b, which, 1), X.bn_fwd(b, which, 0), X.bn_fwd(b, which, 1),
                                X.bn_mode(b, which), -1, F, c.H, gblk(b, which ? L.g2 : L.g1), gblk(b, which ? L.b2 : L.b1),
                                X.bn_coef(b, which, 0), X.bn_coef(b, which, 1), s));
    return run_bn_bwd_apply(dn, z, alpha, X.bn_fwd(b, which, 2), X.bn_fwd(b, which, 0), X.bn_coef(b, which, 0),
                            X.bn_coef(b, which, 1), F, c.H, dalpha, s);
  };

  float* w = X.at<float>(p.w);
  float* g_buf[2] = {X.at<float>(p.g), X.at<float>(p.g + al256(F * c.B * 4))};
  float* dn2 = X.at<float>(p.dn2);
  float* dn1 = X.at<float>(p.dn1);
  float* d_score = X.at<float>(p.d_score);
  float* d_w = X.at<float>(p.d_w);
  float* dn0 = X.at<float>(p.dn0);

  if (stage_lo <= 0 && 0 < stage_hi) {
  float* g_cur = g_buf[0];
  if (!accumulate) CTN_CUDA(cudaMemsetAsync(grads, 0, (size_t)L.total * 4, s));
  CTN_CUDA(cudaMemsetAsync(X.at<char>(p.red), 0, (size_t)(nblk * 2 + 1) * M * 2 * 8, s));
  if (bn) CTN_CUDA(cudaMemsetAsync(X.at<char>(p.bn_ab), 0, (size_t)nblk * 2 * 2 * c.H * 4, s));
  // transposed bf16 hi/lo weight planes for the data-gradient GEMMs
  CTN_TRY(run_split_planes(X.blk(0, L.W1), c.H, c.B, nblk, L.blk_stride, X.at<char>(p.pl_W1T),
                           X.at<char>(p.pl_W1T + p.pl_lo), (int64_t)c.H * c.B, 1, s));
  CTN_TRY(run_split_planes(X.blk(0, L.W2), c.B, c.H, nblk, L.blk_stride, X.at<char>(p.pl_W2T),
                           X.at<char>(p.pl_W2T + p.pl_lo), (int64_t)c.B * c.H, 1, s));
  CTN_TRY(run_split_planes(X.params + L.Wb, c.B, c.N, 1, 0, X.at<char>(p.pl_WbT), X.at<char>(p.pl_WbT + p.pl_lo), 0, 1,
                           s));
  CTN_TRY(run_split_planes(X.params + L.Wm, c.C * c.N, c.B, 1, 0, X.at<char>(p.pl_WmT), X.at<char>(p.pl_WmT + p.pl_lo),
                           0, 1, s));
  CTN_TRY(run_decoder_bwd(d_est, X.at<float>(p.score), w, X.params + L.V, M, K, c.C, c.N, c.L, p.T,
                          c.mask_nonlinear == CTN_MASK_SOFTMAX, d_score, d_w, grads + L.V, s));
  {  // mask conv: dWm, g = d_score Wm
    WgradArgs wa = {};
    wa.G = d_score; wa.Act = X.x(nblk); wa.dW = grads + L.Wm; wa.F = F; wa.O = c.C * c.N; wa.I = c.B; wa.K = K;
    CTN_TRY(launch_wgrad(wa, s));
    GemmArgs a = {};
    a.A = d_score; a.W = X.params + L.Wm; a.w_is_kn = 1; a.D = g_cur; a.F = F; a.O = c.B; a.Kd = c.C * c.N; a.K = K;
    a.W_hi = X.at<char>(p.pl_WmT); a.W_lo = X.at<char>(p.pl_WmT + p.pl_lo);
    CTN_TRY(launch_gemm(a, s));
  }
  }
  for (int b = nblk - 1; b >= 0; --b) {
    const int stage = c.R - b / c.X;
    if (stage < stage_lo || stage >= stage_hi) continue;
    const int swaps = nblk - 1 - b;  // ping-pong state is a function of the block index => stages are restartable
    float* g_cur = g_buf[swaps & 1];
    float* g_nxt = g_buf[(swaps + 1) & 1];
    const int dil = 1 << (b % c.X);
    const NormStats st1 = X.stats(b, 0), st2 = X.stats(b, 1);
    {  // dn2 = g W2
      GemmArgs a = {};
      a.A = g_cur; a.W = X.blk(b, L.W2); a.w_is_kn = 1; a.D = dn2; a.F = F; a.O = c.H; a.Kd = c.B; a.K = K;
      a.W_hi = X.at<char>(p.pl_W2T) + (int64_t)b * c.B * c.H * 2;
      a.W_lo = X.at<char>(p.pl_W2T + p.pl_lo) + (int64_t)b * c.B * c.H * 2;
      // + norm2 backward reduction over dn2 (dgamma2, dbeta2, per-sample sums) in the same kernel
      // (BatchNorm: the sums land in the block's (A, B) slot and are folded at once)
      a.st = st2; a.nred_z = X.z2(b); a.nred_alpha = X.blk(b, L.a2); a.nred_gamma = X.gam(b, 1);
      a.nred_dgamma = bn ? X.bn_ab(b, 1, 1) : gblk(b, L.g2);
      a.nred_dbeta = bn ? X.bn_ab(b, 1, 0) : gblk(b, L.b2);
      a.nred_red = gln ? X.red(b, 1) : nullptr;
      a.nred_part = X.at<float>(p.part) + (int64_t)b * (p.part_dw + p.part_nr) + p.part_dw;
      CTN_TRY(launch_gemm_nred(a, M, bn ? 0 : 1, s));
    }
    {  // dW2 = g^T norm2(prelu(z2))
      WgradArgs wa = {};
      wa.G = g_cur; wa.Act = X.z2(b); wa.dW = gblk(b, L.W2); wa.F = F; wa.O = c.B; wa.I = c.H; wa.K = K;
      wa.alpha = X.blk(b, L.a2); wa.gamma = X.gam(b, 1); wa.beta = X.bet(b, 1); wa.st = st2;
      CTN_TRY(launch_wgrad(wa, s));
    }
    // gLN: the backward of norm2 (+ PReLU) is elementwise given the per-sample sums, so the depthwise backward can apply
    // it as the dn2 / z2 rows enter its window (no gln_bwd_apply pass over [F, H]).  Built, parity-tested and measured in
    // three forms (register window, two-phase shared-memory tile, every tile staged by bulk copies): the last one is the
    // fastest (30.3 us in place against 19.8 + 12.6 us for the two kernels) but at 104 KB of tiles only two blocks fit an
    // SM and the graph-replayed step is 5.91 against 5.80 ms, so it stays opt-in (CTN_APPLY_FUSION=1).  cLN needs
    // per-frame means and BatchNorm per-channel coefficients: separate pass.
    static const bool apply_fusion = env_flag("CTN_APPLY_FUSION");
    if (gln && apply_fusion) {
      CTN_TRY(run_dwconv_bwd_gln_fused(dn2, X.z2(b), X.blk(b, L.a2), st2, X.blk(b, L.g2), X.red(b, 1), gblk(b, L.a2),
                                       X.z1(b), X.blk(b, L.a1), st1, X.gam(b, 0), X.bet(b, 0), X.blk(b, L.Wd), M, K, c.H,
                                       c.P, dil, c.causal, dn1, gblk(b, L.Wd), gblk(b, L.g1), gblk(b, L.b1), X.red(b, 0),
                                       X.at<float>(p.part) + (int64_t)b * (p.part_dw + p.part_nr), 1, s));
    } else {
      CTN_TRY(norm_apply(b, 1, dn2, X.z2(b)));
      CTN_TRY(run_dwconv_bwd(dn2, X.z1(b), X.blk(b, L.a1), st1, X.gam(b, 0), X.bet(b, 0), X.blk(b, L.Wd), M, K, c.H,
                             c.P, dil, c.causal, dn1, gblk(b, L.Wd), bn ? X.bn_ab(b, 0, 1) : gblk(b, L.g1),
                             bn ? X.bn_ab(b, 0, 0) : gblk(b, L.b1), gln ? X.red(b, 0) : nullptr,
                             X.at<float>(p.part) + (int64_t)b * (p.part_dw + p.part_nr), bn ? 0 : 1, s));
    }
    CTN_TRY(norm_apply(b, 0, dn1, X.z1(b)));
    {  // dW1 = dz1^T x
      WgradArgs wa = {};
      wa.G = dn1; wa.Act = X.x(b); wa.dW = gblk(b, L.W1); wa.F = F; wa.O = c.H; wa.I = c.B; wa.K = K;
      CTN_TRY(launch_wgrad(wa, s));
    }
    {  // g_prev = g + dz1 W1
      GemmArgs a = {};
      a.A = dn1; a.W = X.blk(b, L.W1); a.w_is_kn = 1; a.D = g_nxt; a.F = F; a.O = c.B; a.Kd = c.H; a.K = K;
      a.W_hi = X.at<char>(p.pl_W1T) + (int64_t)b * c.H * c.B * 2;
      a.W_lo = X.at<char>(p.pl_W1T + p.pl_lo) + (int64_t)b * c.H * c.B * 2;
      a.res = g_cur;
      CTN_TRY(launch_gemm(a, s));
    }
    if (b % c.X == 0 && !bn) {  // first block of the repeat = last of its stage: fold the stage's partial rows in one launch
      FoldBatch fb;
      int n = 0;
      for (int bb = b; bb < b + c.X; ++bb) {
        const float* pb = X.at<float>(p.part) + (int64_t)bb * (p.part_dw + p.part_nr);
        if (n + 2 > FOLD_MAX) {
          CTN_TRY(run_fold_batch(fb, n, s));
          n = 0;
        }
        fb.e[n++] = FoldEntry{pb, dwconv_bwd_blocks(M, K, 1 << (bb % c.X)), c.H, c.P, gblk(bb, L.Wd), gblk(bb, L.g1),
                              gblk(bb, L.b1)};
        const bool nr_fused = nred_fusion_enabled() && c.B % 64 == 0 && c.H % 128 == 0 && F >= 16;  // see launch_gemm_nred
        if (!nr_fused)
          fb.e[n++] = FoldEntry{pb + p.part_dw, norm_bwd_blocks(M, K), c.H, 0, nullptr, gblk(bb, L.g2), gblk(bb, L.b2)};
      }
      CTN_TRY(run_fold_batch(fb, n, s));
    }
  }
  if (!(stage_lo <= c.R + 1 && c.R + 1 < stage_hi)) return 0;
  // bottleneck + first cLN + encoder
  float* g_cur = g_buf[nblk & 1];
  NormStats st0 = {};
  st0.row = X.at<float>(p.rowstat0);
  {
    WgradArgs wa = {};
    wa.G = g_cur; wa.Act = w; wa.dW = grads + L.Wb; wa.F = F; wa.O = c.B; wa.I = c.N; wa.K = K;
    wa.gamma = X.params + L.g0; wa.beta = X.params + L.b0; wa.st = st0;
    CTN_TRY(launch_wgrad(wa, s));
    GemmArgs a = {};
    a.A = g_cur; a.W = X.params + L.Wb; a.w_is_kn = 1; a.D = dn0; a.F = F; a.O = c.N; a.Kd = c.B; a.K = K;
    a.W_hi = X.at<char>(p.pl_WbT); a.W_lo = X.at<char>(p.pl_WbT + p.pl_lo);
    CTN_TRY(launch_gemm(a, s));
  }
  CTN_TRY(run_norm_bwd_reduce(dn0, w, nullptr, st0, X.params + L.g0, M, K, c.N, grads + L.g0, grads + L.b0, nullptr,
                              X.at<float>(p.part) + (int64_t)nblk * (p.part_dw + p.part_nr), 0, s));
  CTN_TRY(run_norm_bwd_apply(dn0, w, nullptr, st0, X.params + L.g0, nullptr, M, K, c.N, nullptr, s));
  return run_encoder_bwd(mixture, w, dn0, d_w, M, p.T, c.N, c.L, grads + L.U, s);
}

}  // namespace ctn

// =============================================================================================
// C ABI
// =============================================================================================
using namespace ctn;

extern "C" {

int32_t ctn_version(void) { return 100; }

// debug: per-kernel time between consecutive launch events recorded since the last report (CTN_TIMING=1); a launch's
// interval starts at the previous launch's event, so memsets / gaps are charged to the kernel that follows them
int32_t ctn_timing_report(int32_t reset_only) {
  if (!timing_enabled()) return 0;
  cudaDeviceSynchronize();
  if (!reset_only && g_ntrec > 1) {
    struct Agg { const char* name; double ms; int n; };
    Agg agg[64];
    int na = 0;
    double total = 0.0;
    for (int i = 1; i < g_ntrec; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, g_trec[i - 1].ev, g_trec[i].ev) != cudaSuccess) continue;
      int j = 0;
      for (; j < na; ++j)
        if (strcmp(agg[j].name, g_trec[i].name) == 0) break;
      if (j == na && na < 64) agg[na++] = Agg{g_trec[i].name, 0.0, 0};
      if (j < 64) { agg[j].ms += ms; agg[j].n += 1; }
      total += ms;
    }
    printf("ctn timing: %d launches, %.3f ms\n", g_ntrec - 1, total);
    for (int j = 0; j < na; ++j)
      printf("  %-32s n=%5d  total %8.3f ms  avg %7.2f us  %5.1f %%\n", agg[j].name, agg[j].n, agg[j].ms,
             1000.0 * agg[j].ms / agg[j].n, 100.0 * agg[j].ms / total);
    fflush(stdout);
  }
  for (int i = 0; i < g_ntrec; ++i) cudaEventDestroy(g_trec[i].ev);
  g_ntrec = 0;
  return 0;
}
int64_t ctn_launch_count(void) { return (int64_t)__atomic_load_n(&ctn::g_launches, __ATOMIC_RELAXED); }
const char* ctn_last_error(void) { return ctn::last_error(); }

int32_t ctn_param_tensors(const ctn_config* cfg) { return cfg ? 4 + 9 * cfg->R * cfg->X + 2 : 0; }

int64_t ctn_param_floats(const ctn_config* cfg) {
  if (validate(cfg)) return -1;
  return make_layout(*cfg).total;
}

int32_t ctn_param_layout(const ctn_config* cfg, int64_t* offsets, int64_t* numels, int32_t n) {
  CTN_TRY(validate(cfg));
  CTN_REQUIRE(n == ctn_param_tensors(cfg), "param_layout: expected %d entries, got %d", ctn_param_tensors(cfg), n);
  const ctn_config& c = *cfg;
  const ParamLayout L = make_layout(c);
  int i = 0;
  auto put = [&](int64_t off, int64_t ne) { offsets[i] = off; numels[i] = ne; ++i; };
  put(L.U, (int64_t)c.N * c.L);
  put(L.g0, c.N);
  put(L.b0, c.N);
  put(L.Wb, (int64_t)c.B * c.N);
  for (int b = 0; b < c.R * c.X; ++b) {
    const int64_t o = L.blk0 + (int64_t)b * L.blk_stride;
    put(o + L.W1, (int64_t)c.H * c.B);
    put(o + L.a1, 1);
    put(o + L.g1, c.H);
    put(o + L.b1, c.H);
    put(o + L.Wd, (int64_t)c.H * c.P);
    put(o + L.a2, 1);
    put(o + L.g2, c.H);
    put(o + L.b2, c.H);
    put(o + L.W2, (int64_t)c.B * c.H);
  }
  put(L.Wm, (int64_t)c.C * c.N * c.B);
  put(L.V, (int64_t)c.L * c.N);
  return 0;
}

int32_t ctn_num_frames(const ctn_config* cfg, int32_t T) {
  if (!cfg || cfg->L < 2 || T < cfg->L) return 0;
  return (T - cfg->L) / (cfg->L / 2) + 1;
}

int64_t ctn_workspace_bytes(const ctn_config* cfg, int32_t M, int32_t T, int32_t training) {
  if (validate(cfg) || M < 1 || T < cfg->L) return -1;
  return make_plan(*cfg, M, T, training == 1 ? 1 : 0).total;  // (mode 2, bf16 inference, uses the inference plan)
}

int32_t ctn_model_forward(const ctn_config* cfg, const float* params, const float* mixture, int32_t M, int32_t T,
                          float* est, void* workspace, int64_t workspace_bytes, int32_t training, cudaStream_t stream) {
  CTN_TRY(check_io(cfg, M, T, workspace, workspace_bytes, training));
  CTN_REQUIRE(params && mixture && est, "model_forward: null pointer");
  CTN_REQUIRE(cfg->norm_type != CTN_NORM_BN, "model_forward: norm_type BN carries running statistics, call "
              "ctn_model_forward_bn");
  Ctx X = {*cfg, make_layout(*cfg), make_plan(*cfg, M, T, training == 1 ? 1 : 0), params, reinterpret_cast<char*>(workspace),
           stream};
  X.bf16_act = training == 2 ? 1 : 0;
  return model_forward(X, mixture, est);
}

int64_t ctn_norm_state_floats(const ctn_config* cfg) {
  if (validate(cfg)) return -1;
  return cfg->norm_type == CTN_NORM_BN ? (int64_t)cfg->R * cfg->X * 4 * cfg->H : 0;
}

int32_t ctn_model_forward_bn(const ctn_config* cfg, const float* params, float* norm_state, const float* mixture,
                             int32_t M, int32_t T, float* est, void* workspace, int64_t workspace_bytes,
                             int32_t training, int32_t batch_stats, cudaStream_t stream) {
  CTN_TRY(check_io(cfg, M, T, workspace, workspace_bytes, training));
  CTN_REQUIRE(cfg->norm_type == CTN_NORM_BN, "model_forward_bn: the configuration is not norm_type BN");
  CTN_REQUIRE(params && norm_state && mixture && est, "model_forward_bn: null pointer");
  Ctx X = {*cfg, make_layout(*cfg), make_plan(*cfg, M, T, training), params, reinterpret_cast<char*>(workspace), stream};
  X.bn_state = norm_state;
  X.bn_batch = batch_stats ? 1 : 0;
  return model_forward(X, mixture, est);
}

int32_t ctn_model_backward(const ctn_config* cfg, const float* params, const float* mixture, int32_t M, int32_t T,
                           const float* d_est, float* grads, void* workspace, int64_t workspace_bytes,
                           int32_t accumulate, cudaStream_t stream) {
  CTN_TRY(check_io(cfg, M, T, workspace, workspace_bytes, 1));
  CTN_REQUIRE(params && mixture && d_est && grads, "model_backward: null pointer");
  Ctx X = {*cfg, make_layout(*cfg), make_plan(*cfg, M, T, 1), params, reinterpret_cast<char*>(workspace), stream};
  return model_backward(X, mixture, d_est, grads, accumulate, 0, cfg->R + 2);
}

int32_t ctn_model_backward_stage(const ctn_config* cfg, const float* params, const float* mixture, int32_t M, int32_t T,
                                 const float* d_est, float* grads, void* workspace, int64_t workspace_bytes,
                                 int32_t accumulate, int32_t stage, cudaStream_t stream) {
  CTN_TRY(check_io(cfg, M, T, workspace, workspace_bytes, 1));
  CTN_REQUIRE(params && mixture && d_est && grads, "model_backward_stage: null pointer");
  CTN_REQUIRE(stage >= 0 && stage < cfg->R + 2, "model_backward_stage: stage %d out of [0,%d)", stage, cfg->R + 2);
  Ctx X = {*cfg, make_layout(*cfg), make_plan(*cfg, M, T, 1), params, reinterpret_cast<char*>(workspace), stream};
  return model_backward(X, mixture, d_est, grads, accumulate, stage, stage + 1);
}

int32_t ctn_grad_bucket(const ctn_config* cfg, int32_t stage, int64_t* offset, int64_t* count) {
  CTN_TRY(validate(cfg));
  CTN_REQUIRE(stage >= 0 && stage < cfg->R + 2 && offset && count, "grad_bucket: bad arguments");
  const ParamLayout L = make_layout(*cfg);
  if (stage == 0) {
    *offset = L.Wm; *count = L.total - L.Wm;
  } else if (stage <= cfg->R) {
    const int r = cfg->R - stage;
    *offset = L.blk0 + (int64_t)r * cfg->X * L.blk_stride; *count = (int64_t)cfg->X * L.blk_stride;
  } else {
    *offset = 0; *count = L.blk0;
  }
  return 0;
}

}  // extern "C"
