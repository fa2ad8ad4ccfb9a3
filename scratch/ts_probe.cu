// ts_probe.cu — probes for tcgen05.mma with the A operand in TMEM (".ts" form):
//   1. layout: A[128 x K] written with tcgen05.st.32x32b (lane = row, columns = K, 16-bit elements packed two per
//      column, low half first?) x B[N x K] K-major in 128B-swizzled shared memory -> compare with a host matmul
//   2. throughput: cycles per MMA, SS vs TS, kind::f16 vs kind::tf32, N = 128 / 256
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scratch/ts_probe.cu -o scratch/ts_probe -lcuda
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra "
      "WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <bool TF32>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (TF32)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
template <bool TF32>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (TF32)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// 1. layout test.  A fp32 [128][KE] (KE = 64 for bf16, 32 for tf32), B raw 128-byte rows [N][128 B], D [128][N]
// ---------------------------------------------------------------------------------------------------------------
template <bool TF32, bool HIGH_FIRST>
__global__ void __launch_bounds__(160) layout_kernel(const float* A, const uint8_t* Brows, float* D, int N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // B rows into 128B-swizzled K-major layout
  for (int q = threadIdx.x; q < N * 8; q += blockDim.x) {
    const int row = q >> 3, c = q & 7;
    const uint4 v = reinterpret_cast<const uint4*>(Brows)[q];
    *reinterpret_cast<uint4*>(smem + row * 128 + ((c ^ (row & 7)) << 4)) = v;
  }
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  const uint32_t a_col = 256;
  if (warp < 4) {
    const int m = warp * 32 + lane;
    uint32_t r[32];
    if (TF32) {
      for (int i = 0; i < 32; ++i) {
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(A[m * 32 + i]));
        r[i] = h;
      }
    } else {
      for (int i = 0; i < 32; ++i) {
        const uint32_t e0 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(A[m * 64 + 2 * i]));
        const uint32_t e1 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(A[m * 64 + 2 * i + 1]));
        r[i] = HIGH_FIRST ? ((e0 << 16) | e1) : ((e1 << 16) | e0);
      }
    }
    tmem_st32(tb + ((uint32_t)(warp * 32) << 16) + a_col, r);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = make_idesc(128, N, TF32 ? 2u : 1u);
    for (int k = 0; k < 4; ++k) {
      const uint64_t bd = make_desc(smem_u32(smem) + k * 32, 16, 1024);
      mma_ts<TF32>(tb, tb + a_col + 8 * k, bd, idesc, k != 0);
    }
    umma_commit(&bar);
  }
  if (warp < 4) {
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int m = warp * 32 + lane;
    for (int j = 0; j < N; j += 8) {
      float v[8];
      tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + j, v);
      for (int i = 0; i < 8; ++i) D[m * N + j + i] = v[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// 2. throughput: one thread issues iters x 4 MMAs into one accumulator; cycles from first issue to completion
// ---------------------------------------------------------------------------------------------------------------
template <bool TF32, bool TS>
__global__ void __launch_bounds__(160) rate_kernel(int N, int iters, long long* out, int extra_smem_traffic) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = threadIdx.x; q < (128 + 256) * 128 / 4; q += blockDim.x) reinterpret_cast<uint32_t*>(smem)[q] = 0x3c003c00u;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp < 4) {
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = 0x3c003c00u;
    tmem_st32(tb + ((uint32_t)(warp * 32) << 16) + 256, r);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = make_idesc(128, N, TF32 ? 2u : 1u);
    const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem) + 128 * 128;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < 4; ++k) {
        const uint64_t bd = make_desc(b_s + k * 32, 16, 1024);
        if (TS) mma_ts<TF32>(tb, tb + 256 + 8 * k, bd, idesc, 1);
        else mma_ss<TF32>(tb, make_desc(a_s + k * 32, 16, 1024), bd, idesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  } else if (warp < 4 && extra_smem_traffic) {
    // background shared-memory traffic from the other warps (what converter warps would do): ld + st 16 B per thread
    uint32_t base = smem_u32(smem) + (128 + 256) * 128 + threadIdx.x * 16;
    uint4 v = make_uint4(1, 2, 3, 4);
    for (int it = 0; it < iters * extra_smem_traffic; ++it) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (it & 7) * 2048), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + ((it + 3) & 7) * 2048));
    }
    if (v.x == 0xdeadbeef) out[1000] = v.y;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
static float tf32_round(float x) {  // round to nearest, ties away (cvt.rna)
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x1000u;
  u &= 0xffffe000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}

template <bool TF32, bool HF>
static void run_layout(const char* name) {
  const int N = 128, KE = TF32 ? 32 : 64;
  std::vector<float> A(128 * KE), Bf(N * KE), D(128 * N);
  srand(1);
  for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f);
  for (auto& v : Bf) v = (rand() / (float)RAND_MAX - 0.5f);
  std::vector<uint8_t> Brows(N * 128);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < KE; ++k) {
      if (TF32) {
        float v = tf32_round(Bf[n * KE + k]);
        Bf[n * KE + k] = v;
        memcpy(&Brows[n * 128 + k * 4], &v, 4);
      } else {
        __nv_bfloat16 h = __float2bfloat16_rn(Bf[n * KE + k]);
        Bf[n * KE + k] = __bfloat162float(h);
        memcpy(&Brows[n * 128 + k * 2], &h, 2);
      }
    }
  float *dA, *dD;
  uint8_t* dB;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMalloc(&dB, Brows.size()));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Brows.data(), Brows.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, D.size() * 4));
  CK(cudaFuncSetAttribute(layout_kernel<TF32, HF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  layout_kernel<TF32, HF><<<1, 160, 64 * 1024>>>(dA, dB, dD, N);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < KE; ++k) ref += (double)(TF32 ? tf32_round(A[m * KE + k]) : bf16_round(A[m * KE + k])) * Bf[n * KE + k];
      maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("layout %-28s max|err| %.3e  max|ref| %.3e  -> %s\n", name, maxerr, maxref, maxerr < 1e-4 * maxref ? "MATCH" : "mismatch");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

template <bool TF32, bool TS>
static void run_rate(const char* name, int N, int traffic) {
  long long* d;
  CK(cudaMalloc(&d, 2048 * 8));
  CK(cudaFuncSetAttribute(rate_kernel<TF32, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  const int iters = 500;
  for (int rep = 0; rep < 2; ++rep) {
    rate_kernel<TF32, TS><<<148, 160, 96 * 1024>>>(N, iters, d, traffic);
    CK(cudaDeviceSynchronize());
  }
  long long h[148];
  CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double per = avg / (iters * 4);
  const double macs = 128.0 * N * (TF32 ? 8 : 16);
  printf("rate %-8s N=%3d bg-traffic=%d : %.1f cycles / MMA, %.0f MAC/clk/SM (all 148 SMs busy)\n", name, N, traffic, per, macs / per);
  cudaFree(d);
}

int main() {
  run_layout<false, false>("bf16 TS low-half-first");
  run_layout<false, true>("bf16 TS high-half-first");
  run_layout<true, false>("tf32 TS");
  for (int traffic = 0; traffic <= 8; traffic += 8) {
    for (int N = 128; N <= 256; N += 128) {
      run_rate<false, false>("bf16 SS", N, traffic);
      run_rate<false, true>("bf16 TS", N, traffic);
      run_rate<true, false>("tf32 SS", N, traffic);
      run_rate<true, true>("tf32 TS", N, traffic);
    }
  }
  return 0;
}
