// peer_reduce.cu — the gradient all-reduce of the data-parallel step as ONE kernel over NVLink peer memory.
//
// The reference averages the replicas' gradients inside nn.DataParallel's backward (src/train.py:83-85, one hub GPU).
// Here every rank's flat gradient buffer is a CUDA-IPC allocation that every other rank of the node has mapped
// (ctn_peer_alloc / ctn_peer_export / ctn_peer_open), and one kernel per step does the whole exchange:
//   1. barrier: "my gradients are complete" flags written into every peer's flag block (st.release.sys), every block
//      waits for all of them in its own flag block (ld.acquire.sys);
//   2. reduce-scatter + all-gather fused: the rank owns the slice [rank n/W, (rank+1) n/W); it loads that slice from every
//      peer's buffer over NVLink (peer loads, fixed rank order: every replica ends with bit-identical gradients), scales,
//      and stores the result into the same slice of EVERY peer's buffer (peer stores) — in place: nobody else reads or
//      writes that slice in this phase;
//   3. the last block to finish tells every peer "my slice has landed" and waits for theirs; the kernel ends, and the
//      clip + Adam kernels that follow on the stream read a complete, averaged local buffer.
// No NCCL call, no host involvement: the kernel is captured into the step's CUDA graph like any other, so the
// data-parallel step is one graph launch exactly like the single-GPU step.  34.8 MB of gradients move as W-1 slices in
// and W-1 slices out per GPU, all links busy in both directions at once.
//
// Flag block (one per rank, zero-initialised, uint32): [0, 8) barrier-1 slots written by peers, [8, 16) barrier-2 slots,
// [16] local epoch, [17] local finished-block counter, [18] sticky error: a wait gave up after the time-out (default 600 s
// like NCCL's watchdog, CTN_PEER_TIMEOUT_S) — the kernel then traps, so the step fails loudly instead of applying
// gradients that miss a rank.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ctn {
namespace {

constexpr int PEER_MAX = 8;

struct PeerArgs {
  float* buf[PEER_MAX];
  uint32_t* flags[PEER_MAX];
  int rank, world;
  int64_t offset4, count4;  // in float4 units
  float scale;
  unsigned long long timeout_ns;  // a rank that never arrives: give up (and trap)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {  // never from a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float4* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// wait until the slot written by a peer reaches `want` (epochs only grow; the comparison is wrap-safe)
__device__ __forceinline__ void wait_slot(uint32_t* slot, uint32_t want, uint32_t* err, unsigned long long timeout_ns) {
  const unsigned long long t0 = global_ns();
  while ((int32_t)(ld_acquire_sys(slot) - want) < 0) {
    if (global_ns() - t0 > timeout_ns) {
      *err = 1u;
      __threadfence_system();
      __trap();
    }
  }
}

template <int W>
__global__ void __launch_bounds__(512) peer_all_reduce_kernel(PeerArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ uint32_t s_epoch, s_last;
  uint32_t* mine = a.flags[a.rank];
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(mine + 16);
  __syncthreads();
  const uint32_t e1 = 2u * s_epoch + 1u, e2 = e1 + 1u;
  // ---- 1. every rank's gradients are complete (their kernels precede this one on each rank's stream) ----
  if (blockIdx.x == 0 && threadIdx.x < W) st_release_sys(a.flags[threadIdx.x] + a.rank, e1);
  if (threadIdx.x < W) wait_slot(mine + threadIdx.x, e1, mine + 18, a.timeout_ns);
  __syncthreads();
  // ---- 2. my slice: sum over the ranks in rank order, scale, store into every rank's buffer ----
  const int64_t per = (a.count4 + W - 1) / W;
  const int64_t lo = a.offset4 + per * a.rank;
  int64_t hi = lo + per;
  if (hi > a.offset4 + a.count4) hi = a.offset4 + a.count4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += 2 * stride) {
    const int64_t i2 = i + stride;
    const bool two = i2 < hi;
    float4 v[W], w[W];
#pragma unroll
    for (int p = 0; p < W; ++p) {
      v[p] = ld_peer(reinterpret_cast<const float4*>(a.buf[p]) + i);
      if (two) w[p] = ld_peer(reinterpret_cast<const float4*>(a.buf[p]) + i2);
    }
    float4 s = v[0], t = two ? w[0] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 1; p < W; ++p) {
      s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w;
      if (two) { t.x += w[p].x; t.y += w[p].y; t.z += w[p].z; t.w += w[p].w; }
    }
    s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
    t.x *= a.scale; t.y *= a.scale; t.z *= a.scale; t.w *= a.scale;
#pragma unroll
    for (int p = 0; p < W; ++p) {
      st_peer(reinterpret_cast<float4*>(a.buf[p]) + i, s);
      if (two) st_peer(reinterpret_cast<float4*>(a.buf[p]) + i2, t);
    }
  }
  // ---- 3. the last block to finish announces the slice and waits for everybody else's ----
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(mine + 17, 1u);
    __threadfence();
    s_last = prev == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < W) {
      st_release_sys(a.flags[threadIdx.x] + 8 + a.rank, e2);
      wait_slot(mine + 8 + threadIdx.x, e2, mine + 18, a.timeout_ns);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mine[17] = 0u;
      mine[16] = s_epoch + 1u;
      __threadfence();
    }
  }
}

}  // namespace

int run_peer_all_reduce(float* const* bufs, uint32_t* const* flags, int rank, int world, int64_t offset, int64_t count,
                        float scale, cudaStream_t s) {
  CTN_REQUIRE(world >= 2 && world <= PEER_MAX, "peer_all_reduce: world size %d not in [2, %d]", world, PEER_MAX);
  CTN_REQUIRE(rank >= 0 && rank < world, "peer_all_reduce: rank %d out of range", rank);
  CTN_REQUIRE(bufs != nullptr && flags != nullptr, "peer_all_reduce: null pointer table");
  CTN_REQUIRE(offset >= 0 && count > 0 && offset % 4 == 0 && count % 4 == 0,
              "peer_all_reduce: offset and count must be multiples of 4 floats (got %lld, %lld)", (long long)offset,
              (long long)count);
  PeerArgs a = {};
  for (int p = 0; p < world; ++p) {
    CTN_REQUIRE(bufs[p] != nullptr && flags[p] != nullptr, "peer_all_reduce: rank %d's buffer is not mapped", p);
    CTN_REQUIRE((reinterpret_cast<uintptr_t>(bufs[p]) & 15) == 0, "peer_all_reduce: buffers must be 16-byte aligned");
    a.buf[p] = bufs[p];
    a.flags[p] = flags[p];
  }
  a.rank = rank; a.world = world; a.offset4 = offset / 4; a.count4 = count / 4; a.scale = scale;
  static const long long timeout_s = getenv("CTN_PEER_TIMEOUT_S") ? atoll(getenv("CTN_PEER_TIMEOUT_S")) : 600;
  a.timeout_ns = (unsigned long long)(timeout_s > 0 ? timeout_s : 600) * 1000000000ull;
  // enough blocks to keep every link busy, few enough that all of them are resident while they wait at the barrier
  int dev = 0, sms = 148;
  CTN_CUDA(cudaGetDevice(&dev));
  CTN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t per = (a.count4 + world - 1) / world;
  int grid = (int)((per + 2 * 512 - 1) / (2 * 512));
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  switch (world) {
#define CTN_PEER_CASE(W_) case W_: launch_kernel(peer_all_reduce_kernel<W_>, grid, 512, 0, s, a); break;
    CTN_PEER_CASE(2) CTN_PEER_CASE(3) CTN_PEER_CASE(4) CTN_PEER_CASE(5) CTN_PEER_CASE(6) CTN_PEER_CASE(7) CTN_PEER_CASE(8)
#undef CTN_PEER_CASE
    default: break;
  }
  return check_launch("peer_all_reduce_kernel");
}

}  // namespace ctn

using namespace ctn;

extern "C" {

int32_t ctn_peer_alloc(int64_t bytes, void** ptr) {
  CTN_REQUIRE(ptr != nullptr && bytes > 0, "peer_alloc: bad arguments");
  CTN_CUDA(cudaMalloc(ptr, (size_t)bytes));  // a whole allocation of its own: its IPC handle maps exactly this buffer
  CTN_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
  CTN_CUDA(cudaDeviceSynchronize());
  return 0;
}

int32_t ctn_peer_free(void* ptr) {
  if (ptr != nullptr) CTN_CUDA(cudaFree(ptr));
  return 0;
}

int32_t ctn_peer_export(const void* ptr, uint8_t* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CTN_REQUIRE(ptr != nullptr && handle64 != nullptr, "peer_export: null pointer");
  cudaIpcMemHandle_t h;
  CTN_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle64, &h, 64);
  return 0;
}

int32_t ctn_peer_open(const uint8_t* handle64, void** ptr) {
  CTN_REQUIRE(ptr != nullptr && handle64 != nullptr, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CTN_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int32_t ctn_peer_close(void* ptr) {
  if (ptr != nullptr) CTN_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int32_t ctn_peer_all_reduce(float* const* bufs_host, uint32_t* const* flags_host, int32_t rank, int32_t world,
                            int64_t offset, int64_t count, float scale, cudaStream_t stream) {
  return run_peer_all_reduce(bufs_host, flags_host, rank, world, offset, count, scale, stream);
}

}  // extern "C"
