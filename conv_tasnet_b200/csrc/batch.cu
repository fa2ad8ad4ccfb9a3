// batch.cu — device-side batch assembly: what _collate_fn + pad_list + the three .cuda() copies do on
// the host in the reference (src/data.py:159-183,322-331; src/solver.py:184-187), and the inverse (utils.remove_pad,
// src/utils.py:50-67) for the evaluation / separation loops.
//
// The host packs the ragged utterances back to back (no padding, no permute) into ONE pinned staging buffer
// [offsets int64 (B+1) | mixtures | sources], one async H2D moves it, and this kernel writes the zero-padded
// padded_mixture [B,T], padded_source [B,C,T] (transposing the loader's [T_i, C] layout) and the lengths.
#include "common.cuh"

namespace ctn {
namespace {

__global__ void __launch_bounds__(256) assemble_batch_kernel(const float* __restrict__ packed_mix,
                                                             const float* __restrict__ packed_src,
                                                             const int64_t* __restrict__ offsets, int C, int T,
                                                             float* __restrict__ mix_out, float* __restrict__ src_out,
                                                             int64_t* __restrict__ lengths_out) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int64_t off = offsets[b], len = offsets[b + 1] - off;
  if (blockIdx.x == 0 && threadIdx.x == 0) lengths_out[b] = len;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const bool valid = t < len;
  mix_out[(int64_t)b * T + t] = valid ? packed_mix[off + t] : 0.f;
  if (packed_src != nullptr) {
    for (int c = 0; c < C; ++c)
      src_out[((int64_t)b * C + c) * T + t] = valid ? packed_src[(off + t) * C + c] : 0.f;
  }
}

// remove_pad: packed[out_off[b] * C + c * len_b + t] = in[b, c, t] for t < len_b  (item b = a dense [C, len_b] block)
__global__ void __launch_bounds__(256) pack_valid_kernel(const float* __restrict__ in, const int64_t* __restrict__ lengths,
                                                         const int64_t* __restrict__ out_off, int C, int T,
                                                         float* __restrict__ packed) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  int64_t len = lengths[b];
  len = len < 0 ? 0 : (len > T ? T : len);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= len) return;
  for (int c = 0; c < C; ++c) packed[out_off[b] * C + (int64_t)c * len + t] = in[((int64_t)b * C + c) * T + t];
}

}  // namespace

int run_assemble_batch(const float* packed_mix, const float* packed_src, const int64_t* offsets, int B, int C, int T,
                       float* mix_out, float* src_out, int64_t* lengths_out, cudaStream_t s) {
  CTN_REQUIRE(B >= 1 && B <= 65535 && T >= 1 && C >= 1, "assemble_batch: bad shape B=%d C=%d T=%d", B, C, T);
  CTN_REQUIRE((packed_src == nullptr) == (src_out == nullptr), "assemble_batch: packed_src and src_out go together");
  launch_kernel(assemble_batch_kernel, dim3(cdiv(T, 256), B), 256, 0, s, packed_mix, packed_src, offsets, C, T, mix_out,
                src_out, lengths_out);
  return check_launch("assemble_batch_kernel");
}

int run_pack_valid(const float* in, const int64_t* lengths, const int64_t* out_off, int B, int C, int T, float* packed,
                   cudaStream_t s) {
  CTN_REQUIRE(B >= 1 && B <= 65535 && T >= 1 && C >= 1, "pack_valid: bad shape B=%d C=%d T=%d", B, C, T);
  launch_kernel(pack_valid_kernel, dim3(cdiv(T, 256), B), 256, 0, s, in, lengths, out_off, C, T, packed);
  return check_launch("pack_valid_kernel");
}

}  // namespace ctn
