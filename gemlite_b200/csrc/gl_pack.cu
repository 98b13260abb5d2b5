// Bit-packing kernels: N-bit codes (one per byte) -> packed words along the K ("cols") axis.
// Layout contract: gemlite/bitpack.py:36-60 -- element j of a row lands in word j / eps at bit
// offset (j % eps) * nbits, eps = pack_bits / nbits; GemLiteLinear stores the transpose
// (cols/eps, rows) with `rows` (= out_features) contiguous (core.py:387-392,478-480).
//
// HBM-bound byte shuffling: each block converts a 32(words) x 32(rows) tile; reads are contiguous
// along the code row, the transposed store goes through a padded shared tile so that both sides
// are coalesced.
#include "gl_common.cuh"

namespace glb {
namespace {

constexpr int PT = 32;

template <typename WordT>
__global__ void __launch_bounds__(PT * 8) pack_cols_kernel(const uint8_t* __restrict__ src, WordT* __restrict__ dst,
                                                           int64_t rows, int64_t words_per_row, int nbits, int eps,
                                                           int transpose) {
  __shared__ unsigned long long tile[PT][PT + 1];
  const int64_t w0 = (int64_t)blockIdx.x * PT, r0 = (int64_t)blockIdx.y * PT;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int rr = ty; rr < PT; rr += 8) {
    const int64_t r = r0 + rr, w = w0 + tx;
    unsigned long long acc = 0ull;
    if (r < rows && w < words_per_row) {
      const uint8_t* p = src + r * (words_per_row * eps) + w * eps;
      for (int j = 0; j < eps; ++j) acc |= (unsigned long long)p[j] << (j * nbits);
    }
    tile[rr][tx] = acc;
  }
  __syncthreads();
  if (transpose) {
    for (int ww = ty; ww < PT; ww += 8) {
      const int64_t w = w0 + ww, r = r0 + tx;
      if (r < rows && w < words_per_row) dst[w * rows + r] = (WordT)tile[tx][ww];
    }
  } else {
    for (int rr = ty; rr < PT; rr += 8) {
      const int64_t r = r0 + rr, w = w0 + tx;
      if (r < rows && w < words_per_row) dst[r * words_per_row + w] = (WordT)tile[rr][tx];
    }
  }
}

template <typename WordT>
__global__ void __launch_bounds__(PT * 8) unpack_cols_kernel(const WordT* __restrict__ src, uint8_t* __restrict__ dst,
                                                             int64_t rows, int64_t words_per_row, int nbits, int eps,
                                                             int transpose) {
  __shared__ unsigned long long tile[PT][PT + 1];
  const int64_t w0 = (int64_t)blockIdx.x * PT, r0 = (int64_t)blockIdx.y * PT;
  const int tx = threadIdx.x, ty = threadIdx.y;
  using U = typename std::make_unsigned<WordT>::type;
  if (transpose) {
    for (int ww = ty; ww < PT; ww += 8) {
      const int64_t w = w0 + ww, r = r0 + tx;
      tile[tx][ww] = (r < rows && w < words_per_row) ? (unsigned long long)(U)src[w * rows + r] : 0ull;
    }
  } else {
    for (int rr = ty; rr < PT; rr += 8) {
      const int64_t r = r0 + rr, w = w0 + tx;
      tile[rr][tx] = (r < rows && w < words_per_row) ? (unsigned long long)(U)src[r * words_per_row + w] : 0ull;
    }
  }
  __syncthreads();
  const unsigned long long mask = (nbits == 64) ? ~0ull : ((1ull << nbits) - 1ull);
  for (int rr = ty; rr < PT; rr += 8) {
    const int64_t r = r0 + rr, w = w0 + tx;
    if (r < rows && w < words_per_row) {
      const unsigned long long v = tile[rr][tx];
      uint8_t* p = dst + r * (words_per_row * eps) + w * eps;
      for (int j = 0; j < eps; ++j) p[j] = (uint8_t)((v >> (j * nbits)) & mask);
    }
  }
}

}  // namespace
}  // namespace glb

using namespace glb;

extern "C" int glb_pack_cols(const uint8_t* W_q, int64_t rows, int64_t cols, int nbits, int pack_bits, void* out,
                             int transpose, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GLB_REQUIRE(pack_bits == 8 || pack_bits == 16 || pack_bits == 32 || pack_bits == 64, GLB_ERR_INVALID,
              "Unsuported bitpacking width %d", pack_bits);
  GLB_REQUIRE(nbits == 8 || nbits == 4 || nbits == 2 || nbits == 1, GLB_ERR_INVALID, "Unsuported nbits %d", nbits);
  GLB_REQUIRE(pack_bits >= nbits, GLB_ERR_INVALID, "pack_bits < nbits");
  const int eps = pack_bits / nbits;
  GLB_REQUIRE(W_q && out && rows > 0 && cols > 0 && cols % eps == 0, GLB_ERR_INVALID,
              "pack_cols: bad shape (%lld, %lld) for eps=%d", (long long)rows, (long long)cols, eps);
  const int64_t wpr = cols / eps;
  dim3 block(PT, 8), grid((unsigned)((wpr + PT - 1) / PT), (unsigned)((rows + PT - 1) / PT));
  GLB_REQUIRE(grid.y <= 65535u, GLB_ERR_UNSUPPORTED, "pack_cols: too many rows");
  set_kernel_name("pack_cols");
  switch (pack_bits) {
    case 8: pack_cols_kernel<uint8_t><<<grid, block, 0, stream>>>(W_q, (uint8_t*)out, rows, wpr, nbits, eps, transpose); break;
    case 16: pack_cols_kernel<int16_t><<<grid, block, 0, stream>>>(W_q, (int16_t*)out, rows, wpr, nbits, eps, transpose); break;
    case 32: pack_cols_kernel<int32_t><<<grid, block, 0, stream>>>(W_q, (int32_t*)out, rows, wpr, nbits, eps, transpose); break;
    default: pack_cols_kernel<long long><<<grid, block, 0, stream>>>(W_q, (long long*)out, rows, wpr, nbits, eps, transpose); break;
  }
  count_launch();
  GLB_CHECK_CUDA(cudaGetLastError());
  return GLB_OK;
}

extern "C" int glb_unpack_cols(const void* packed, int64_t rows, int64_t cols, int nbits, int pack_bits, uint8_t* out,
                               int transpose, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GLB_REQUIRE(pack_bits == 8 || pack_bits == 16 || pack_bits == 32 || pack_bits == 64, GLB_ERR_INVALID,
              "Unsuported bitpacking width %d", pack_bits);
  GLB_REQUIRE(nbits == 8 || nbits == 4 || nbits == 2 || nbits == 1, GLB_ERR_INVALID, "Unsuported nbits %d", nbits);
  GLB_REQUIRE(pack_bits >= nbits, GLB_ERR_INVALID, "pack_bits < nbits");
  const int eps = pack_bits / nbits;
  GLB_REQUIRE(packed && out && rows > 0 && cols > 0 && cols % eps == 0, GLB_ERR_INVALID, "unpack_cols: bad shape");
  const int64_t wpr = cols / eps;
  dim3 block(PT, 8), grid((unsigned)((wpr + PT - 1) / PT), (unsigned)((rows + PT - 1) / PT));
  GLB_REQUIRE(grid.y <= 65535u, GLB_ERR_UNSUPPORTED, "unpack_cols: too many rows");
  set_kernel_name("unpack_cols");
  switch (pack_bits) {
    case 8: unpack_cols_kernel<uint8_t><<<grid, block, 0, stream>>>((const uint8_t*)packed, out, rows, wpr, nbits, eps, transpose); break;
    case 16: unpack_cols_kernel<int16_t><<<grid, block, 0, stream>>>((const int16_t*)packed, out, rows, wpr, nbits, eps, transpose); break;
    case 32: unpack_cols_kernel<int32_t><<<grid, block, 0, stream>>>((const int32_t*)packed, out, rows, wpr, nbits, eps, transpose); break;
    default: unpack_cols_kernel<long long><<<grid, block, 0, stream>>>((const long long*)packed, out, rows, wpr, nbits, eps, transpose); break;
  }
  count_launch();
  GLB_CHECK_CUDA(cudaGetLastError());
  return GLB_OK;
}
