// Dynamic per-token activation quantisation (INT8 / FP8), one CTA per row.
// Arithmetic contract: gemlite/quant_utils.py:268-305 -- fp32 absmax / qmax, clamp 1e-6, x / s with
// the same div.full.f32 the reference's Triton `/` lowers to, clamp to [min, max], INT8 rounds half
// away from zero (libdevice.round, :255-257), FP8 converts round-to-nearest-even (saturating).
// HBM-bound: the row is read once with 128-bit loads into registers, reduced, and written once.
#include <type_traits>
#include "gl_common.cuh"

namespace glb {
namespace {

constexpr int AQ_THREADS = 256;
constexpr int AQ_MAX_PER_THREAD = 64;  // register-resident up to K = 16384

template <typename XT> __device__ __forceinline__ float x_to_float(XT v);
template <> __device__ __forceinline__ float x_to_float<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float x_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float x_to_float<float>(float v) { return v; }

template <int QD> __device__ __forceinline__ uint8_t quant_one(float v, float s, float lo, float hi) {
  float t = div_full(v, s);
  t = fminf(fmaxf(t, lo), hi);
  if (QD == GLB_INT8) return (uint8_t)(int8_t)(int)roundf(t);   // half away from zero
  if (QD == GLB_FP8E4) return float_to_fp8e4(t);
  return float_to_fp8e5(t);
}

template <typename XT, int QD>
__global__ void __launch_bounds__(AQ_THREADS) act_quant_kernel(const XT* __restrict__ x, uint8_t* __restrict__ q,
                                                               float* __restrict__ s_out, int64_t K, float qmax,
                                                               float lo, float hi) {
  __shared__ float red[AQ_THREADS / 32];
  __shared__ float s_sh;
  const int64_t row = blockIdx.x;
  const XT* xr = x + row * K;
  uint8_t* qr = q + row * K;
  pdl_wait_prior_grids();
  // pass 1: absmax (values stay in registers when the row is short enough; otherwise re-read from L2)
  float vals[AQ_MAX_PER_THREAD];
  const bool in_regs = K <= (int64_t)AQ_THREADS * AQ_MAX_PER_THREAD;
  float amax = 0.f;
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < AQ_MAX_PER_THREAD; ++i) {
      const int64_t k = (int64_t)i * AQ_THREADS + threadIdx.x;
      vals[i] = (k < K) ? x_to_float<XT>(xr[k]) : 0.f;
      amax = fmaxf(amax, fabsf(vals[i]));
    }
  } else {
    for (int64_t k = threadIdx.x; k < K; k += AQ_THREADS) amax = fmaxf(amax, fabsf(x_to_float<XT>(xr[k])));
  }
  amax = warp_max(amax);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < AQ_THREADS / 32) ? red[threadIdx.x] : 0.f;
    v = warp_max(v);
    if (threadIdx.x == 0) {
      float s = fmaxf(div_full(v, qmax), 1e-6f);
      s_sh = s;
      s_out[row] = s;
    }
  }
  __syncthreads();
  const float s = s_sh;
  pdl_launch_dependents();
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < AQ_MAX_PER_THREAD; ++i) {
      const int64_t k = (int64_t)i * AQ_THREADS + threadIdx.x;
      if (k < K) qr[k] = quant_one<QD>(vals[i], s, lo, hi);
    }
  } else {
    for (int64_t k = threadIdx.x; k < K; k += AQ_THREADS) qr[k] = quant_one<QD>(x_to_float<XT>(xr[k]), s, lo, hi);
  }
}

// 16-bit activations, K % 8 == 0, 16-byte aligned rows: 128-bit loads (8 values) and 64-bit stores, VPT vectors per
// thread register-resident between the absmax and the quantise pass (K <= 2048 * VPT).  Same arithmetic as above.
// (r1: the scalar kernel above took 55 us for 4096 x 4096 -- 0.9 TB/s -- a third of the whole A8W8 call.)
template <typename XT, int QD, int VPT>
__global__ void __launch_bounds__(AQ_THREADS) act_quant_vec_kernel(const XT* __restrict__ x, uint8_t* __restrict__ q,
                                                                   float* __restrict__ s_out, int64_t K, float qmax,
                                                                   float lo, float hi) {
  static_assert(sizeof(XT) == 2, "vector path is for fp16 / bf16 rows");
  __shared__ float red[AQ_THREADS / 32];
  __shared__ float s_sh;
  const int64_t row = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * K);
  uint2* qr = reinterpret_cast<uint2*>(q + row * K);
  const int nvec = (int)(K >> 3);
  pdl_wait_prior_grids();
  uint4 raw[VPT];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int v = i * AQ_THREADS + threadIdx.x;
    raw[i] = (v < nvec) ? xr[v] : make_uint4(0u, 0u, 0u, 0u);
    const XT* e = reinterpret_cast<const XT*>(&raw[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(x_to_float<XT>(e[j])));
  }
  amax = warp_max(amax);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < AQ_THREADS / 32) ? red[threadIdx.x] : 0.f;
    v = warp_max(v);
    if (threadIdx.x == 0) {
      float s = fmaxf(div_full(v, qmax), 1e-6f);
      s_sh = s;
      s_out[row] = s;
    }
  }
  __syncthreads();
  const float s = s_sh;
  pdl_launch_dependents();
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int v = i * AQ_THREADS + threadIdx.x;
    if (v < nvec) {
      const XT* e = reinterpret_cast<const XT*>(&raw[i]);
      uint32_t w[2] = {0u, 0u};
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j >> 2] |= (uint32_t)quant_one<QD>(x_to_float<XT>(e[j]), s, lo, hi) << (8 * (j & 3));
      qr[v] = make_uint2(w[0], w[1]);
    }
  }
}

template <typename XT, int QD, int VPT>
void launch_aq_vec(const void* x, int64_t M, int64_t K, void* q, float* s, float qmax, cudaStream_t stream) {
  act_quant_vec_kernel<XT, QD, VPT><<<dim3((unsigned)M), dim3(AQ_THREADS), 0, stream>>>((const XT*)x, (uint8_t*)q, s, K, qmax, -qmax - (QD == GLB_INT8 ? 1.f : 0.f), qmax);
}
template <typename XT, int QD>
bool try_aq_vec(const void* x, int64_t M, int64_t K, void* q, float* s, float qmax, cudaStream_t stream) {
  if ((K & 7) || K > 2048 * 8) return false;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(q) & 7)) return false;
  if (K <= 2048) launch_aq_vec<XT, QD, 1>(x, M, K, q, s, qmax, stream);
  else if (K <= 4096) launch_aq_vec<XT, QD, 2>(x, M, K, q, s, qmax, stream);
  else if (K <= 8192) launch_aq_vec<XT, QD, 4>(x, M, K, q, s, qmax, stream);
  else launch_aq_vec<XT, QD, 8>(x, M, K, q, s, qmax, stream);
  return true;
}

template <typename XT>
int launch_aq(const void* x, int64_t M, int64_t K, int q_dtype, void* q, float* s, cudaStream_t stream) {
  dim3 grid((unsigned)M), block(AQ_THREADS);
  if constexpr (sizeof(XT) == 2) {
    bool done = false;
    if (q_dtype == GLB_INT8) done = try_aq_vec<XT, GLB_INT8>(x, M, K, q, s, 127.f, stream);
    else if (q_dtype == GLB_FP8E4) done = try_aq_vec<XT, GLB_FP8E4>(x, M, K, q, s, 448.f, stream);
    else if (q_dtype == GLB_FP8E5) done = try_aq_vec<XT, GLB_FP8E5>(x, M, K, q, s, 57344.f, stream);
    if (done) {
      count_launch();
      GLB_CHECK_CUDA(cudaGetLastError());
      return GLB_OK;
    }
  }
  switch (q_dtype) {
    case GLB_INT8: act_quant_kernel<XT, GLB_INT8><<<grid, block, 0, stream>>>((const XT*)x, (uint8_t*)q, s, K, 127.f, -128.f, 127.f); break;
    case GLB_FP8E4: act_quant_kernel<XT, GLB_FP8E4><<<grid, block, 0, stream>>>((const XT*)x, (uint8_t*)q, s, K, 448.f, -448.f, 448.f); break;
    case GLB_FP8E5: act_quant_kernel<XT, GLB_FP8E5><<<grid, block, 0, stream>>>((const XT*)x, (uint8_t*)q, s, K, 57344.f, -57344.f, 57344.f); break;
    default: set_error("quant_act_per_token: unsupported q_dtype %d", q_dtype); return GLB_ERR_UNSUPPORTED;
  }
  count_launch();
  GLB_CHECK_CUDA(cudaGetLastError());
  return GLB_OK;
}

// =================================================================================================
// MX / NV activation quantisers: per row and per group of 32 (MXFP8, MXFP4; e8m0 scale) or 16
// (NVFP4; e4m3 scale x 0.05) consecutive k.  Reference: gemlite/quant_utils.py:381-391 (bitwise next power of
// two), :503-591 (scale_activations_mxfp8_triton_v2), :770-857 (mxfp4), :860-954 (nvfp4).  fp32 arithmetic with
// IEEE division, so codes and scales equal the CPU oracle's bit for bit (the reference's GPU build divides with
// div.full.f32 and can fall on the other side of an exact tie).
// One thread owns 8 consecutive k of one row: 4 threads form an MX group, 2 an NV group (shuffle max); the scale
// rows of the padding (M rounded up to the group size, as the reference allocates them) are written too: 97 = 2^-30
// for MX, 0 for NV, what the reference's masked loads (other = 0) produce there.
// STATUS: validated on a B200 in round 2 against the reference's fixture and the oracle (MXFP8 / MXFP4 bit for bit at the
// first visit; NVFP4 after the uint8 carry semantics of the code packing were reproduced, see below).
// =================================================================================================
__device__ __forceinline__ uint32_t next_pow2_exponent(float val) {          // quant_utils.py:381-391, eps_exp = -30
  const uint32_t xi = __float_as_uint(val);
  uint32_t e = (xi >> 23) & 0xffu;
  if (xi & 0x7fffffu) e += 1u;
  return min(max(e, 97u), 254u);
}

// e2m1 index of |wq| = number of the thresholds {0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5, 7} strictly below it (quant_utils.py:806-808;
// 8 = beyond the grid, kept: the reference's byte packing lets it carry).  All eight thresholds have zero low 21 bits, so
// |wq| > T  <=>  (bits(|wq|) - 1) >> 21 >= bits(T) >> 21 = {500, 506, 509, 511, 513, 515, 517, 519}: one subtraction and a
// few selects instead of eight compares (the chain made the quantiser instruction-bound: 380 instructions per thread).
__device__ __forceinline__ uint32_t e2m1_code(float wq) {
  const uint32_t a = __float_as_uint(wq) & 0x7fffffffu;
  const int d = (int)((a - 1u) >> 21) - 500;
  uint32_t idx = (d >= 9) ? min(3u + ((uint32_t)(d - 9) >> 1), 8u) : ((d >= 6) ? 2u : ((d >= 0) ? 1u : 0u));
  if (a == 0u || a > 0x7f800000u) idx = 0u;          // zero (a - 1 wraps) and NaN (every comparison false)
  return (wq >= 0.f) ? idx : idx + 8u;
}

// KIND: GLB_MXFP8 / GLB_MXFP4 / GLB_NVFP4
template <typename XT, int KIND>
__global__ void __launch_bounds__(256) act_quant_mx_kernel(const XT* __restrict__ x, uint8_t* __restrict__ q,
                                                           uint8_t* __restrict__ scales, int64_t M, int64_t M_pad, int64_t K) {
  constexpr int GS = (KIND == GLB_NVFP4) ? 16 : 32;
  constexpr int LANES = GS / 8;                       // threads per group
  const uint32_t per_row = (uint32_t)(K >> 3);        // 32-bit index math (the host checks the range): a 64-bit division costs ~100 instructions
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < (uint32_t)M_pad * per_row;    // whole groups are live or dead together (per_row % LANES == 0)
  const uint32_t r32 = live ? i / per_row : 0u;
  const int64_t r = r32, t = live ? (int64_t)(i - r32 * per_row) : 0;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  if (live && r < M) {
    const XT* src = x + r * K + t * 8;
    if constexpr (sizeof(XT) == 2) {
      const uint4 raw = *reinterpret_cast<const uint4*>(src);
      const XT* e = reinterpret_cast<const XT*>(&raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = x_to_float<XT>(e[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = x_to_float<XT>(src[j]);
    }
  }
  float amax = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(v[j]));
#pragma unroll
  for (int o = 1; o < LANES; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (!live) return;
  const int64_t g = t / LANES;
  float inv_den;                                       // elements are divided by this (IEEE)
  if (KIND == GLB_NVFP4) {
    float s = __fdiv_rn(amax, (float)(6.0 * 0.05));    // the reference's constexpr 6. * 0.05 (a Python double) as fp32
    s = fminf(s, 448.f);
    const uint8_t s8 = float_to_fp8e4(s);
    if ((t % LANES) == 0) scales[r * (K / GS) + g] = s8;
    inv_den = fmaxf(__fmul_rn(fp8e4_to_float(s8), 0.05f), 1e-6f);
  } else {
    const uint32_t e = next_pow2_exponent(__fdiv_rn(amax, (KIND == GLB_MXFP8) ? 448.f : 6.f));
    if ((t % LANES) == 0) scales[r * (K / GS) + g] = (uint8_t)e;
    inv_den = __uint_as_float(e << 23);
  }
  if (r >= M) return;
  // The MX scale is a power of two 2^(e-127): dividing by it and multiplying by 2^(127-e) round the same real number the
  // same way (both exact unless the result is subnormal, and then both round it once) -- eight IEEE divisions per thread
  // made this kernel instruction-bound (58 us for 4096 x 8192 bf16).  e = 254 needs the subnormal 2^-127; 0 and 255 keep
  // the division (0 / 0 and NaN propagation as before).  NVFP4 divides by an arbitrary e4m3 x 0.05: IEEE division stays.
  float mul = 0.f;
  bool use_mul = false;
  if (KIND != GLB_NVFP4) {
    const uint32_t e = __float_as_uint(inv_den) >> 23;
    if (e >= 1u && e <= 254u) { mul = (e <= 253u) ? __uint_as_float((254u - e) << 23) : __uint_as_float(0x00400000u); use_mul = true; }
  }
  if (KIND == GLB_MXFP8) {
    uint32_t w[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float qv = use_mul ? __fmul_rn(v[j], mul) : __fdiv_rn(v[j], inv_den);
      const float o = fminf(fmaxf(qv, -448.f), 448.f);
      w[j >> 2] |= (uint32_t)float_to_fp8e4(o) << (8 * (j & 3));
    }
    *reinterpret_cast<uint2*>(q + r * K + t * 8) = make_uint2(w[0], w[1]);
  } else {
    // two codes per byte, even k = low nibble, in the reference's uint8 arithmetic (`lo | (hi << 4)`,
    // quant_utils.py:810-812 / 903-905): a saturated NVFP4 group (e4m3 scale clamped at 448) can produce |wq| > 7,
    // i.e. index 8 (16 for negatives) -- the carry of an even code lands in the odd nibble of the SAME byte, the carry
    // of an odd code falls off the byte.  Reproduced as is (found on hardware against the reference fixture).
    uint32_t w = 0u;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const float q0 = use_mul ? __fmul_rn(v[j], mul) : __fdiv_rn(v[j], inv_den);
      const float q1 = use_mul ? __fmul_rn(v[j + 1], mul) : __fdiv_rn(v[j + 1], inv_den);
      const uint32_t lo = e2m1_code(q0), hi = e2m1_code(q1);
      w |= ((lo | (hi << 4)) & 0xffu) << (4 * j);
    }
    *reinterpret_cast<uint32_t*>(q + r * (K >> 1) + t * 4) = w;
  }
}

// 16-bit inputs, 16 consecutive k per thread (two 128-bit loads): the 8-per-thread kernel above is instruction-bound
// (profiles/mx_dynamic_breakdown_r2.txt: 45 us for 4096 x 8192 bf16 -> e2m1, a quarter of an A4W4 layer call) -- index
// arithmetic, the group reduction and the scale computation are per thread, so twice the elements per thread halves them;
// an MX group is 2 threads, an NV group 1.  Same arithmetic, same outputs (bit for bit: tests/test_gpu_parity.py).
template <typename XT, int KIND>
__global__ void __launch_bounds__(256) act_quant_mx16_kernel(const XT* __restrict__ x, uint8_t* __restrict__ q,
                                                             uint8_t* __restrict__ scales, uint32_t M, uint32_t M_pad, uint32_t K) {
  static_assert(sizeof(XT) == 2, "16-bit activations only");
  constexpr int GS = (KIND == GLB_NVFP4) ? 16 : 32;
  constexpr int LANES = GS / 16;
  const uint32_t per_row = K >> 4;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < M_pad * per_row;              // whole groups are live or dead together (per_row % LANES == 0)
  const uint32_t r = live ? i / per_row : 0u;
  const uint32_t t = live ? i - r * per_row : 0u;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0.f;
  if (live && r < M) {
    const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)r * K + (size_t)t * 16);
    const uint4 a = __ldg(src), b = __ldg(src + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (sizeof(XT) == 2 && std::is_same<XT, __nv_bfloat16>::value) {
        v[2 * j] = __uint_as_float(w[j] << 16);
        v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
      } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
        v[2 * j] = f.x; v[2 * j + 1] = f.y;
      }
    }
  }
  float amax = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) amax = fmaxf(amax, fabsf(v[j]));
#pragma unroll
  for (int o = 1; o < LANES; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (!live) return;
  const uint32_t g = t / LANES;
  float mul = 0.f, inv_den = 1.f;
  if (KIND == GLB_NVFP4) {
    float sc = __fdiv_rn(amax, (float)(6.0 * 0.05));
    sc = fminf(sc, 448.f);
    const uint8_t s8 = float_to_fp8e4(sc);
    scales[(size_t)r * (K / GS) + g] = s8;
    inv_den = fmaxf(__fmul_rn(fp8e4_to_float(s8), 0.05f), 1e-6f);
  } else {
    const uint32_t e = next_pow2_exponent(__fdiv_rn(amax, (KIND == GLB_MXFP8) ? 448.f : 6.f));     // 97 .. 254
    if ((t % LANES) == 0) scales[(size_t)r * (K / GS) + g] = (uint8_t)e;
    mul = (e <= 253u) ? __uint_as_float((254u - e) << 23) : __uint_as_float(0x00400000u);           // exact reciprocal of 2^(e-127)
  }
  if (r >= M) return;
  float qv[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) qv[j] = (KIND == GLB_NVFP4) ? __fdiv_rn(v[j], inv_den) : __fmul_rn(v[j], mul);
  if (KIND == GLB_MXFP8) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float o = fminf(fmaxf(qv[j], -448.f), 448.f);
      w[j >> 2] |= (uint32_t)float_to_fp8e4(o) << (8 * (j & 3));
    }
    *reinterpret_cast<uint4*>(q + (size_t)r * K + (size_t)t * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    uint32_t w[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const uint32_t lo = e2m1_code(qv[j]), hi = e2m1_code(qv[j + 1]);
      w[j >> 3] |= ((lo | (hi << 4)) & 0xffu) << (4 * (j & 7));      // uint8 arithmetic of the reference (see above)
    }
    *reinterpret_cast<uint2*>(q + (size_t)r * (K >> 1) + (size_t)t * 8) = make_uint2(w[0], w[1]);
  }
}

template <typename XT>
int launch_aq_mx(const void* x, int64_t M, int64_t K, int mx_dtype, void* q, void* scales, cudaStream_t stream) {
  const int gs = (mx_dtype == GLB_NVFP4) ? 16 : 32;
  const int64_t M_pad = (M + gs - 1) / gs * gs;
  if (sizeof(XT) == 2 && (reinterpret_cast<uintptr_t>(q) & 15) == 0 && M_pad * (K >> 4) < 0xffffff00LL && K < 0x7fffffffLL) {
    // 16 elements per thread (see act_quant_mx16_kernel)
    static int v16 = -1;
    if (v16 < 0) { const char* e = getenv("GLB_AQ_MX16"); v16 = e ? atoi(e) : 1; }
    if (v16) {
      const unsigned grid16 = (unsigned)((M_pad * (K >> 4) + 255) / 256);
      using X16 = typename std::conditional<sizeof(XT) == 2, XT, __half>::type;
      switch (mx_dtype) {
        case GLB_MXFP8: act_quant_mx16_kernel<X16, GLB_MXFP8><<<grid16, 256, 0, stream>>>((const X16*)x, (uint8_t*)q, (uint8_t*)scales, (uint32_t)M, (uint32_t)M_pad, (uint32_t)K); break;
        case GLB_MXFP4: act_quant_mx16_kernel<X16, GLB_MXFP4><<<grid16, 256, 0, stream>>>((const X16*)x, (uint8_t*)q, (uint8_t*)scales, (uint32_t)M, (uint32_t)M_pad, (uint32_t)K); break;
        case GLB_NVFP4: act_quant_mx16_kernel<X16, GLB_NVFP4><<<grid16, 256, 0, stream>>>((const X16*)x, (uint8_t*)q, (uint8_t*)scales, (uint32_t)M, (uint32_t)M_pad, (uint32_t)K); break;
        default: set_error("quant_act_mx: unsupported mx_dtype %d", mx_dtype); return GLB_ERR_UNSUPPORTED;
      }
      count_launch();
      GLB_CHECK_CUDA(cudaGetLastError());
      return GLB_OK;
    }
  }
  const int64_t threads = M_pad * (K >> 3);
  const unsigned grid = (unsigned)((threads + 255) / 256);
  switch (mx_dtype) {
    case GLB_MXFP8: act_quant_mx_kernel<XT, GLB_MXFP8><<<grid, 256, 0, stream>>>((const XT*)x, (uint8_t*)q, (uint8_t*)scales, M, M_pad, K); break;
    case GLB_MXFP4: act_quant_mx_kernel<XT, GLB_MXFP4><<<grid, 256, 0, stream>>>((const XT*)x, (uint8_t*)q, (uint8_t*)scales, M, M_pad, K); break;
    case GLB_NVFP4: act_quant_mx_kernel<XT, GLB_NVFP4><<<grid, 256, 0, stream>>>((const XT*)x, (uint8_t*)q, (uint8_t*)scales, M, M_pad, K); break;
    default: set_error("quant_act_mx: unsupported mx_dtype %d", mx_dtype); return GLB_ERR_UNSUPPORTED;
  }
  count_launch();
  GLB_CHECK_CUDA(cudaGetLastError());
  return GLB_OK;
}

// MX-quantised activations (e4m3 bytes or e2m1 nibbles + one e8m0 scale per 32) -> their exact values in bf16.
// e4m3 has 4 significant bits, e2m1 2, the scale is a power of two >= 2^-30: every product is representable in bf16, so
// a bf16 tensor-core GEMM over the expanded activations accumulates exactly the products tl.dot_scaled would
// (gemm_kernels.py:508-524) -- used to run A8W4 / A4W4 MXFP dynamic layers on the MXFP4 tcgen05 kernels (gl_api.cu).
template <bool FP4>
__global__ void __launch_bounds__(256) expand_act_mx_kernel(const uint8_t* __restrict__ q, const uint8_t* __restrict__ scales,
                                                            __nv_bfloat16* __restrict__ out, int64_t M, int64_t K) {
  const int64_t per_row = K >> 3;                                  // 8 k per thread
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * per_row) return;
  const int64_t r = i / per_row, t = i - r * per_row;
  const float sc = e8m0_to_float(scales[r * (K >> 5) + (t >> 2)]);
  float v[8];
  if (FP4) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(q + r * (K >> 1) + t * 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fp4_e2m1_to_float((w >> (4 * j)) & 15u);
  } else {
    const uint2 w = *reinterpret_cast<const uint2*>(q + r * K + t * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fp8e4_to_float((uint8_t)(((j < 4) ? w.x : w.y) >> (8 * (j & 3))));
  }
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j] * sc, v[2 * j + 1] * sc);      // exact
    o[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(out + r * K + t * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

}  // namespace

int launch_expand_act_mx(const void* q, const void* scales, void* out_bf16, int64_t M, int64_t K, int mx_dtype, cudaStream_t stream) {
  const int64_t threads = M * (K >> 3);
  const unsigned grid = (unsigned)((threads + 255) / 256);
  if (mx_dtype == GLB_MXFP4)
    expand_act_mx_kernel<true><<<grid, 256, 0, stream>>>((const uint8_t*)q, (const uint8_t*)scales, (__nv_bfloat16*)out_bf16, M, K);
  else
    expand_act_mx_kernel<false><<<grid, 256, 0, stream>>>((const uint8_t*)q, (const uint8_t*)scales, (__nv_bfloat16*)out_bf16, M, K);
  count_launch();
  GLB_CHECK_CUDA(cudaGetLastError());
  return GLB_OK;
}

}  // namespace glb

using namespace glb;

extern "C" int glb_quant_act_mx(const void* x, int64_t M, int64_t K, int x_dtype, int mx_dtype, void* q, void* scales,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GLB_REQUIRE(x && q && scales && M > 0 && K > 0, GLB_ERR_INVALID, "quant_act_mx: bad arguments");
  GLB_REQUIRE(K % 32 == 0, GLB_ERR_UNSUPPORTED, "quant_act_mx: K must be a multiple of 32 (got %lld)", (long long)K);
  GLB_REQUIRE(M <= 0x7fffffffLL && (M + 32) * (K >> 3) < 0xffffff00LL, GLB_ERR_UNSUPPORTED, "quant_act_mx: too large");
  GLB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(q) & 7) == 0, GLB_ERR_INVALID,
              "quant_act_mx: x must be 16-byte and q 8-byte aligned");
  set_kernel_name("act_quant_mx");
  switch (x_dtype) {
    case GLB_FP16: return launch_aq_mx<__half>(x, M, K, mx_dtype, q, scales, stream);
    case GLB_BF16: return launch_aq_mx<__nv_bfloat16>(x, M, K, mx_dtype, q, scales, stream);
    case GLB_FP32: return launch_aq_mx<float>(x, M, K, mx_dtype, q, scales, stream);
    default: set_error("quant_act_mx: unsupported x_dtype %d", x_dtype); return GLB_ERR_UNSUPPORTED;
  }
}

extern "C" int glb_quant_act_per_token(const void* x, int64_t M, int64_t K, int x_dtype, int q_dtype, void* q, float* s,
                                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GLB_REQUIRE(x && q && s && M > 0 && K > 0, GLB_ERR_INVALID, "quant_act_per_token: bad arguments");
  GLB_REQUIRE(M <= 0x7fffffffLL, GLB_ERR_UNSUPPORTED, "quant_act_per_token: M too large");
  set_kernel_name("act_quant_per_token");
  switch (x_dtype) {
    case GLB_FP16: return launch_aq<__half>(x, M, K, q_dtype, q, s, stream);
    case GLB_BF16: return launch_aq<__nv_bfloat16>(x, M, K, q_dtype, q, s, stream);
    case GLB_FP32: return launch_aq<float>(x, M, K, q_dtype, q, s, stream);
    default: set_error("quant_act_per_token: unsupported x_dtype %d", x_dtype); return GLB_ERR_UNSUPPORTED;
  }
}
