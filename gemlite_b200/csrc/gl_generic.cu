// General-case fused dequant + matmul kernel.
//
// Covers every (W_nbits, packing word, W_group_mode, channel_scale_mode, activation dtype, meta
// dtype, output dtype) combination GemLiteLinear.pack() can produce, with the reference's rounding
// points (dequantize(): gemlite/triton_kernels/utils.py:57-89; epilogue: gemm_kernels.py:392-406).
// It is a plain CUDA-core kernel: one thread per output column, a tile of up to 8 rows of x staged
// in shared memory.  The specialised decode (gl_gemv.cu) and tcgen05 (gl_gemm_tc.cu) kernels take
// the shapes that matter for throughput; this one is what forward() uses for everything else.
#include "gl_common.cuh"

namespace glb {
namespace {

constexpr int GEN_COLS = 128;   // threads per block == output columns per block
constexpr int GEN_ROWS = 8;     // rows of x per block
constexpr int GEN_KC = 256;     // k-chunk staged in shared memory

// ---- arithmetic in the "meta" (scale) dtype, each op individually rounded like the reference ----
template <typename T> struct Meta;
template <> struct Meta<__half> {
  using T = __half;
  static __device__ T from_int(int v) { return __int2half_rn(v); }
  static __device__ T from_float(float v) { return __float2half_rn(v); }
  static __device__ float to_float(T v) { return __half2float(v); }
  static __device__ T sub(T a, T b) { return __hsub(a, b); }
  static __device__ T mul(T a, T b) { return __hmul(a, b); }
  static __device__ T fma(T a, T b, T c) { return __hfma(a, b, c); }
  static __device__ T load(const void* p, int64_t i) { return reinterpret_cast<const __half*>(p)[i]; }
};
template <> struct Meta<__nv_bfloat16> {
  using T = __nv_bfloat16;
  static __device__ T from_int(int v) { return __int2bfloat16_rn(v); }
  static __device__ T from_float(float v) { return __float2bfloat16_rn(v); }
  static __device__ float to_float(T v) { return __bfloat162float(v); }
  static __device__ T sub(T a, T b) { return __hsub(a, b); }
  static __device__ T mul(T a, T b) { return __hmul(a, b); }
  static __device__ T fma(T a, T b, T c) { return __hfma(a, b, c); }
  static __device__ T load(const void* p, int64_t i) { return reinterpret_cast<const __nv_bfloat16*>(p)[i]; }
};
template <> struct Meta<float> {
  using T = float;
  static __device__ T from_int(int v) { return (float)v; }
  static __device__ T from_float(float v) { return v; }
  static __device__ float to_float(T v) { return v; }
  static __device__ T sub(T a, T b) { return __fsub_rn(a, b); }
  static __device__ T mul(T a, T b) { return __fmul_rn(a, b); }
  static __device__ T fma(T a, T b, T c) { return __fmaf_rn(a, b, c); }
  static __device__ T load(const void* p, int64_t i) { return reinterpret_cast<const float*>(p)[i]; }
};

struct GenParams {
  glb_linear L;
  const void* x;
  const float* scales_x;
  void* y;
  const void* bias;
  int64_t M;
  int act_dtype;     // storage dtype of x
  int out_dtype;     // storage dtype of y (MX tags resolved)
};

__device__ __forceinline__ float load_act(const void* x, int dt, int64_t i) {
  switch (dt) {
    case GLB_FP16: return __half2float(reinterpret_cast<const __half*>(x)[i]);
    case GLB_BF16: return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[i]);
    case GLB_FP32: return reinterpret_cast<const float*>(x)[i];
    case GLB_INT8: return (float)reinterpret_cast<const int8_t*>(x)[i];
    case GLB_FP8E4: return fp8e4_to_float(reinterpret_cast<const uint8_t*>(x)[i]);
    case GLB_FP8E5: return fp8e5_to_float(reinterpret_cast<const uint8_t*>(x)[i]);
    default: return 0.f;
  }
}

// value of a float after `.to(input_dtype)`, returned as float (exact embedding)
__device__ __forceinline__ float cast_to_act(float v, int dt) {
  switch (dt) {
    case GLB_FP16: return __half2float(__float2half_rn(v));
    case GLB_BF16: return __bfloat162float(__float2bfloat16_rn(v));
    case GLB_INT8: { float t = truncf(v); return fminf(fmaxf(t, -128.f), 127.f); }
    case GLB_FP8E4: return fp8e4_to_float(float_to_fp8e4(v));
    case GLB_FP8E5: return fp8e5_to_float(float_to_fp8e5(v));
    default: return v;
  }
}

__device__ __forceinline__ void store_out(void* y, int dt, int64_t i, float v) {
  switch (dt) {
    case GLB_FP16: reinterpret_cast<__half*>(y)[i] = __float2half_rn(v); break;
    case GLB_BF16: reinterpret_cast<__nv_bfloat16*>(y)[i] = __float2bfloat16_rn(v); break;
    case GLB_FP32: reinterpret_cast<float*>(y)[i] = v; break;
    case GLB_INT32: reinterpret_cast<int32_t*>(y)[i] = (int32_t)v; break;
    default: break;
  }
}
__device__ __forceinline__ float round_out(float v, int dt) {
  switch (dt) {
    case GLB_FP16: return __half2float(__float2half_rn(v));
    case GLB_BF16: return __bfloat162float(__float2bfloat16_rn(v));
    default: return v;
  }
}

// raw weight element (k, n): integer code (is_code) or a float value of an unquantised weight
__device__ __forceinline__ void load_raw(const glb_linear& L, int64_t k, int64_t n, int& code, float& fval, bool& is_code) {
  const int eps = L.elements_per_sample;
  const int64_t idx = (k / eps) * L.wq_stride0 + n * L.wq_stride1;
  is_code = true;
  fval = 0.f;
  uint64_t word = 0;
  switch (L.wq_dtype) {
    case GLB_UINT8: word = reinterpret_cast<const uint8_t*>(L.W_q)[idx]; break;
    case GLB_INT8:
      if (eps == 1) { code = reinterpret_cast<const int8_t*>(L.W_q)[idx]; return; }
      word = reinterpret_cast<const uint8_t*>(L.W_q)[idx]; break;
    case GLB_INT16: case GLB_UINT16: word = reinterpret_cast<const uint16_t*>(L.W_q)[idx]; break;
    case GLB_INT32: case GLB_UINT32: word = reinterpret_cast<const uint32_t*>(L.W_q)[idx]; break;
    case GLB_INT64: word = reinterpret_cast<const uint64_t*>(L.W_q)[idx]; break;
    case GLB_FP8E4: is_code = false; fval = fp8e4_to_float(reinterpret_cast<const uint8_t*>(L.W_q)[idx]); return;
    case GLB_FP8E5: is_code = false; fval = fp8e5_to_float(reinterpret_cast<const uint8_t*>(L.W_q)[idx]); return;
    case GLB_FP16: is_code = false; fval = __half2float(reinterpret_cast<const __half*>(L.W_q)[idx]); return;
    case GLB_BF16: is_code = false; fval = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(L.W_q)[idx]); return;
    case GLB_FP32: is_code = false; fval = reinterpret_cast<const float*>(L.W_q)[idx]; return;
    default: break;
  }
  if (eps > 1) {
    const int sh = (int)(k % eps) * L.W_nbits;
    code = (int)((word >> sh) & (uint64_t)(uint32_t)L.unpack_mask);
  } else {
    code = (int)word;
  }
}

template <typename MT>
__device__ __forceinline__ float dequant_one(const GenParams& P, int zscalar, int64_t k, int64_t n) {
  using M_ = Meta<MT>;
  const glb_linear& L = P.L;
  int code; float fval; bool is_code;
  load_raw(L, k, n, code, fval, is_code);
  const int mode = L.W_group_mode;
  if (mode == 0) return cast_to_act(is_code ? (float)code : fval, P.act_dtype);
  const int64_t g = (L.group_size >= L.K) ? 0 : (k / L.group_size);
  const bool zs = (L.zeros_numel == 1);
  MT b = is_code ? M_::from_int(code) : M_::from_float(fval);
  MT r;
  if (mode == 1) {
    MT z = zs ? M_::from_int(zscalar) : M_::load(L.zeros, g * L.z_stride0 + n * L.z_stride1);
    r = M_::sub(b, z);
  } else if (mode == 2) {
    r = M_::mul(b, M_::load(L.scales, g * L.s_stride0 + n * L.s_stride1));
  } else if (mode == 3) {
    MT s = M_::load(L.scales, g * L.s_stride0 + n * L.s_stride1);
    if (zs) r = M_::mul(M_::from_int(code - zscalar), s);
    else r = M_::mul(M_::sub(b, M_::load(L.zeros, g * L.z_stride0 + n * L.z_stride1)), s);
  } else {  // 4: fma(b, s, z')
    r = M_::fma(b, M_::load(L.scales, g * L.s_stride0 + n * L.s_stride1),
                M_::load(L.zeros, g * L.z_stride0 + n * L.z_stride1));
  }
  return cast_to_act(M_::to_float(r), P.act_dtype);
}

// integer-meta variant (no scales tensor: modes 0/1 with an integer zero-point)
__device__ __forceinline__ float dequant_one_int(const GenParams& P, int zscalar, int64_t k, int64_t n) {
  int code; float fval; bool is_code;
  load_raw(P.L, k, n, code, fval, is_code);
  float v = is_code ? (float)code : fval;
  if (P.L.W_group_mode == 1) v -= (float)zscalar;
  return cast_to_act(v, P.act_dtype);
}

// MX weights: e2m1 nibbles (2 per byte, even k low) or e4m3 bytes, times 2^(e8m0 - 127) per group
__device__ __forceinline__ float dequant_one_mx(const GenParams& P, int64_t k, int64_t n) {
  const glb_linear& L = P.L;
  float v;
  if (L.elements_per_sample == 2) {
    uint8_t byte = reinterpret_cast<const uint8_t*>(L.W_q)[(k >> 1) * L.wq_stride0 + n * L.wq_stride1];
    v = fp4_e2m1_to_float((k & 1) ? (byte >> 4) : (byte & 15));
  } else {
    v = fp8e4_to_float(reinterpret_cast<const uint8_t*>(L.W_q)[k * L.wq_stride0 + n * L.wq_stride1]);
  }
  const int64_t g = k / L.group_size;
  const uint8_t sb = reinterpret_cast<const uint8_t*>(L.scales)[n * L.s_stride0 + g * L.s_stride1];
  const float sc = (L.scales_dtype == GLB_FP8E4) ? fp8e4_to_float(sb) : e8m0_to_float(sb);
  return v * sc;
}

// MX / NV quantised activations (channel_scale_mode 4; tl.dot_scaled lhs semantics, gemm_kernels.py:508-524): e4m3
// bytes (MXFP8) or e2m1 nibbles (MXFP4 / NVFP4; even k = low nibble) times the scale of (row m, group k / gs):
// scales_x is (M padded, K/gs) e8m0 bytes, or e4m3 bytes for NVFP4.  fp32 holds the product exactly.
__device__ __forceinline__ float load_act_mx(const GenParams& P, int64_t m, int64_t k) {
  const glb_linear& L = P.L;
  const uint8_t* xb = reinterpret_cast<const uint8_t*>(P.x);
  float v;
  if (L.input_dtype == GLB_MXFP8) {
    v = fp8e4_to_float(xb[m * L.K + k]);
  } else {
    const uint8_t byte = xb[m * (L.K >> 1) + (k >> 1)];
    v = fp4_e2m1_to_float((k & 1) ? (byte >> 4) : (byte & 15));
  }
  const uint8_t sb = reinterpret_cast<const uint8_t*>(P.scales_x)[m * (L.K / L.group_size) + k / L.group_size];
  return v * ((L.input_dtype == GLB_NVFP4) ? fp8e4_to_float(sb) : e8m0_to_float(sb));
}

template <typename MT, int KIND /*0 float-meta, 1 int-meta, 2 mx*/, bool INT_ACC>
__global__ void __launch_bounds__(GEN_COLS) generic_kernel(const GenParams P) {
  __shared__ float xs[GEN_ROWS][GEN_KC];
  const glb_linear& L = P.L;
  const int64_t n = (int64_t)blockIdx.x * GEN_COLS + threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * GEN_ROWS;
  const int rows = (int)min((int64_t)GEN_ROWS, P.M - m0);
  const bool active = n < L.N;
  // scalar zero-point (a 0-d int32 tensor in the reference, core.py:470-471) is read here, not on the host
  const int zscalar = (L.zeros_numel == 1) ? reinterpret_cast<const int32_t*>(L.zeros)[0] : 0;

  float accf[GEN_ROWS];
  int acci[GEN_ROWS];
#pragma unroll
  for (int r = 0; r < GEN_ROWS; ++r) { accf[r] = 0.f; acci[r] = 0; }

  for (int64_t k0 = 0; k0 < L.K; k0 += GEN_KC) {
    const int kc = (int)min((int64_t)GEN_KC, L.K - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < GEN_ROWS * GEN_KC; i += GEN_COLS) {
      const int r = i / GEN_KC, c = i % GEN_KC;
      float xv = 0.f;
      if (r < rows && c < kc) {
        if (KIND == 2 && L.channel_scale_mode == 4) xv = load_act_mx(P, m0 + r, k0 + c);
        else xv = load_act(P.x, P.act_dtype, (m0 + r) * L.K + k0 + c);
      }
      xs[r][c] = xv;
    }
    __syncthreads();
    if (active) {
      for (int c = 0; c < kc; ++c) {
        float w;
        if (KIND == 0) w = dequant_one<MT>(P, zscalar, k0 + c, n);
        else if (KIND == 1) w = dequant_one_int(P, zscalar, k0 + c, n);
        else w = dequant_one_mx(P, k0 + c, n);
        if (INT_ACC) {
          const int wi = (int)w;
#pragma unroll
          for (int r = 0; r < GEN_ROWS; ++r) acci[r] += (int)xs[r][c] * wi;
        } else {
#pragma unroll
          for (int r = 0; r < GEN_ROWS; ++r) accf[r] = fmaf(xs[r][c], w, accf[r]);
        }
      }
    }
  }
  if (!active) return;

  // ---- epilogue (gemm_kernels.py:392-406), then bias (core.py:192-193) ----
  using M_ = Meta<MT>;
  const int csm = L.channel_scale_mode;
  for (int r = 0; r < rows; ++r) {
    float acc = INT_ACC ? (float)acci[r] : accf[r];   // cvt.rn.f32.s32 for the int32 accumulator
    float out;
    if (KIND == 2) {
      if (L.group_size == 16) acc *= 0.0025f;          // NVFP4 meta scale 0.05^2 (gemm_kernels.py:461,530)
      out = acc;
      // MXFP8 activations quantised per token: acc.to(out dtype) * scales_a in fp32 (gemm_kernels.py:533-537)
      if (csm == 2 && P.scales_x) out = __fmul_rn(round_out(acc, P.out_dtype), P.scales_x[m0 + r]);
    } else if (csm == 0) {
      out = acc;
    } else if (csm == 1) {
      MT a_m = M_::from_float(acc);
      out = M_::to_float(M_::mul(a_m, M_::load(L.scales, n * L.s_stride1)));
    } else {
      const float a_m = M_::to_float(M_::from_float(acc));
      const float sa = P.scales_x ? P.scales_x[m0 + r] : 1.f;
      const float sb = (csm == 3) ? M_::to_float(M_::load(L.scales, n * L.s_stride1)) : 1.f;
      out = __fmul_rn(a_m, __fmul_rn(sa, sb));
    }
    const int64_t oi = (m0 + r) * L.N + n;
    if (INT_ACC && csm == 0 && KIND != 2 && P.out_dtype == GLB_INT32 && !P.bias) {
      reinterpret_cast<int32_t*>(P.y)[oi] = acci[r];   // raw accumulator, exact
      continue;
    }
    if (P.bias) {
      float b;
      switch (P.out_dtype) {
        case GLB_FP16: b = __half2float(reinterpret_cast<const __half*>(P.bias)[n]); break;
        case GLB_BF16: b = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.bias)[n]); break;
        default: b = reinterpret_cast<const float*>(P.bias)[n]; break;
      }
      out = __fadd_rn(round_out(out, P.out_dtype), b);
    }
    store_out(P.y, P.out_dtype, oi, out);
  }
}

}  // namespace

int launch_generic(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M,
                   const void* bias, cudaStream_t stream) {
  GenParams P;
  P.L = L; P.x = x; P.scales_x = scales_x; P.y = y; P.bias = bias; P.M = M;
  P.act_dtype = act_storage_dtype(L.input_dtype);
  P.out_dtype = act_storage_dtype(L.output_dtype);
  if (L.zeros_numel == 1)
    GLB_REQUIRE(L.zeros_dtype == GLB_INT32, GLB_ERR_UNSUPPORTED, "scalar zero-point must be int32");
  const bool mx = is_mx(L.input_dtype);
  const bool int_acc = (P.act_dtype == GLB_INT8);
  dim3 grid((unsigned)((L.N + GEN_COLS - 1) / GEN_COLS), (unsigned)((M + GEN_ROWS - 1) / GEN_ROWS));
  GLB_REQUIRE(grid.y <= 65535u, GLB_ERR_UNSUPPORTED, "generic kernel: M too large (%lld)", (long long)M);

  const int md = L.meta_dtype;
  const bool float_meta = (md == GLB_FP16 || md == GLB_BF16 || md == GLB_FP32);
  set_kernel_name("generic_cuda_core");
#define GLB_LAUNCH_GEN(MT, KIND, IA) \
  generic_kernel<MT, KIND, IA><<<grid, GEN_COLS, 0, stream>>>(P)
  if (mx) {
    GLB_LAUNCH_GEN(float, 2, false);
  } else if (!float_meta) {
    if (int_acc) GLB_LAUNCH_GEN(float, 1, true); else GLB_LAUNCH_GEN(float, 1, false);
  } else if (md == GLB_FP16) {
    if (int_acc) GLB_LAUNCH_GEN(__half, 0, true); else GLB_LAUNCH_GEN(__half, 0, false);
  } else if (md == GLB_BF16) {
    if (int_acc) GLB_LAUNCH_GEN(__nv_bfloat16, 0, true); else GLB_LAUNCH_GEN(__nv_bfloat16, 0, false);
  } else {
    if (int_acc) GLB_LAUNCH_GEN(float, 0, true); else GLB_LAUNCH_GEN(float, 0, false);
  }
#undef GLB_LAUNCH_GEN
  count_launch();
  GLB_CHECK_CUDA(cudaGetLastError());
  return GLB_OK;
}

}  // namespace glb
