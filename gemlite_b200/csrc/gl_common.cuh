// Shared device/host helpers for the gemlite_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/gemlite_b200.h"

namespace glb {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
void set_kernel_name(const char* name);
void count_launch();
extern std::atomic<int> g_pdl_enabled;

#define GLB_CHECK_CUDA(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      glb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return GLB_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define GLB_REQUIRE(cond, code, ...)    \
  do {                                  \
    if (!(cond)) {                      \
      glb::set_error(__VA_ARGS__);      \
      return (code);                    \
    }                                   \
  } while (0)

inline bool is_mx(int dt) { return dt >= GLB_MXFP16 && dt <= GLB_NVFP4; }

// storage dtype an `input_dtype` tag implies for the activation pointer
inline int act_storage_dtype(int input_dtype) {
  switch (input_dtype) {
    case GLB_MXFP16: return GLB_FP16;
    case GLB_MXBF16: return GLB_BF16;
    case GLB_MXFP8: return GLB_FP8E4;
    default: return input_dtype;
  }
}
inline int dtype_bytes(int dt) {
  switch (dt) {
    case GLB_FP32: case GLB_INT32: case GLB_UINT32: return 4;
    case GLB_FP16: case GLB_BF16: case GLB_INT16: case GLB_UINT16: return 2;
    case GLB_INT64: return 8;
    default: return 1;
  }
}

// host helpers shared by the TMA-based kernels (defined in gl_api.cu)
int make_tmap_2d(void* out_tmap /*CUtensorMap*/, int glb_elem_dtype, const void* base, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle128);
int num_sms();
// opt the kernel `func` in to `bytes` of dynamic shared memory on the CURRENT device (remembered per kernel and device)
int ensure_dyn_smem(const void* func, size_t bytes);
#define GLB_ENSURE_SMEM(kern, bytes)                                                       \
  do {                                                                                     \
    const int _rc = glb::ensure_dyn_smem(reinterpret_cast<const void*>(kern), (size_t)(bytes)); \
    if (_rc != GLB_OK) return _rc;                                                         \
  } while (0)

// sub-kernel entry points (one translation unit each)
int launch_generic(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M,
                   const void* bias, cudaStream_t stream);
bool gemv_supported(const glb_linear& L, int64_t M);
int launch_gemv(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M,
                const void* bias, void* workspace, size_t ws_bytes, cudaStream_t stream);
size_t gemv_workspace(const glb_linear& L, int64_t M);
bool gemv_bcast_supported(const glb_linear& L, int64_t M);
int launch_gemv_bcast(const glb_linear& L, const void* x, const float* scales_x, void* y, void* const* y_dst, int n_dst,
                      long long y_col0, int64_t M, const void* bias, void* workspace, size_t ws_bytes, cudaStream_t stream);
int debug_read_timeline(unsigned long long* out, int max_words);
bool gemm_tc_supported(const glb_linear& L, int64_t M);
size_t gemm_tc_workspace(const glb_linear& L, int64_t M);
int launch_gemm_tc(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M,
                   void* workspace, size_t ws_bytes, const void* bias, cudaStream_t stream);
bool gemm_tc2_supported(const glb_linear& L, int64_t M);
bool gemm_mx_tc2_supported(const glb_linear& L, int64_t M);
bool gemm_w8_tc2_supported(const glb_linear& L, int64_t M);
bool gemm_w8_mid_supported(const glb_linear& L, int64_t M);
bool gemv_packed_supported(const glb_linear& L, int64_t M);
int launch_gemv_packed(const glb_linear& L, const void* x, void* y, int64_t M, const void* bias, cudaStream_t stream);
bool gemm_bs_supported(const glb_linear& L, int64_t M);
int launch_gemm_bs(const glb_linear& L, const void* x, const void* scales_x, void* y, int64_t M, const void* bias, cudaStream_t stream);
int launch_gemm_w8_mid(const glb_linear& L, const void* x, void* y, int64_t M, const void* bias, cudaStream_t stream);
bool gemv_mx_supported(const glb_linear& L, int64_t M);
int launch_gemv_mx(const glb_linear& L, const void* x, void* y, int64_t M, const void* bias, cudaStream_t stream);
bool gemm_a8_tc2_supported(const glb_linear& L, int64_t M);
int launch_gemm_a8_tc2(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M, const void* bias,
                       cudaStream_t stream);
bool gemv_a8_enabled();
bool gemv_a16w8_supported(const glb_linear& L, int64_t M);
int launch_gemv_a16w8(const glb_linear& L, const void* x, void* y, int64_t M, const void* bias, cudaStream_t stream);
bool gemv_a8_supported(const glb_linear& L, int64_t M);
int launch_gemv_a8(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M, const void* bias,
                   cudaStream_t stream);
size_t gemm_tc2_workspace(const glb_linear& L, int64_t M);
int launch_expand_act_mx(const void* q, const void* scales, void* out_bf16, int64_t M, int64_t K, int mx_dtype, cudaStream_t stream);
bool gemm_mid_supported(const glb_linear& L, int64_t M);
bool gemm_a8_mid_supported(const glb_linear& L, int64_t M);
int launch_gemm_a8_mid(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M, const void* bias,
                       cudaStream_t stream);
int launch_gemm_mid(const glb_linear& L, const void* x, void* y, int64_t M, const void* bias, cudaStream_t stream);
int launch_gemm_tc2(const glb_linear& L, const void* x, void* y, int64_t M, void* workspace, size_t ws_bytes, const void* bias,
                    cudaStream_t stream);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ float e8m0_to_float(uint32_t e) {
  // 2^(e-127); e == 0 is 2^-127 (an fp32 subnormal), e == 255 is NaN
  uint32_t bits = (e == 0) ? 0x00400000u : ((e == 255) ? 0x7fc00000u : (e << 23));
  return __uint_as_float(bits);
}

__device__ __forceinline__ float fp4_e2m1_to_float(uint32_t code) {
  // index = bit pattern, sign = bit 3 (reference table quant_utils.py:31-37)
  uint32_t m = code & 7u;
  // magnitudes 0, .5, 1, 1.5, 2, 3, 4, 6
  float v = (m < 2u) ? 0.5f * (float)m : __uint_as_float(((126u + (m >> 1)) << 23) | ((m & 1u) << 22));
  return (code & 8u) ? -v : v;
}

__device__ __forceinline__ float fp8e4_to_float(uint8_t v) {
  __half_raw hr = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)v, __NV_E4M3);
  return __half2float(__half(hr));
}
__device__ __forceinline__ float fp8e5_to_float(uint8_t v) {
  __half_raw hr = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)v, __NV_E5M2);
  return __half2float(__half(hr));
}
__device__ __forceinline__ uint8_t float_to_fp8e4(float f) {
  return (uint8_t)__nv_cvt_float_to_fp8(f, __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ uint8_t float_to_fp8e5(float f) {
  return (uint8_t)__nv_cvt_float_to_fp8(f, __NV_SATFINITE, __NV_E5M2);
}

// fp32 division as Triton emits it for `/` (div.full.f32, <= 2 ulp) -- keeps the activation
// quantiser bit-compatible with the reference kernel on the same hardware.
__device__ __forceinline__ float div_full(float a, float b) {
  float r;
  asm("div.full.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// programmatic dependent launch (no-ops when the kernel was not launched with the attribute)
__device__ __forceinline__ void pdl_wait_prior_grids() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

#endif  // __CUDACC__

}  // namespace glb
