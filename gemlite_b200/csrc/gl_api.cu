// C-ABI entry points + kernel-family dispatch (see include/gemlite_b200.h).
#include <stdarg.h>
#include <string.h>

#include <cuda.h>
#include <mutex>

#include <string.h>
#include <mutex>
#include <map>
#include <unordered_map>
#include <utility>

#include "gl_common.cuh"

namespace glb {

static thread_local char tl_error[512] = "";
static thread_local const char* tl_kernel = "";
static std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_pdl_enabled{1};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
}
void set_kernel_name(const char* name) { tl_kernel = name; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- TMA descriptors: cuTensorMapEncodeTiled is fetched from the driver at run time (no libcuda link) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Encoded tensor maps are cached: a decode call needs three (weights, scales, zeros) and cuTensorMapEncodeTiled costs
// ~1 us each on the host -- more than the kernel's own launch.  Key = every argument; bounded, cleared when full.
namespace {
struct TmapKey {
  uint64_t v[8];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 8; ++i) { h ^= k.v[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
struct TmapVal { alignas(64) unsigned char b[128]; };
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, TmapVal, TmapKeyHash> g_tmap_cache;
}  // namespace

int make_tmap_2d(void* out_tmap, int dt, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer, int swizzle128) {
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
  int dev = 0;
  cudaGetDevice(&dev);
  TmapKey key = {{(uint64_t)dt | ((uint64_t)dev << 32), (uint64_t)(uintptr_t)base, inner, outer, row_stride_bytes, box_inner,
                  box_outer, (uint64_t)swizzle128}};
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      memcpy(out_tmap, it->second.b, 128);
      return GLB_OK;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  GLB_REQUIRE(fn != nullptr, GLB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMapDataType cdt;
  switch (dt) {
    case GLB_FP16: cdt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; break;
    case GLB_BF16: cdt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; break;
    case GLB_INT32: cdt = CU_TENSOR_MAP_DATA_TYPE_INT32; break;
    case GLB_FP32: cdt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
    default: cdt = CU_TENSOR_MAP_DATA_TYPE_UINT8; break;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_tmap), cdt, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GLB_REQUIRE(r == CUDA_SUCCESS, GLB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();
    TmapVal v;
    memcpy(v.b, out_tmap, 128);
    g_tmap_cache[key] = v;
  }
  return GLB_OK;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to ONE device: a process that drives several GPUs from one
// thread must opt in on each of them.  Remembered per (kernel, device); the per-thread one-entry cache keeps the mutex off
// the path of a decode loop that launches the same kernel over and over.
int ensure_dyn_smem(const void* func, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  static thread_local const void* last_func = nullptr;
  static thread_local int last_dev = -1;
  static thread_local size_t last_bytes = 0;
  if (func == last_func && dev == last_dev && bytes <= last_bytes) return GLB_OK;
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;      // (kernel, device) -> bytes configured
  const std::pair<const void*, int> key(func, dev);
  size_t have = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = done.find(key);
    if (it != done.end()) have = it->second;
  }
  if (have < bytes) {
    GLB_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    std::lock_guard<std::mutex> lk(mu);
    done[key] = bytes;
    have = bytes;
  }
  last_func = func; last_dev = dev; last_bytes = have;
  return GLB_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int validate(const glb_linear* L, int64_t M) {
  GLB_REQUIRE(L != nullptr, GLB_ERR_INVALID, "layer descriptor is NULL");
  GLB_REQUIRE(L->W_q != nullptr && L->K > 0 && L->N > 0, GLB_ERR_INVALID, "layer has no weights");
  GLB_REQUIRE(M > 0, GLB_ERR_INVALID, "M must be positive (got %lld)", (long long)M);
  GLB_REQUIRE(L->elements_per_sample >= 1, GLB_ERR_INVALID, "elements_per_sample must be >= 1");
  GLB_REQUIRE(L->wq_rows * L->elements_per_sample == L->K && L->wq_cols == L->N, GLB_ERR_INVALID,
              "W_q shape (%lld, %lld) does not match K=%lld N=%lld eps=%d", (long long)L->wq_rows,
              (long long)L->wq_cols, (long long)L->K, (long long)L->N, L->elements_per_sample);
  GLB_REQUIRE(L->W_group_mode >= 0 && L->W_group_mode <= 4, GLB_ERR_INVALID, "bad W_group_mode %d", L->W_group_mode);
  GLB_REQUIRE(L->channel_scale_mode >= 0 && L->channel_scale_mode <= 4, GLB_ERR_INVALID, "bad channel_scale_mode %d",
              L->channel_scale_mode);
  if (!is_mx(L->input_dtype)) {
    GLB_REQUIRE(L->channel_scale_mode != 4, GLB_ERR_UNSUPPORTED, "channel_scale_mode 4 needs an MX input dtype");
    if (L->W_group_mode >= 2) GLB_REQUIRE(L->scales_numel > 0, GLB_ERR_INVALID, "W_group_mode %d needs scales", L->W_group_mode);
    if (L->W_group_mode == 1 || L->W_group_mode >= 3)
      GLB_REQUIRE(L->zeros_numel > 0, GLB_ERR_INVALID, "W_group_mode %d needs zeros", L->W_group_mode);
    if (L->channel_scale_mode == 1 || L->channel_scale_mode == 3)
      GLB_REQUIRE(L->scales_numel == L->N, GLB_ERR_INVALID, "channel-wise scaling needs N scales");
  } else {
    GLB_REQUIRE(L->scales_numel > 0 && L->group_size > 0, GLB_ERR_INVALID, "MX layer needs group scales");
    // weight-only MX layers (fp16 / bf16 activations): mode 0.  Dynamically quantised activations (helper.py:660-945):
    // MXFP8 with per-token post-scale (2) or e8m0 micro-scales (4); MXFP4 / NVFP4 with micro-scales (4)
    const int csm = L->channel_scale_mode;
    const bool dyn = (L->input_dtype == GLB_MXFP8 || L->input_dtype == GLB_MXFP4 || L->input_dtype == GLB_NVFP4);
    GLB_REQUIRE(csm == 0 || (dyn && csm == 4) || (L->input_dtype == GLB_MXFP8 && csm == 2), GLB_ERR_UNSUPPORTED,
                "MX layer: channel_scale_mode %d is not valid for input dtype %d", csm, L->input_dtype);
    if (dyn) GLB_REQUIRE(L->group_size == ((L->input_dtype == GLB_NVFP4) ? 16 : 32), GLB_ERR_UNSUPPORTED,
                         "MX layer: group_size %d does not match input dtype %d", L->group_size, L->input_dtype);
  }
  return GLB_OK;
}

// MXFP dynamic layers with MXFP4 weights and MX micro-scaled activations (A8W4_MXFP_dynamic post_scale=False,
// A4W4_MXFP_dynamic; helper.py:732-893): the quantised activations are expanded to their exact bf16 values
// (gl_actquant.cu) into the workspace and the layer then runs as the MXFP4 weight-only layer it also is -- on the
// tcgen05 / streaming MXFP4 kernels instead of the CUDA-core general kernel.  Same products as tl.dot_scaled, fp32
// accumulation.  (Native block-scaled tcgen05.mma would skip the expansion; not built yet.)
static bool mx_expand_view(const glb_linear& L, glb_linear* as_weight_only);
// Native block-scaled tcgen05 kernel (gl_gemm_bs.cu) for the dynamic MX / NV layers it takes -- except at M <= 4 where an
// MXFP4 layer has the streaming kernel behind the expansion (weights read once at HBM speed, no 256-token tile to fill)
static bool use_block_scaled(const glb_linear& L, int64_t M) {
  if (!gemm_bs_supported(L, M)) return false;
  glb_linear W;
  if (mx_expand_view(L, &W) && gemv_mx_supported(W, M)) return false;
  return true;
}
static bool mx_expand_view(const glb_linear& L, glb_linear* as_weight_only) {
  if (!(L.input_dtype == GLB_MXFP8 || L.input_dtype == GLB_MXFP4) || L.channel_scale_mode != 4) return false;
  if (L.output_dtype != GLB_BF16) return false;
  if (L.W_nbits != 4 || L.elements_per_sample != 2 || L.group_size != 32) return false;
  glb_linear W = L;
  W.input_dtype = GLB_MXBF16; W.output_dtype = GLB_MXBF16; W.scaled_activations = 0; W.channel_scale_mode = 0; W.W_group_mode = 2;
  if (as_weight_only) *as_weight_only = W;
  return true;
}
// the expanded activations live BEHIND whatever the inner launch uses of the (shared, per-stream) workspace, and never inside
// its first 64 KB: those are the split-K arrival counters of every layer on the stream and must stay zero between launches
static size_t expand_offset(size_t inner_ws) { const size_t a = (inner_ws + 255) & ~(size_t)255; return a < 65536 ? 65536 : a; }

// reference dispatch: core.py:100-114
static int pick_matmul_type(const glb_linear& L, int64_t M) {
  if (M > 64) return GLB_MM_GEMM;
  if (M > 1) return GLB_MM_GEMM_SPLITK;
  if (is_mx(L.input_dtype)) return GLB_MM_GEMM_SPLITK;
  return (L.W_nbits < 8) ? GLB_MM_GEMV_REVSPLITK : GLB_MM_GEMV_SPLITK;
}

}  // namespace glb

using namespace glb;

extern "C" int glb_version(void) { return GLB_ABI_VERSION; }
extern "C" const char* glb_last_error(void) { return tl_error; }
extern "C" const char* glb_last_kernel(void) { return tl_kernel; }
extern "C" uint64_t glb_launch_count(void) { return g_launches.load(); }
extern "C" void glb_set_pdl(int enabled) { g_pdl_enabled.store(enabled ? 1 : 0); }
extern "C" int glb_get_pdl(void) { return g_pdl_enabled.load(); }
extern "C" int glb_debug_timeline(unsigned long long* out, int max_words) { return debug_read_timeline(out, max_words); }

extern "C" int glb_matmul_bcast(const glb_linear* L, const void* x, const void* scales_x, void* const* y_dst, int n_dst, int64_t col0,
                                int64_t M, const void* bias, void* stream_) {
  int rc = validate(L, M);
  if (rc != GLB_OK) return rc;
  GLB_REQUIRE(x && y_dst && n_dst >= 1 && n_dst <= 8, GLB_ERR_INVALID, "glb_matmul_bcast: bad arguments");
  for (int d = 0; d < n_dst; ++d) GLB_REQUIRE(y_dst[d] != nullptr, GLB_ERR_INVALID, "glb_matmul_bcast: destination %d is NULL", d);
  GLB_REQUIRE(gemv_bcast_supported(*L, M), GLB_ERR_UNSUPPORTED,
              "glb_matmul_bcast: only A16W4 decode layers at M == 1 store to several destinations (use glb_matmul + a collective)");
  return launch_gemv_bcast(*L, x, reinterpret_cast<const float*>(scales_x), y_dst[0], y_dst, n_dst, (long long)col0, M, bias, nullptr, 0,
                           (cudaStream_t)stream_);
}

extern "C" unsigned long long glb_stream_capture_id(void* stream_) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  unsigned long long id = 0;
  if (cudaStreamGetCaptureInfo((cudaStream_t)stream_, &st, &id) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return (st == cudaStreamCaptureStatusActive) ? id : 0;
}

extern "C" size_t glb_workspace_bytes(const glb_linear* L, int64_t M, int matmul_type) {
  if (!L || M <= 0) return 0;
  if (use_block_scaled(*L, M)) return 0;
  {
    glb_linear W;
    if (mx_expand_view(*L, &W) && (gemm_mx_tc2_supported(W, M) || gemv_mx_supported(W, M)))
      return expand_offset(gemm_tc_workspace(W, M)) + (size_t)M * (size_t)L->K * 2;
  }
  const bool was_auto = (matmul_type == GLB_MM_AUTO);
  if (was_auto) matmul_type = pick_matmul_type(*L, M);
  // same routing as glb_matmul: the decode kernel first (its chained hand-off keeps arrival counters in the workspace)
  const bool gemv_name = (matmul_type <= GLB_MM_GEMV_SPLITK);
  const bool can_gemv = gemv_supported(*L, M), can_tc = gemm_tc_supported(*L, M);
  if (can_gemv && (gemv_name || (was_auto && M <= 16) || !can_tc)) return gemv_workspace(*L, M);
  if (matmul_type != GLB_MM_GEMM && gemv_packed_supported(*L, M)) return 0;
  if ((was_auto || matmul_type != GLB_MM_GEMM) && (gemm_mid_supported(*L, M) || gemm_w8_mid_supported(*L, M))) return 0;
  if (can_tc) return gemm_tc_workspace(*L, M);
  return 0;
}

extern "C" int glb_matmul(const glb_linear* L, const void* x, const void* scales_x, void* y, int64_t M, int matmul_type,
                          void* workspace, size_t workspace_bytes, const void* bias, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = validate(L, M);
  if (rc != GLB_OK) return rc;
  GLB_REQUIRE(x && y, GLB_ERR_INVALID, "x / y is NULL");
  GLB_REQUIRE(matmul_type >= GLB_MM_AUTO && matmul_type <= GLB_MM_GEMM, GLB_ERR_INVALID, "bad matmul_type %d", matmul_type);
  if (L->channel_scale_mode >= 2)
    GLB_REQUIRE(scales_x != nullptr, GLB_ERR_INVALID, "channel_scale_mode %d needs activation scales", L->channel_scale_mode);
  if (use_block_scaled(*L, M)) return launch_gemm_bs(*L, x, scales_x, y, M, bias, stream);
  {
    glb_linear W;
    if (mx_expand_view(*L, &W) && (gemm_mx_tc2_supported(W, M) || gemv_mx_supported(W, M))) {
      const size_t inner_ws = gemm_tc_workspace(W, M), inner = expand_offset(inner_ws), need = inner + (size_t)M * (size_t)L->K * 2;
      GLB_REQUIRE(workspace && workspace_bytes >= need, GLB_ERR_WORKSPACE, "MX dynamic layer: workspace too small: need %zu bytes, got %zu",
                  need, workspace_bytes);
      void* xb = reinterpret_cast<unsigned char*>(workspace) + inner;
      rc = launch_expand_act_mx(x, scales_x, xb, M, L->K, L->input_dtype, stream);
      if (rc != GLB_OK) return rc;
      if (gemv_mx_supported(W, M)) return launch_gemv_mx(W, xb, y, M, bias, stream);
      return launch_gemm_tc(W, xb, nullptr, y, M, workspace, inner_ws, bias, stream);
    }
  }
  const bool was_auto = (matmul_type == GLB_MM_AUTO);
  if (was_auto) matmul_type = pick_matmul_type(*L, M);
  const bool gemv_name = (matmul_type <= GLB_MM_GEMV_SPLITK);
  // the GEMV names are M == 1 kernels in the reference (tests skip them for M > 1)
  GLB_REQUIRE(!(gemv_name && M > 16), GLB_ERR_UNSUPPORTED, "GEMV kernels take at most 16 rows (got %lld)", (long long)M);
  const float* sx = reinterpret_cast<const float*>(scales_x);
  const bool can_gemv = gemv_supported(*L, M), can_tc = gemm_tc_supported(*L, M);

  // 1. decode kernel (M <= 16, packed 4-bit): the automatic choice for small batches, and what the
  //    GEMV names select; 2. tcgen05 kernel: the automatic choice above that, and what the GEMM
  //    names select whenever the layer qualifies.
  // 0. MXFP4 weight-only layers at M <= 4: streaming CUDA-core kernel (the reference sends every MX dtype at M == 1 to
  //    GEMM_SPLITK, core.py:100-114); an explicit "GEMM" request keeps the tensor-core kernel
  if (matmul_type != GLB_MM_GEMM && gemv_mx_supported(*L, M)) return launch_gemv_mx(*L, x, y, M, bias, stream);
  // 0a. A8W8 at 16 < M <= 128: clustered streaming kernel (gl_gemm_a8_mid.cu); "GEMM" by name keeps the tcgen05 tile
  if ((was_auto || matmul_type != GLB_MM_GEMM) && gemm_a8_mid_supported(*L, M))
    return launch_gemm_a8_mid(*L, x, sx, y, M, bias, stream);
  // 0d. A16W8 weight-only at 16 < M <= 128: clustered streaming kernel with in-register widening (gl_gemm_a8_mid.cu)
  if ((was_auto || matmul_type != GLB_MM_GEMM) && gemm_w8_mid_supported(*L, M)) return launch_gemm_w8_mid(*L, x, y, M, bias, stream);
  // 0c. A16W8 weight-only (fp16 x int8 / e4m3) at M <= 16: streaming kernel with in-register widening (gl_gemv_a8.cu)
  if (matmul_type != GLB_MM_GEMM && gemv_a16w8_supported(*L, M)) return launch_gemv_a16w8(*L, x, y, M, bias, stream);
  // 0b. A8W8 at M <= 16: 16-feature-row streaming kernel (gl_gemv_a8.cu) (validated in round 2)
  if (matmul_type != GLB_MM_GEMM && gemv_a8_enabled() && gemv_a8_supported(*L, M))
    return launch_gemv_a8(*L, x, sx, y, M, bias, stream);
  const bool prefer_gemv = can_gemv && (gemv_name || (was_auto && M <= 16) || !can_tc);
  if (prefer_gemv) return launch_gemv(*L, x, sx, y, M, bias, workspace, workspace_bytes, stream);
  // 1a. packed 8 / 2 / 1-bit grouped layers at M <= 8: streaming CUDA-core kernel (gl_gemv_packed.cu); "GEMM" by name keeps tcgen05
  if (matmul_type != GLB_MM_GEMM && gemv_packed_supported(*L, M)) return launch_gemv_packed(*L, x, y, M, bias, stream);
  // 1b. batched decode, 16 < M <= 128 (the reference's GEMM_SPLITK range and the bottom of its GEMM range): streaming
  //     mma.sync kernel (gl_gemm_mid.cu); an explicit "GEMM" request keeps the tcgen05 kernel
  if ((was_auto || matmul_type != GLB_MM_GEMM) && gemm_mid_supported(*L, M)) return launch_gemm_mid(*L, x, y, M, bias, stream);
  if (can_tc) {
    const size_t need = gemm_tc_workspace(*L, M);
    GLB_REQUIRE(need == 0 || (workspace && workspace_bytes >= need), GLB_ERR_WORKSPACE,
                "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return launch_gemm_tc(*L, x, sx, y, M, workspace, workspace_bytes, bias, stream);
  }
  // 3. general-case CUDA-core kernel
  return launch_generic(*L, x, sx, y, M, bias, stream);
}
