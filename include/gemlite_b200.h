/* gemlite_b200.h -- C ABI of the B200-native low-bit fused dequant + matmul library.
 *
 * This is the drop-in boundary.  The reference (mobiusml/gemlite) is a Python package whose hot
 * path ends in Triton launches; each entry point below replaces one of those launch sites and is
 * what a maintainer would bind (ctypes stub shown in INTEGRATION.md):
 *
 *   glb_pack_cols / glb_unpack_cols ... gemlite/bitpack.py:112-144 (pack_weights_over_cols_triton),
 *                                       :146-205 (unpack_over_cols)          [called from core.py:384-392]
 *   glb_quant_act_per_token ........... gemlite/quant_utils.py:308-347 (scale_activations_per_token_triton)
 *                                                                           [called from core.py:162-163]
 *   glb_quant_act_mx .................. gemlite/quant_utils.py:546-591, :819-857, :917-954 (scale_activations_mxfp8 /
 *                                       _mxfp4 / _nvfp4, the *_triton_v2 kernels)  [called from core.py:165-175]
 *   glb_matmul ........................ the `.forward(...)` of the five kernel classes
 *                                       gemv_kernels.py:553, gemv_revsplitK_kernels.py:391,
 *                                       gemv_splitK_kernels.py:423, gemm_splitK_kernels.py:595,
 *                                       gemm_kernels.py:549                 [called from core.py:184-190]
 *                                       plus the bias add of core.py:192-193 (fused).
 *   glb_workspace_bytes ............... replaces the torch.zeros / fp32 staging buffers those
 *                                       forwards allocate (e.g. gemv_revsplitK_kernels.py:402-422).
 *
 * Conventions: plain pointers and sizes only (no torch types); all pointers are DEVICE pointers
 * borrowed for the duration of the call unless stated; nothing is allocated, nothing synchronises;
 * every launch goes to `stream` (a cudaStream_t passed as void*), so calls are CUDA-graph
 * capturable and re-entrant across streams.  Return value: 0 on success, negative glb_status on
 * failure (never throws); glb_last_error() returns a thread-local description.
 */
#ifndef GEMLITE_B200_H
#define GEMLITE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLB_ABI_VERSION 1

/* Only the entry points below are exported (the library is built with -fvisibility=hidden). */
#if defined(__GNUC__)
#define GLB_API __attribute__((visibility("default")))
#else
#define GLB_API
#endif

typedef enum glb_status {
  GLB_OK = 0,
  GLB_ERR_INVALID = -1,      /* bad argument / shape / alignment */
  GLB_ERR_UNSUPPORTED = -2,  /* combination the library does not implement */
  GLB_ERR_CUDA = -3,         /* a CUDA runtime/driver call failed (message has the code) */
  GLB_ERR_WORKSPACE = -4     /* workspace missing or too small */
} glb_status;

/* Numeric tags: identical to the reference's DType enum values (gemlite/dtypes.py:8-29),
 * because they travel inside `meta_args` / the serialized `metadata` tensor. */
typedef enum glb_dtype {
  GLB_FP32 = 0, GLB_FP16 = 1, GLB_BF16 = 2, GLB_FP8E4 = 3, GLB_INT8 = 4, GLB_UINT8 = 5,
  GLB_INT32 = 6, GLB_UINT32 = 7, GLB_FP8E5 = 8, GLB_INT16 = 9, GLB_UINT16 = 10, GLB_INT64 = 11,
  GLB_MXFP16 = 14, GLB_MXBF16 = 15, GLB_MXFP8 = 16, GLB_MXFP4 = 17, GLB_NVFP4 = 18, GLB_E8M0 = 19
} glb_dtype;

/* Kernel selector: index into the reference's GEMLITE_MATMUL_TYPES (gemlite/core.py:56-66).
 * The names are kept for forward_manual(); each selects the B200 kernel family that covers the
 * same batch regime.  GLB_MM_AUTO applies the reference's M-based dispatch (core.py:100-114). */
typedef enum glb_matmul_type {
  GLB_MM_AUTO = -1,
  GLB_MM_GEMV = 0,
  GLB_MM_GEMV_REVSPLITK = 1,
  GLB_MM_GEMV_SPLITK = 2,
  GLB_MM_GEMM_SPLITK = 3,
  GLB_MM_GEMM = 4
} glb_matmul_type;

/* POD mirror of GemLiteLinear.get_tensor_args() + get_meta_args() (gemlite/core.py:522-538).
 * Strides are in ELEMENTS of the respective tensor dtype. */
typedef struct glb_linear {
  /* W_q: packed words, logical shape (K / elements_per_sample, N)  [core.py:387-392]
   *      or raw 8/16/32-bit weights, logical shape (K, N) view of an (N, K) tensor [core.py:377] */
  const void* W_q;
  int64_t wq_rows, wq_cols;
  int64_t wq_stride0, wq_stride1;
  int32_t wq_dtype;                 /* glb_dtype of the storage word */
  /* scales: (K/group_size, N) [core.py:420]; MX: uint8 e8m0 viewed (N, K/group_size) [core.py:490-497];
   *         numel == 0 when absent */
  const void* scales;
  int64_t scales_numel;
  int64_t s_stride0, s_stride1;
  int32_t scales_dtype;
  /* zeros: (K/group_size, N) in the scale dtype, or ONE int32 (scalar zero-point), or absent */
  const void* zeros;
  int64_t zeros_numel;
  int64_t z_stride0, z_stride1;
  int32_t zeros_dtype;
  int32_t _pad0;
  int64_t K, N;                     /* in_features, out_features */
  /* meta_args in the reference's order (core.py:525-538) */
  int32_t scaled_activations;
  int32_t W_nbits;
  int32_t group_size;
  int32_t unpack_mask;
  int32_t elements_per_sample;
  int32_t input_dtype;
  int32_t output_dtype;
  int32_t acc_dtype;
  int32_t meta_dtype;
  int32_t channel_scale_mode;
  int32_t W_group_mode;
  int32_t data_contiguous;
} glb_linear;

/* Library / ABI identification. */
GLB_API int glb_version(void);
GLB_API const char* glb_last_error(void);
/* Name of the kernel family the last glb_matmul on this thread launched (for tests / profiling). */
GLB_API const char* glb_last_kernel(void);
/* Number of kernel launches issued by this library since load (all threads). */
GLB_API uint64_t glb_launch_count(void);

/* Bit-pack `nbits`-wide codes (one per byte) along the columns of a (rows, cols) uint8 matrix into
 * `pack_bits`-wide words: element j of a row -> word j / eps, bit offset (j % eps) * nbits,
 * eps = pack_bits / nbits.  Output is (rows, cols/eps) row-major, or its transpose
 * (cols/eps, rows) row-major when `transpose` != 0 (the layout GemLiteLinear stores). */
GLB_API int glb_pack_cols(const uint8_t* W_q, int64_t rows, int64_t cols, int nbits, int pack_bits,
                  void* out, int transpose, void* stream);
/* Inverse of glb_pack_cols: `packed` has the layout glb_pack_cols wrote; `out` is (rows, cols) uint8. */
GLB_API int glb_unpack_cols(const void* packed, int64_t rows, int64_t cols, int nbits, int pack_bits,
                    uint8_t* out, int transpose, void* stream);

/* Dynamic per-token activation quantisation: for each of the M rows of x (M, K; x_dtype in
 * {FP16, BF16, FP32}), s = max(absmax/qmax, 1e-6) in fp32, q = clamp(x / s) rounded half away from
 * zero (INT8) or converted round-to-nearest-even (FP8E4 / FP8E5).  q is (M, K) of q_dtype, s is (M) fp32. */
GLB_API int glb_quant_act_per_token(const void* x, int64_t M, int64_t K, int x_dtype, int q_dtype,
                            void* q, float* s, void* stream);

/* MX / NV dynamic activation quantisation (gemlite/quant_utils.py:503-954, the *_triton_v2 functions): x (M, K;
 * x_dtype in {FP16, BF16, FP32}; K % 32 == 0) is quantised per row and per group of 32 (mx_dtype GLB_MXFP8: q (M, K)
 * e4m3 bytes; GLB_MXFP4: q (M, K/2) e2m1 codes, even k in the low nibble; scales = e8m0 exponent bytes) or 16
 * (GLB_NVFP4: q (M, K/2) e2m1 codes; scales = e4m3 bytes, effective scale = e4m3 * 0.05) consecutive k.
 * `scales` is (M_pad, K/group) with M_pad = M rounded up to the group size, as the reference allocates it; the rows
 * M..M_pad-1 are written too (97 = 2^-30 for MX, 0 for NV).  fp32 arithmetic, IEEE division. */
GLB_API int glb_quant_act_mx(const void* x, int64_t M, int64_t K, int x_dtype, int mx_dtype,
                     void* q, void* scales, void* stream);

/* Unique id of the CUDA-graph capture `stream` is recording, 0 when it is not capturing.  The workspace contract
 * ("zero-initialised once") needs it on the caller's side: a buffer first allocated and zeroed INSIDE a capture is only
 * zero when that graph replays, so a caller must not hand it to launches outside that capture (gemlite_b200/core.py
 * keeps one workspace per capture id). */
GLB_API unsigned long long glb_stream_capture_id(void* stream);

/* Scratch the matmul needs for (layer, M, matmul_type); 0 when none.  Non-zero only for the split-K launches of
 * the CTA-pair tcgen05 kernel (16 < M <= 128 on layers with few 256-feature tiles): per-slice arrival counters
 * followed by fp32 partial tiles.  The buffer must be ZERO-INITIALISED once by the caller (the reference's Triton
 * forwards use torch.zeros for the same purpose, gemv_revsplitK_kernels.py:422); every launch leaves the counters
 * at zero again, so one buffer per stream can be reused by consecutive calls without re-clearing it. */
GLB_API size_t glb_workspace_bytes(const glb_linear* layer, int64_t M, int matmul_type);

/* y(M, N) = epilogue( x(M, K) . dequant(W_q)(K, N) ) [+ bias]
 *   x        : (M, K) row-major, dtype = storage type of layer->input_dtype
 *              (already quantised when layer->scaled_activations != 0)
 *   scales_x : (M) fp32 per-token activation scales; MX / NV layers with channel_scale_mode 4: (M, K / group) e8m0 / e4m3 bytes; or NULL
 *   y        : (M, N) row-major, layer->output_dtype
 *   bias     : (N) in the output dtype, or NULL
 *   workspace: >= glb_workspace_bytes(...) bytes, 16-byte aligned, zero-initialised once (may be NULL when that is 0);
 *              not shared between streams that run concurrently */
GLB_API int glb_matmul(const glb_linear* layer, const void* x, const void* scales_x, void* y, int64_t M,
               int matmul_type, void* workspace, size_t workspace_bytes, const void* bias, void* stream);

/* Global knobs (host side, cheap): enable programmatic dependent launch on the decode kernels so that
 * a layer's weight prefetch overlaps the tail of the previous layer's kernel. Default: on. */
/* Fused output all-gather for column-sharded decode (no reference equivalent -- the reference is single-GPU).
 * Like glb_matmul at M == 1, but the result row is stored at column `col0` of EVERY destination in `y_dst`
 * (device-accessible pointers to row 0 / column 0 of each rank's gathered output: this GPU's buffer and the peers'
 * symmetric buffers mapped over NVLink).  The caller synchronises the ranks (one barrier per step).  Returns
 * GLB_ERR_UNSUPPORTED for anything but an A16W4 decode layer at M == 1: use glb_matmul + a collective there. */
GLB_API int glb_matmul_bcast(const glb_linear* layer, const void* x, const void* scales_x, void* const* y_dst, int n_dst,
                             int64_t col0, int64_t M, const void* bias, void* stream);

GLB_API void glb_set_pdl(int enabled);
GLB_API int glb_get_pdl(void);

/* Diagnostics (GLB_GEMV_DEBUG=4): copies the decode kernel's recorded phase timestamps (9 x uint64 per record:
 * CTA index, then 8 %globaltimer stamps in ns) to host memory `out`, clears the buffer, returns the word count. */
GLB_API int glb_debug_timeline(unsigned long long* out, int max_words);

#ifdef __cplusplus
}
#endif
#endif /* GEMLITE_B200_H */
