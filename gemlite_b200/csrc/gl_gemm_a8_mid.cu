// A8W8 batched-decode kernel, 16 < M <= 128: INT8 x INT8 -> INT32 (bit-exact) and e4m3 x e4m3 -> FP32, weights (N, K)
// K-contiguous bytes -- the reference's GEMM_SPLITK range for 8-bit activations (core.py:108-114,
// triton_kernels/gemm_splitK_kernels.py:277-450; per-channel / per-token scaling of :330-350 fused in the epilogue).
//
// Why (profiles/vs_reference_timing_r2_visit1.txt): at M = 64 the 128-feature tcgen05 tile streams a 4096^2 layer
// through 32 CTAs (19.9 us per call in a CUDA graph, the reference's Triton build 12.0 us); an SM can only pull ~100 GB/s
// by itself, so 512 KB per CTA is ~5 us before any math.  Same decomposition as gl_gemm_mid.cu: a CLUSTER of
// KS = K / 1024 CTAs owns a 128-feature slab, CTA `rank` the k range [1024 rank, +1024): 128 KB of weights pulled into
// shared memory by descriptor-free bulk copies BEFORE griddepcontrol.wait (the activation quantiser that produces x is
// still running then), the x tile (<= 64 tokens x 1024 bytes) after it.  Warp w = (16-feature block w % 8, k half w / 8);
// mma.sync.m16n8k32 with the weights as the 16-row operand and 8 tokens per instruction -- 8-bit operands need no
// unpacking, every lane feeds 16 contiguous bytes of its row to two MMAs (same k permutation on both operands).
// Partial sums: k halves through shared memory, cluster ranks through distributed shared memory, both in fixed order
// (integer sums are exact in any order; the fp32 order is fixed => deterministic).  M > 64: second pass over the
// resident weights.  Epilogue arithmetic = the tcgen05 A8W8 kernels' (gl_gemm_tc.cu): cvt.rn.f32.s32(acc) * (s_x * s_w).
#include <cuda.h>
#include <stdlib.h>

#include "gl_common.cuh"
#include "gl_ptx.cuh"

namespace glb {
namespace {

constexpr int AM_ROWS = 128;                     // features per cluster
constexpr int AM_KC = 1024;                      // k bytes per CTA
constexpr int AM_PITCH = AM_KC + 64;             // row pitch: 68 x 16 B => the 8 lanes of a quarter-warp hit 8 bank groups
constexpr int AM_NW = 16;
constexpr int AM_TOK = 64;
constexpr int AM_NT = AM_TOK / 8;
constexpr int AM_WBYTES = AM_ROWS * AM_PITCH;    // 136 KB
constexpr int AM_XBYTES = AM_TOK * AM_PITCH;     // 68 KB; reused as the [2][64][128] 32-bit reduction buffer
constexpr int AM_SMEM = AM_WBYTES + AM_XBYTES + 128;
static_assert(2 * AM_TOK * AM_ROWS * 4 <= AM_XBYTES, "reduction buffer must fit in the x tile");

struct A8MidParams {
  const uint8_t* W; const uint8_t* x; const float* scales_w; const float* scales_x; void* y; const void* bias;
  int M, N, K, csm, out_dtype, ks;
};

__device__ __forceinline__ void mma_s8(int32_t (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_e4m3(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   ptx::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t ld_dsmem_u32(uint32_t cluster_addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(cluster_addr) : "memory");
  return v;
}

template <bool IS_INT>
__global__ void __launch_bounds__(AM_NW * 32, 1)
gemm_a8w8_mid_kernel(const A8MidParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* wbuf = smem;                                           // [128][AM_PITCH]
  unsigned char* xbuf = wbuf + AM_WBYTES;                               // [64][AM_PITCH]
  uint32_t* red = reinterpret_cast<uint32_t*>(xbuf);                    // [2 k halves][64][128] after the main loop
  uint64_t* bars = reinterpret_cast<uint64_t*>(xbuf + AM_XBYTES);       // [4] weights (32 rows each), [1] x

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ks = P.ks;
  const uint32_t rank = (ks > 1) ? ptx::cluster_ctarank() : 0u;
  const int slab = blockIdx.x / ks;
  const int feat0 = slab * AM_ROWS, kbeg = (int)rank * AM_KC;
  const int fb = warp & 7, kh = warp >> 3;

  // ---- weights: independent of the previous kernel ----
  if (threadIdx.x == 0) {
    for (int c = 0; c < 5; ++c) ptx::mbar_init(&bars[c], 1);
    ptx::fence_mbar_init();
    for (int c = 0; c < 4; ++c) ptx::mbar_arrive_expect_tx(&bars[c], 32u * AM_KC);
  }
  __syncthreads();
  if (threadIdx.x < AM_ROWS)
    bulk_g2s(wbuf + (size_t)threadIdx.x * AM_PITCH, P.W + (size_t)(feat0 + threadIdx.x) * P.K + kbeg, AM_KC, &bars[threadIdx.x >> 5]);
  pdl_wait_prior_grids();
  pdl_launch_dependents();

  const unsigned char* wa = wbuf + (size_t)(fb * 16 + g) * AM_PITCH + (size_t)kh * 512 + 16 * t;
  const unsigned char* wb = wa + (size_t)8 * AM_PITCH;
  const unsigned char* xa = xbuf + (size_t)g * AM_PITCH + (size_t)kh * 512 + 16 * t;

  int pass = 0;
  for (int tok0 = 0; tok0 < P.M; tok0 += AM_TOK, ++pass) {
    const int ntok = (P.M - tok0 < AM_TOK) ? (P.M - tok0) : AM_TOK;
    // ---- x tile: one bulk copy per token row (rows >= M keep whatever the buffer holds; their outputs are never stored) ----
    if (threadIdx.x == 0) ptx::mbar_arrive_expect_tx(&bars[4], (uint32_t)ntok * AM_KC);
    if (pass > 0) ptx::fence_proxy_async_smem();                         // generic accesses of the reduction buffer before async-proxy writes
    __syncthreads();                                                    // expect_tx precedes every complete_tx
    if (threadIdx.x < ntok)
      bulk_g2s(xbuf + (size_t)threadIdx.x * AM_PITCH, P.x + (size_t)(tok0 + threadIdx.x) * P.K + kbeg, AM_KC, &bars[4]);
    if (pass == 0) ptx::mbar_wait(&bars[fb >> 1], 0);
    ptx::mbar_wait(&bars[4], (uint32_t)(pass & 1));

    int32_t acci[AM_NT][4];
    float accf[AM_NT][4];
#pragma unroll
    for (int n = 0; n < AM_NT; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) { acci[n][i] = 0; accf[n][i] = 0.f; }
#pragma unroll 2
    for (int o = 0; o < 512; o += 64) {
      const uint4 p = *reinterpret_cast<const uint4*>(wa + o);
      const uint4 q = *reinterpret_cast<const uint4*>(wb + o);
#pragma unroll
      for (int n = 0; n < AM_NT; ++n) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xa + (size_t)n * 8 * AM_PITCH + o);
        if (IS_INT) {
          mma_s8(acci[n], p.x, q.x, p.y, q.y, xv.x, xv.y);
          mma_s8(acci[n], p.z, q.z, p.w, q.w, xv.z, xv.w);
        } else {
          mma_e4m3(accf[n], p.x, q.x, p.y, q.y, xv.x, xv.y);
          mma_e4m3(accf[n], p.z, q.z, p.w, q.w, xv.z, xv.w);
        }
      }
    }

    // ---- reduction: k halves through shared memory, ranks through DSMEM ----
    __syncthreads();                                                    // every warp is done reading the x tile
#pragma unroll
    for (int n = 0; n < AM_NT; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // d0 / d1: (feature g, tokens 2t, 2t+1); d2 / d3: (feature g + 8, same tokens)
        const int f = fb * 16 + g + ((i & 2) ? 8 : 0), tok = n * 8 + 2 * t + (i & 1);
        red[((size_t)kh * AM_TOK + tok) * AM_ROWS + f] = IS_INT ? (uint32_t)acci[n][i] : __float_as_uint(accf[n][i]);
      }
    __syncthreads();
    for (int i = threadIdx.x; i < AM_TOK * AM_ROWS; i += AM_NW * 32) {
      const uint32_t a = red[i], b = red[AM_TOK * AM_ROWS + i];
      red[i] = IS_INT ? (uint32_t)((int32_t)a + (int32_t)b) : __float_as_uint(__uint_as_float(a) + __uint_as_float(b));
    }
    if (ks > 1) ptx::cluster_sync(); else __syncthreads();
    {
      const int per = AM_TOK * AM_ROWS / ks;
      const uint32_t red_addr = ptx::smem_u32(red);
      for (int j = threadIdx.x; j < per; j += AM_NW * 32) {
        const int i = (int)rank * per + j;
        const int tok = i >> 7, fl = i & (AM_ROWS - 1);
        const int m = tok0 + tok;
        int32_t si = 0;
        float sf = 0.f;
        if (ks > 1) {
          for (int r = 0; r < ks; ++r) {
            const uint32_t v = ld_dsmem_u32(ptx::mapa(red_addr + (uint32_t)i * 4u, (uint32_t)r));
            if (IS_INT) si += (int32_t)v; else sf += __uint_as_float(v);
          }
        } else {
          if (IS_INT) si = (int32_t)red[i]; else sf = __uint_as_float(red[i]);
        }
        if (m >= P.M) continue;
        const int f = feat0 + fl;
        const int64_t oi = (int64_t)m * P.N + f;
        if (IS_INT && P.out_dtype == GLB_INT32 && P.csm == 0 && !P.bias) {
          reinterpret_cast<int32_t*>(P.y)[oi] = si;
          continue;
        }
        float v = IS_INT ? (float)si : sf;                               // cvt.rn.f32.s32
        const float sb = (P.csm == 1 || P.csm == 3) ? P.scales_w[f] : 1.f;
        if (P.csm == 1) v = __fmul_rn(v, sb);
        else if (P.csm == 2) v = __fmul_rn(v, P.scales_x[m]);
        else if (P.csm == 3) v = __fmul_rn(v, __fmul_rn(P.scales_x[m], sb));
        float bv = 0.f;
        if (P.bias) {
          if (P.out_dtype == GLB_FP16) bv = __half2float(reinterpret_cast<const __half*>(P.bias)[f]);
          else if (P.out_dtype == GLB_BF16) bv = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.bias)[f]);
          else bv = reinterpret_cast<const float*>(P.bias)[f];
        }
        if (P.out_dtype == GLB_FP16) {
          if (P.bias) v = __half2float(__float2half_rn(v)) + bv;
          reinterpret_cast<__half*>(P.y)[oi] = __float2half_rn(v);
        } else if (P.out_dtype == GLB_BF16) {
          if (P.bias) v = __bfloat162float(__float2bfloat16_rn(v)) + bv;
          reinterpret_cast<__nv_bfloat16*>(P.y)[oi] = __float2bfloat16_rn(v);
        } else if (P.out_dtype == GLB_FP32) {
          reinterpret_cast<float*>(P.y)[oi] = v + bv;
        } else {
          reinterpret_cast<int32_t*>(P.y)[oi] = (int32_t)(v + bv);
        }
      }
    }
    if (ks > 1) ptx::cluster_sync(); else __syncthreads();             // peers are done with this CTA's partials
  }
}

}  // namespace

bool gemm_a8_mid_supported(const glb_linear& L, int64_t M) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("GLB_MID_KERNEL"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return false;
  if (M <= 16 || M > 128) return false;
  if (is_mx(L.input_dtype)) return false;
  const int xd = act_storage_dtype(L.input_dtype);
  const bool is_int = (xd == GLB_INT8), is_f8 = (xd == GLB_FP8E4);
  if (!is_int && !is_f8) return false;
  if (L.W_group_mode != 0 || L.elements_per_sample != 1) return false;
  if (is_int && L.wq_dtype != GLB_INT8) return false;
  if (is_f8 && L.wq_dtype != GLB_FP8E4) return false;
  if (L.wq_stride0 != 1 || L.wq_stride1 != L.K) return false;            // K-contiguous weights
  if (L.channel_scale_mode < 0 || L.channel_scale_mode > 3) return false;
  if ((L.channel_scale_mode == 1 || L.channel_scale_mode == 3) && (L.scales_dtype != GLB_FP32 || L.scales_numel != L.N)) return false;
  if (L.channel_scale_mode >= 1 && L.meta_dtype != GLB_FP32) return false;
  if (!(L.output_dtype == GLB_FP16 || L.output_dtype == GLB_BF16 || L.output_dtype == GLB_FP32 ||
        (is_int && L.output_dtype == GLB_INT32)))
    return false;
  if (!(L.K == 1024 || L.K == 2048 || L.K == 4096 || L.K == 8192)) return false;
  if (L.N % AM_ROWS != 0) return false;
  if (reinterpret_cast<uintptr_t>(L.W_q) & 15) return false;
  return true;
}

int launch_gemm_a8_mid(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M, const void* bias,
                       cudaStream_t stream) {
  GLB_REQUIRE(gemm_a8_mid_supported(L, M), GLB_ERR_UNSUPPORTED, "A8W8 batched-decode kernel: unsupported layer configuration");
  GLB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, GLB_ERR_INVALID, "A8W8 batched-decode kernel: x must be 16-byte aligned");
  A8MidParams P;
  P.W = reinterpret_cast<const uint8_t*>(L.W_q);
  P.x = reinterpret_cast<const uint8_t*>(x);
  P.scales_w = (L.channel_scale_mode == 1 || L.channel_scale_mode == 3) ? reinterpret_cast<const float*>(L.scales) : nullptr;
  P.scales_x = scales_x; P.y = y; P.bias = bias;
  P.M = (int)M; P.N = (int)L.N; P.K = (int)L.K; P.csm = L.channel_scale_mode; P.out_dtype = L.output_dtype;
  P.ks = (int)(L.K / AM_KC);
  const bool is_int = (act_storage_dtype(L.input_dtype) == GLB_INT8);
  auto kern = is_int ? gemm_a8w8_mid_kernel<true> : gemm_a8w8_mid_kernel<false>;
  GLB_ENSURE_SMEM(kern, AM_SMEM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(L.N / AM_ROWS) * (unsigned)P.ks);
  cfg.blockDim = dim3(AM_NW * 32);
  cfg.dynamicSmemBytes = AM_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (P.ks > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)P.ks; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_pdl_enabled.load()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  set_kernel_name(is_int ? "gemm_a8w8_int8_mma_mid" : "gemm_a8w8_fp8_mma_mid");
  GLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  count_launch();
  return GLB_OK;
}

}  // namespace glb
