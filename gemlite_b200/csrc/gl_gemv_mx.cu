// Decode-regime kernel for MXFP4 weight-only layers (A16W4_MXFP, config C5 at M <= 4): HBM-bound streaming of the
// e2m1 codes with the dot products on the CUDA cores.
//
// Reference counterpart: gemm_splitK_MX_kernel / gemv_MX_kernel at M == 1 (core.py:100-114 routes every MX dtype at
// M == 1 to GEMM_SPLITK; gemv_kernels.py:440-550 decodes e2m1 through a LUT gather and e8m0 by exp2(e - 127)).
//
// Layout consumed (unchanged): codes (N, K/2) bytes, K contiguous, even k = low nibble; scales (K/32, N) e8m0 bytes.
// Algorithmic bytes per launch: K*N/2 + K*N/32 + 2*M*K + 2*M*N (C5: 142.6 MB -> 21.7 us at the measured HBM peak).
//
// Design: one CTA per 64 feature rows (N = 16384 -> 256 CTAs, two per SM); a producer lane streams 512 k of every row
// per stage (256 contiguous bytes per row: the first version read 64 B per row and stage and reached only 1.7 TB/s --
// one DRAM page activation per 64 bytes) as two 128B-swizzled [128 B x 64 rows] TMA boxes through a 4-deep ring;
// 16 consumer warps = 2 row halves x 8 group selectors, thread = (feature row, 2 of the stage's 16 groups): per group
// one conflict-free LDS.128 of codes, ONE coalesced e8m0 byte (requested a stage ahead),
// cvt.rn.f16x2.e2m1x2 per byte (exact), fp32 accumulation with fma.rn.f32.f16 against x broadcast from shared memory,
// scale applied once per 32-k group in fp32.  x is staged as fp16 (bf16 activations are converted: exact unless
// |x| > 65504 or < 6e-8).  Cross-warp (k quarter) reduction through shared memory, bias fused, coalesced stores.
#include <stdlib.h>

#include "gl_common.cuh"
#include "gl_ptx.cuh"

namespace glb {
namespace {

constexpr int MXV_ROWS = 64;
constexpr int MXV_BK = 512;                        // k per stage (256 bytes of codes per row, 16 groups)
constexpr int MXV_STAGES = 4;
constexpr int MXV_BOX = MXV_ROWS * 128;            // one swizzled box: 128 B x 64 rows = 8 KB
constexpr int MXV_TILE = 2 * MXV_BOX;              // 16 KB of codes per stage
constexpr int MXV_WARPS = 16;                      // consumers; warp 16 = producer
constexpr int MXV_THREADS = (MXV_WARPS + 1) * 32;

struct MxvParams {
  const uint8_t* scales;   // (K/32, N) e8m0
  const void* x;           // (M, K) fp16 / bf16
  void* y;                 // (M, N)
  const void* bias;
  int M, N, K;
};

__device__ __forceinline__ uint32_t mxv_e2m1x2(uint32_t byte) {
  uint32_t r;
  asm("{\n\t.reg .b8 t;\n\tcvt.u8.u32 t, %1;\n\tcvt.rn.f16x2.e2m1x2 %0, t;\n\t}" : "=r"(r) : "r"(byte));
  return r;
}
__device__ __forceinline__ float mxv_fma(uint32_t a16x2, uint32_t b16x2, float c, bool hi) {
  const unsigned short ha = hi ? (unsigned short)(a16x2 >> 16) : (unsigned short)(a16x2 & 0xffffu);
  const unsigned short hb = hi ? (unsigned short)(b16x2 >> 16) : (unsigned short)(b16x2 & 0xffffu);
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(c) : "h"(ha), "h"(hb));
  return c;
}

template <bool BF16, int MT>
__global__ void __launch_bounds__(MXV_THREADS, (MT == 1) ? 2 : 1) gemv_mxfp4_kernel(const __grid_constant__ CUtensorMap tmap_w, const MxvParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  unsigned char* ring = smem;                                                // STAGES x 8 KB
  __half* xs = reinterpret_cast<__half*>(smem + MXV_STAGES * MXV_TILE);      // [MT][K] fp16
  float* red = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(xs) + (size_t)MT * P.K * 2);   // [8 gsel][MT][64 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 8 * MT * MXV_ROWS);
  uint64_t* pfull = bars;
  uint64_t* pfree = bars + MXV_STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * MXV_ROWS;
  const int stages = P.K / MXV_BK;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmap_w);
    for (int s = 0; s < MXV_STAGES; ++s) { ptx::mbar_init(&pfull[s], 1); ptx::mbar_init(&pfree[s], MXV_WARPS); }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  if (warp == MXV_WARPS) {
    // ================= producer: stream the CTA's rows, 128 k (64 bytes per row) per stage =================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int kb = 0; kb < stages; ++kb) {
        ptx::mbar_wait(&pfree[s], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&pfull[s], MXV_TILE);
        ptx::tma_load_2d(ring + s * MXV_TILE, &tmap_w, &pfull[s], kb * (MXV_BK / 2), row0);
        ptx::tma_load_2d(ring + s * MXV_TILE + MXV_BOX, &tmap_w, &pfull[s], kb * (MXV_BK / 2) + 128, row0);
        if (++s == MXV_STAGES) { s = 0; ph ^= 1u; }
      }
    }
    return;
  }

  // ================= consumers =================
  // x -> shared memory as fp16 (weights are already streaming)
  {
    const int total = MT * P.K;
    for (int i = threadIdx.x * 8; i < total; i += MXV_WARPS * 32 * 8) {
      const int m = i / P.K, k = i - m * P.K;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (m < P.M) {
        const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(P.x) + (size_t)m * P.K + k);
        if (BF16) {
          const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&v);
          __half2 h[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = __float22half2_rn(__bfloat1622float2(b[j]));
          o = *reinterpret_cast<uint4*>(h);
        } else {
          o = v;
        }
      }
      *reinterpret_cast<uint4*>(xs + i) = o;
    }
  }
  asm volatile("bar.sync 1, %0;" ::"r"(MXV_WARPS * 32) : "memory");      // consumers only (the producer warp has left)

  const int rh = warp & 1, gsel = warp >> 1;           // row half, group selector (0..7): this thread's groups are gsel, gsel + 8
  const int row = rh * 32 + lane;                      // feature row inside the CTA
  const int n = row0 + row;
  float out[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) out[m] = 0.f;
  // scales of (stage kb, group g): row (16*kb + g) of the (K/32, N) matrix; consecutive lanes = consecutive n (coalesced)
  const uint8_t* mxp = P.scales + (size_t)gsel * P.N + n;
  uint8_t mxraw0 = mxp[0], mxraw1 = mxp[(size_t)8 * P.N];
  // swizzled code chunk of group g: box g / 8, 16-byte chunk (g % 8) ^ (row % 8)   (TMA 128B swizzle)
  const uint32_t coff = (uint32_t)row * 128u + (uint32_t)(((gsel & 7) ^ (row & 7)) << 4);
  uint32_t s = 0, ph = 0;
  for (int kb = 0; kb < stages; ++kb) {
    const float sf0 = e8m0_to_float(mxraw0), sf1 = e8m0_to_float(mxraw1);
    if (kb + 1 < stages) { mxp += (size_t)16 * P.N; mxraw0 = mxp[0]; mxraw1 = mxp[(size_t)8 * P.N]; }
    ptx::mbar_wait(&pfull[s], ph);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint4 wv = *reinterpret_cast<const uint4*>(ring + s * MXV_TILE + u * MXV_BOX + coff);
      const uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
      float acc[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) acc[m] = 0.f;
      const __half* xk = xs + kb * MXV_BK + (gsel + 8 * u) * 32;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = mxv_e2m1x2(w[cc] >> (8 * j));       // (k 8cc+2j, 8cc+2j+1) as fp16x2
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const uint4 xv = *reinterpret_cast<const uint4*>(xk + (size_t)m * P.K + cc * 8);   // warp-uniform address: broadcast
          const uint32_t xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[m] = mxv_fma(v[j], xx[j], acc[m], false);
            acc[m] = mxv_fma(v[j], xx[j], acc[m], true);
          }
        }
      }
      const float sf = u ? sf1 : sf0;
#pragma unroll
      for (int m = 0; m < MT; ++m) out[m] = fmaf(acc[m], sf, out[m]);
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&pfree[s]);
    if (++s == MXV_STAGES) { s = 0; ph ^= 1u; }
  }

  // ---- reduce the eight group selectors, bias, store (consecutive lanes = consecutive features: coalesced) ----
#pragma unroll
  for (int m = 0; m < MT; ++m) red[(gsel * MT + m) * MXV_ROWS + row] = out[m];
  asm volatile("bar.sync 1, %0;" ::"r"(MXV_WARPS * 32) : "memory");
  if (gsel == 0) {
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (m < P.M) {
        float v = 0.f;
#pragma unroll
        for (int g2 = 0; g2 < 8; ++g2) v += red[(g2 * MT + m) * MXV_ROWS + row];
        if (BF16) {
          if (P.bias) v = __bfloat162float(__float2bfloat16_rn(v)) + __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.bias)[n]);
          reinterpret_cast<__nv_bfloat16*>(P.y)[(size_t)m * P.N + n] = __float2bfloat16_rn(v);
        } else {
          if (P.bias) v = __half2float(__float2half_rn(v)) + __half2float(reinterpret_cast<const __half*>(P.bias)[n]);
          reinterpret_cast<__half*>(P.y)[(size_t)m * P.N + n] = __float2half_rn(v);
        }
      }
    }
  }
}

template <bool BF16, int MT>
int launch_mxv(const glb_linear& L, const void* x, void* y, int64_t M, const void* bias, cudaStream_t stream) {
  CUtensorMap tw;
  int rc = make_tmap_2d(&tw, GLB_UINT8, L.W_q, (uint64_t)(L.K / 2), (uint64_t)L.N, (uint64_t)(L.K / 2), 128, MXV_ROWS, 1);
  if (rc != GLB_OK) return rc;
  MxvParams P;
  P.scales = reinterpret_cast<const uint8_t*>(L.scales);
  P.x = x; P.y = y; P.bias = bias; P.M = (int)M; P.N = (int)L.N; P.K = (int)L.K;
  const size_t smem = (size_t)MXV_STAGES * MXV_TILE + (size_t)MT * L.K * 2 + (size_t)8 * MT * MXV_ROWS * 4 + 2 * MXV_STAGES * 8 + 64;
  auto kern = gemv_mxfp4_kernel<BF16, MT>;
  GLB_ENSURE_SMEM(kern, smem);
  kern<<<dim3((unsigned)(L.N / MXV_ROWS)), dim3(MXV_THREADS), smem, stream>>>(tw, P);
  count_launch();
  GLB_CHECK_CUDA(cudaGetLastError());
  return GLB_OK;
}

}  // namespace

bool gemv_mx_supported(const glb_linear& L, int64_t M) {
  if (M < 1 || M > 4) return false;
  if (!gemm_mx_tc2_supported(L, M)) return false;                 // same layout checks (MXFP4 weight-only, e8m0 / 32)
  if (L.N % MXV_ROWS != 0 || L.K % MXV_BK != 0) return false;
  const size_t mt = (M == 1) ? 1 : ((M == 2) ? 2 : 4);
  const size_t smem = (size_t)MXV_STAGES * MXV_TILE + mt * L.K * 2 + 8 * mt * MXV_ROWS * 4 + 2 * MXV_STAGES * 8 + 64;
  return smem <= 220 * 1024;
}

int launch_gemv_mx(const glb_linear& L, const void* x, void* y, int64_t M, const void* bias, cudaStream_t stream) {
  GLB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, GLB_ERR_INVALID, "MX decode kernel: x must be 16-byte aligned");
  const bool bf = (L.input_dtype == GLB_MXBF16);
  set_kernel_name("gemv_mxfp4");
  if (M == 1) return bf ? launch_mxv<true, 1>(L, x, y, M, bias, stream) : launch_mxv<false, 1>(L, x, y, M, bias, stream);
  if (M == 2) return bf ? launch_mxv<true, 2>(L, x, y, M, bias, stream) : launch_mxv<false, 2>(L, x, y, M, bias, stream);
  return bf ? launch_mxv<true, 4>(L, x, y, M, bias, stream) : launch_mxv<false, 4>(L, x, y, M, bias, stream);
}

}  // namespace glb
