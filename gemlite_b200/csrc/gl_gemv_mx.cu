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

// -------------------------------------------------------------------------------------------------
// Tensor-path flavour of the same kernel (GLB_MXV_VARIANT=1; STATUS: written at the end of round 1 without GPU time
// left to validate it).  The FHFMA version above issues ~3.2 instructions per weight and is issue-bound at 47 % of the
// HBM peak; here the dot products go through mma.sync.m16n8k16 (weights = 16-row operand, up to 8 tokens = the
// 8-column operand), which leaves the e2m1 -> f16 conversion (1 shift + 1 F2FP per byte) as the dominant cost.
// Same ring, producer and x staging.  Consumer warp = (row tile rt = warp & 3: 16 of the CTA's 64 rows, group quarter
// gq = warp >> 2: 4 of the stage's 16 groups).  Lane (g, t) reads, per group, ONE 32-bit word of row g and one of row
// g + 8 (bytes 4t..4t+3 of the group's 16 = k 8t..8t+7; with the TMA 128B swizzle the 32 lanes hit 32 different banks)
// and the matching 16 bytes of x; the k order inside a group is permuted identically for both operands.  The group's
// 16x8 fp32 tile is scaled by the two rows' e8m0 factors and added to the running output -- the factors of a stage are
// fetched by 32 lanes (row, group pair), converted once and exchanged with shuffles.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mxv_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <bool BF16, int MT>
__global__ void __launch_bounds__(MXV_THREADS, (MT <= 2) ? 2 : 1) gemv_mxfp4_mma_kernel(const __grid_constant__ CUtensorMap tmap_w, const MxvParams P) {
  static_assert(MT <= 8, "one 8-token column tile");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  unsigned char* ring = smem;
  __half* xs = reinterpret_cast<__half*>(smem + MXV_STAGES * MXV_TILE);      // [MT][K] fp16
  float* red = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(xs) + (size_t)MT * P.K * 2);   // [4 gq][MT][64 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 4 * MT * MXV_ROWS);
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
  asm volatile("bar.sync 1, %0;" ::"r"(MXV_WARPS * 32) : "memory");

  const int rt = warp & 3, gq = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  const bool tok_ok = g < MT;                                     // lane's B column = token g
  float out[4] = {0.f, 0.f, 0.f, 0.f};                            // rows (g, g + 8) of the tile x tokens (2t, 2t + 1)
  // scale fetch: lane -> (row rt*16 + (lane & 15), groups gq*4 + 2*(lane >> 4) + {0, 1}) of the stage
  const uint8_t* mxp = P.scales + (size_t)(gq * 4 + 2 * (lane >> 4)) * P.N + row0 + rt * 16 + (lane & 15);
  uint8_t mxraw0 = mxp[0], mxraw1 = mxp[P.N];
  // this lane's code words: CTA rows rt*16 + g and + 8 (same row % 8 = g), byte 4t of the group's 16-byte chunk
  const uint32_t roff = (uint32_t)(rt * 16 + g) * 128u + 4u * (uint32_t)t;
  const __half* xrow = xs + (size_t)(tok_ok ? g : 0) * P.K + 8 * t;
  uint32_t s = 0, ph = 0;
  for (int kb = 0; kb < stages; ++kb) {
    const float sf0 = e8m0_to_float(mxraw0), sf1 = e8m0_to_float(mxraw1);
    if (kb + 1 < stages) { mxp += (size_t)16 * P.N; mxraw0 = mxp[0]; mxraw1 = mxp[P.N]; }
    ptx::mbar_wait(&pfull[s], ph);
    const unsigned char* tile = ring + s * MXV_TILE;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gidx = gq * 4 + j;                                 // group of the stage: box gidx / 8, chunk gidx % 8
      const unsigned char* base = tile + (gidx >> 3) * MXV_BOX + roff;
      const uint32_t sw = (uint32_t)(((gidx & 7) ^ g) << 4);
      const uint32_t w0 = *reinterpret_cast<const uint32_t*>(base + sw);
      const uint32_t w1 = *reinterpret_cast<const uint32_t*>(base + 8 * 128 + sw);
      uint4 xv = make_uint4(0u, 0u, 0u, 0u);
      if (tok_ok) xv = *reinterpret_cast<const uint4*>(xrow + kb * MXV_BK + gidx * 32);
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      mxv_mma(d, mxv_e2m1x2(w0), mxv_e2m1x2(w1), mxv_e2m1x2(w0 >> 8), mxv_e2m1x2(w1 >> 8), xv.x, xv.y);
      mxv_mma(d, mxv_e2m1x2(w0 >> 16), mxv_e2m1x2(w1 >> 16), mxv_e2m1x2(w0 >> 24), mxv_e2m1x2(w1 >> 24), xv.z, xv.w);
      // factors of (row g, group) and (row g + 8, group): held by lanes g / g + 8 (+16 for the second group pair)
      const float fsel = (j & 1) ? sf1 : sf0;
      const float fa = __shfl_sync(0xffffffffu, fsel, g + 16 * (j >> 1));
      const float fb = __shfl_sync(0xffffffffu, fsel, g + 8 + 16 * (j >> 1));
      out[0] = fmaf(d[0], fa, out[0]); out[1] = fmaf(d[1], fa, out[1]);
      out[2] = fmaf(d[2], fb, out[2]); out[3] = fmaf(d[3], fb, out[3]);
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&pfree[s]);
    if (++s == MXV_STAGES) { s = 0; ph ^= 1u; }
  }

  // ---- reduce the four group quarters, bias, store ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = 2 * t + (i & 1), row = rt * 16 + g + ((i & 2) ? 8 : 0);
    if (m < MT) red[(gq * MT + m) * MXV_ROWS + row] = out[i];
  }
  asm volatile("bar.sync 1, %0;" ::"r"(MXV_WARPS * 32) : "memory");
  for (int i = threadIdx.x; i < MT * MXV_ROWS; i += MXV_WARPS * 32) {
    const int m = i / MXV_ROWS, row = i - m * MXV_ROWS;            // consecutive threads = consecutive features: coalesced
    if (m >= P.M) continue;
    const int n = row0 + row;
    float v = 0.f;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) v += red[(q4 * MT + m) * MXV_ROWS + row];
    if (BF16) {
      if (P.bias) v = __bfloat162float(__float2bfloat16_rn(v)) + __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.bias)[n]);
      reinterpret_cast<__nv_bfloat16*>(P.y)[(size_t)m * P.N + n] = __float2bfloat16_rn(v);
    } else {
      if (P.bias) v = __half2float(__float2half_rn(v)) + __half2float(reinterpret_cast<const __half*>(P.bias)[n]);
      reinterpret_cast<__half*>(P.y)[(size_t)m * P.N + n] = __float2half_rn(v);
    }
  }
}

int mxv_variant() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GLB_MXV_VARIANT"); v = e ? atoi(e) : 0; }
  return v;
}

template <bool BF16, int MT>
int launch_mxv_mma(const glb_linear& L, const void* x, void* y, int64_t M, const void* bias, cudaStream_t stream) {
  CUtensorMap tw;
  int rc = make_tmap_2d(&tw, GLB_UINT8, L.W_q, (uint64_t)(L.K / 2), (uint64_t)L.N, (uint64_t)(L.K / 2), 128, MXV_ROWS, 1);
  if (rc != GLB_OK) return rc;
  MxvParams P;
  P.scales = reinterpret_cast<const uint8_t*>(L.scales);
  P.x = x; P.y = y; P.bias = bias; P.M = (int)M; P.N = (int)L.N; P.K = (int)L.K;
  const size_t smem = (size_t)MXV_STAGES * MXV_TILE + (size_t)MT * L.K * 2 + (size_t)4 * MT * MXV_ROWS * 4 + 2 * MXV_STAGES * 8 + 64;
  auto kern = gemv_mxfp4_mma_kernel<BF16, MT>;
  GLB_ENSURE_SMEM(kern, smem);
  kern<<<dim3((unsigned)(L.N / MXV_ROWS)), dim3(MXV_THREADS), smem, stream>>>(tw, P);
  count_launch();
  GLB_CHECK_CUDA(cudaGetLastError());
  return GLB_OK;
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
  if (mxv_variant() == 1) {          // opt-in tensor-path flavour (see gemv_mxfp4_mma_kernel)
    set_kernel_name("gemv_mxfp4_mma");
    if (M == 1) return bf ? launch_mxv_mma<true, 1>(L, x, y, M, bias, stream) : launch_mxv_mma<false, 1>(L, x, y, M, bias, stream);
    if (M == 2) return bf ? launch_mxv_mma<true, 2>(L, x, y, M, bias, stream) : launch_mxv_mma<false, 2>(L, x, y, M, bias, stream);
    return bf ? launch_mxv_mma<true, 4>(L, x, y, M, bias, stream) : launch_mxv_mma<false, 4>(L, x, y, M, bias, stream);
  }
  set_kernel_name("gemv_mxfp4");
  if (M == 1) return bf ? launch_mxv<true, 1>(L, x, y, M, bias, stream) : launch_mxv<false, 1>(L, x, y, M, bias, stream);
  if (M == 2) return bf ? launch_mxv<true, 2>(L, x, y, M, bias, stream) : launch_mxv<false, 2>(L, x, y, M, bias, stream);
  return bf ? launch_mxv<true, 4>(L, x, y, M, bias, stream) : launch_mxv<false, 4>(L, x, y, M, bias, stream);
}

}  // namespace glb
