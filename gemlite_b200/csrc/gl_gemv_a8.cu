// A8W8 decode kernel (M <= 16): INT8 x INT8 -> INT32 (bit-exact) and FP8 x FP8 -> FP32, weights (N, K) K-contiguous
// bytes (reference: gemlite/core.py:100-114 sends these shapes to GEMV_SPLITK / GEMM_SPLITK; the per-channel /
// per-token scaling of triton_kernels/gemm_splitK_kernels.py:330-350 is fused in the epilogue).
//
// *Bound: HBM* -- N*K weight bytes per launch (16.8 MB at 4096^2 -> 2.6 us at the measured peak).  The tcgen05 A8W8
// kernel tiles 128 features per CTA, i.e. 32 CTAs at N = 4096: fine for prefill, but a decode call then streams its
// weights through 32 SMs only.  This kernel gives every CTA 16 feature rows over the full K (256 CTAs at N = 4096,
// two resident per SM):
//   * the 16 rows (K bytes each, contiguous in HBM) are pulled into shared memory by descriptor-free bulk copies
//     (cp.async.bulk, one per row and K-quarter, four mbarriers) issued BEFORE griddepcontrol.wait, so under
//     programmatic dependent launch the weights stream while the activation quantiser that produces x still runs;
//   * x (M rows of K bytes) is bulk-copied after the wait;
//   * the dot products run on the legacy tensor path, mma.sync.m16n8k32 (weights = 16-row operand, tokens = 8-column
//     operand).  Shared-memory rows are padded to K + 64 bytes so that the 128-bit fragment loads of a quarter-warp
//     (2 rows x 4 k-chunks) hit 8 different 16-byte bank groups; every lane reads 16 contiguous bytes of its row and
//     feeds them to two MMAs -- the k order inside a 64-byte block is permuted identically for weights and x, which a
//     dot product does not see.
//   * warp w owns K-eighth w; the 8 partial tiles are summed through shared memory in warp order (integer sums are
//     exact in any order; the fp32 order is fixed, so the result is deterministic).
// SASS (cuobjdump, sm_100a): the int8 flavour issues native IMMA.16832.S8.S8; there is NO legacy fp8 MMA on sm_100a -- ptxas
// lowers mma.sync...e4m3/e5m2 to F2FP.F16.E4M3.UNPACK_B conversions + HMMA.16816.F32 (exact products, fp32 accumulate),
// about 0.75 conversions per weight, so the fp8 flavour is expected to sit near the issue limit rather than at the HBM
// floor.  Weights arrive by UBLKCP (cp.async.bulk).
// STATUS: validated on a B200 in round 2 (parity + timing in profiles/optin_validation_r2.txt); default for M <= 16.
#include <cuda.h>
#include <stdlib.h>

#include "gl_common.cuh"
#include "gl_ptx.cuh"

namespace glb {
namespace {

constexpr int A8V_ROWS = 16;      // feature rows per CTA (the MMA's M)
constexpr int A8V_WARPS = 8;
constexpr int A8V_NCH = 4;        // K quarters, one mbarrier each
constexpr int A8V_PAD = 64;       // row pitch = K + 64 bytes

struct A8vParams {
  const uint8_t* W;        // (N, K) bytes
  const uint8_t* x;        // (M, K) bytes
  const float* scales_w;   // (N) or null
  const float* scales_x;   // (M) or null
  void* y;
  const void* bias;
  int M, N, K;
  int csm;                 // channel_scale_mode 0..3
  int out_dtype;           // GLB_FP16 / BF16 / FP32 / INT32
};

// D(16x8) += A(16x32, row) * B(32x8, col); fragment layout (PTX ISA, m16n8k32 8-bit types): lane = 4*g + t,
// a0: (row g, k 4t..4t+3) a1: (row g+8, same k) a2: (row g, k 16+4t..) a3: (row g+8, k 16+4t..),
// b0: (k 4t..4t+3, col g) b1: (k 16+4t.., col g), d0/d1: (row g, cols 2t, 2t+1) d2/d3: (row g+8, cols 2t, 2t+1)
__device__ __forceinline__ void mma_s8(int32_t (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int WF, int XF>   // 0 = e4m3, 1 = e5m2
__device__ __forceinline__ void mma_f8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if (WF == 0 && XF == 0)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else if (WF == 0 && XF == 1)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e5m2.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else if (WF == 1 && XF == 0)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e5m2.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e5m2.e5m2.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   ptx::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

// NT: 8-token column tiles (1: M <= 8, 2: M <= 16).  FMT: 0 int8; 1 + 2*WF + XF for the fp8 format pairs.
template <int FMT, int NT>
__global__ void __launch_bounds__(A8V_WARPS * 32, (NT == 1) ? 2 : 1)
gemv_a8w8_kernel(const A8vParams P) {
  constexpr bool IS_INT = (FMT == 0);
  constexpr int WF = IS_INT ? 0 : ((FMT - 1) >> 1), XF = IS_INT ? 0 : ((FMT - 1) & 1);
  extern __shared__ __align__(128) unsigned char smem[];
  const int K = P.K, pitch = K + A8V_PAD;
  unsigned char* wbuf = smem;                                            // [16][pitch]
  unsigned char* xbuf = wbuf + (size_t)A8V_ROWS * pitch;                 // [8 * NT][pitch]
  uint32_t* red = reinterpret_cast<uint32_t*>(xbuf + (size_t)8 * NT * pitch);   // [8 warps][NT][16 rows][8 cols]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + A8V_WARPS * NT * 128);     // [4] weights, [1] x

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int feat0 = blockIdx.x * A8V_ROWS;
  const int kq = K / A8V_NCH;                  // bytes per row and K quarter

  // ---- 1. weights -> shared memory (independent of the previous kernel) ----
  if (warp == 0) {
    if (lane == 0) {
      for (int c = 0; c <= A8V_NCH; ++c) ptx::mbar_init(&bars[c], 1);
      ptx::fence_mbar_init();
      for (int c = 0; c < A8V_NCH; ++c) ptx::mbar_arrive_expect_tx(&bars[c], (uint32_t)(A8V_ROWS * kq));
    }
    __syncwarp();
    if (lane < A8V_ROWS) {
      const uint8_t* src = P.W + (size_t)(feat0 + lane) * K;
#pragma unroll
      for (int c = 0; c < A8V_NCH; ++c) bulk_copy_g2s(wbuf + (size_t)lane * pitch + c * kq, src + c * kq, (uint32_t)kq, &bars[c]);
    }
  }
  __syncthreads();                 // barrier initialisation visible to every thread (still ahead of the dependency wait)

  // ---- 2. x depends on the previous kernel (the activation quantiser) ----
  pdl_wait_prior_grids();
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bars[A8V_NCH], (uint32_t)P.M * (uint32_t)K);
    for (int m = 0; m < P.M; ++m) bulk_copy_g2s(xbuf + (size_t)m * pitch, P.x + (size_t)m * K, (uint32_t)K, &bars[A8V_NCH]);
  }

  // ---- 3. main loop: warp w owns K eighth w (quarter w / 2) ----
  int32_t acci[NT][4];
  float accf[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) { acci[n][i] = 0; accf[n][i] = 0.f; }
  const int kw = K / A8V_WARPS;                                     // bytes of every row this warp owns (multiple of 64)
  const unsigned char* wa = wbuf + (size_t)g * pitch + (size_t)warp * kw + 16 * t;
  const unsigned char* wb = wa + (size_t)8 * pitch;
  const unsigned char* xa[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const int tok = n * 8 + g;
    xa[n] = xbuf + (size_t)(tok < P.M ? tok : 0) * pitch + (size_t)warp * kw + 16 * t;   // rows >= M: any valid row (discarded)
  }
  ptx::mbar_wait(&bars[warp >> 1], 0);
  ptx::mbar_wait(&bars[A8V_NCH], 0);
#pragma unroll 4
  for (int o = 0; o < kw; o += 64) {
    const uint4 p = *reinterpret_cast<const uint4*>(wa + o);
    const uint4 q = *reinterpret_cast<const uint4*>(wb + o);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const uint4 xv = *reinterpret_cast<const uint4*>(xa[n] + o);
      if (IS_INT) {
        mma_s8(acci[n], p.x, q.x, p.y, q.y, xv.x, xv.y);
        mma_s8(acci[n], p.z, q.z, p.w, q.w, xv.z, xv.w);
      } else {
        mma_f8<WF, XF>(accf[n], p.x, q.x, p.y, q.y, xv.x, xv.y);
        mma_f8<WF, XF>(accf[n], p.z, q.z, p.w, q.w, xv.z, xv.w);
      }
    }
  }

  // ---- 4. cross-warp reduction in warp order, scaling / bias epilogue ----
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    uint32_t* r = red + ((size_t)warp * NT + n) * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = g + ((i & 2) ? 8 : 0), col = 2 * t + (i & 1);
      r[row * 8 + col] = IS_INT ? (uint32_t)acci[n][i] : __float_as_uint(accf[n][i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NT * 128; i += A8V_WARPS * 32) {
    const int n = i >> 7, col = (i >> 4) & 7, row = i & 15;          // consecutive threads -> consecutive features
    const int m = n * 8 + col;
    if (m >= P.M) continue;
    int32_t si = 0;
    float sf = 0.f;
#pragma unroll
    for (int w = 0; w < A8V_WARPS; ++w) {
      const uint32_t v = red[((size_t)w * NT + n) * 128 + row * 8 + col];
      if (IS_INT) si += (int32_t)v; else sf += __uint_as_float(v);
    }
    const int f = feat0 + row;
    const int64_t oi = (int64_t)m * P.N + f;
    if (IS_INT && P.out_dtype == GLB_INT32 && P.csm == 0 && !P.bias) {
      reinterpret_cast<int32_t*>(P.y)[oi] = si;
      continue;
    }
    // same arithmetic as the tcgen05 A8W8 epilogue (gl_gemm_tc.cu)
    float v = IS_INT ? (float)si : sf;
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

// =================================================================================================
// A16W8 weight-only decode (fp16 activations x int8 / e4m3 weights, M <= 16): helper.A16W8_INT8 / A16W8_FP8
// (gemlite/helper.py:88-170; the reference sends these to GEMV_SPLITK at M == 1, core.py:100-114).  Same streaming
// skeleton as gemv_a8w8_kernel (16 feature rows per CTA over the full K, rows bulk-copied ahead of griddepcontrol.wait);
// the weights are widened to fp16 in registers -- int8: byte ^ 0x80 -> prmt with 0x64 -> (1024 + u) -> hsub2 1152 = the
// exact signed value (one LOP3 per 4, one PRMT + one HSUB2 per 2 weights); e4m3: cvt.rn.f16x2.e4m3x2, exact -- and fed
// to mma.sync.m16n8k16 with fp32 accumulation.  The per-channel scale is applied to the accumulator (W_group_mode 2:
// the reference rounds every q * s to fp16 first, so this path differs from the oracle by weight-rounding noise like the
// 4-bit decode kernel; channel_scale_mode 1 is identical to the reference's post-scale).
// =================================================================================================
struct A16vParams {
  const uint8_t* W;        // (N, K) bytes
  const void* x;           // (M, K) fp16
  const float* scales_w;   // (N)
  void* y;
  const void* bias;
  int M, N, K;
};

__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 4 weight bytes (k .. k+3 of one row) -> two fp16x2 registers (k, k+1) and (k+2, k+3)
template <int WF> __device__ __forceinline__ void widen4(uint32_t w, uint32_t& lo, uint32_t& hi) {
  if (WF == 0) {
    const uint32_t u = w ^ 0x80808080u;                                   // signed -> offset-128
    uint32_t m0, m1;
    asm("prmt.b32 %0, %1, %2, 0x5140;" : "=r"(m0) : "r"(u), "r"(0x64646464u));     // (1024 + u0, 1024 + u1)
    asm("prmt.b32 %0, %1, %2, 0x7362;" : "=r"(m1) : "r"(u), "r"(0x64646464u));     // (1024 + u2, 1024 + u3)
    const __half2 k1152 = __half2half2(__ushort_as_half((unsigned short)0x6480));  // 1152
    __half2 r0 = __hsub2(*reinterpret_cast<const __half2*>(&m0), k1152), r1 = __hsub2(*reinterpret_cast<const __half2*>(&m1), k1152);
    lo = *reinterpret_cast<uint32_t*>(&r0); hi = *reinterpret_cast<uint32_t*>(&r1);
  } else {
    asm("{\n\t.reg .b16 t;\n\tcvt.u16.u32 t, %1;\n\tcvt.rn.f16x2.e4m3x2 %0, t;\n\t}" : "=r"(lo) : "r"(w & 0xffffu));
    asm("{\n\t.reg .b16 t;\n\tcvt.u16.u32 t, %1;\n\tcvt.rn.f16x2.e4m3x2 %0, t;\n\t}" : "=r"(hi) : "r"(w >> 16));
  }
}

template <int WF, int NT>
__global__ void __launch_bounds__(A8V_WARPS * 32, 1)
gemv_a16w8_kernel(const A16vParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int K = P.K, pitch = K + A8V_PAD, xpitch = 2 * K + 16;          // (2K + 16) / 16 odd => conflict-free token rows
  const int xrows = (P.M < 8 * NT) ? P.M : 8 * NT;
  unsigned char* wbuf = smem;                                            // [16][pitch]
  unsigned char* xbuf = wbuf + (size_t)A8V_ROWS * pitch;                 // [xrows][xpitch]
  float* red = reinterpret_cast<float*>(xbuf + (((size_t)xrows * xpitch + 127) & ~(size_t)127));   // [8 warps][NT][16][8]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + A8V_WARPS * NT * 128);                       // [4] weights, [1] x

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int feat0 = blockIdx.x * A8V_ROWS;
  const int kq = K / A8V_NCH;

  if (warp == 0) {
    if (lane == 0) {
      for (int c = 0; c <= A8V_NCH; ++c) ptx::mbar_init(&bars[c], 1);
      ptx::fence_mbar_init();
      for (int c = 0; c < A8V_NCH; ++c) ptx::mbar_arrive_expect_tx(&bars[c], (uint32_t)(A8V_ROWS * kq));
    }
    __syncwarp();
    if (lane < A8V_ROWS) {
      const uint8_t* src = P.W + (size_t)(feat0 + lane) * K;
#pragma unroll
      for (int c = 0; c < A8V_NCH; ++c) bulk_copy_g2s(wbuf + (size_t)lane * pitch + c * kq, src + c * kq, (uint32_t)kq, &bars[c]);
    }
  }
  __syncthreads();
  pdl_wait_prior_grids();
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bars[A8V_NCH], (uint32_t)xrows * (uint32_t)K * 2u);
    for (int m = 0; m < xrows; ++m)
      bulk_copy_g2s(xbuf + (size_t)m * xpitch, reinterpret_cast<const unsigned char*>(P.x) + (size_t)m * K * 2, (uint32_t)K * 2u, &bars[A8V_NCH]);
  }

  float acc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
  const int kw = K / A8V_WARPS;                                         // k (= bytes) of every row this warp owns
  const unsigned char* wa = wbuf + (size_t)g * pitch + (size_t)warp * kw + 16 * t;
  const unsigned char* wb = wa + (size_t)8 * pitch;
  const unsigned char* xa[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const int tok = n * 8 + g;
    xa[n] = xbuf + (size_t)(tok < xrows ? tok : 0) * xpitch + ((size_t)warp * kw + 16 * t) * 2;    // rows >= M: any valid row (discarded)
  }
  ptx::mbar_wait(&bars[warp >> 1], 0);
  ptx::mbar_wait(&bars[A8V_NCH], 0);
#pragma unroll 2
  for (int o = 0; o < kw; o += 64) {
    const uint4 p = *reinterpret_cast<const uint4*>(wa + o);             // 16 weights of row g:     k = o + 16 t .. + 15
    const uint4 q = *reinterpret_cast<const uint4*>(wb + o);             // 16 weights of row g + 8
    uint32_t pl[4], ph[4], ql[4], qh[4];
    widen4<WF>(p.x, pl[0], ph[0]); widen4<WF>(p.y, pl[1], ph[1]); widen4<WF>(p.z, pl[2], ph[2]); widen4<WF>(p.w, pl[3], ph[3]);
    widen4<WF>(q.x, ql[0], qh[0]); widen4<WF>(q.y, ql[1], qh[1]); widen4<WF>(q.z, ql[2], qh[2]); widen4<WF>(q.w, ql[3], qh[3]);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const uint4 x0 = *reinterpret_cast<const uint4*>(xa[n] + 2 * o);        // x of the same 16 k: (k, k+1) (k+2, k+3) ...
      const uint4 x1 = *reinterpret_cast<const uint4*>(xa[n] + 2 * o + 16);
      mma_f16(acc[n], pl[0], ql[0], ph[0], qh[0], x0.x, x0.y);
      mma_f16(acc[n], pl[1], ql[1], ph[1], qh[1], x0.z, x0.w);
      mma_f16(acc[n], pl[2], ql[2], ph[2], qh[2], x1.x, x1.y);
      mma_f16(acc[n], pl[3], ql[3], ph[3], qh[3], x1.z, x1.w);
    }
  }

#pragma unroll
  for (int n = 0; n < NT; ++n) {
    float* r = red + ((size_t)warp * NT + n) * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) r[(g + ((i & 2) ? 8 : 0)) * 8 + 2 * t + (i & 1)] = acc[n][i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NT * 128; i += A8V_WARPS * 32) {
    const int n = i >> 7, col = (i >> 4) & 7, row = i & 15;
    const int m = n * 8 + col;
    if (m >= P.M) continue;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < A8V_WARPS; ++w) v += red[((size_t)w * NT + n) * 128 + row * 8 + col];
    const int f = feat0 + row;
    v = __fmul_rn(v, P.scales_w[f]);
    if (P.bias) v = __half2float(__float2half_rn(v)) + __half2float(reinterpret_cast<const __half*>(P.bias)[f]);
    reinterpret_cast<__half*>(P.y)[(int64_t)m * P.N + f] = __float2half_rn(v);
  }
}

size_t a16v_smem_bytes(int K, int M, int nt) {
  const int xrows = (M < 8 * nt) ? M : 8 * nt;
  return (size_t)A8V_ROWS * (K + A8V_PAD) + ((((size_t)xrows * (2 * K + 16)) + 127) & ~(size_t)127) + (size_t)A8V_WARPS * nt * 128 * 4 +
         (A8V_NCH + 1) * 8 + 16;
}

template <int WF, int NT>
int launch_a16v(const A16vParams& P, cudaStream_t stream) {
  const size_t smem = a16v_smem_bytes(P.K, P.M, NT);
  auto kern = gemv_a16w8_kernel<WF, NT>;
  if (smem > 48 * 1024) GLB_ENSURE_SMEM(kern, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(P.N / A8V_ROWS));
  cfg.blockDim = dim3(A8V_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (g_pdl_enabled.load()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  GLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  count_launch();
  return GLB_OK;
}

size_t a8v_smem_bytes(int K, int nt) {
  return (size_t)(A8V_ROWS + 8 * nt) * (K + A8V_PAD) + (size_t)A8V_WARPS * nt * 128 * 4 + (A8V_NCH + 1) * 8 + 16;
}

template <int FMT, int NT>
int launch_a8v(const A8vParams& P, cudaStream_t stream) {
  const size_t smem = a8v_smem_bytes(P.K, NT);
  auto kern = gemv_a8w8_kernel<FMT, NT>;
  if (smem > 48 * 1024) GLB_ENSURE_SMEM(kern, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(P.N / A8V_ROWS));
  cfg.blockDim = dim3(A8V_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (g_pdl_enabled.load()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  GLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  count_launch();
  return GLB_OK;
}

}  // namespace

// validated on hardware in round 2 (profiles/optin_validation_r2.txt: all A8W8 tests bit-exact / in tolerance through it,
// C4 M <= 8 23.6 -> 13.3 us per eager call incl. the quantiser launch): the default for M <= 16
bool gemv_a8_enabled() { return true; }

bool gemv_a8_supported(const glb_linear& L, int64_t M) {
  if (M < 1 || M > 16) return false;
  const int xd = act_storage_dtype(L.input_dtype);
  const bool is_int = (xd == GLB_INT8), is_f8 = (xd == GLB_FP8E4 || xd == GLB_FP8E5);
  if (!is_int && !is_f8) return false;
  if (is_mx(L.input_dtype)) return false;
  if (L.W_group_mode != 0 || L.elements_per_sample != 1) return false;
  if (is_int && L.wq_dtype != GLB_INT8) return false;
  if (is_f8 && !(L.wq_dtype == GLB_FP8E4 || L.wq_dtype == GLB_FP8E5)) return false;
  if (L.wq_stride0 != 1 || L.wq_stride1 != L.K) return false;            // K-contiguous weights
  if (L.channel_scale_mode < 0 || L.channel_scale_mode > 3) return false;
  if ((L.channel_scale_mode == 1 || L.channel_scale_mode == 3) && (L.scales_dtype != GLB_FP32 || L.scales_numel != L.N)) return false;
  if (L.channel_scale_mode >= 1 && L.meta_dtype != GLB_FP32) return false;
  if (!(L.output_dtype == GLB_FP16 || L.output_dtype == GLB_BF16 || L.output_dtype == GLB_FP32 ||
        (is_int && L.output_dtype == GLB_INT32)))
    return false;
  if (L.N % A8V_ROWS != 0 || L.K % (A8V_WARPS * 64) != 0) return false;  // whole 64-byte blocks per warp
  if (reinterpret_cast<uintptr_t>(L.W_q) & 15) return false;
  if (a8v_smem_bytes((int)L.K, M <= 8 ? 1 : 2) > 220 * 1024) return false;
  return true;
}

// A16W8 weight-only (fp16 activations; int8 or e4m3 weights; per-channel fp32 scales applied before (W_group_mode 2) or
// after (channel_scale_mode 1) the product -- helper.py:156-163)
bool gemv_a16w8_supported(const glb_linear& L, int64_t M) {
  if (M < 1 || M > 16) return false;
  if (L.input_dtype != GLB_FP16 || L.output_dtype != GLB_FP16) return false;
  if (L.W_nbits != 8 || L.elements_per_sample != 1) return false;
  if (!(L.wq_dtype == GLB_INT8 || L.wq_dtype == GLB_FP8E4)) return false;
  if (L.wq_stride0 != 1 || L.wq_stride1 != L.K) return false;
  if (!((L.W_group_mode == 2 && L.channel_scale_mode == 0) || (L.W_group_mode == 0 && L.channel_scale_mode == 1))) return false;
  if (L.scales_dtype != GLB_FP32 || L.scales_numel != L.N || L.group_size != L.K) return false;
  if (L.zeros_numel > 0 && L.W_group_mode != 0 && L.W_group_mode != 2) return false;
  if (L.N % A8V_ROWS != 0 || L.K % (A8V_WARPS * 64) != 0) return false;
  if (reinterpret_cast<uintptr_t>(L.W_q) & 15) return false;
  if (a16v_smem_bytes((int)L.K, (int)M, M <= 8 ? 1 : 2) > 220 * 1024) return false;
  return true;
}

int launch_gemv_a16w8(const glb_linear& L, const void* x, void* y, int64_t M, const void* bias, cudaStream_t stream) {
  GLB_REQUIRE(gemv_a16w8_supported(L, M), GLB_ERR_UNSUPPORTED, "A16W8 decode kernel: unsupported layer configuration");
  GLB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, GLB_ERR_INVALID, "A16W8 decode kernel: x must be 16-byte aligned");
  A16vParams P;
  P.W = reinterpret_cast<const uint8_t*>(L.W_q); P.x = x; P.scales_w = reinterpret_cast<const float*>(L.scales);
  P.y = y; P.bias = bias; P.M = (int)M; P.N = (int)L.N; P.K = (int)L.K;
  const bool f8 = (L.wq_dtype == GLB_FP8E4);
  set_kernel_name(f8 ? "gemv_a16w8_fp8" : "gemv_a16w8_int8");
  if (f8) return (M <= 8) ? launch_a16v<1, 1>(P, stream) : launch_a16v<1, 2>(P, stream);
  return (M <= 8) ? launch_a16v<0, 1>(P, stream) : launch_a16v<0, 2>(P, stream);
}

int launch_gemv_a8(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M, const void* bias,
                   cudaStream_t stream) {
  GLB_REQUIRE(gemv_a8_supported(L, M), GLB_ERR_UNSUPPORTED, "A8W8 decode kernel: unsupported layer configuration");
  GLB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, GLB_ERR_INVALID, "A8W8 decode kernel: x must be 16-byte aligned");
  A8vParams P;
  P.W = reinterpret_cast<const uint8_t*>(L.W_q);
  P.x = reinterpret_cast<const uint8_t*>(x);
  P.scales_w = (L.channel_scale_mode == 1 || L.channel_scale_mode == 3) ? reinterpret_cast<const float*>(L.scales) : nullptr;
  P.scales_x = scales_x; P.y = y; P.bias = bias;
  P.M = (int)M; P.N = (int)L.N; P.K = (int)L.K; P.csm = L.channel_scale_mode; P.out_dtype = L.output_dtype;
  const bool is_int = (act_storage_dtype(L.input_dtype) == GLB_INT8);
  set_kernel_name(is_int ? "gemv_a8w8_int8" : "gemv_a8w8_fp8");
  const int fmt = is_int ? 0 : (1 + 2 * ((L.wq_dtype == GLB_FP8E5) ? 1 : 0) + ((act_storage_dtype(L.input_dtype) == GLB_FP8E5) ? 1 : 0));
#define GLB_A8V(F)                                                                             \
  case F: return (M <= 8) ? launch_a8v<F, 1>(P, stream) : launch_a8v<F, 2>(P, stream);
  switch (fmt) {
    GLB_A8V(0) GLB_A8V(1) GLB_A8V(2) GLB_A8V(3) GLB_A8V(4)
    default: break;
  }
#undef GLB_A8V
  set_error("A8W8 decode kernel: bad format %d", fmt);
  return GLB_ERR_UNSUPPORTED;
}

}  // namespace glb
