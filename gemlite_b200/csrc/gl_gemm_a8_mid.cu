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
__device__ __forceinline__ uint4 ld_dsmem_u32x4(uint32_t cluster_addr) {
  uint4 v;
  asm volatile("ld.shared::cluster.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(cluster_addr) : "memory");
  return v;
}

// NT = 8-token tiles per pass (8: 64 tokens, 4: 32 tokens for M <= 32)
template <bool IS_INT, int NT>
__global__ void __launch_bounds__(AM_NW * 32, 1)
gemm_a8w8_mid_kernel(const A8MidParams P) {
  constexpr int TOK = NT * 8, NREG = 4 * NT, PLANE = 8 * NREG * 32;    // 32-bit words per k-half plane (fragment order)
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* wbuf = smem;                                           // [128][AM_PITCH]
  unsigned char* xbuf = wbuf + AM_WBYTES;                               // [64][AM_PITCH]
  uint32_t* red = reinterpret_cast<uint32_t*>(xbuf);                    // [2 k halves][8 feature blocks][4 NT regs][32 lanes] after the main loop
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
  for (int tok0 = 0; tok0 < P.M; tok0 += TOK, ++pass) {
    const int ntok = (P.M - tok0 < TOK) ? (P.M - tok0) : TOK;
    // ---- x tile: one bulk copy per token row (rows >= M keep whatever the buffer holds; their outputs are never stored) ----
    if (threadIdx.x == 0) ptx::mbar_arrive_expect_tx(&bars[4], (uint32_t)ntok * AM_KC);
    if (pass > 0) ptx::fence_proxy_async_smem();                         // generic accesses of the reduction buffer before async-proxy writes
    __syncthreads();                                                    // expect_tx precedes every complete_tx
    if (threadIdx.x < ntok)
      bulk_g2s(xbuf + (size_t)threadIdx.x * AM_PITCH, P.x + (size_t)(tok0 + threadIdx.x) * P.K + kbeg, AM_KC, &bars[4]);
    if (pass == 0) ptx::mbar_wait(&bars[fb >> 1], 0);
    ptx::mbar_wait(&bars[4], (uint32_t)(pass & 1));

    int32_t acci[NT][4];
    float accf[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) { acci[n][i] = 0; accf[n][i] = 0.f; }
#pragma unroll 2
    for (int o = 0; o < 512; o += 64) {
      const uint4 p = *reinterpret_cast<const uint4*>(wa + o);
      const uint4 q = *reinterpret_cast<const uint4*>(wb + o);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
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

    // ---- reduction: k halves through shared memory, ranks through DSMEM; partial tiles stay in FRAGMENT order
    //      ([feature block][register][lane]: conflict-free stores, 128-bit loads) until the store decodes them ----
    __syncthreads();                                                    // every warp is done reading the x tile
    {
      uint32_t* plane = red + (size_t)kh * PLANE + (size_t)fb * NREG * 32 + lane;
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) plane[(n * 4 + i) * 32] = IS_INT ? (uint32_t)acci[n][i] : __float_as_uint(accf[n][i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PLANE / 4; i += AM_NW * 32) {
      uint4* p0 = reinterpret_cast<uint4*>(red) + i;
      const uint4 a = p0[0], b = p0[PLANE / 4];
      if (IS_INT) p0[0] = make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);          // two's complement: exact
      else p0[0] = make_uint4(__float_as_uint(__uint_as_float(a.x) + __uint_as_float(b.x)), __float_as_uint(__uint_as_float(a.y) + __uint_as_float(b.y)),
                              __float_as_uint(__uint_as_float(a.z) + __uint_as_float(b.z)), __float_as_uint(__uint_as_float(a.w) + __uint_as_float(b.w)));
    }
    if (ks > 1) ptx::cluster_sync(); else __syncthreads();
    {
      const int per = (PLANE / 4) / ks;
      const uint32_t red_addr = ptx::smem_u32(red);
      for (int j = threadIdx.x; j < per; j += AM_NW * 32) {
        const int q4 = (int)rank * per + j;                             // 4 consecutive lanes of one register
        uint4 v;
        if (ks > 1) {
          v = ld_dsmem_u32x4(ptx::mapa(red_addr + (uint32_t)q4 * 16u, 0u));
          for (int r = 1; r < ks; ++r) {
            const uint4 u = ld_dsmem_u32x4(ptx::mapa(red_addr + (uint32_t)q4 * 16u, (uint32_t)r));
            if (IS_INT) { v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
            else {
              v.x = __float_as_uint(__uint_as_float(v.x) + __uint_as_float(u.x)); v.y = __float_as_uint(__uint_as_float(v.y) + __uint_as_float(u.y));
              v.z = __float_as_uint(__uint_as_float(v.z) + __uint_as_float(u.z)); v.w = __float_as_uint(__uint_as_float(v.w) + __uint_as_float(u.w));
            }
          }
        } else {
          v = reinterpret_cast<const uint4*>(red)[q4];
        }
        // decode: element e = q4 * 4 + t -> (feature block, register = n * 4 + i, lane = 4 g + t)
        const int e = q4 * 4;
        const int blk = e / (NREG * 32), reg = (e >> 5) % NREG, gq = (e & 31) >> 2;
        const int i = reg & 3, n = reg >> 2;
        const int f = feat0 + blk * 16 + gq + ((i & 2) ? 8 : 0);
        const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
        const float sb = (P.csm == 1 || P.csm == 3) ? P.scales_w[f] : 1.f;
        float bv = 0.f;
        if (P.bias) {
          if (P.out_dtype == GLB_FP16) bv = __half2float(reinterpret_cast<const __half*>(P.bias)[f]);
          else if (P.out_dtype == GLB_BF16) bv = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.bias)[f]);
          else bv = reinterpret_cast<const float*>(P.bias)[f];
        }
#pragma unroll
        for (int tq = 0; tq < 4; ++tq) {                                // the four lanes of the group: tokens 2 t + (i & 1)
          const int m = tok0 + n * 8 + 2 * tq + (i & 1);
          if (m >= P.M) continue;
          const int64_t oi = (int64_t)m * P.N + f;
          if (IS_INT && P.out_dtype == GLB_INT32 && P.csm == 0 && !P.bias) {
            reinterpret_cast<int32_t*>(P.y)[oi] = (int32_t)vv[tq];
            continue;
          }
          float o = IS_INT ? (float)(int32_t)vv[tq] : __uint_as_float(vv[tq]);        // cvt.rn.f32.s32
          if (P.csm == 1) o = __fmul_rn(o, sb);
          else if (P.csm == 2) o = __fmul_rn(o, P.scales_x[m]);
          else if (P.csm == 3) o = __fmul_rn(o, __fmul_rn(P.scales_x[m], sb));
          if (P.out_dtype == GLB_FP16) {
            if (P.bias) o = __half2float(__float2half_rn(o)) + bv;
            reinterpret_cast<__half*>(P.y)[oi] = __float2half_rn(o);
          } else if (P.out_dtype == GLB_BF16) {
            if (P.bias) o = __bfloat162float(__float2bfloat16_rn(o)) + bv;
            reinterpret_cast<__nv_bfloat16*>(P.y)[oi] = __float2bfloat16_rn(o);
          } else if (P.out_dtype == GLB_FP32) {
            reinterpret_cast<float*>(P.y)[oi] = o + bv;
          } else {
            reinterpret_cast<int32_t*>(P.y)[oi] = (int32_t)(o + bv);
          }
        }
      }
    }
    if (ks > 1) ptx::cluster_sync(); else __syncthreads();             // peers are done with this CTA's partials
  }
}


// ------------------------------------------------------------------------------------------------
// A16W8 weight-only (fp16 / bf16 activations x int8 / e4m3 weights, helper.py:88-170) at 16 < M <= 128 on the same
// skeleton: the 8-bit weights stay resident (128 features x 1024 k per CTA), the activation tile is 32 tokens x 1024 k of
// 16-bit values (pitch 2064 B: conflict-free 128-bit loads), the weights are widened to the compute type in registers
// (prmt + hsub2 / cvt.f16x2.e4m3x2, bf16 through fp32) and fed to mma.sync.m16n8k16 with the k permutation both operands
// share.  Pre-scaling (W_group_mode 2) reproduces the reference's rounding point round16(float(q) * s) in the widening
// (a first version scaled the accumulator instead and sat 1e-4 .. 8e-4 away from the live reference's outputs -- its
// weight-rounding noise); post-scaling (channel_scale_mode 1) multiplies the fp32 accumulator.  M > 32: further passes
// over the resident weights.
constexpr int WM_TOK = 32, WM_NT = 4;
constexpr int WM_XPITCH = 2 * AM_KC + 16;          // 129 x 16 B
constexpr int WM_XBYTES = WM_TOK * WM_XPITCH;      // 66 KB; reused as the [2][8][16][32] fp32 reduction buffer
constexpr int WM_SMEM = AM_WBYTES + WM_XBYTES + 128;
static_assert(2 * WM_TOK * AM_ROWS * 4 <= WM_XBYTES, "reduction buffer must fit in the x tile");

struct W8MidParams {
  const uint8_t* W; const void* x; const float* scales_w; void* y; const void* bias;
  int M, N, K, ks, prescale;
};

template <bool BF16> __device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if (BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 4 weight bytes (k .. k+3) -> two 16-bit-float pairs (k, k+1), (k+2, k+3): exact values, or -- pre-scaling, W_group_mode 2 --
// round16(float(q) * s), the reference's rounding point (b.to(fp32) * scales -> compute dtype)
template <bool BF16, bool F8> __device__ __forceinline__ void widen4_mid(uint32_t w, float sc, bool prescale, uint32_t& lo, uint32_t& hi) {
  uint32_t l16, h16;
  if (!F8) {
    const uint32_t u = w ^ 0x80808080u;
    uint32_t m0, m1;
    asm("prmt.b32 %0, %1, %2, 0x5140;" : "=r"(m0) : "r"(u), "r"(0x64646464u));
    asm("prmt.b32 %0, %1, %2, 0x7362;" : "=r"(m1) : "r"(u), "r"(0x64646464u));
    const __half2 k1152 = __half2half2(__ushort_as_half((unsigned short)0x6480));
    __half2 r0 = __hsub2(*reinterpret_cast<const __half2*>(&m0), k1152), r1 = __hsub2(*reinterpret_cast<const __half2*>(&m1), k1152);
    l16 = *reinterpret_cast<uint32_t*>(&r0); h16 = *reinterpret_cast<uint32_t*>(&r1);
  } else {
    asm("{\n\t.reg .b16 t;\n\tcvt.u16.u32 t, %1;\n\tcvt.rn.f16x2.e4m3x2 %0, t;\n\t}" : "=r"(l16) : "r"(w & 0xffffu));
    asm("{\n\t.reg .b16 t;\n\tcvt.u16.u32 t, %1;\n\tcvt.rn.f16x2.e4m3x2 %0, t;\n\t}" : "=r"(h16) : "r"(w >> 16));
  }
  if (!BF16 && !prescale) { lo = l16; hi = h16; return; }
  float2 a = __half22float2(*reinterpret_cast<__half2*>(&l16)), b = __half22float2(*reinterpret_cast<__half2*>(&h16));
  if (prescale) { a.x *= sc; a.y *= sc; b.x *= sc; b.y *= sc; }
  if (BF16) {
    __nv_bfloat162 r0 = __floats2bfloat162_rn(a.x, a.y), r1 = __floats2bfloat162_rn(b.x, b.y);
    lo = *reinterpret_cast<uint32_t*>(&r0); hi = *reinterpret_cast<uint32_t*>(&r1);
  } else {
    __half2 r0 = __floats2half2_rn(a.x, a.y), r1 = __floats2half2_rn(b.x, b.y);
    lo = *reinterpret_cast<uint32_t*>(&r0); hi = *reinterpret_cast<uint32_t*>(&r1);
  }
}

template <bool BF16, bool F8>
__global__ void __launch_bounds__(AM_NW * 32, 1)
gemm_a16w8_mid_kernel(const W8MidParams P) {
  constexpr int NT = WM_NT, TOK = WM_TOK, NREG = 4 * NT, PLANE = 8 * NREG * 32;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* wbuf = smem;                                           // [128][AM_PITCH]
  unsigned char* xbuf = wbuf + AM_WBYTES;                               // [32][WM_XPITCH]
  uint32_t* red = reinterpret_cast<uint32_t*>(xbuf);
  uint64_t* bars = reinterpret_cast<uint64_t*>(xbuf + WM_XBYTES);       // [4] weights (32 rows each), [1] x

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ks = P.ks;
  const uint32_t rank = (ks > 1) ? ptx::cluster_ctarank() : 0u;
  const int slab = blockIdx.x / ks;
  const int feat0 = slab * AM_ROWS, kbeg = (int)rank * AM_KC;
  const int fb = warp & 7, kh = warp >> 3;

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
  const unsigned char* xa = xbuf + (size_t)g * WM_XPITCH + (size_t)kh * 1024 + 32 * t;
  const unsigned char* xg = reinterpret_cast<const unsigned char*>(P.x);
  const bool pre = (P.prescale != 0);                                  // W_group_mode 2: scale inside the dequantisation
  const float s_lo = pre ? P.scales_w[feat0 + fb * 16 + g] : 1.f, s_hi = pre ? P.scales_w[feat0 + fb * 16 + g + 8] : 1.f;

  int pass = 0;
  for (int tok0 = 0; tok0 < P.M; tok0 += TOK, ++pass) {
    const int ntok = (P.M - tok0 < TOK) ? (P.M - tok0) : TOK;
    if (threadIdx.x == 0) ptx::mbar_arrive_expect_tx(&bars[4], (uint32_t)ntok * 2u * AM_KC);
    if (pass > 0) ptx::fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x < ntok)
      bulk_g2s(xbuf + (size_t)threadIdx.x * WM_XPITCH, xg + ((size_t)(tok0 + threadIdx.x) * P.K + kbeg) * 2, 2u * AM_KC, &bars[4]);
    if (pass == 0) ptx::mbar_wait(&bars[fb >> 1], 0);
    ptx::mbar_wait(&bars[4], (uint32_t)(pass & 1));

    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
#pragma unroll 2
    for (int o = 0; o < 512; o += 64) {
      const uint4 p = *reinterpret_cast<const uint4*>(wa + o);          // 16 weights of row g
      const uint4 q = *reinterpret_cast<const uint4*>(wb + o);          // 16 weights of row g + 8
      uint32_t pl[4], ph[4], ql[4], qh[4];
      widen4_mid<BF16, F8>(p.x, s_lo, pre, pl[0], ph[0]); widen4_mid<BF16, F8>(p.y, s_lo, pre, pl[1], ph[1]);
      widen4_mid<BF16, F8>(p.z, s_lo, pre, pl[2], ph[2]); widen4_mid<BF16, F8>(p.w, s_lo, pre, pl[3], ph[3]);
      widen4_mid<BF16, F8>(q.x, s_hi, pre, ql[0], qh[0]); widen4_mid<BF16, F8>(q.y, s_hi, pre, ql[1], qh[1]);
      widen4_mid<BF16, F8>(q.z, s_hi, pre, ql[2], qh[2]); widen4_mid<BF16, F8>(q.w, s_hi, pre, ql[3], qh[3]);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const unsigned char* xr = xa + (size_t)n * 8 * WM_XPITCH + 2 * o;
        const uint4 x0 = *reinterpret_cast<const uint4*>(xr);           // x of the same 16 k: (k, k+1) (k+2, k+3) ...
        const uint4 x1 = *reinterpret_cast<const uint4*>(xr + 16);
        mma_16816<BF16>(acc[n], pl[0], ql[0], ph[0], qh[0], x0.x, x0.y);
        mma_16816<BF16>(acc[n], pl[1], ql[1], ph[1], qh[1], x0.z, x0.w);
        mma_16816<BF16>(acc[n], pl[2], ql[2], ph[2], qh[2], x1.x, x1.y);
        mma_16816<BF16>(acc[n], pl[3], ql[3], ph[3], qh[3], x1.z, x1.w);
      }
    }

    __syncthreads();                                                    // every warp is done reading the x tile
    {
      uint32_t* plane = red + (size_t)kh * PLANE + (size_t)fb * NREG * 32 + lane;
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) plane[(n * 4 + i) * 32] = __float_as_uint(acc[n][i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PLANE / 4; i += AM_NW * 32) {
      uint4* p0 = reinterpret_cast<uint4*>(red) + i;
      const uint4 a = p0[0], b = p0[PLANE / 4];
      p0[0] = make_uint4(__float_as_uint(__uint_as_float(a.x) + __uint_as_float(b.x)), __float_as_uint(__uint_as_float(a.y) + __uint_as_float(b.y)),
                         __float_as_uint(__uint_as_float(a.z) + __uint_as_float(b.z)), __float_as_uint(__uint_as_float(a.w) + __uint_as_float(b.w)));
    }
    if (ks > 1) ptx::cluster_sync(); else __syncthreads();
    {
      const int per = (PLANE / 4) / ks;
      const uint32_t red_addr = ptx::smem_u32(red);
      for (int j = threadIdx.x; j < per; j += AM_NW * 32) {
        const int q4 = (int)rank * per + j;
        uint4 v;
        if (ks > 1) {
          v = ld_dsmem_u32x4(ptx::mapa(red_addr + (uint32_t)q4 * 16u, 0u));
          for (int r = 1; r < ks; ++r) {                                // fixed order => deterministic
            const uint4 u = ld_dsmem_u32x4(ptx::mapa(red_addr + (uint32_t)q4 * 16u, (uint32_t)r));
            v.x = __float_as_uint(__uint_as_float(v.x) + __uint_as_float(u.x)); v.y = __float_as_uint(__uint_as_float(v.y) + __uint_as_float(u.y));
            v.z = __float_as_uint(__uint_as_float(v.z) + __uint_as_float(u.z)); v.w = __float_as_uint(__uint_as_float(v.w) + __uint_as_float(u.w));
          }
        } else {
          v = reinterpret_cast<const uint4*>(red)[q4];
        }
        const int e = q4 * 4;
        const int blk = e / (NREG * 32), reg = (e >> 5) % NREG, gq = (e & 31) >> 2;
        const int i = reg & 3, n = reg >> 2;
        const int f = feat0 + blk * 16 + gq + ((i & 2) ? 8 : 0);
        const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
        const float sb = pre ? 1.f : P.scales_w[f];                   // post-scaling: channel_scale_mode 1
        float bv = 0.f;
        if (P.bias) bv = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.bias)[f]) : __half2float(reinterpret_cast<const __half*>(P.bias)[f]);
#pragma unroll
        for (int tq = 0; tq < 4; ++tq) {
          const int m = tok0 + n * 8 + 2 * tq + (i & 1);
          if (m >= P.M) continue;
          const int64_t oi = (int64_t)m * P.N + f;
          float o = pre ? __uint_as_float(vv[tq]) : __fmul_rn(__uint_as_float(vv[tq]), sb);
          if (BF16) {
            if (P.bias) o = __bfloat162float(__float2bfloat16_rn(o)) + bv;
            reinterpret_cast<__nv_bfloat16*>(P.y)[oi] = __float2bfloat16_rn(o);
          } else {
            if (P.bias) o = __half2float(__float2half_rn(o)) + bv;
            reinterpret_cast<__half*>(P.y)[oi] = __float2half_rn(o);
          }
        }
      }
    }
    if (ks > 1) ptx::cluster_sync(); else __syncthreads();
  }
}

}  // namespace

bool gemm_w8_mid_supported(const glb_linear& L, int64_t M) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("GLB_MID_KERNEL"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return false;
  if (M <= 16 || M > 128) return false;
  if (!(L.input_dtype == GLB_FP16 || L.input_dtype == GLB_BF16) || L.output_dtype != L.input_dtype) return false;
  if (L.W_nbits != 8 || L.elements_per_sample != 1) return false;
  if (!(L.wq_dtype == GLB_INT8 || L.wq_dtype == GLB_FP8E4)) return false;
  if (L.wq_stride0 != 1 || L.wq_stride1 != L.K) return false;
  if (!((L.W_group_mode == 2 && L.channel_scale_mode == 0) || (L.W_group_mode == 0 && L.channel_scale_mode == 1))) return false;
  if (L.scales_dtype != GLB_FP32 || L.scales_numel != L.N || L.group_size != L.K) return false;
  if (!(L.K == 1024 || L.K == 2048 || L.K == 4096 || L.K == 8192)) return false;
  if (L.N % AM_ROWS != 0) return false;
  if (reinterpret_cast<uintptr_t>(L.W_q) & 15) return false;
  return true;
}

int launch_gemm_w8_mid(const glb_linear& L, const void* x, void* y, int64_t M, const void* bias, cudaStream_t stream) {
  GLB_REQUIRE(gemm_w8_mid_supported(L, M), GLB_ERR_UNSUPPORTED, "A16W8 batched-decode kernel: unsupported layer configuration");
  GLB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, GLB_ERR_INVALID, "A16W8 batched-decode kernel: x must be 16-byte aligned");
  W8MidParams P;
  P.W = reinterpret_cast<const uint8_t*>(L.W_q); P.x = x; P.scales_w = reinterpret_cast<const float*>(L.scales);
  P.y = y; P.bias = bias; P.M = (int)M; P.N = (int)L.N; P.K = (int)L.K; P.ks = (int)(L.K / AM_KC);
  P.prescale = (L.W_group_mode == 2) ? 1 : 0;
  const bool bf = (L.input_dtype == GLB_BF16), f8 = (L.wq_dtype == GLB_FP8E4);
  auto kern = bf ? (f8 ? gemm_a16w8_mid_kernel<true, true> : gemm_a16w8_mid_kernel<true, false>)
                 : (f8 ? gemm_a16w8_mid_kernel<false, true> : gemm_a16w8_mid_kernel<false, false>);
  GLB_ENSURE_SMEM(kern, WM_SMEM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(L.N / AM_ROWS) * (unsigned)P.ks);
  cfg.blockDim = dim3(AM_NW * 32);
  cfg.dynamicSmemBytes = WM_SMEM;
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
  set_kernel_name(f8 ? "gemm_a16w8_fp8_mma_mid" : "gemm_a16w8_int8_mma_mid");
  GLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  count_launch();
  return GLB_OK;
}

namespace {
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
  auto kern = (M <= 32) ? (is_int ? gemm_a8w8_mid_kernel<true, 4> : gemm_a8w8_mid_kernel<false, 4>)
                        : (is_int ? gemm_a8w8_mid_kernel<true, 8> : gemm_a8w8_mid_kernel<false, 8>);
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
