// Decode-regime kernels (small M, HBM-bound): fused unpack + dequant + dot.
//
// Replaces gemlite/triton_kernels/gemv_revsplitK_kernels.py (packed W_nbits < 8, the reference's
// default at M == 1, core.py:105) and serves forward_manual("GEMV"/"GEMV_REVSPLITK").
//
// Layout consumed (unchanged reference layout, core.py:387-392): W_q int32 (K/8, N), N contiguous,
// word (r, n) holds codes k = 8r .. 8r+7 of column n in nibbles 0..7; scales / zeros (K/gs, N).
//
// Design for B200 (148 SMs, ~6.5 TB/s measured => ~23 B/clk/SM => ~46 four-bit weights/clk/SM):
//  * a 32-column slab of the weight matrix is owned by a CLUSTER of KSPLIT (1 or 2) CTAs that split
//    K; partial sums meet in the leader's shared memory through DSMEM (no atomics, no memset,
//    deterministic). N=4096 => 256 CTAs of 8 warps: two CTAs per SM, every SM busy;
//  * a warp-load is 4 packed rows x 128 contiguous bytes (uint4 per lane, ld.global.nc no-allocate);
//    every lane issues ALL its weight loads (8 x 128-bit per batch) and the scale / zero vectors of
//    its groups before it consumes anything;
//  * weights and metadata do not depend on the previous layer: they are requested BEFORE
//    griddepcontrol.wait, x (which does) after it => with programmatic dependent launch the next
//    layer's HBM traffic overlaps this layer's tail;
//  * the multiply-accumulate is issue-bound on CUDA cores at this bandwidth (>= 1.9 issue slots per
//    weight for lop3 + sub + fma + flush), so the dot products run on the tensor pipe instead:
//    nibbles -> fp16 integers by the 0x6400 magic (1 shift + 4 lop3 per 8 weights), exact
//    (h - 1024) / (h/16 - 64), then mma.sync.m16n8k16 with fp32 accumulation; the k order inside a
//    word is permuted consistently on the x side at staging time, so no shuffles are needed;
//  * per group g:  y += s_g * sum_k x_k q_k  +  z'_g * sum_k x_k   (mode 4; modes 2/3 analogous) --
//    the group-wise factorisation of sum_k x_k * fma(q_k, s_g, z'_g), applied in fp32; bias fused.
// The same kernel handles 1 <= M <= 16 rows of x (rows of the MMA A operand).
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include <unordered_map>

#include "gl_common.cuh"
#include "gl_ptx.cuh"

namespace glb {
namespace {

constexpr int GV_THREADS = 256;
constexpr int GV_WARPS = 8;
constexpr int GV_COLS = 32;        // columns per cluster
constexpr int GV_R = 8;            // uint4 loads a lane keeps in flight per batch
#ifndef GLB_GEMV_FAST
#define GLB_GEMV_FAST 1            // 1: offset-carrying nibble unpack (no subtract), fp16 only
#endif
#ifndef GLB_GEMV_TIMELINE
#define GLB_GEMV_TIMELINE 0        // 1: record %globaltimer phase stamps (tools/decode_timeline.py)
#endif
#ifndef GLB_GEMV_TMA
#define GLB_GEMV_TMA 1             // 1: weights via TMA into shared memory (co-resident layers under PDL)
#endif

struct GemvParams {
  void* const* y_dst = nullptr;   // glb_matmul_bcast: destinations (device-accessible pointers), else null
  int n_dst = 0;
  long long y_col0 = 0;
  const uint32_t* W;     // (K/8, N) int32
  const void* scales;    // (G, N)
  const void* zeros;     // (G, N) or null
  const void* x;         // (M, K)
  void* y;               // (M, N)
  const void* bias;      // (N) or null
  int K, N, M;
  int group_size;        // 32 * RPG
  int mode;              // 2, 3, 4
};

template <bool BF16> __device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (BF16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t mask, uint32_t orv) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(mask), "r"(orv));  // (a & mask) | orv
  return r;
}

// 8 nibbles of one word -> 4 registers of fp16x2 / bf16x2 values.
// Pairing: h[0] = (k0,k4), h[1] = (k1,k5), h[2] = (k2,k6), h[3] = (k3,k7).
//   exact form : the integers n themselves (magic 0x6400 = 1024, then h-1024 resp. h/16-64)
//   FAST form  : fp16 only -- the raw magic values 1024+n (h[0],h[2]) and 1024+16n (h[1],h[3]); the x side
//                carries x/16 in the (k1,k5),(k3,k7) slots and 1024*sum(x') is subtracted per group.
template <bool BF16, bool FAST> __device__ __forceinline__ void nibbles_to_f16x2(uint32_t q, uint32_t (&h)[4]) {
  if (!BF16) {
    const uint32_t t = q >> 8;
    const uint32_t h0 = lop3_and_or(q, 0x000F000Fu, 0x64006400u);   // 1024 + n
    const uint32_t h1 = lop3_and_or(q, 0x00F000F0u, 0x64006400u);   // 1024 + 16 n
    const uint32_t h2 = lop3_and_or(t, 0x000F000Fu, 0x64006400u);
    const uint32_t h3 = lop3_and_or(t, 0x00F000F0u, 0x64006400u);
    if (FAST) { h[0] = h0; h[1] = h1; h[2] = h2; h[3] = h3; return; }
    const __half2 k1024 = __half2half2(__ushort_as_half((unsigned short)0x6400));
    const __half2 k16th = __half2half2(__ushort_as_half((unsigned short)0x2C00));   // 1/16
    const __half2 km64 = __half2half2(__ushort_as_half((unsigned short)0xD400));    // -64
    __half2 r0 = __hsub2(*reinterpret_cast<const __half2*>(&h0), k1024);
    __half2 r1 = __hfma2(*reinterpret_cast<const __half2*>(&h1), k16th, km64);
    __half2 r2 = __hsub2(*reinterpret_cast<const __half2*>(&h2), k1024);
    __half2 r3 = __hfma2(*reinterpret_cast<const __half2*>(&h3), k16th, km64);
    h[0] = *reinterpret_cast<uint32_t*>(&r0); h[1] = *reinterpret_cast<uint32_t*>(&r1);
    h[2] = *reinterpret_cast<uint32_t*>(&r2); h[3] = *reinterpret_cast<uint32_t*>(&r3);
  } else {
    // bf16 has 7 mantissa bits: 0x4300 = 128.0 with ulp 1 -> one nibble per shift
    const __nv_bfloat162 k128 = __bfloat162bfloat162(__ushort_as_bfloat16((unsigned short)0x4300));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t v = lop3_and_or(q >> (4 * i), 0x000F000Fu, 0x43004300u);
      __nv_bfloat162 r = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&v), k128);
      h[i] = *reinterpret_cast<uint32_t*>(&r);
    }
  }
}


// Same unpack with the one shift per word issued as mul.hi (FMA pipe) instead of SHF (ALU pipe, which the four
// LOP3 already saturate: both pipes issue one warp instruction every two cycles per scheduler).
template <bool BF16, bool FAST> __device__ __forceinline__ void nibbles_to_f16x2_m1(uint32_t q, uint32_t (&h)[4]) {
  if (!BF16 && FAST) {
    uint32_t t;
    asm("mul.hi.u32 %0, %1, 0x01000000;" : "=r"(t) : "r"(q));     // q >> 8
    h[0] = lop3_and_or(q, 0x000F000Fu, 0x64006400u);
    h[1] = lop3_and_or(q, 0x00F000F0u, 0x64006400u);
    h[2] = lop3_and_or(t, 0x000F000Fu, 0x64006400u);
    h[3] = lop3_and_or(t, 0x00F000F0u, 0x64006400u);
  } else {
    nibbles_to_f16x2<BF16, FAST>(q, h);
  }
}

template <bool BF16> __device__ __forceinline__ float cvt_meta(uint16_t raw) {
  if (BF16) return __bfloat162float(__ushort_as_bfloat16(raw));
  return __half2float(__ushort_as_half(raw));
}
template <bool BF16> __device__ __forceinline__ void st_out(void* p, int64_t i, float v) {
  if (BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}
template <bool BF16> __device__ __forceinline__ float rnd_out(float v) {
  if (BF16) return __bfloat162float(__float2bfloat16_rn(v));
  return __half2float(__float2half_rn(v));
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f32(float* local_smem_ptr, uint32_t rank, float v) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local_smem_ptr), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

// Shared memory (per CTA, Kc = K / KSPLIT):
//   xs[MT][Kc] permuted 16-bit activations | zero[4] (16 B of zeros)
//   | red[8][MT][32] fp32 per-warp partials | peer[MT][32] fp32 partials of the other cluster rank
//
// MMA mapping (mma.sync.m16n8k16, D[16x8] = A[16x16] B[16x8]):
//   A = weights^T : rows 0-7 <-> column col0+4g+cA, rows 8-15 <-> column col0+4g+cB of lane group g
//                   (one MMA consumes TWO of the lane's four words -> 256 weights per instruction);
//   B = x^T       : column j <-> token j (8 tokens per MMA; NB = 2 token blocks for M <= 16);
//   k slots (2t,2t+1 | 2t+8,2t+9) of lane t <-> nibble pairs of ITS packed row (run*4 + t), the same
//   permutation being applied to x when it is staged, so no data ever moves between lanes.
template <bool BF16, int MT, int RPG /*4-row runs per group = gs/32*/, int KSPLIT, bool FAST>
__global__ void __launch_bounds__(GV_THREADS, 2) gemv_w4_mma_kernel(const GemvParams P) {
  static_assert(!(BF16 && FAST), "the offset-carrying unpack needs fp16's 10-bit mantissa");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int GPB = (GV_R / RPG) > 0 ? (GV_R / RPG) : 1;   // groups a warp covers per batch
  constexpr int RUNS = GPB * RPG;                            // uint4 loads per batch
  constexpr int NB = (MT > 8) ? 2 : 1;                       // token blocks of 8
  const int K = P.K, N = P.N;
  const int G = K / P.group_size;
  const int Kc = K / KSPLIT, Gc = G / KSPLIT;
  const uint32_t rank = (KSPLIT > 1) ? cluster_ctarank() : 0u;
  const int slab = (KSPLIT > 1) ? (blockIdx.x / KSPLIT) : blockIdx.x;
  const int kbeg = (int)rank * Kc, gbeg = (int)rank * Gc;

  uint16_t* xs = reinterpret_cast<uint16_t*>(smem_raw);
  uint32_t* zero16 = reinterpret_cast<uint32_t*>(smem_raw + (size_t)MT * Kc * 2);
  float* red = reinterpret_cast<float*>(smem_raw + (size_t)MT * Kc * 2 + 16);
  float* peer = red + GV_WARPS * MT * GV_COLS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g8 = lane >> 2, t4 = lane & 3;        // mma lane coordinates
  const int col0 = slab * GV_COLS;
  constexpr int rpg = RPG * 4;                     // packed rows per group
  const int nbatches = (Gc + GV_WARPS * GPB - 1) / (GV_WARPS * GPB);

  // ---- 1. everything that does not depend on the previous kernel: weights + scale/zero vectors ----
  // lane (g8, t4) loads packed row (run*4 + t4), columns col0 + 4*g8 .. +3, and owns the outputs of exactly
  // those four columns (tokens 2*t4, 2*t4+1 of each token block): their scales/zeros are one 8-byte vector.
  const uint32_t* wbase = P.W + (size_t)col0 + 4 * g8 + (size_t)t4 * N;
  const uint16_t* sbase = reinterpret_cast<const uint16_t*>(P.scales) + col0 + 4 * g8;
  const uint16_t* zbase = reinterpret_cast<const uint16_t*>(P.zeros) + col0 + 4 * g8;
  uint4 q[RUNS];
  uint2 sv[GPB], zv[GPB];
  auto issue = [&](int batch) {
    const int gl0 = (batch * GV_WARPS + warp) * GPB;           // first group (CTA-local) of this warp in the batch
    if (gl0 < Gc) {
      const uint32_t* wp = wbase + (size_t)(gbeg + gl0) * rpg * N;
#pragma unroll
      for (int i = 0; i < RUNS; ++i) q[i] = ldg_nc_v4(wp + (size_t)(i * 4) * N);   // consecutive 4-row runs
#pragma unroll
      for (int j = 0; j < GPB; ++j) {
        sv[j] = __ldg(reinterpret_cast<const uint2*>(sbase + (size_t)(gbeg + gl0 + j) * N));
        if (P.mode != 2) zv[j] = __ldg(reinterpret_cast<const uint2*>(zbase + (size_t)(gbeg + gl0 + j) * N));
      }
    }
  };
  issue(0);

  // ---- 2. x depends on the previous kernel: wait for it only now (PDL), then stage it permuted ----
  pdl_wait_prior_grids();
  {
    // word r of row m: source k = 8r + j -> slot order [0,4,1,5,2,6,3,7] so that an LDS.128 yields
    // {(k0,k4), (k1,k5), (k2,k6), (k3,k7)} = the B-fragment pairs matching nibbles_to_f16x2
    const uint16_t* xg = reinterpret_cast<const uint16_t*>(P.x) + kbeg;
    const int words = Kc >> 3;
    for (int i = threadIdx.x; i < MT * words; i += GV_THREADS) {
      const int m = (MT == 1) ? 0 : (i / words), r = (MT == 1) ? i : (i - m * words);
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (m < P.M) {
        const uint4 v = *reinterpret_cast<const uint4*>(xg + (size_t)m * K + (size_t)r * 8);
        o.x = __byte_perm(v.x, v.z, 0x5410);   // (k0,k4)
        o.y = __byte_perm(v.x, v.z, 0x7632);   // (k1,k5)
        o.z = __byte_perm(v.y, v.w, 0x5410);   // (k2,k6)
        o.w = __byte_perm(v.y, v.w, 0x7632);   // (k3,k7)
        if (FAST) {                            // the 1024+16n slots see x/16 (exact unless subnormal)
          const __half2 k16th = __half2half2(__ushort_as_half((unsigned short)0x2C00));
          __half2 a = __hmul2(*reinterpret_cast<__half2*>(&o.y), k16th), b2 = __hmul2(*reinterpret_cast<__half2*>(&o.w), k16th);
          o.y = *reinterpret_cast<uint32_t*>(&a); o.w = *reinterpret_cast<uint32_t*>(&b2);
        }
      }
      *reinterpret_cast<uint4*>(xs + (size_t)m * Kc + (size_t)r * 8) = o;
    }
    if (threadIdx.x < 4) zero16[threadIdx.x] = 0u;
  }
  __syncthreads();
  pdl_launch_dependents();

  // ---- 3. main loop ----
  // out[nb][c][tk]: column col0 + 4*g8 + c, token nb*8 + 2*t4 + tk
  float out[NB][4][2];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int c = 0; c < 4; ++c) out[nb][c][0] = out[nb][c][1] = 0.f;
  // B-fragment source: x row of token (nb*8 + g8); lanes whose token does not exist read 16 bytes of zeros (stride 0)
  const unsigned char* xb[NB];
  int xst[NB];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    const bool ok = (nb * 8 + g8) < MT;
    xb[nb] = ok ? reinterpret_cast<const unsigned char*>(xs + (size_t)(nb * 8 + g8) * Kc) + t4 * 16
                : reinterpret_cast<const unsigned char*>(zero16);
    xst[nb] = ok ? 64 : 0;                                        // bytes per 4-row run
  }
  // "ones" A operand: one extra MMA per B fragment yields the group sums of x in the accumulator layout:
  //   rows 0-7  (d0,d1): sum_k x_k   (in FAST mode the slots that carry x/16 are weighted by 16)
  //   rows 8-15 (d2,d3): sum_k x'_k  (FAST only: what the 1024 offset multiplies)
  const uint32_t one2 = BF16 ? 0x3F803F80u : 0x3C003C00u;
  const uint32_t ones_a[4] = {one2, one2, FAST ? 0x4C004C00u /* (16,16) */ : one2, one2};

  for (int batch = 0; batch < nbatches; ++batch) {
    const int gl0 = (batch * GV_WARPS + warp) * GPB;
    if (gl0 < Gc) {
#pragma unroll
      for (int j = 0; j < GPB; ++j) {
        float dacc[NB][2][4], dsum[NB][4];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
          for (int pr = 0; pr < 2; ++pr) dacc[nb][pr][0] = dacc[nb][pr][1] = dacc[nb][pr][2] = dacc[nb][pr][3] = 0.f;
          dsum[nb][0] = dsum[nb][1] = dsum[nb][2] = dsum[nb][3] = 0.f;
        }
#pragma unroll
        for (int r = 0; r < RPG; ++r) {
          const int i = j * RPG + r;
          const int run = (gl0 + j) * RPG + r;                           // 4-row run inside this CTA's K range
          uint32_t h[4][4];
          nibbles_to_f16x2<BF16, FAST>(q[i].x, h[0]);
          nibbles_to_f16x2<BF16, FAST>(q[i].y, h[1]);
          nibbles_to_f16x2<BF16, FAST>(q[i].z, h[2]);
          nibbles_to_f16x2<BF16, FAST>(q[i].w, h[3]);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const uint4 xv = *reinterpret_cast<const uint4*>(xb[nb] + (size_t)run * xst[nb]);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {                             // hh = 0: (k0,k4),(k1,k5)   hh = 1: (k2,k6),(k3,k7)
              const uint32_t b0 = hh ? xv.z : xv.x, b1 = hh ? xv.w : xv.y;
              const uint32_t a01[4] = {h[0][2 * hh], h[1][2 * hh], h[0][2 * hh + 1], h[1][2 * hh + 1]};
              const uint32_t a23[4] = {h[2][2 * hh], h[3][2 * hh], h[2][2 * hh + 1], h[3][2 * hh + 1]};
              mma16816<BF16>(dacc[nb][0], a01, b0, b1);
              mma16816<BF16>(dacc[nb][1], a23, b0, b1);
              if (P.mode != 2 || FAST) mma16816<BF16>(dsum[nb], ones_a, b0, b1);
            }
          }
        }
        // ---- end of group: y += s * sum(x q) + z' * sum(x), z' = zeros (mode 4) | -zeros*s (mode 3) | 0 (mode 2)
        float sc[4], zc[4];
        {
          const uint32_t sw[2] = {sv[j].x, sv[j].y}, zw[2] = {zv[j].x, zv[j].y};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            sc[c] = cvt_meta<BF16>((uint16_t)((c & 1) ? (sw[c >> 1] >> 16) : (sw[c >> 1] & 0xffffu)));
            zc[c] = 0.f;
            if (P.mode != 2) {
              zc[c] = cvt_meta<BF16>((uint16_t)((c & 1) ? (zw[c >> 1] >> 16) : (zw[c >> 1] & 0xffffu)));
              if (P.mode == 3) zc[c] = -zc[c] * sc[c];
            }
          }
        }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
          for (int tk = 0; tk < 2; ++tk) {
            const float sx = dsum[nb][tk];
            const float off = FAST ? 1024.f * dsum[nb][2 + tk] : 0.f;
            // pair 0 = columns (c0 rows 0-7 -> d0,d1 ; c1 rows 8-15 -> d2,d3), pair 1 = (c2, c3)
            out[nb][0][tk] = fmaf(sc[0], dacc[nb][0][tk] - off, fmaf(zc[0], sx, out[nb][0][tk]));
            out[nb][1][tk] = fmaf(sc[1], dacc[nb][0][2 + tk] - off, fmaf(zc[1], sx, out[nb][1][tk]));
            out[nb][2][tk] = fmaf(sc[2], dacc[nb][1][tk] - off, fmaf(zc[2], sx, out[nb][2][tk]));
            out[nb][3][tk] = fmaf(sc[3], dacc[nb][1][2 + tk] - off, fmaf(zc[3], sx, out[nb][3][tk]));
          }
        }
      }
    }
    if (batch + 1 < nbatches) issue(batch + 1);
  }

  // ---- 4. cross-warp reduction, cluster (split-K) reduction through DSMEM, epilogue ----
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int tk = 0; tk < 2; ++tk) {
      const int m = nb * 8 + 2 * t4 + tk;
      if (m < MT) {
#pragma unroll
        for (int c = 0; c < 4; ++c) red[(warp * MT + m) * GV_COLS + 4 * g8 + c] = out[nb][c][tk];
      }
    }
  __syncthreads();
  float part[(MT * GV_COLS + GV_THREADS - 1) / GV_THREADS];
#pragma unroll
  for (int u = 0; u < (MT * GV_COLS + GV_THREADS - 1) / GV_THREADS; ++u) {
    const int i = u * GV_THREADS + threadIdx.x;
    float v = 0.f;
    if (i < MT * GV_COLS) {
      const int m = i / GV_COLS, c = i - m * GV_COLS;
#pragma unroll
      for (int w = 0; w < GV_WARPS; ++w) v += red[(w * MT + m) * GV_COLS + c];
      if (KSPLIT > 1 && rank != 0) st_cluster_f32(&peer[i], 0u, v);
    }
    part[u] = v;
  }
  if (KSPLIT > 1) cluster_sync_all();
  if (rank == 0) {
#pragma unroll
    for (int u = 0; u < (MT * GV_COLS + GV_THREADS - 1) / GV_THREADS; ++u) {
      const int i = u * GV_THREADS + threadIdx.x;
      if (i < MT * GV_COLS) {
        const int m = i / GV_COLS, c = i - m * GV_COLS;
        if (m < P.M) {
          float v = part[u];
          if (KSPLIT > 1) v += peer[i];
          if (P.bias) v = rnd_out<BF16>(v) + cvt_meta<BF16>(reinterpret_cast<const uint16_t*>(P.bias)[col0 + c]);
          st_out<BF16>(P.y, (int64_t)m * N + col0 + c, v);
        }
      }
    }
  }
}

template <int MT, int KSPLIT> size_t gemv_smem_bytes(int K) {
  const int Kc = K / KSPLIT;
  return (size_t)MT * Kc * 2 + 16 + (size_t)GV_WARPS * MT * GV_COLS * 4 + (size_t)MT * GV_COLS * 4;
}

template <bool BF16, int MT, int RPG, int KSPLIT, bool FAST>
int launch_w4(const GemvParams& P, cudaStream_t stream) {
  const size_t smem = gemv_smem_bytes<MT, KSPLIT>(P.K);
  auto kern = gemv_w4_mma_kernel<BF16, MT, RPG, KSPLIT, FAST>;
  if (smem > 48 * 1024) GLB_ENSURE_SMEM(kern, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(P.N / GV_COLS) * KSPLIT);
  cfg.blockDim = dim3(GV_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (KSPLIT > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = KSPLIT; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
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


// =================================================================================================
// Decode kernel, TMA flavour (the default for M <= 8): same arithmetic and MMA mapping as
// gemv_w4_mma_kernel above, but the weights / scales / zeros of the CTA are pulled into SHARED MEMORY
// by TMA (cp.async.bulk.tensor) right at kernel start -- before griddepcontrol.wait -- instead of into
// registers.  With <= 64 registers per thread four CTAs fit on an SM, so under programmatic dependent
// launch the CTAs of layer i+1 are resident (their 32 KB of weights landing in shared memory) while
// layer i is still computing: HBM streaming of the next layer overlaps the dependent x -> y chain.
// =================================================================================================
constexpr int GT_NCH = 4;   // weight chunks, one mbarrier each

// diagnostics: per-launch phase timestamps (ns, %globaltimer) of CTA 0 / last CTA, warp 0 -- enabled by GLB_GEMV_DEBUG bit 2
__device__ unsigned long long g_timeline[8192];
__device__ unsigned int g_timeline_pos;
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// mbarrier wait for a whole warp.  With -DGLB_GEMV_LEADER_WAIT=1 lane 0 polls and the warp reconverges on
// __syncwarp() (the leader's acquire + the warp barrier order the other lanes' reads after the TMA writes): one
// try_wait stream per warp instead of 32 lanes hitting the same shared-memory word.  Compile-time experiment, not yet
// measured; the default build polls with every lane.
#ifndef GLB_GEMV_LEADER_WAIT
#define GLB_GEMV_LEADER_WAIT 0
#endif
#ifndef GLB_GEMV_PREWAIT_ALL
#define GLB_GEMV_PREWAIT_ALL 0      // 1: all mbarrier waits before griddepcontrol.wait (compile-time experiment, see the kernel)
#endif
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
#if GLB_GEMV_LEADER_WAIT
  if ((threadIdx.x & 31) == 0) ptx::mbar_wait(bar, parity);
  __syncwarp();
#else
  ptx::mbar_wait(bar, parity);
#endif
}

struct GemvTmaParams {
  const void* x; void* y; const void* bias;
  int K, N, M, group_size, mode;
  int late_trigger;      // dependents are signalled 0: at kernel start, 1: after the main loop, 2: right after griddepcontrol.wait
  int bulk1d;            // 1: descriptor-free cp.async.bulk row copies instead of tensor-map TMA for the weights
  const uint32_t* W;     // (K/8, N) int32 (for bulk1d)
  const void* scales;    // (G, N)   (META_REGS flavour reads them with plain loads)
  const void* zeros;     // (G, N) or null
  int chunk_shift;       // log2(rows per weight chunk)
  int debug;             // diagnostics only: bit 0 skip the main loop, bit 1 skip the weight copies
  int issue_mode;        // who issues the TMA copies (see the kernel)
};

constexpr int GT_GPW = 4;    // META_REGS: groups a warp may own (their scale / zero vectors live in registers)

template <bool BF16, int MT, int RPG, int KSPLIT, bool FAST, int NW, bool META_REGS>
__global__ void __launch_bounds__(NW * 32, (NW == 8) ? (META_REGS ? 3 : 4) : 2)
gemv_w4_tma_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_s,
                   const __grid_constant__ CUtensorMap tmap_z, const GemvTmaParams P) {
  static_assert(MT <= 8, "one token block");
  static_assert(!(BF16 && FAST), "the offset-carrying unpack needs fp16's 10-bit mantissa");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int rpg = RPG * 4;                     // packed rows per group
  const int K = P.K, N = P.N;
  const int G = K / P.group_size;
  const int Kc = K / KSPLIT, Gc = G / KSPLIT, rows_c = Kc >> 3;
  const uint32_t rank = (KSPLIT > 1) ? cluster_ctarank() : 0u;
  const int slab = (KSPLIT > 1) ? (blockIdx.x / KSPLIT) : blockIdx.x;
  const int kbeg = (int)rank * Kc, gbeg = (int)rank * Gc;
  const int col0 = slab * GV_COLS;
  const int nch = (rows_c % (GT_NCH * rpg) == 0) ? GT_NCH : 1;
  const int chunk_rows = rows_c / nch;

  // shared memory carve-up (weight buffer 1024-byte aligned: it is written by TMA with the 128B swizzle)
  unsigned char* wbuf = smem_raw;                                        // rows_c x 128 B (no integer round-trip: keeps LDS)
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0u) __trap();
  const size_t meta_bytes_smem = META_REGS ? 0 : (((size_t)Gc * 64 + 127) & ~(size_t)127);
  unsigned char* sbuf = wbuf + (size_t)rows_c * 128;                     // Gc x 64 B (padded to 128); absent with META_REGS
  unsigned char* zbuf = sbuf + meta_bytes_smem;
  uint16_t* xs = reinterpret_cast<uint16_t*>(zbuf + meta_bytes_smem);
  uint32_t* zero16 = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(xs) + (size_t)MT * Kc * 2);
  float* red = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(zero16) + 16);
  float* peer = red + NW * MT * GV_COLS;
  float* gsum = peer + MT * GV_COLS;                                     // [MT][Gc][2] group sums of x
  uint64_t* bars = reinterpret_cast<uint64_t*>(gsum + (size_t)MT * Gc * 2);   // [GT_NCH] weights, [1] metadata

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g8 = lane >> 2, t4 = lane & 3;

#if GLB_GEMV_TIMELINE
  unsigned long long ts[8];
  const bool rec = (P.debug & 4) && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
#else
  unsigned long long ts[1];
  constexpr bool rec = false;
#endif
  if (rec) ts[GLB_GEMV_TIMELINE ? 0 : 0] = gtimer();
  // ---- 1. weights + metadata -> shared memory by TMA (independent of the previous kernel) ----
  // P.issue_mode 0: thread 0 issues everything (one serial chain of ~12 asynchronous operations);
  //          bit 1: wait for the metadata before griddepcontrol.wait;
//          bit 0: the LAST GT_NCH + 1 warps issue one copy each, in parallel (each lane initialises the barrier it
  //                 uses): shorter start -> first-byte latency, and high warp ids are favoured by the scheduler while
  //                 the SM is still busy with the previous layer's main loop.
  if ((P.issue_mode & 1) == 0 || P.bulk1d) {
    if (threadIdx.x == 0) {
      ptx::prefetch_tmap(&tmap_w);
      ptx::prefetch_tmap(&tmap_s);
      for (int c = 0; c <= GT_NCH; ++c) ptx::mbar_init(&bars[c], 1);
      ptx::fence_mbar_init();
      if (!META_REGS) {
        const uint32_t meta_bytes = (uint32_t)Gc * 64u * (P.mode != 2 ? 2u : 1u);
        ptx::mbar_arrive_expect_tx(&bars[GT_NCH], meta_bytes);
        ptx::tma_load_2d(sbuf, &tmap_s, &bars[GT_NCH], col0, gbeg);
        if (P.mode != 2) ptx::tma_load_2d(zbuf, &tmap_z, &bars[GT_NCH], col0, gbeg);
      }
      for (int c = 0; c < ((P.debug & 2) ? 0 : nch); ++c) {
        ptx::mbar_arrive_expect_tx(&bars[c], (uint32_t)chunk_rows * 128u);
        if (!P.bulk1d) ptx::tma_load_2d(wbuf + (size_t)c * chunk_rows * 128, &tmap_w, &bars[c], col0, (kbeg >> 3) + c * chunk_rows);
      }
    }
  } else if (lane == 0 && warp >= NW - (GT_NCH + 1)) {
    const int c = warp - (NW - (GT_NCH + 1));            // 0 .. GT_NCH-1: weight chunk c, GT_NCH: metadata
    ptx::mbar_init(&bars[c], 1);
    ptx::fence_mbar_init();
    if (c == GT_NCH) {
      if (!META_REGS) {
        ptx::prefetch_tmap(&tmap_s);
        const uint32_t meta_bytes = (uint32_t)Gc * 64u * (P.mode != 2 ? 2u : 1u);
        ptx::mbar_arrive_expect_tx(&bars[GT_NCH], meta_bytes);
        ptx::tma_load_2d(sbuf, &tmap_s, &bars[GT_NCH], col0, gbeg);
        if (P.mode != 2) ptx::tma_load_2d(zbuf, &tmap_z, &bars[GT_NCH], col0, gbeg);
      }
    } else if (c < nch && !(P.debug & 2)) {
      ptx::mbar_arrive_expect_tx(&bars[c], (uint32_t)chunk_rows * 128u);
      ptx::tma_load_2d(wbuf + (size_t)c * chunk_rows * 128, &tmap_w, &bars[c], col0, (kbeg >> 3) + c * chunk_rows);
    }
  }
  if (P.bulk1d) {
    // one 128-byte bulk copy per packed row, issued by as many threads as there are rows (no descriptor fetch)
    __syncthreads();                                       // barrier init + expect_tx visible before any complete_tx
    for (int r = threadIdx.x; r < rows_c; r += NW * 32) {
      const uint32_t* src = P.W + (size_t)((kbeg >> 3) + r) * N + col0;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       ptx::smem_u32(wbuf + (size_t)r * 128)),
                   "l"(src), "r"(128u), "r"(ptx::smem_u32(&bars[r / chunk_rows]))
                   : "memory");
    }
  }
  // META_REGS: the scale / zero vectors of this lane's four columns, for the (<= GT_GPW) groups its warp owns
  uint2 sreg[GT_GPW], zreg[GT_GPW];
  if (META_REGS) {
#pragma unroll
    for (int i = 0; i < GT_GPW; ++i) {
      const int gl = warp + NW * i;
      sreg[i] = make_uint2(0u, 0u); zreg[i] = make_uint2(0u, 0u);
      if (gl < Gc) {
        const size_t mo = (size_t)(gbeg + gl) * N + col0 + 4 * g8;
        sreg[i] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(P.scales) + mo));
        if (P.mode != 2) zreg[i] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(P.zeros) + mo));
      }
    }
  }
  if (P.late_trigger == 0) pdl_launch_dependents();  // everything this CTA can do ahead of its input has been issued
#if GLB_GEMV_PREWAIT_ALL
  {
    // experiment: take EVERY mbarrier wait (metadata and all weight chunks) before griddepcontrol.wait.  In a chain of
    // dependent launches the copies land long before the previous layer finishes, so the post-wait path then holds no
    // barrier polls at all; an isolated call pays for it by not overlapping its first chunks' compute with the rest
    // of the copy.
    __syncthreads();
    if (!META_REGS) mbar_wait_warp(&bars[GT_NCH], 0);
    if (!(P.debug & 2))
      for (int c = 0; c < nch; ++c) mbar_wait_warp(&bars[c], 0);
  }
#endif
  if (!GLB_GEMV_PREWAIT_ALL && (P.issue_mode & 2) && !META_REGS) {
    // the metadata wait does not depend on the previous layer either: take it (and the barrier-init visibility it
    // needs) before griddepcontrol.wait, off the x -> y critical path
    __syncthreads();
    mbar_wait_warp(&bars[GT_NCH], 0);
  }

  // ---- 2. x depends on the previous kernel: wait for it only now, then stage it permuted ----
  if (rec) ts[GLB_GEMV_TIMELINE ? 1 : 0] = gtimer();
  pdl_wait_prior_grids();
  if (rec) ts[GLB_GEMV_TIMELINE ? 2 : 0] = gtimer();
  if (P.late_trigger == 2) pdl_launch_dependents();  // prefetch depth 1: the next layer streams while this one computes
  {
    // Each thread stages one 8-element word: permuted to {(k0,k4),(k1,k5),(k2,k6),(k3,k7)} (the B-fragment pairs
    // matching nibbles_to_f16x2) and, in FAST mode, with the (k1,k5),(k3,k7) slots divided by 16.  The rpg
    // consecutive threads of one group also reduce the group sums the epilogue needs:
    //   gsum[.][0] = sum over slots that pair with 1024+n   (plain x)
    //   gsum[.][1] = sum over slots that pair with 1024+16n (x/16 in FAST mode, plain x otherwise)
    const uint16_t* xg = reinterpret_cast<const uint16_t*>(P.x) + kbeg;
    const int words = Kc >> 3;
    const int total = MT * words;
    for (int i0 = 0; i0 < total; i0 += NW * 32) {
      const int i = i0 + threadIdx.x;
      const bool live = i < total;
      const int m = (MT == 1) ? 0 : (i / words), r = (MT == 1) ? i : (i - m * words);
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (live && m < P.M) {
        const uint4 v = *reinterpret_cast<const uint4*>(xg + (size_t)m * K + (size_t)r * 8);
        o.x = __byte_perm(v.x, v.z, 0x5410);   // (k0,k4)
        o.y = __byte_perm(v.x, v.z, 0x7632);   // (k1,k5)
        o.z = __byte_perm(v.y, v.w, 0x5410);   // (k2,k6)
        o.w = __byte_perm(v.y, v.w, 0x7632);   // (k3,k7)
        if (FAST) {
          const __half2 k16th = __half2half2(__ushort_as_half((unsigned short)0x2C00));
          __half2 a = __hmul2(*reinterpret_cast<__half2*>(&o.y), k16th), b2 = __hmul2(*reinterpret_cast<__half2*>(&o.w), k16th);
          o.y = *reinterpret_cast<uint32_t*>(&a); o.w = *reinterpret_cast<uint32_t*>(&b2);
        }
      }
      if (live) *reinterpret_cast<uint4*>(xs + (size_t)m * Kc + (size_t)r * 8) = o;
      if (P.mode != 2 || FAST) {
        float slo = cvt_meta<BF16>((uint16_t)(o.x & 0xffffu)) + cvt_meta<BF16>((uint16_t)(o.x >> 16)) +
                    cvt_meta<BF16>((uint16_t)(o.z & 0xffffu)) + cvt_meta<BF16>((uint16_t)(o.z >> 16));
        float shi = cvt_meta<BF16>((uint16_t)(o.y & 0xffffu)) + cvt_meta<BF16>((uint16_t)(o.y >> 16)) +
                    cvt_meta<BF16>((uint16_t)(o.w & 0xffffu)) + cvt_meta<BF16>((uint16_t)(o.w >> 16));
#pragma unroll
        for (int off = rpg / 2; off > 0; off >>= 1) {          // rpg (4..32) consecutive lanes hold one group
          slo += __shfl_xor_sync(0xffffffffu, slo, off);
          shi += __shfl_xor_sync(0xffffffffu, shi, off);
        }
        if (live && (r % rpg) == 0) {
          gsum[(m * Gc + r / rpg) * 2 + 0] = slo;
          gsum[(m * Gc + r / rpg) * 2 + 1] = shi;
        }
      }
    }
    if (threadIdx.x < 4) zero16[threadIdx.x] = 0u;
  }
  __syncthreads();              // also publishes the mbarrier initialisation to every thread
  if (rec) ts[GLB_GEMV_TIMELINE ? 3 : 0] = gtimer();

  // ---- 3. main loop: warp w walks groups w, w + NW, ... of this CTA's K range ----
  float out[4][2];
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c][0] = out[c][1] = 0.f;
  // Row assignment inside an 8-row block (two MMA "runs" a = 0, 1): lane t4 owns packed row 2*t4 + a.  With the
  // TMA 128-byte swizzle (16-byte chunk index ^= row % 8) the 8 lanes of every quarter-warp then read 8 different
  // chunks {g8, g8+1} ^ {a, 2+a, 4+a, 6+a}: conflict-free 128-bit shared-memory loads.  x is staged word-per-row,
  // so the matching B fragment of the lane is word (block*8 + 2*t4 + a) of its token's row.
  const bool tok_ok = (g8 < MT);
  const unsigned char* xb = reinterpret_cast<const unsigned char*>(xs + (size_t)(tok_ok ? g8 : 0) * Kc) + (2 * t4) * 16;
  const unsigned char* wl = wbuf + (2 * t4) * 128;
  const int sw0 = (g8 ^ (2 * t4)) << 4, sw1 = (g8 ^ (2 * t4 + 1)) << 4;   // swizzled chunk offsets for a = 0 / 1
  if (!GLB_GEMV_PREWAIT_ALL && !META_REGS && !(P.issue_mode & 2)) mbar_wait_warp(&bars[GT_NCH], 0);   // scales / zeros landed
  if (rec) ts[GLB_GEMV_TIMELINE ? 4 : 0] = gtimer();

#pragma unroll
  for (int gi = 0; gi < (META_REGS ? GT_GPW : (1 << 20)); ++gi) {
    const int gl = warp + NW * gi;
    if (gl >= ((P.debug & 1) ? 0 : Gc)) break;
    if (!GLB_GEMV_PREWAIT_ALL && !(P.debug & 2)) mbar_wait_warp(&bars[(gl * rpg) >> P.chunk_shift], 0);   // this group's weight chunk landed
    if (rec && gl == warp) ts[GLB_GEMV_TIMELINE ? 5 : 0] = gtimer();
    // four independent accumulator chains per run: [column pair][k half]
    float dacc[2][2][4];
#pragma unroll
    for (int pr = 0; pr < 2; ++pr)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) dacc[pr][hh][0] = dacc[pr][hh][1] = dacc[pr][hh][2] = dacc[pr][hh][3] = 0.f;
#pragma unroll
    for (int r = 0; r < RPG; ++r) {
      const int blk = gl * (RPG / 2) + (r >> 1), a = r & 1;            // 8-row block of this CTA's slice, half-run
      const uint4 q = *reinterpret_cast<const uint4*>(wl + (size_t)blk * 1024 + a * 128 + (a ? sw1 : sw0));
      uint4 xv = make_uint4(0u, 0u, 0u, 0u);
      if (tok_ok) xv = *reinterpret_cast<const uint4*>(xb + (size_t)blk * 128 + a * 16);
      uint32_t ha[4], hb[4];
      nibbles_to_f16x2<BF16, FAST>(q.x, ha);
      nibbles_to_f16x2<BF16, FAST>(q.y, hb);
      {
        const uint32_t a0[4] = {ha[0], hb[0], ha[1], hb[1]}, a1[4] = {ha[2], hb[2], ha[3], hb[3]};
        mma16816<BF16>(dacc[0][0], a0, xv.x, xv.y);
        mma16816<BF16>(dacc[0][1], a1, xv.z, xv.w);
      }
      nibbles_to_f16x2<BF16, FAST>(q.z, ha);
      nibbles_to_f16x2<BF16, FAST>(q.w, hb);
      {
        const uint32_t a0[4] = {ha[0], hb[0], ha[1], hb[1]}, a1[4] = {ha[2], hb[2], ha[3], hb[3]};
        mma16816<BF16>(dacc[1][0], a0, xv.x, xv.y);
        mma16816<BF16>(dacc[1][1], a1, xv.z, xv.w);
      }
    }
    // ---- end of group: y += s * (sum(x q)) + z' * sum(x) ----
    uint2 sw, zw = make_uint2(0u, 0u);
    if (META_REGS) { sw = sreg[gi < GT_GPW ? gi : 0]; zw = zreg[gi < GT_GPW ? gi : 0]; }
    else {
      sw = *reinterpret_cast<const uint2*>(sbuf + (size_t)gl * 64 + g8 * 8);
      if (P.mode != 2) zw = *reinterpret_cast<const uint2*>(zbuf + (size_t)gl * 64 + g8 * 8);
    }
    const uint32_t swa[2] = {sw.x, sw.y}, zwa[2] = {zw.x, zw.y};
    float sx[2] = {0.f, 0.f}, off[2] = {0.f, 0.f};
    if (P.mode != 2 || FAST) {
#pragma unroll
      for (int tk = 0; tk < 2; ++tk) {
        const int m = 2 * t4 + tk;
        if (m < MT) {
          const float2 gs2 = *reinterpret_cast<const float2*>(&gsum[(m * Gc + gl) * 2]);
          sx[tk] = FAST ? fmaf(16.f, gs2.y, gs2.x) : (gs2.x + gs2.y);     // sum_k x_k
          off[tk] = FAST ? 1024.f * (gs2.x + gs2.y) : 0.f;                // 1024 * sum_k x'_k
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float sc = cvt_meta<BF16>((uint16_t)((c & 1) ? (swa[c >> 1] >> 16) : (swa[c >> 1] & 0xffffu)));
      float zc = cvt_meta<BF16>((uint16_t)((c & 1) ? (zwa[c >> 1] >> 16) : (zwa[c >> 1] & 0xffffu)));
      if (P.mode == 3) zc = -zc * sc;
      const int pr = c >> 1, hi = (c & 1) * 2;        // c0,c2 -> rows 0-7 (d0,d1) ; c1,c3 -> rows 8-15 (d2,d3)
#pragma unroll
      for (int tk = 0; tk < ((MT == 1) ? 1 : 2); ++tk) {     // MT == 1: token 1 does not exist, skip its arithmetic
        const float d = dacc[pr][0][hi + tk] + dacc[pr][1][hi + tk];
        out[c][tk] = fmaf(sc, d - off[tk], fmaf(zc, sx[tk], out[c][tk]));
      }
    }
  }

  if (P.late_trigger == 1) pdl_launch_dependents();
  if (rec) ts[GLB_GEMV_TIMELINE ? 6 : 0] = gtimer();
  // ---- 4. cross-warp reduction, cluster (split-K) reduction through DSMEM, epilogue ----
#pragma unroll
  for (int tk = 0; tk < 2; ++tk) {
    const int m = 2 * t4 + tk;
    if (m < MT) {
#pragma unroll
      for (int c = 0; c < 4; ++c) red[(warp * MT + m) * GV_COLS + 4 * g8 + c] = out[c][tk];
    }
  }
  __syncthreads();
  if (rec) {
    ts[GLB_GEMV_TIMELINE ? 7 : 0] = gtimer();
    const unsigned int pos = atomicAdd(&g_timeline_pos, 1u);
    if (pos < 8192 / 9) {
      g_timeline[pos * 9] = blockIdx.x;
      for (int k = 0; k < (GLB_GEMV_TIMELINE ? 8 : 1); ++k) g_timeline[pos * 9 + 1 + k] = ts[k];
    }
  }
  float part = 0.f;
  const int i = threadIdx.x;
  if (i < MT * GV_COLS) {
    const int m = i / GV_COLS, c = i - m * GV_COLS;
#pragma unroll
    for (int w = 0; w < NW; ++w) part += red[(w * MT + m) * GV_COLS + c];
    if (KSPLIT > 1 && rank != 0) st_cluster_f32(&peer[i], 0u, part);
  }
  if (KSPLIT > 1) cluster_sync_all();
  if (rank == 0 && i < MT * GV_COLS) {
    const int m = i / GV_COLS, c = i - m * GV_COLS;
    if (m < P.M) {
      float v = part;
      if (KSPLIT > 1) v += peer[i];
      if (P.bias) v = rnd_out<BF16>(v) + cvt_meta<BF16>(reinterpret_cast<const uint16_t*>(P.bias)[col0 + c]);
      st_out<BF16>(P.y, (int64_t)m * N + col0 + c, v);
    }
  }
}

template <int MT, int KSPLIT, int NW, bool META_REGS> size_t gemv_tma_smem_bytes(int K, int group_size) {
  const int Kc = K / KSPLIT, Gc = K / group_size / KSPLIT;
  const size_t meta = META_REGS ? 0 : (((size_t)Gc * 64 + 127) & ~(size_t)127);
  // (dynamic shared memory starts 1024-byte aligned -- the kernel traps otherwise -- so no alignment slack: with
  // META_REGS the K = 4096 slab takes 74.9 KB and THREE layers fit on an SM)
  return (size_t)(Kc / 8) * 128 + 2 * meta + (size_t)MT * Kc * 2 + 16 + (size_t)NW * MT * GV_COLS * 4 +
         (size_t)MT * GV_COLS * 4 + (size_t)MT * Gc * 8 + (GT_NCH + 1) * 8 + 64;
}

template <bool BF16, int MT, int RPG, int KSPLIT, bool FAST, int NW, bool META_REGS = false>
int launch_w4_tma(const glb_linear& L, const GemvParams& P0, int late_trigger, cudaStream_t stream) {
  const int Kc = P0.K / KSPLIT, Gc = P0.K / P0.group_size / KSPLIT, rows_c = Kc / 8;
  constexpr int rpg = RPG * 4;
  const int nch = (rows_c % (GT_NCH * rpg) == 0) ? GT_NCH : 1;
  CUtensorMap tw, ts, tz;
  int rc = make_tmap_2d(&tw, GLB_INT32, L.W_q, (uint64_t)L.N, (uint64_t)(L.K / 8), (uint64_t)L.N * 4, GV_COLS, rows_c / nch, 1);
  if (rc != GLB_OK) return rc;
  rc = make_tmap_2d(&ts, GLB_FP16, L.scales, (uint64_t)L.N, (uint64_t)(L.K / L.group_size), (uint64_t)L.N * 2, GV_COLS, Gc, 0);
  if (rc != GLB_OK) return rc;
  rc = make_tmap_2d(&tz, GLB_FP16, (P0.mode != 2) ? L.zeros : L.scales, (uint64_t)L.N, (uint64_t)(L.K / L.group_size),
                    (uint64_t)L.N * 2, GV_COLS, Gc, 0);
  if (rc != GLB_OK) return rc;
  GemvTmaParams P;
  P.x = P0.x; P.y = P0.y; P.bias = P0.bias; P.K = P0.K; P.N = P0.N; P.M = P0.M; P.group_size = P0.group_size; P.mode = P0.mode;
  P.late_trigger = late_trigger % 10;
  P.bulk1d = late_trigger / 10;
  P.W = reinterpret_cast<const uint32_t*>(L.W_q);
  P.scales = P0.scales; P.zeros = P0.zeros;
  {
    int cr = rows_c / nch, sh = 0;
    while ((1 << sh) < cr) ++sh;
    P.chunk_shift = sh;
    static int dbg = -1, issue = -1;
    if (dbg < 0) { const char* e = getenv("GLB_GEMV_DEBUG"); dbg = e ? atoi(e) : 0; }
    if (issue < 0) { const char* e = getenv("GLB_GEMV_ISSUE"); issue = e ? atoi(e) : 0; }   // measured r1 (profiles/decode_issue_sweep_r1.txt): 0 -> 3.59, 1 -> 3.61, 2 -> 3.71, 3 -> 3.89 us/layer
    P.debug = dbg;
    P.issue_mode = (NW > GT_NCH + 1) ? issue : (issue & 2);
  }
  const size_t smem = gemv_tma_smem_bytes<MT, KSPLIT, NW, META_REGS>(P0.K, P0.group_size);
  auto kern = gemv_w4_tma_kernel<BF16, MT, RPG, KSPLIT, FAST, NW, META_REGS>;
  if (smem > 48 * 1024) GLB_ENSURE_SMEM(kern, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(P0.N / GV_COLS) * KSPLIT);
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (KSPLIT > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = KSPLIT; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_pdl_enabled.load()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  GLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tw, ts, tz, P));
  count_launch();
  return GLB_OK;
}


// =================================================================================================
// M == 1 decode kernel, round 2 ("v7"): same slab / TMA / swizzle scheme as gemv_w4_tma_kernel with
//   * per-warp x staging: a warp only ever touches the x words of its own groups, so it loads them itself (one
//     16-byte word per lane and group), permutes / pre-scales them in registers, reduces the group sums with shuffles
//     and parks the words in its private rows of the x buffer -- no block barrier between the hand-off and the first HMMA;
//   * a leaner main loop (profiles/decode_v7_r2.txt): two accumulator chains per group, the nibble shift on the FMA pipe
//     (mul.hi), x fragments only in the 4 lanes of token 0: ~148 issued instructions per group and warp instead of ~215;
// Nothing spins before griddepcontrol.wait (the co-resident CTA of the previous layer is in its main loop then).
// Tried on hardware in round 2 and removed (profiles/decode_experiments_r2.txt):
//   * a device-side arrival counter instead of griddepcontrol.wait for provably adjacent launches (store y -> fence ->
//     red / ld.acquire polling): 2.50 us per EMPTY chained launch against 1.66 us with the kernel-boundary hand-off --
//     the hardware path is faster than two L2 round trips plus a gpu-scope fence;
//   * an L2 prefetch hint (every CTA asks the L2 for its share of the tensors of the layer 1 / 2 / 3 launches ahead,
//     cp.async.bulk.prefetch.L2): 3.97 us per layer against 3.59 without -- the prefetch competes with the fill the
//     next launch is waiting for;
//   * one polling lane per warp (+ __syncwarp) on the copy barriers: ptxas then wraps every later shuffle in
//     WARPSYNC.COLLECTIVE sequences, 2x slower end to end; all weight waits ahead of the hand-off: +0.4 us.
// =================================================================================================
struct GemvM1Params {
  const void* x; void* y; const void* bias;
  // fused output all-gather (glb_matmul_bcast): the 32 outputs of a CTA are stored to EVERY destination (this GPU's
  // buffer and the peers' symmetric buffers mapped over NVLink), at column y_col0 of each; n_dst == 0: plain store to y
  void* y_dst[8];
  int n_dst;
  long long y_col0;
  int K, N, group_size, mode;
  int trigger;               // dependents are signalled 0: at kernel start, 2: right after the hand-off
  int debug;                 // bit 0: skip main loop, bit 1: skip weight copies, bit 2: timeline stamps (diagnostics)
  int chunk_shift;           // log2(packed rows per weight chunk)
};

constexpr int M1_NW = 16;
constexpr int M1_MAXGPW = 8;     // groups per warp (K / group_size / 16)

__device__ __forceinline__ uint4 ld_relaxed_gpu_v4(const void* p) {
  uint4 r;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}

template <bool BF16, int RPG, bool FAST, int GPW>
__global__ void __launch_bounds__(M1_NW * 32, 2)
gemv_w4_m1_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_s,
                  const __grid_constant__ CUtensorMap tmap_z, const GemvM1Params P) {
  static_assert(!(BF16 && FAST), "the offset-carrying unpack needs fp16's 10-bit mantissa");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int rpg = RPG * 4;                     // packed rows (= 8-element x words) per group
  const int K = P.K, N = P.N;
  const int G = K / P.group_size, rows = K >> 3;
  const int col0 = blockIdx.x * GV_COLS;
  const int nch = (rows % (GT_NCH * rpg) == 0 && rows / GT_NCH >= 8) ? GT_NCH : 1;
  const int chunk_rows = rows / nch;

  unsigned char* wbuf = smem_raw;                                        // rows x 128 B, TMA 128B swizzle
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0u) __trap();
  const size_t meta_bytes = ((size_t)G * 64 + 127) & ~(size_t)127;
  unsigned char* sbuf = wbuf + (size_t)rows * 128;                       // G x 64 B
  unsigned char* zbuf = sbuf + meta_bytes;
  unsigned char* xs = zbuf + meta_bytes;                                 // rows x 16 B: permuted x words (warp-private rows)
  float* red = reinterpret_cast<float*>(xs + (size_t)rows * 16);         // [M1_NW][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + M1_NW * GV_COLS);   // [GT_NCH] weights, [1] metadata

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g8 = lane >> 2, t4 = lane & 3;

  // diagnostics (GLB_GEMV_DEBUG bit 2): %globaltimer phase stamps of thread 0 of the first / last CTA
  const bool rec = (P.debug & 4) && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
  unsigned long long ts[8];
  if (rec) ts[0] = gtimer();
  // ---- 1. everything that does not depend on the previous layer ----
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmap_w);
    ptx::prefetch_tmap(&tmap_s);
    for (int c = 0; c <= GT_NCH; ++c) ptx::mbar_init(&bars[c], 1);
    ptx::fence_mbar_init();
    const uint32_t mb = (uint32_t)G * 64u * (P.mode != 2 ? 2u : 1u);
    ptx::mbar_arrive_expect_tx(&bars[GT_NCH], mb);
    ptx::tma_load_2d(sbuf, &tmap_s, &bars[GT_NCH], col0, 0);
    if (P.mode != 2) ptx::tma_load_2d(zbuf, &tmap_z, &bars[GT_NCH], col0, 0);
    for (int c = 0; c < ((P.debug & 2) ? 0 : nch); ++c) {
      ptx::mbar_arrive_expect_tx(&bars[c], (uint32_t)chunk_rows * 128u);
      ptx::tma_load_2d(wbuf + (size_t)c * chunk_rows * 128, &tmap_w, &bars[c], col0, c * chunk_rows);
    }
  }
  if (rec) ts[1] = gtimer();                         // copies issued
  if (P.trigger == 0) pdl_launch_dependents();
  __syncthreads();                                   // barrier initialisation visible to every thread

  // ---- 2. hand-off: the previous layer's y is this layer's x ----
  if (rec) ts[2] = gtimer();
  pdl_wait_prior_grids();
  if (rec) ts[3] = gtimer();                         // hand-off done
  if (P.trigger == 2) pdl_launch_dependents();

  // ---- 3. this warp's x words: load, permute, pre-scale, group sums ----
  float sxv[GPW], offv[GPW];
  {
    const uint16_t* xg = reinterpret_cast<const uint16_t*>(P.x);
    uint4 v[GPW];
#pragma unroll
    for (int gi = 0; gi < GPW; ++gi) {
      const int gl = warp + M1_NW * gi;
      v[gi] = make_uint4(0u, 0u, 0u, 0u);
      if (gl < G && lane < rpg) v[gi] = ld_relaxed_gpu_v4(xg + ((size_t)gl * rpg + lane) * 8);
    }
#pragma unroll
    for (int gi = 0; gi < GPW; ++gi) {
      const int gl = warp + M1_NW * gi;
      uint4 o;
      o.x = __byte_perm(v[gi].x, v[gi].z, 0x5410);   // (k0,k4)
      o.y = __byte_perm(v[gi].x, v[gi].z, 0x7632);   // (k1,k5)
      o.z = __byte_perm(v[gi].y, v[gi].w, 0x5410);   // (k2,k6)
      o.w = __byte_perm(v[gi].y, v[gi].w, 0x7632);   // (k3,k7)
      if (FAST) {
        const __half2 k16th = __half2half2(__ushort_as_half((unsigned short)0x2C00));
        __half2 a = __hmul2(*reinterpret_cast<__half2*>(&o.y), k16th), b2 = __hmul2(*reinterpret_cast<__half2*>(&o.w), k16th);
        o.y = *reinterpret_cast<uint32_t*>(&a); o.w = *reinterpret_cast<uint32_t*>(&b2);
      }
      if (gl < G && lane < rpg) *reinterpret_cast<uint4*>(xs + ((size_t)gl * rpg + lane) * 16) = o;
      float slo = cvt_meta<BF16>((uint16_t)(o.x & 0xffffu)) + cvt_meta<BF16>((uint16_t)(o.x >> 16)) +
                  cvt_meta<BF16>((uint16_t)(o.z & 0xffffu)) + cvt_meta<BF16>((uint16_t)(o.z >> 16));
      float shi = cvt_meta<BF16>((uint16_t)(o.y & 0xffffu)) + cvt_meta<BF16>((uint16_t)(o.y >> 16)) +
                  cvt_meta<BF16>((uint16_t)(o.w & 0xffffu)) + cvt_meta<BF16>((uint16_t)(o.w >> 16));
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {       // lanes >= rpg carry zeros
        slo += __shfl_xor_sync(0xffffffffu, slo, off);
        shi += __shfl_xor_sync(0xffffffffu, shi, off);
      }
      sxv[gi] = FAST ? fmaf(16.f, shi, slo) : (slo + shi);      // sum_k x_k
      offv[gi] = FAST ? 1024.f * (slo + shi) : 0.f;             // 1024 * sum_k x'_k
    }
  }
  __syncwarp();
  if (rec) ts[4] = gtimer();                         // x staged (warp 0)

  // ---- 4. main loop ----
  float out[4] = {0.f, 0.f, 0.f, 0.f};
  const bool tok_ok = (g8 == 0);
  const unsigned char* xb = xs + (2 * t4) * 16;
  const unsigned char* wl = wbuf + (2 * t4) * 128;
  const int sw0 = (g8 ^ (2 * t4)) << 4, sw1 = (g8 ^ (2 * t4 + 1)) << 4;
  ptx::mbar_wait(&bars[GT_NCH], 0);                  // scales / zeros landed (long ago, in a chain)
#pragma unroll
  for (int gi = 0; gi < GPW; ++gi) {
    const int gl = warp + M1_NW * gi;
    if (gl >= ((P.debug & 1) ? 0 : G)) break;
    if (!(P.debug & 2)) ptx::mbar_wait(&bars[(gl * rpg) >> P.chunk_shift], 0);   // this group's weight chunk landed
    float dacc[2][4];
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) dacc[pr][0] = dacc[pr][1] = dacc[pr][2] = dacc[pr][3] = 0.f;
#pragma unroll
    for (int r = 0; r < RPG; ++r) {
      const int blk = gl * (RPG / 2) + (r >> 1), a = r & 1;
      const uint4 q = *reinterpret_cast<const uint4*>(wl + (size_t)blk * 1024 + a * 128 + (a ? sw1 : sw0));
      uint4 xv = make_uint4(0u, 0u, 0u, 0u);
      if (tok_ok) xv = *reinterpret_cast<const uint4*>(xb + (size_t)blk * 128 + a * 16);
      uint32_t ha[4], hb[4];
      nibbles_to_f16x2_m1<BF16, FAST>(q.x, ha);
      nibbles_to_f16x2_m1<BF16, FAST>(q.y, hb);
      {
        const uint32_t a0[4] = {ha[0], hb[0], ha[1], hb[1]}, a1[4] = {ha[2], hb[2], ha[3], hb[3]};
        mma16816<BF16>(dacc[0], a0, xv.x, xv.y);
        mma16816<BF16>(dacc[0], a1, xv.z, xv.w);
      }
      nibbles_to_f16x2_m1<BF16, FAST>(q.z, ha);
      nibbles_to_f16x2_m1<BF16, FAST>(q.w, hb);
      {
        const uint32_t a0[4] = {ha[0], hb[0], ha[1], hb[1]}, a1[4] = {ha[2], hb[2], ha[3], hb[3]};
        mma16816<BF16>(dacc[1], a0, xv.x, xv.y);
        mma16816<BF16>(dacc[1], a1, xv.z, xv.w);
      }
    }
    const uint2 sw = *reinterpret_cast<const uint2*>(sbuf + (size_t)gl * 64 + g8 * 8);
    uint2 zw = make_uint2(0u, 0u);
    if (P.mode != 2) zw = *reinterpret_cast<const uint2*>(zbuf + (size_t)gl * 64 + g8 * 8);
    const uint32_t swa[2] = {sw.x, sw.y}, zwa[2] = {zw.x, zw.y};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float sc = cvt_meta<BF16>((uint16_t)((c & 1) ? (swa[c >> 1] >> 16) : (swa[c >> 1] & 0xffffu)));
      float zc = cvt_meta<BF16>((uint16_t)((c & 1) ? (zwa[c >> 1] >> 16) : (zwa[c >> 1] & 0xffffu)));
      if (P.mode == 3) zc = -zc * sc;
      // c0,c2 -> rows 0-7 of the tile (d0); c1,c3 -> rows 8-15 (d2); token 0 = accumulator column 0 (lanes t4 == 0)
      const float d = dacc[c >> 1][(c & 1) * 2];
      out[c] = fmaf(sc, d - offv[gi], fmaf(zc, sxv[gi], out[c]));
    }
  }

  // ---- 5. cross-warp reduction, bias, store ----
  if (rec) ts[5] = gtimer();                         // loop done (warp 0)
  if (t4 == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) red[warp * GV_COLS + 4 * g8 + c] = out[c];
  }
  __syncthreads();
  if (threadIdx.x < GV_COLS) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < M1_NW; ++w) v += red[w * GV_COLS + threadIdx.x];
    if (P.bias) v = rnd_out<BF16>(v) + cvt_meta<BF16>(reinterpret_cast<const uint16_t*>(P.bias)[col0 + threadIdx.x]);
    if (P.n_dst == 0) {
      st_out<BF16>(P.y, (int64_t)col0 + threadIdx.x, v);
    } else {
#pragma unroll 1
      for (int d = 0; d < P.n_dst; ++d) st_out<BF16>(P.y_dst[d], (int64_t)P.y_col0 + col0 + threadIdx.x, v);
    }
  }
  if (rec) {
    ts[6] = gtimer();
    ts[7] = ts[0];
    const unsigned int pos = atomicAdd(&g_timeline_pos, 1u);
    if (pos < 8192 / 9) {
      g_timeline[pos * 9] = blockIdx.x;
      for (int k = 0; k < 8; ++k) g_timeline[pos * 9 + 1 + k] = ts[k];
    }
  }
}

template <int GPW> size_t gemv_m1_smem_bytes(int K, int group_size) {
  const int G = K / group_size;
  const size_t meta = ((size_t)G * 64 + 127) & ~(size_t)127;
  return (size_t)(K / 8) * 128 + 2 * meta + (size_t)(K / 8) * 16 + (size_t)M1_NW * GV_COLS * 4 + (GT_NCH + 2) * 8 + 64;
}

bool gemv_m1_ok(const GemvParams& P) {
  if (P.M != 1) return false;
  const int G = P.K / P.group_size, rows = P.K / 8, rpg = P.group_size / 8;
  if (rows * 128 > 64 * 1024 || rows < 32) return false;
  if (G > M1_NW * M1_MAXGPW || G > 256) return false;
  const int nch = (rows % (GT_NCH * rpg) == 0 && rows / GT_NCH >= 8) ? GT_NCH : 1;
  const int cr = rows / nch;
  if (cr > 256 || cr % 8 != 0 || (cr & (cr - 1)) != 0 || cr % rpg != 0) return false;   // chunk index is a shift; whole groups per chunk
  if (rpg % 8 != 0 || rpg > 32) return false;           // swizzled 8-row blocks must not straddle groups
  return true;
}

template <bool BF16, int RPG, bool FAST, int GPW>
int launch_w4_m1(const glb_linear& L, const GemvParams& P0, cudaStream_t stream) {
  const int rows = P0.K / 8, G = P0.K / P0.group_size;
  constexpr int rpg = RPG * 4;
  const int nch = (rows % (GT_NCH * rpg) == 0 && rows / GT_NCH >= 8) ? GT_NCH : 1;
  CUtensorMap tw, ts, tz;
  int rc = make_tmap_2d(&tw, GLB_INT32, L.W_q, (uint64_t)L.N, (uint64_t)rows, (uint64_t)L.N * 4, GV_COLS, rows / nch, 1);
  if (rc != GLB_OK) return rc;
  rc = make_tmap_2d(&ts, GLB_FP16, L.scales, (uint64_t)L.N, (uint64_t)G, (uint64_t)L.N * 2, GV_COLS, G, 0);
  if (rc != GLB_OK) return rc;
  rc = make_tmap_2d(&tz, GLB_FP16, (P0.mode != 2) ? L.zeros : L.scales, (uint64_t)L.N, (uint64_t)G, (uint64_t)L.N * 2, GV_COLS, G, 0);
  if (rc != GLB_OK) return rc;
  GemvM1Params P;
  P.x = P0.x; P.y = P0.y; P.bias = P0.bias; P.K = P0.K; P.N = P0.N; P.group_size = P0.group_size; P.mode = P0.mode;
  P.n_dst = P0.n_dst; P.y_col0 = P0.y_col0;
  for (int d = 0; d < 8; ++d) P.y_dst[d] = (d < P0.n_dst) ? P0.y_dst[d] : nullptr;
  {
    static int dbg = -1, trig = -1;
    if (dbg < 0) { const char* e = getenv("GLB_GEMV_DEBUG"); dbg = e ? atoi(e) : 0; }
    if (trig < 0) { const char* e = getenv("GLB_GEMV_TRIGGER"); trig = e ? atoi(e) : 2; }
    P.debug = dbg; P.trigger = trig;
    int cr = rows / nch, sh = 0;
    while ((1 << sh) < cr) ++sh;
    P.chunk_shift = sh;
  }
  const size_t smem = gemv_m1_smem_bytes<GPW>(P0.K, P0.group_size);
  auto kern = gemv_w4_m1_kernel<BF16, RPG, FAST, GPW>;
  if (smem > 48 * 1024) GLB_ENSURE_SMEM(kern, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(P0.N / GV_COLS));
  cfg.blockDim = dim3(M1_NW * 32);
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
  GLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tw, ts, tz, P));
  count_launch();
  return GLB_OK;
}

// Kernel flavour (GLB_GEMV_VARIANT, read once): 3 = register prefetch + 2-CTA cluster; 4 = TMA + 2-CTA cluster,
// 256 threads (default); 41 = same with the dependent-launch trigger after the main loop; 5 / 51 = TMA, one
// 512-thread CTA per slab over the full K (exactly two CTAs per SM), early / late trigger; 6 / 62 = TMA, one
// 256-thread CTA per slab, scale / zero vectors in registers -> 74.9 KB of shared memory, THREE layers per SM.
int gemv_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GLB_GEMV_VARIANT");
    v = e ? atoi(e) : 7;
  }
  return v;
}

// shapes the TMA flavour takes: <= 8 tokens, the CTA's weight slice fits 64 KB of shared memory
bool gemv_tma_ok(const GemvParams& P, int ksplit) {
  const int Kc = P.K / ksplit, Gc = P.K / P.group_size / ksplit;
  if (P.M > 8) return false;
  if ((Kc / 8) * 128 > 64 * 1024 || Gc > 256 || (Kc / 8) > 256 * GT_NCH) return false;
  const int rows_c = Kc / 8, rpg = P.group_size / 8;
  const int nch = (rows_c % (GT_NCH * rpg) == 0) ? GT_NCH : 1;
  if (rows_c / nch > 256) return false;
  const int cr = rows_c / nch;
  if (cr & (cr - 1)) return false;                     // chunk index is a shift
  if (rpg % 8 != 0 || cr % 8 != 0) return false;       // swizzled 8-row blocks must not straddle groups / chunks
  return true;
}

template <bool BF16, int MT>
int dispatch_w4(const glb_linear& L, const GemvParams& P, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  const int G = P.K / P.group_size;
  const int rpgi = P.group_size / 32;
  // split K over a 2-CTA cluster when that still leaves whole groups per warp-batch
#define GLB_GV(RPG)                                                                                   \
  {                                                                                                   \
    constexpr int GPB = (GV_R / RPG) > 0 ? (GV_R / RPG) : 1;                                          \
    constexpr bool F = !BF16 && GLB_GEMV_FAST;                                                        \
    constexpr int MT8 = (MT <= 8) ? MT : 8;                                                           \
    const bool split = (G % (2 * GPB) == 0 && P.N / GV_COLS < 296);                                   \
    const int variant = gemv_variant();                                                               \
    if (MT == 1 && RPG >= 2 && variant == 7 && gemv_m1_ok(P) && (reinterpret_cast<uintptr_t>(L.scales) & 15) == 0 && \
        (P.mode == 2 || (reinterpret_cast<uintptr_t>(L.zeros) & 15) == 0) && (P.N * 2) % 16 == 0) {   \
      constexpr int R2 = (RPG >= 2) ? RPG : 2;                                                        \
      const int gpw = (G + M1_NW - 1) / M1_NW;                                                        \
      if (gpw <= 1) return launch_w4_m1<BF16, R2, F, 1>(L, P, stream);           \
      if (gpw <= 2) return launch_w4_m1<BF16, R2, F, 2>(L, P, stream);           \
      if (gpw <= 4) return launch_w4_m1<BF16, R2, F, 4>(L, P, stream);           \
      return launch_w4_m1<BF16, R2, F, 8>(L, P, stream);                         \
    }                                                                                                 \
    if (MT <= 8 && variant != 3 && (reinterpret_cast<uintptr_t>(L.scales) & 15) == 0 &&               \
        (P.mode == 2 || (reinterpret_cast<uintptr_t>(L.zeros) & 15) == 0) && (P.N * 2) % 16 == 0) {   \
      /* x1: late trigger, x2: trigger after wait, x3: bulk1d + trigger after wait */               \
      const int late = (variant == 7) ? 2 : (variant % 10 == 1 && variant > 10) ? 1 : ((variant % 10 == 2 && variant > 10) ? 2 : ((variant % 10 == 3 && variant > 10) ? 12 : 0)); \
      if ((variant == 6 || variant / 10 == 6) && MT8 == 1 && gemv_tma_ok(P, 1) && G <= GT_GPW * 8)     \
        return launch_w4_tma<BF16, MT8, RPG, 1, F, 8, true>(L, P, late, stream);                      \
      if ((variant == 7 || variant == 5 || variant / 10 == 5 || variant == 6 || variant / 10 == 6) && gemv_tma_ok(P, 1)) \
        return launch_w4_tma<BF16, MT8, RPG, 1, F, 16>(L, P, late, stream);                           \
      if (split && gemv_tma_ok(P, 2)) return launch_w4_tma<BF16, MT8, RPG, 2, F, 8>(L, P, late, stream); \
      if (gemv_tma_ok(P, 1)) return launch_w4_tma<BF16, MT8, RPG, 1, F, 8>(L, P, late, stream);       \
    }                                                                                                 \
    if (split) return launch_w4<BF16, MT, RPG, 2, F>(P, stream);                                      \
    return launch_w4<BF16, MT, RPG, 1, F>(P, stream);                                                 \
  }
  switch (rpgi) {
    case 1: GLB_GV(1)
    case 2: GLB_GV(2)
    case 4: GLB_GV(4)
    case 8: GLB_GV(8)
    default: break;
  }
#undef GLB_GV
  set_error("decode kernel: unsupported group_size %d", P.group_size);
  return GLB_ERR_UNSUPPORTED;
}

}  // namespace

int debug_read_timeline(unsigned long long* out, int max_words) {
  unsigned int n = 0;
  GLB_CHECK_CUDA(cudaMemcpyFromSymbol(&n, g_timeline_pos, sizeof(n)));
  int words = (int)n * 9;
  if (words > max_words) words = max_words;
  if (words > 8192) words = 8192;
  if (words > 0) GLB_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_timeline, (size_t)words * 8));
  n = 0;
  GLB_CHECK_CUDA(cudaMemcpyToSymbol(g_timeline_pos, &n, sizeof(n)));
  return words;
}

bool gemv_supported(const glb_linear& L, int64_t M) {
  if (M < 1 || M > 16) return false;
  if (is_mx(L.input_dtype)) return false;
  if (!(L.input_dtype == GLB_FP16 || L.input_dtype == GLB_BF16)) return false;
  if (L.output_dtype != L.input_dtype || L.meta_dtype != L.input_dtype) return false;
  if (L.W_nbits != 4 || L.elements_per_sample != 8 || L.wq_dtype != GLB_INT32) return false;
  if (L.wq_stride1 != 1 || L.wq_stride0 != L.N) return false;
  if (L.channel_scale_mode != 0) return false;
  if (!(L.W_group_mode == 2 || L.W_group_mode == 3 || L.W_group_mode == 4)) return false;
  if (L.W_group_mode != 2 && (L.zeros_numel != L.scales_numel || L.zeros_dtype != L.scales_dtype)) return false;
  const int gs = L.group_size;
  if (!(gs == 32 || gs == 64 || gs == 128 || gs == 256) || L.K % gs != 0) return false;
  const int rpg = gs / 32, gpb = (GV_R / rpg) > 0 ? (GV_R / rpg) : 1;
  if ((L.K / gs) % gpb != 0) return false;
  if (L.scales_numel != (L.K / gs) * L.N || L.s_stride1 != 1 || L.s_stride0 != L.N) return false;
  if (L.W_group_mode != 2 && (L.z_stride1 != 1 || L.z_stride0 != L.N)) return false;
  if (L.N % GV_COLS != 0 || L.K % 8 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(L.W_q) & 15) || (reinterpret_cast<uintptr_t>(L.scales) & 7)) return false;
  if (L.W_group_mode != 2 && (reinterpret_cast<uintptr_t>(L.zeros) & 7)) return false;
  const int mt = (M == 1) ? 1 : (M <= 8 ? 8 : 16);
  const size_t smem = (size_t)mt * L.K * 2 + (size_t)mt * (L.K / gs) * 4 + (size_t)(GV_WARPS + 1) * mt * GV_COLS * 4 + 16;
  if (smem > 200 * 1024) return false;
  return true;
}

size_t gemv_workspace(const glb_linear& L, int64_t M) {
  (void)L; (void)M;
  return 0;                    // the decode kernels need no scratch
}

int launch_gemv(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M, const void* bias,
                void* workspace, size_t ws_bytes, cudaStream_t stream) {
  return launch_gemv_bcast(L, x, scales_x, y, nullptr, 0, 0, M, bias, workspace, ws_bytes, stream);
}

// fused output all-gather: only the M == 1 kernel stores to several destinations
bool gemv_bcast_supported(const glb_linear& L, int64_t M) {
  if (M != 1 || !gemv_supported(L, M) || gemv_variant() != 7) return false;
  GemvParams P;
  P.K = (int)L.K; P.N = (int)L.N; P.M = 1; P.group_size = L.group_size; P.mode = L.W_group_mode;
  return gemv_m1_ok(P) && L.group_size >= 64 && (reinterpret_cast<uintptr_t>(L.scales) & 15) == 0 &&
         (L.W_group_mode == 2 || (reinterpret_cast<uintptr_t>(L.zeros) & 15) == 0) && (L.N * 2) % 16 == 0;
}

int launch_gemv_bcast(const glb_linear& L, const void* x, const float* scales_x, void* y, void* const* y_dst, int n_dst,
                      long long y_col0, int64_t M, const void* bias, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  (void)scales_x;
  if (n_dst > 0) GLB_REQUIRE(n_dst <= 8 && gemv_bcast_supported(L, M), GLB_ERR_UNSUPPORTED, "fused output gather: M == 1 A16W4 decode layers only");
  GLB_REQUIRE(gemv_supported(L, M), GLB_ERR_UNSUPPORTED, "decode kernel: unsupported layer configuration");
  GLB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, GLB_ERR_INVALID, "decode kernel: x must be 16-byte aligned");
  GemvParams P;
  P.W = reinterpret_cast<const uint32_t*>(L.W_q);
  P.scales = L.scales; P.zeros = (L.W_group_mode == 2) ? nullptr : L.zeros;
  P.x = x; P.y = y; P.bias = bias;
  P.y_dst = y_dst; P.n_dst = n_dst; P.y_col0 = y_col0;
  P.K = (int)L.K; P.N = (int)L.N; P.M = (int)M;
  P.group_size = L.group_size; P.mode = L.W_group_mode;
  const bool bf = (L.input_dtype == GLB_BF16);
  set_kernel_name("gemv_w4_mma");
  if (M == 1) return bf ? dispatch_w4<true, 1>(L, P, workspace, ws_bytes, stream) : dispatch_w4<false, 1>(L, P, workspace, ws_bytes, stream);
  if (M <= 8) return bf ? dispatch_w4<true, 8>(L, P, workspace, ws_bytes, stream) : dispatch_w4<false, 8>(L, P, workspace, ws_bytes, stream);
  return bf ? dispatch_w4<true, 16>(L, P, workspace, ws_bytes, stream) : dispatch_w4<false, 16>(L, P, workspace, ws_bytes, stream);
}

}  // namespace glb
