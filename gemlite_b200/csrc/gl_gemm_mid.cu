// Batched-decode kernel for A16W4 (grouped INT4, fp16 / bf16), 16 < M <= 128 -- the reference's GEMM_SPLITK range
// (gemlite/triton_kernels/gemm_splitK_kernels.py:277-450, dispatch core.py:108-114) and the lower end of its GEMM range.
//
// Why a third kernel (profiles/vs_reference_timing_r2_visit1.txt): at these batch sizes the layer is still a
// weight-streaming problem (8.9 MB at 4096^2 -> 1.4 us at the HBM roof) but the CTA-pair tcgen05 kernel needs
// 17.8 / 22.8 / 33.2 us for M = 32 / 64 / 128 -- its 8-stage work items are all pipeline fill, dequant latency and
// split-K parking -- and the reference's Triton build is within 1.0-1.5x of it.  The decode kernel's recipe scales
// better here: the whole weight slice of a CTA is pulled into shared memory up front by TMA, unpacked with one
// LOP3 per two weights, and fed to the legacy tensor path (mma.sync.m16n8k16: weights = the 16-row operand, 8 tokens
// per instruction), which at M <= 128 is bound by the HMMA issue rate (8 clk per instruction and scheduler): 2.1 us per
// 32 tokens of a 4096^2 layer.
//
// Decomposition: a CLUSTER of KS = K / 1024 CTAs owns a 128-column slab; CTA `rank` owns k in [1024 rank, 1024 rank + 1024)
// (64 KB of packed words = four 128B-swizzled 32-column sub-slabs, one TMA box each).  Warp w = (sub-slab w % 4, k
// quarter w / 4) walks its 32 packed rows; per 8-row block and half it loads one uint4 of words (4 columns x 8 k),
// dequantises them EXACTLY like the reference (utils.py:57-89: integers -> fma / sub-mul in the scale dtype, one
// rounding per op) and issues 4 MMAs per 8-token tile against the x tile (<= 64 tokens, permuted to the nibble order at
// staging time, padded rows => conflict-free 128-bit loads).  fp32 accumulation in registers over the warp's k quarter;
// the four k-quarter warps meet in shared memory, the KS ranks through distributed shared memory in rank order
// (deterministic, no atomics, no workspace).  M > 64 runs a second pass over the same resident weights.
// Arithmetic contract = the tcgen05 kernels' (weights rounded to the activation dtype, fp32 accumulate).
#include <cuda.h>
#include <stdlib.h>

#include "gl_common.cuh"
#include "gl_ptx.cuh"

namespace glb {
namespace {

constexpr int MID_COLS = 128;                     // columns per cluster
constexpr int MID_SUB = 32;                       // columns per sub-slab (one 128-byte swizzled row)
constexpr int MID_NW = 16;
constexpr int MID_KC = 1024;                      // k per CTA
constexpr int MID_ROWS = MID_KC / 8;              // packed rows per CTA (128)
constexpr int MID_TOK = 64;                       // tokens per pass (the NT = 4 instantiation uses 32)
constexpr int MID_XSTRIDE = MID_ROWS * 16 + 16;   // bytes per token of the x tile: odd number of 16-byte units
constexpr int MID_WBYTES = MID_ROWS * MID_COLS * 4;          // 64 KB
constexpr int MID_XBYTES = MID_TOK * MID_XSTRIDE;            // 129 KB; reused as the [4 k quarters] x 32 KB reduction planes
constexpr int MID_SMEM = MID_WBYTES + MID_XBYTES + 64 + 1024;
static_assert(4 * MID_TOK * MID_COLS * 4 <= MID_XBYTES, "reduction buffer must fit in the x tile");

struct MidParams {
  const void* scales; const void* zeros; const void* x; void* y; const void* bias;
  int M, N, K, group_size, mode, ks;
};

template <bool BF16> __device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if (BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t mask, uint32_t orv) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(mask), "r"(orv));
  return r;
}

// one packed word (8 codes of one column) -> 4 registers of dequantised pairs (k0,k4) (k1,k5) (k2,k6) (k3,k7)
template <bool BF16> __device__ __forceinline__ void dequant_word(uint32_t q, uint32_t s2, uint32_t z2, int mode, uint32_t (&w)[4]) {
  uint32_t e[4];
  if (!BF16) {
    uint32_t t;
    asm("mul.hi.u32 %0, %1, 0x01000000;" : "=r"(t) : "r"(q));     // q >> 8 on the FMA pipe
    const uint32_t h0 = lop3_and_or(q, 0x000F000Fu, 0x64006400u), h1 = lop3_and_or(q, 0x00F000F0u, 0x64006400u);
    const uint32_t h2 = lop3_and_or(t, 0x000F000Fu, 0x64006400u), h3 = lop3_and_or(t, 0x00F000F0u, 0x64006400u);
    const __half2 k1024 = __half2half2(__ushort_as_half((unsigned short)0x6400));
    const __half2 k16th = __half2half2(__ushort_as_half((unsigned short)0x2C00));
    const __half2 km64 = __half2half2(__ushort_as_half((unsigned short)0xD400));
    __half2 r0 = __hsub2(*reinterpret_cast<const __half2*>(&h0), k1024);
    __half2 r1 = __hfma2(*reinterpret_cast<const __half2*>(&h1), k16th, km64);
    __half2 r2 = __hsub2(*reinterpret_cast<const __half2*>(&h2), k1024);
    __half2 r3 = __hfma2(*reinterpret_cast<const __half2*>(&h3), k16th, km64);
    e[0] = *reinterpret_cast<uint32_t*>(&r0); e[1] = *reinterpret_cast<uint32_t*>(&r1);
    e[2] = *reinterpret_cast<uint32_t*>(&r2); e[3] = *reinterpret_cast<uint32_t*>(&r3);
    const __half2 s = *reinterpret_cast<const __half2*>(&s2), z = *reinterpret_cast<const __half2*>(&z2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 v = *reinterpret_cast<const __half2*>(&e[i]);
      __half2 r = (mode == 4) ? __hfma2(v, s, z) : ((mode == 3) ? __hmul2(__hsub2(v, z), s) : __hmul2(v, s));
      w[i] = *reinterpret_cast<uint32_t*>(&r);
    }
  } else {
    const __nv_bfloat162 k128 = __bfloat162bfloat162(__ushort_as_bfloat16((unsigned short)0x4300));
    const __nv_bfloat162 s = *reinterpret_cast<const __nv_bfloat162*>(&s2), z = *reinterpret_cast<const __nv_bfloat162*>(&z2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t hv = lop3_and_or(q >> (4 * i), 0x000F000Fu, 0x43004300u);          // (k_i, k_{i+4})
      const __nv_bfloat162 v = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&hv), k128);
      __nv_bfloat162 r = (mode == 4) ? __hfma2(v, s, z) : ((mode == 3) ? __hmul2(__hsub2(v, z), s) : __hmul2(v, s));
      w[i] = *reinterpret_cast<uint32_t*>(&r);
    }
  }
}

__device__ __forceinline__ float4 ld_dsmem_f32x4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr) : "memory");
  return v;
}

// NT = 8-token tiles per pass (8: 64 tokens, 4: 32 tokens for M <= 32)
template <bool BF16, int NT>
__global__ void __launch_bounds__(MID_NW * 32, 1)
gemm_a16w4_mid_kernel(const __grid_constant__ CUtensorMap tmap_w, const MidParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* wbuf = smem_raw;                                   // 4 sub-slabs x [128 rows][128 B], TMA 128B swizzle
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0u) __trap();
  unsigned char* xs = wbuf + MID_WBYTES;                            // [64 tokens][MID_XSTRIDE]
  float* red = reinterpret_cast<float*>(xs);                        // [4 k quarters][4 sub-slabs][16 NT regs][32 lanes] after the main loop
  constexpr int TOK = NT * 8, NREG = 8 * NT, PLANE = 4 * NREG * 32;   // floats per k-quarter plane
  uint64_t* bars = reinterpret_cast<uint64_t*>(xs + MID_XBYTES);    // [4] one per sub-slab

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g8 = lane >> 2, t4 = lane & 3;
  const int ks = P.ks;
  const uint32_t rank = (ks > 1) ? ptx::cluster_ctarank() : 0u;
  const int slab = blockIdx.x / ks;
  const int col0 = slab * MID_COLS, kbeg = (int)rank * MID_KC;
  const int sl = warp & 3, kq = warp >> 2;

  // ---- weights: independent of the previous kernel ----
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmap_w);
    for (int c = 0; c < 4; ++c) ptx::mbar_init(&bars[c], 1);
    ptx::fence_mbar_init();
    for (int c = 0; c < 4; ++c) {
      ptx::mbar_arrive_expect_tx(&bars[c], (uint32_t)(MID_ROWS * 128));
      ptx::tma_load_2d(wbuf + (size_t)c * MID_ROWS * 128, &tmap_w, &bars[c], col0 + c * MID_SUB, kbeg >> 3);
    }
  }
  __syncthreads();
  pdl_wait_prior_grids();
  pdl_launch_dependents();

  const int gs = P.group_size, rpg = gs >> 3;                       // packed rows per group (8, 16 or 32)
  const int row_beg = kq * 32;                                      // this warp's packed rows inside the CTA
  const unsigned char* wsub = wbuf + (size_t)sl * MID_ROWS * 128;
  const uint16_t* sp = reinterpret_cast<const uint16_t*>(P.scales);
  const uint16_t* zp = reinterpret_cast<const uint16_t*>(P.zeros);
  const int ncol = col0 + sl * MID_SUB + 4 * g8;                    // first of this lane's 4 columns
  auto dup = [](uint32_t v) { return v | (v << 16); };

  for (int tok0 = 0; tok0 < P.M; tok0 += TOK) {
    // ---- stage the x tile: tokens tok0 .. tok0+TOK-1, k in [kbeg, kbeg + 1024), one permuted 16-byte word per packed row;
    //      eight loads in flight per thread (one L2 round trip per batch instead of one per word) ----
    {
      const uint16_t* xg = reinterpret_cast<const uint16_t*>(P.x) + kbeg;
      for (int i0 = threadIdx.x; i0 < TOK * MID_ROWS; i0 += 8 * MID_NW * 32) {
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = i0 + j * MID_NW * 32;
          const int m = tok0 + (i >> 7);
          v[j] = make_uint4(0u, 0u, 0u, 0u);
          if (i < TOK * MID_ROWS && m < P.M) v[j] = *reinterpret_cast<const uint4*>(xg + (size_t)m * P.K + (size_t)(i & (MID_ROWS - 1)) * 8);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = i0 + j * MID_NW * 32;
          if (i >= TOK * MID_ROWS) continue;
          uint4 o;
          o.x = __byte_perm(v[j].x, v[j].z, 0x5410);   // (k0,k4)
          o.y = __byte_perm(v[j].x, v[j].z, 0x7632);   // (k1,k5)
          o.z = __byte_perm(v[j].y, v[j].w, 0x5410);   // (k2,k6)
          o.w = __byte_perm(v[j].y, v[j].w, 0x7632);   // (k3,k7)
          *reinterpret_cast<uint4*>(xs + (size_t)(i >> 7) * MID_XSTRIDE + (size_t)(i & (MID_ROWS - 1)) * 16) = o;
        }
      }
    }
    __syncthreads();
    if (tok0 == 0) {
      if (lane == 0) ptx::mbar_wait(&bars[sl], 0);
      __syncwarp();
    }

    float acc[2][NT][4];
#pragma unroll
    for (int pr = 0; pr < 2; ++pr)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[pr][nt][0] = acc[pr][nt][1] = acc[pr][nt][2] = acc[pr][nt][3] = 0.f;

    for (int rg = 0; rg < 32; rg += rpg) {                            // groups of this warp's k quarter
      const int g = (kbeg >> 3) / rpg + (row_beg + rg) / rpg;         // global group index
      const uint2 sw = *reinterpret_cast<const uint2*>(sp + (size_t)g * P.N + ncol);
      uint2 zw = make_uint2(0u, 0u);
      if (P.mode != 2) zw = *reinterpret_cast<const uint2*>(zp + (size_t)g * P.N + ncol);
      const uint32_t s2[4] = {dup(sw.x & 0xffffu), dup(sw.x >> 16), dup(sw.y & 0xffffu), dup(sw.y >> 16)};
      const uint32_t z2[4] = {dup(zw.x & 0xffffu), dup(zw.x >> 16), dup(zw.y & 0xffffu), dup(zw.y >> 16)};
      for (int rb = 0; rb < rpg; rb += 8) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int row = row_beg + rg + rb + 2 * t4 + a;             // packed row inside the CTA; row % 8 == 2 t4 + a
          const uint4 q = *reinterpret_cast<const uint4*>(wsub + (size_t)row * 128 + ((g8 ^ (2 * t4 + a)) << 4));
          uint32_t w[4][4];
          dequant_word<BF16>(q.x, s2[0], z2[0], P.mode, w[0]);
          dequant_word<BF16>(q.y, s2[1], z2[1], P.mode, w[1]);
          dequant_word<BF16>(q.z, s2[2], z2[2], P.mode, w[2]);
          dequant_word<BF16>(q.w, s2[3], z2[3], P.mode, w[3]);
          const unsigned char* xr = xs + (size_t)g8 * MID_XSTRIDE + (size_t)row * 16;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const uint4 xv = *reinterpret_cast<const uint4*>(xr + (size_t)nt * 8 * MID_XSTRIDE);
            mma16816<BF16>(acc[0][nt], w[0][0], w[1][0], w[0][1], w[1][1], xv.x, xv.y);
            mma16816<BF16>(acc[0][nt], w[0][2], w[1][2], w[0][3], w[1][3], xv.z, xv.w);
            mma16816<BF16>(acc[1][nt], w[2][0], w[3][0], w[2][1], w[3][1], xv.x, xv.y);
            mma16816<BF16>(acc[1][nt], w[2][2], w[3][2], w[2][3], w[3][3], xv.z, xv.w);
          }
        }
      }
    }

    // ---- reduction: k-quarter warps through shared memory, cluster ranks through DSMEM, both in fixed order.  The
    //      partial tiles stay in FRAGMENT order ([sub-slab][register][lane]: conflict-free 32-bit stores, 128-bit loads)
    //      until the very last step decodes (register, lane) -> (token, column) for the store ----
    __syncthreads();                                                  // every warp is done reading the x tile
    {
      float* plane = red + (size_t)kq * PLANE + (size_t)sl * NREG * 32 + lane;
#pragma unroll
      for (int pr = 0; pr < 2; ++pr)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) plane[((pr * NT + nt) * 4 + i) * 32] = acc[pr][nt][i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PLANE / 4; i += MID_NW * 32) {
      float4* p0 = reinterpret_cast<float4*>(red) + i;
      const float4 a = p0[0], b = p0[PLANE / 4], c = p0[2 * (PLANE / 4)], d = p0[3 * (PLANE / 4)];
      p0[0] = make_float4((a.x + b.x) + (c.x + d.x), (a.y + b.y) + (c.y + d.y), (a.z + b.z) + (c.z + d.z), (a.w + b.w) + (c.w + d.w));
    }
    if (ks > 1) ptx::cluster_sync(); else __syncthreads();
    {
      const int per = (PLANE / 4) / ks;                               // float4 groups this rank finishes
      const uint32_t red_addr = ptx::smem_u32(red);
      for (int j = threadIdx.x; j < per; j += MID_NW * 32) {
        const int q4 = (int)rank * per + j;                           // float4 index: 4 consecutive lanes of one register
        float4 v;
        if (ks > 1) {
          v = ld_dsmem_f32x4(ptx::mapa(red_addr + (uint32_t)q4 * 16u, 0u));
          for (int r = 1; r < ks; ++r) {
            const float4 u = ld_dsmem_f32x4(ptx::mapa(red_addr + (uint32_t)q4 * 16u, (uint32_t)r));
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
          }
        } else {
          v = reinterpret_cast<const float4*>(red)[q4];
        }
        // decode: element e = q4 * 4 + t4 -> (sub-slab, register = (pr * NT + nt) * 4 + i, lane = 4 g8 + t4)
        const int e = q4 * 4;
        const int slot = e / (NREG * 32), reg = (e >> 5) % NREG, gq = (e & 31) >> 2;
        const int i = reg & 3, nt = (reg >> 2) % NT, pr = reg / (4 * NT);
        const int col = col0 + slot * MID_SUB + 4 * gq + 2 * pr + (i >> 1);
        const float vv[4] = {v.x, v.y, v.z, v.w};
        float bv = 0.f;
        if (P.bias) bv = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.bias)[col]) : __half2float(reinterpret_cast<const __half*>(P.bias)[col]);
#pragma unroll
        for (int tq = 0; tq < 4; ++tq) {                              // the four lanes of the group: tokens 2 t4 + (i & 1)
          const int m = tok0 + nt * 8 + 2 * tq + (i & 1);
          if (m >= P.M) continue;
          float o = vv[tq];
          const int64_t oi = (int64_t)m * P.N + col;
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
    if (ks > 1) ptx::cluster_sync(); else __syncthreads();           // peers are done with this CTA's partials
  }
}

}  // namespace

bool gemm_mid_supported(const glb_linear& L, int64_t M) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("GLB_MID_KERNEL"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return false;
  if (M <= 16 || M > 128) return false;
  if (!(L.input_dtype == GLB_FP16 || L.input_dtype == GLB_BF16)) return false;
  if (L.output_dtype != L.input_dtype || L.meta_dtype != L.input_dtype) return false;
  if (L.W_nbits != 4 || L.elements_per_sample != 8 || L.wq_dtype != GLB_INT32) return false;
  if (L.wq_stride1 != 1 || L.wq_stride0 != L.N) return false;
  if (L.channel_scale_mode != 0) return false;
  if (!(L.W_group_mode == 2 || L.W_group_mode == 3 || L.W_group_mode == 4)) return false;
  if (L.W_group_mode != 2 && (L.zeros_numel != L.scales_numel || L.zeros_dtype != L.scales_dtype)) return false;
  const int gs = L.group_size;
  if (!(gs == 64 || gs == 128 || gs == 256)) return false;
  if (!(L.K == 1024 || L.K == 2048 || L.K == 4096 || L.K == 8192)) return false;
  if (L.N % MID_COLS != 0) return false;
  if (L.scales_numel != (L.K / gs) * L.N || L.s_stride1 != 1 || L.s_stride0 != L.N) return false;
  if (L.W_group_mode != 2 && (L.z_stride1 != 1 || L.z_stride0 != L.N)) return false;
  if ((reinterpret_cast<uintptr_t>(L.W_q) & 15) || (reinterpret_cast<uintptr_t>(L.scales) & 7)) return false;
  if (L.W_group_mode != 2 && (reinterpret_cast<uintptr_t>(L.zeros) & 7)) return false;
  return true;
}

int launch_gemm_mid(const glb_linear& L, const void* x, void* y, int64_t M, const void* bias, cudaStream_t stream) {
  GLB_REQUIRE(gemm_mid_supported(L, M), GLB_ERR_UNSUPPORTED, "batched-decode kernel: unsupported layer configuration");
  GLB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, GLB_ERR_INVALID, "batched-decode kernel: x must be 16-byte aligned");
  CUtensorMap tw;
  int rc = make_tmap_2d(&tw, GLB_INT32, L.W_q, (uint64_t)L.N, (uint64_t)(L.K / 8), (uint64_t)L.N * 4, MID_SUB, MID_ROWS, 1);
  if (rc != GLB_OK) return rc;
  MidParams P;
  P.scales = L.scales; P.zeros = (L.W_group_mode == 2) ? nullptr : L.zeros; P.x = x; P.y = y; P.bias = bias;
  P.M = (int)M; P.N = (int)L.N; P.K = (int)L.K; P.group_size = L.group_size; P.mode = L.W_group_mode;
  P.ks = (int)(L.K / MID_KC);
  const bool bf = (L.input_dtype == GLB_BF16);
  auto kern = (M <= 32) ? (bf ? gemm_a16w4_mid_kernel<true, 4> : gemm_a16w4_mid_kernel<false, 4>)
                        : (bf ? gemm_a16w4_mid_kernel<true, 8> : gemm_a16w4_mid_kernel<false, 8>);
  GLB_ENSURE_SMEM(kern, MID_SMEM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(L.N / MID_COLS) * (unsigned)P.ks);
  cfg.blockDim = dim3(MID_NW * 32);
  cfg.dynamicSmemBytes = MID_SMEM;
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
  set_kernel_name("gemm_a16w4_mma_mid");
  GLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tw, P));
  count_launch();
  return GLB_OK;
}

}  // namespace glb
