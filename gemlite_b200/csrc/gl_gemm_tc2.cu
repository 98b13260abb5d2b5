// tcgen05 prefill GEMM, CTA-pair flavour (cta_group::2) with the dequantised weights fed to the tensor core
// from TENSOR MEMORY -- the large-M path of A16W4 (rows a9/a10, config C3).
//
// Why (profiles/prefill_tc1_r1_ncu.txt): the single-CTA kernel in gl_gemm_tc.cu keeps the tensor pipe only ~49 %
// busy because the shared-memory data pipe is saturated: per 64-wide k-block a CTA moves 48 KB of operand
// reads (A 16 KB + B 32 KB), 36 KB of TMA writes and 16 KB (x2 bank conflicts) of dequant stores = ~93 % of the
// 128 B/clk port.  Here
//   * a CTA PAIR owns a 256-feature x 256-token tile: tcgen05.mma.cta_group::2, M = 256 (128 TMEM lanes per
//     CTA), N = 256; each CTA stages only ITS 128 tokens of the x tile and the tensor cores of both SMs read
//     both halves -> B traffic per CTA halves;
//   * the dequantised weight tile never touches shared memory: dequant thread == feature row == TMEM lane,
//     16 packed pairs per thread go to tensor memory with one tcgen05.st.32x32b.x16 and the MMA takes its A
//     operand from there ([a_tmem] form).  No STS, no bank conflicts, no async-proxy fence;
//   * a pipeline stage is 128 k wide (two 128B-swizzled x boxes of 16 KB, 4 stages; the A ring in TMEM is
//     4 x 64 columns next to the 256-column fp32 accumulator), so every per-stage cost -- barrier round trips,
//     the MMA issuer's loop body, the producers' bookkeeping -- is paid once per 8 MMAs (1024 tensor-core cycles).
//     The MMA issuer waits on ONE barrier per stage (`ready`: TMA bytes of both CTAs + the arrivals of all 32
//     dequant warps of the pair);
//   * the 8 KB packed-word tiles have their OWN 8-deep ring and producer: they are requested two A-ring
//     revolutions ahead, so when the MMAs free A slot s the dequant warps find their input already in shared
//     memory -- the TMA latency is no longer inside the free-slot -> ready chain (tensor pipe 74 % -> see profiles/).
// Shared-memory traffic per CTA and 128 k: 32 KB operand reads + 40 KB TMA writes + 8 KB packed reads.
//
// Arithmetic contract is unchanged (see gl_gemm_tc.cu): weights are dequantised with the reference's rounding
// (utils.py:57-89) and accumulated in fp32 by the tensor core (gemm_kernels.py:384).
//
// Warp roles per CTA (896 threads): w0-15 dequant (warp % 4 = TMEM lane quarter, warp / 4 = which 32-k quarter of the
// stage), w16-23 epilogue (two per lane quarter: the accumulator is single-buffered, its drain is exposed), w24 x
// producer, w25 packed-word producer, w26 TMEM allocator, w27 MMA issuer (leader CTA only).  Requires group_size % 32 == 0 and K % 128 == 0.
// Barriers: ready / tempty live in the LEADER's shared memory (the peer reports to them remotely);
// pfull / pfree / empty / tfull are per CTA; tcgen05.commit multicasts its arrival to both CTAs.
// History of this kernel's profile: profiles/prefill_tc2_*_r1_ncu.txt.
#include <stdlib.h>
#include <algorithm>

#include "gl_common.cuh"
#include "gl_ptx.cuh"

namespace glb {
namespace {

constexpr int T2_BM = 128;            // features per CTA (256 per pair)
constexpr int T2_BT = 256;            // tokens per pair tile (128 staged per CTA)
constexpr int T2_BK = 128;            // k per pipeline stage
constexpr int T2_STAGES = 4;
constexpr int T2_THREADS = 896;
constexpr int T2_DEQ_WARP0 = 0, T2_DEQ_WARPS = 16, T2_EPI_WARP0 = 16, T2_EPI_WARPS = 8;   // 16 dequant warps: [32-k quarter][lane quarter]
// The latency-critical single-lane roles get the HIGHEST warp ids: the warp scheduler favours high ids among
// eligible warps (B300_MICROARCH.md, "hi-wid-first"); as warp 1 the MMA issuer spent ~45 % of every stage just
// executing its ~55 loop instructions behind the dequant warps of its sub-partition.
constexpr int T2_XPROD_WARP = 24, T2_PPROD_WARP = 25, T2_ALLOC_WARP = 26, T2_MMA_WARP = 27;
constexpr int T2_XBOX_BYTES = (T2_BT / 2) * 64 * 2;        // one 64-k x box: 128 tokens x 128 B = 16 KB
constexpr int T2_X_BYTES = 2 * T2_XBOX_BYTES;              // 32 KB
constexpr int T2_P_BYTES = (T2_BK / 8) * T2_BM * 4;        // 8 KB
constexpr int T2_PSTAGES = 8;                              // packed-word ring (decoupled from the x / A ring, see below)
constexpr int T2_P_OFF = T2_STAGES * T2_X_BYTES;           // packed ring starts after the x ring (128 KB)
constexpr int T2_BAR_OFF = T2_P_OFF + T2_PSTAGES * T2_P_BYTES;   // 192 KB
constexpr int T2_SMEM = T2_BAR_OFF + 1024 /* barriers */ + 1024 /* alignment slack */;
constexpr uint32_t T2_ACOL0 = 256;    // first TMEM column of the A ring (accumulator: columns 0..255)
constexpr uint32_t T2_ACOLS = T2_BK / 2;   // 64 columns of A per stage

struct Tc2Params {
  const void* scales;
  const void* zeros;
  void* y;
  const void* bias;
  int M, N, K;
  int group_size;
  int mode;
  int post_scale;         // 8-bit weight-only: multiply the accumulator by the per-feature fp32 scale (channel_scale_mode 1)
  int tiles_m;            // 256-token blocks
  int tiles_n;            // 256-feature blocks
  int prefetch;           // x stages of L2 prefetch distance (0 = off)
  // split-K (few tiles, e.g. M <= 256 on a 4096-wide layer): work item = (tile, split); every CTA writes its fp32 partial
  // slice to `ws_partial`, the LAST epilogue warp to finish a slice (per-slice counter) sums the partials in split order
  // -> deterministic, no second kernel, counters return to zero
  int splits;             // >= 1
  int kb_per_split;       // 128-k stages per work item
  int n_mma;              // MMA N (tokens per tile actually computed): 64 / 128 / 256
  float* ws_partial;      // [tile][split][rank][128 rows][n_mma]  (splits > 1)
  int* ws_counter;        // [tile][rank][quarter], zero on entry and on exit
};

template <bool BF16> __device__ __forceinline__ void unpack8_nat(uint32_t q, uint32_t (&o)[4]) {
  // 8 nibbles -> exact integers as 16-bit floats in natural k order: o[i] = (k_{2i}, k_{2i+1})
  uint32_t h[4];
  if (!BF16) {
    const uint32_t t = q >> 8;
    uint32_t m0, m1, m2, m3;
    asm("lop3.b32 %0, %1, 0x000F000F, 0x64006400, 0xEA;" : "=r"(m0) : "r"(q));
    asm("lop3.b32 %0, %1, 0x00F000F0, 0x64006400, 0xEA;" : "=r"(m1) : "r"(q));
    asm("lop3.b32 %0, %1, 0x000F000F, 0x64006400, 0xEA;" : "=r"(m2) : "r"(t));
    asm("lop3.b32 %0, %1, 0x00F000F0, 0x64006400, 0xEA;" : "=r"(m3) : "r"(t));
    const __half2 k1024 = __half2half2(__ushort_as_half((unsigned short)0x6400));
    const __half2 k16th = __half2half2(__ushort_as_half((unsigned short)0x2C00));
    const __half2 km64 = __half2half2(__ushort_as_half((unsigned short)0xD400));
    __half2 r0 = __hsub2(*reinterpret_cast<__half2*>(&m0), k1024);           // (k0,k4)
    __half2 r1 = __hfma2(*reinterpret_cast<__half2*>(&m1), k16th, km64);      // (k1,k5)
    __half2 r2 = __hsub2(*reinterpret_cast<__half2*>(&m2), k1024);           // (k2,k6)
    __half2 r3 = __hfma2(*reinterpret_cast<__half2*>(&m3), k16th, km64);      // (k3,k7)
    h[0] = *reinterpret_cast<uint32_t*>(&r0); h[1] = *reinterpret_cast<uint32_t*>(&r1);
    h[2] = *reinterpret_cast<uint32_t*>(&r2); h[3] = *reinterpret_cast<uint32_t*>(&r3);
  } else {
    const __nv_bfloat162 k128 = __bfloat162bfloat162(__ushort_as_bfloat16((unsigned short)0x4300));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t v;
      asm("lop3.b32 %0, %1, 0x000F000F, 0x43004300, 0xEA;" : "=r"(v) : "r"(q >> (4 * i)));
      __nv_bfloat162 r = __hsub2(*reinterpret_cast<__nv_bfloat162*>(&v), k128);   // (k_i, k_{i+4})
      h[i] = *reinterpret_cast<uint32_t*>(&r);
    }
  }
  o[0] = __byte_perm(h[0], h[1], 0x5410);   // (k0,k1)
  o[1] = __byte_perm(h[2], h[3], 0x5410);   // (k2,k3)
  o[2] = __byte_perm(h[0], h[1], 0x7632);   // (k4,k5)
  o[3] = __byte_perm(h[2], h[3], 0x7632);   // (k6,k7)
}

// Packed int32 words of other widths (A16W8 / A16W2 / A16W1 HQQ layers, packing_bitwidth 32; bitpack.py): element j of a
// word sits at bits [NBITS j, NBITS (j+1)).  One word -> 16 / NBITS pairs of exact integers in natural k order.
template <bool BF16, int NBITS> __device__ __forceinline__ void unpack_word_nat(uint32_t q, uint32_t (&o)[16 / NBITS]) {
  if (NBITS == 8) {
    uint32_t m0, m1;
    asm("prmt.b32 %0, %1, %2, 0x5140;" : "=r"(m0) : "r"(q), "r"(0x64646464u));     // (1024 + u0, 1024 + u1)
    asm("prmt.b32 %0, %1, %2, 0x7362;" : "=r"(m1) : "r"(q), "r"(0x64646464u));     // (1024 + u2, 1024 + u3)
    const __half2 k1024 = __half2half2(__ushort_as_half((unsigned short)0x6400));
    __half2 r0 = __hsub2(*reinterpret_cast<const __half2*>(&m0), k1024), r1 = __hsub2(*reinterpret_cast<const __half2*>(&m1), k1024);
    if (BF16) {                                              // 0..255 needs 8 significand bits: exact in bf16
      const float2 a = __half22float2(r0), b = __half22float2(r1);
      __nv_bfloat162 c0 = __floats2bfloat162_rn(a.x, a.y), c1 = __floats2bfloat162_rn(b.x, b.y);
      o[0] = *reinterpret_cast<uint32_t*>(&c0); o[1] = *reinterpret_cast<uint32_t*>(&c1);
    } else {
      o[0] = *reinterpret_cast<uint32_t*>(&r0); o[1] = *reinterpret_cast<uint32_t*>(&r1);
    }
  } else {
    constexpr int H = 16 / NBITS;                            // elements per 16-bit half of the word
    constexpr uint32_t MASK = ((1u << NBITS) - 1u) * 0x00010001u;
    uint32_t h[H];
#pragma unroll
    for (int i = 0; i < H; ++i) {                            // h[i] = (k_i, k_{i+H})
      uint32_t v;
      if (!BF16) {
        asm("lop3.b32 %0, %1, %2, 0x64006400, 0xEA;" : "=r"(v) : "r"(q >> (NBITS * i)), "r"(MASK));
        __half2 r = __hsub2(*reinterpret_cast<__half2*>(&v), __half2half2(__ushort_as_half((unsigned short)0x6400)));
        h[i] = *reinterpret_cast<uint32_t*>(&r);
      } else {
        asm("lop3.b32 %0, %1, %2, 0x43004300, 0xEA;" : "=r"(v) : "r"(q >> (NBITS * i)), "r"(MASK));
        __nv_bfloat162 r = __hsub2(*reinterpret_cast<__nv_bfloat162*>(&v), __bfloat162bfloat162(__ushort_as_bfloat16((unsigned short)0x4300)));
        h[i] = *reinterpret_cast<uint32_t*>(&r);
      }
    }
#pragma unroll
    for (int j = 0; j < H / 2; ++j) {
      o[j] = __byte_perm(h[2 * j], h[2 * j + 1], 0x5410);            // (k_{2j}, k_{2j+1})
      o[H / 2 + j] = __byte_perm(h[2 * j], h[2 * j + 1], 0x7632);    // (k_{H+2j}, k_{H+2j+1})
    }
  }
}

template <bool BF16> __device__ __forceinline__ uint32_t dequant2_ref(uint32_t q2, uint32_t s2, uint32_t z2, int mode) {
  if (!BF16) {
    const __half2 q = *reinterpret_cast<__half2*>(&q2), s = *reinterpret_cast<__half2*>(&s2), z = *reinterpret_cast<__half2*>(&z2);
    __half2 r = (mode == 4) ? __hfma2(q, s, z) : ((mode == 3) ? __hmul2(__hsub2(q, z), s) : __hmul2(q, s));
    return *reinterpret_cast<uint32_t*>(&r);
  } else {
    const __nv_bfloat162 q = *reinterpret_cast<__nv_bfloat162*>(&q2), s = *reinterpret_cast<__nv_bfloat162*>(&s2),
                         z = *reinterpret_cast<__nv_bfloat162*>(&z2);
    __nv_bfloat162 r = (mode == 4) ? __hfma2(q, s, z) : ((mode == 3) ? __hmul2(__hsub2(q, z), s) : __hmul2(q, s));
    return *reinterpret_cast<uint32_t*>(&r);
  }
}

// Tile order: token blocks are walked in groups of T2_GROUP_M; inside a group the feature tile is the slow index.
// The ~74 tiles in flight then share 4 token blocks (16 MB of x at K = 8192) instead of all of them: the x working
// set of a wave stays resident in one die's share of the L2 (the r1 profile of the plain token-fastest order read x
// from DRAM once per wave: 438 MB of DRAM reads for a 104 MB problem, and the x boxes arrived late).
constexpr int T2_GROUP_M = 4;
__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int& tm, int& tn) {
  const int per_group = T2_GROUP_M * tiles_n;
  const int grp = t / per_group, r = t - grp * per_group;
  const int m0 = grp * T2_GROUP_M;
  const int gsz = (tiles_m - m0 < T2_GROUP_M) ? (tiles_m - m0) : T2_GROUP_M;     // last group may be short
  tn = r / gsz;
  tm = m0 + (r - tn * gsz);
}

// MXFP4 weights (A16W4_MXFP, config C5): two e2m1 codes per byte (even k = low nibble), storage (N, K/2) K-contiguous,
// one e8m0 scale byte per 32 codes stored (K/32, N).  A dequant thread's 32-k span of a stage is exactly one group:
// 16 bytes of codes + one scale.  sm_100a converts a byte to two fp16 values in one instruction
// (cvt.rn.f16x2.e2m1x2 -> F2FP.F16.E2M1.UNPACK_B); fp16 compute multiplies by the group scale in fp16 (exact while the
// product is a normal fp16 number), bf16 compute goes through fp32 (e2m1 x 2^e is always exact in bf16).
__device__ __forceinline__ uint32_t e2m1x2_to_f16x2(uint32_t byte) {
  uint32_t r;
  asm("{\n\t.reg .b8 t;\n\tcvt.u8.u32 t, %1;\n\tcvt.rn.f16x2.e2m1x2 %0, t;\n\t}" : "=r"(r) : "r"(byte));
  return r;
}
template <bool BF16> __device__ __forceinline__ uint32_t mx_scale_pair(uint32_t v16x2, float sf, uint32_t s16x2) {
  if (!BF16) {
    __half2 r = __hmul2(*reinterpret_cast<__half2*>(&v16x2), *reinterpret_cast<__half2*>(&s16x2));
    return *reinterpret_cast<uint32_t*>(&r);
  } else {
    const float2 f = __half22float2(*reinterpret_cast<__half2*>(&v16x2));
    __nv_bfloat162 r = __floats2bfloat162_rn(f.x * sf, f.y * sf);
    return *reinterpret_cast<uint32_t*>(&r);
  }
}

constexpr int WFMT_INT4 = 0, WFMT_MXFP4 = 1, WFMT_W8I = 2, WFMT_W8F = 3;
// 8-bit weight-only layers (A16W8 int8 / e4m3, helper.py:88-170): storage (N, K) bytes K-contiguous, one fp32 scale per
// feature.  A stage is 128 rows x 128 bytes (16 KB, TMA 128B swizzle: a thread's two 16-byte chunks land on distinct
// banks across 8 consecutive rows), 4-deep ring.  Pre-scaling (W_group_mode 2): w = round16(float(q) * s), the
// reference's rounding point (b.to(fp32) * scales -> compute dtype); post-scaling (channel_scale_mode 1): the exact
// integer / e4m3 value goes to the tensor core and the accumulator is multiplied by s in the epilogue.
constexpr int T2_P8_BYTES = T2_BK * T2_BM;                 // 16 KB
constexpr int T2_P8_STAGES = 4;
static_assert(T2_P8_STAGES * T2_P8_BYTES == T2_PSTAGES * T2_P_BYTES, "both packed rings fill the same 64 KB");

// 4 weight bytes (k .. k+3) -> two 16-bit-float pairs (k, k+1), (k+2, k+3): exact values, optionally pre-scaled in fp32
template <bool BF16, int WFMT> __device__ __forceinline__ void widen4_w8(uint32_t w, float sc, bool prescale, uint32_t& lo, uint32_t& hi) {
  uint32_t l16, h16;                                         // fp16 pairs holding the exact weight values
  if (WFMT == WFMT_W8I) {
    const uint32_t u = w ^ 0x80808080u;                      // signed -> offset-128
    uint32_t m0, m1;
    asm("prmt.b32 %0, %1, %2, 0x5140;" : "=r"(m0) : "r"(u), "r"(0x64646464u));     // (1024 + u0, 1024 + u1)
    asm("prmt.b32 %0, %1, %2, 0x7362;" : "=r"(m1) : "r"(u), "r"(0x64646464u));     // (1024 + u2, 1024 + u3)
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

template <bool BF16, int WFMT, int NBITS>
__global__ void __launch_bounds__(T2_THREADS, 1)
gemm_a16w4_tc2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const Tc2Params P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T2_BAR_OFF);
  uint64_t* ready = bars;                         // [STAGES]  leader: x bytes of both CTAs landed + all 32 dequant warps wrote A slot s
  uint64_t* empty = bars + T2_STAGES;             // [STAGES]  local : MMAs reading x stage / A slot s completed (multicast commit)
  uint64_t* pfull = bars + 2 * T2_STAGES;         // [PSTAGES] local : packed words landed
  uint64_t* pfree = pfull + T2_PSTAGES;           // [PSTAGES] local : all 16 dequant warps have read the packed words
  uint64_t* tfull = pfree + T2_PSTAGES;           // [1]       local : accumulator complete (multicast commit)
  uint64_t* tempty = tfull + 1;                   // [1]       leader: both CTAs' epilogue warps drained it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = (rank == 0);
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int kblocks = P.kb_per_split;                 // stages per work item
  constexpr bool W8 = (WFMT == WFMT_W8I || WFMT == WFMT_W8F);
  // packed ring depth / stage bytes of this weight format: packed int32 words hold 32 / NBITS elements -> 4 NBITS word
  // rows of 128 features per 128-k stage (NBITS = 4: the 8 KB of T2_P_BYTES)
  constexpr int PBYTES = W8 ? T2_P8_BYTES : ((WFMT == WFMT_INT4) ? 4 * NBITS * T2_BM * 4 : T2_P_BYTES);
  constexpr int PSTAGES = (PBYTES > T2_P_BYTES) ? T2_P8_STAGES : T2_PSTAGES;
  static_assert(PSTAGES * PBYTES <= T2_PSTAGES * T2_P_BYTES, "packed ring overflows its 64 KB");
  const int ntiles = P.tiles_m * P.tiles_n * P.splits;   // work items; item w = (tile w / splits, split w % splits)

  if (warp == T2_XPROD_WARP && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_w);
    for (int s = 0; s < T2_STAGES; ++s) {
      ptx::mbar_init(&ready[s], 1 + 2 * T2_DEQ_WARPS);   // the leader producer's expect_tx arrival + 32 dequant warps
      ptx::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < PSTAGES; ++s) {
      ptx::mbar_init(&pfull[s], 1);
      ptx::mbar_init(&pfree[s], T2_DEQ_WARPS);
    }
    ptx::mbar_init(tfull, 1);
    ptx::mbar_init(tempty, 2 * T2_EPI_WARPS);
    ptx::fence_mbar_init();
  }
  if (warp == T2_ALLOC_WARP) ptx::tmem_alloc_2cta(tmem_slot, 512);
  ptx::tc_fence_before_sync();
  ptx::cluster_sync();                 // barrier inits of BOTH CTAs visible before any remote arrival / TMA completion
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == T2_XPROD_WARP) {
    // ================= x producer (one lane in each CTA): follows the MMA's `empty` signal =================
    if (lane == 0) {
      const uint32_t ready_leader = ptx::mapa(ptx::smem_u32(&ready[0]), 0u);
      const int tok_off = (int)rank * (P.n_mma / 2);     // B rows of the MMA: tokens [0, n/2) from the leader, [n/2, n) from the peer
      uint32_t s = 0, ph = 0;
      for (int t = pair; t < ntiles; t += npairs) {
        int tm, tn;
        tile_coords(t / P.splits, P.tiles_m, P.tiles_n, tm, tn);
        const int tok0 = tm * T2_BT + tok_off;
        const int kb0 = (t % P.splits) * kblocks;
        for (int kb = kb0; kb < kb0 + kblocks; ++kb) {
          ptx::mbar_wait(&empty[s], ph ^ 1u);
          unsigned char* st = smem + s * T2_X_BYTES;
          if (leader) ptx::mbar_arrive_expect_tx(&ready[s], 2 * T2_X_BYTES);
          ptx::tma_load_2d_2cta(st, &tmap_x, ready_leader + s * 8u, kb * T2_BK, tok0);
          ptx::tma_load_2d_2cta(st + T2_XBOX_BYTES, &tmap_x, ready_leader + s * 8u, kb * T2_BK + 64, tok0);
          if (P.prefetch && kb + P.prefetch < kb0 + kblocks) {      // warm the L2 for the stage `prefetch` ahead of this one
            ptx::tma_prefetch_2d(&tmap_x, (kb + P.prefetch) * T2_BK, tok0);
            ptx::tma_prefetch_2d(&tmap_x, (kb + P.prefetch) * T2_BK + 64, tok0);
          }
          if (++s == T2_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == T2_PPROD_WARP) {
    // ================= packed-word producer: runs up to 8 k-blocks ahead, gated only by the dequant warps' reads =================
    if (lane == 0) {
      const int feat_off = (int)rank * T2_BM;
      uint32_t s = 0, ph = 0;
      for (int t = pair; t < ntiles; t += npairs) {
        int tm, tn;
        tile_coords(t / P.splits, P.tiles_m, P.tiles_n, tm, tn);
        const int feat0 = tn * 2 * T2_BM + feat_off;
        const int kb0 = (t % P.splits) * kblocks;
        for (int kb = kb0; kb < kb0 + kblocks; ++kb) {
          ptx::mbar_wait(&pfree[s], ph ^ 1u);
          ptx::mbar_arrive_expect_tx(&pfull[s], PBYTES);
          if (WFMT == WFMT_INT4) ptx::tma_load_2d(smem + T2_P_OFF + s * PBYTES, &tmap_w, &pfull[s], feat0, kb * (4 * NBITS));
          else if (W8) ptx::tma_load_2d(smem + T2_P_OFF + s * PBYTES, &tmap_w, &pfull[s], kb * T2_BK, feat0);   // [128 B][128 rows], swizzled
          else ptx::tma_load_2d(smem + T2_P_OFF + s * PBYTES, &tmap_w, &pfull[s], kb * (T2_BK / 2), feat0);   // [64 B of codes][128 rows]
          if (++s == PSTAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == T2_MMA_WARP) {
    // ================= MMA issuer (leader CTA; the warp stays converged, one elected lane issues) =================
    // The loop body is paid once per 8 MMAs (1024 tensor-core cycles): one barrier wait, one fence, 8 x tcgen05.mma
    // whose operands are stage base + constants, one commit.
    if (leader) {
      const uint32_t idesc = ptx::make_idesc(1u, BF16 ? 1u : 0u, BF16 ? 1u : 0u, 2 * T2_BM, (uint32_t)P.n_mma);
      const uint64_t bdesc0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem));
      uint32_t s = 0, ph = 0;
      int it = 0;
      for (int t = pair; t < ntiles; t += npairs, ++it) {
        ptx::mbar_wait(tempty, (uint32_t)(it & 1) ^ 1u);      // epilogue of the previous tile drained the accumulator
        ptx::tc_fence_after_sync();
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&ready[s], ph);
          ptx::tc_fence_after_sync();
          if (ptx::elect_one()) {
            const uint64_t bd = bdesc0 + (uint64_t)(s * (uint32_t)(T2_X_BYTES >> 4));
            const uint32_t at = tmem_base + T2_ACOL0 + s * T2_ACOLS;
#pragma unroll
            for (int k = 0; k < T2_BK / 16; ++k) {
              // K = 16 per instruction: 8 TMEM columns of A; B: box k/4 (16 KB apart), 32 bytes (2 descriptor units) per step
              const uint64_t bk = bd + (uint64_t)((k >> 2) * (T2_XBOX_BYTES >> 4) + (k & 3) * 2);
              ptx::mma_f16_ts_2cta(tmem_base, at + (uint32_t)(8 * k), bk, idesc, (kb | k) ? 1u : 0u);
            }
            ptx::mma_commit_2cta(&empty[s], (uint16_t)3);
            if (kb == kblocks - 1) ptx::mma_commit_2cta(tfull, (uint16_t)3);
          }
          __syncwarp();
          if (++s == T2_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp >= T2_DEQ_WARP0 && warp < T2_DEQ_WARP0 + T2_DEQ_WARPS) {
    // ================= dequant: thread == feature row == TMEM lane; 4 packed words (32 k) per thread and stage =================
    const int d = warp - T2_DEQ_WARP0;
    const int q4 = d & 3;                                   // TMEM lane quarter this warp may touch (== warp % 4)
    const int row = q4 * 32 + lane;                         // feature row inside this CTA's 128
    const int kq = d >> 2;                                  // which 32-k quarter of the 128-k stage
    const uint32_t taddr0 = tmem_base + ((uint32_t)(q4 * 32) << 16) + T2_ACOL0 + (uint32_t)kq * 16u;
    const uint32_t ready_leader = ptx::mapa(ptx::smem_u32(&ready[0]), 0u);
    const uint32_t* pk0 = reinterpret_cast<const uint32_t*>(smem + T2_P_OFF) + (kq * NBITS) * T2_BM + row;   // this thread's NBITS words of the stage
    auto dup = [](uint16_t v) { return (uint32_t)v | ((uint32_t)v << 16); };
    uint32_t s = 0, ph = 0;                                 // x / A ring position
    uint32_t ps = 0, pph = 0;                               // packed ring position
    for (int t = pair; t < ntiles; t += npairs) {
      int tm, tn;
      tile_coords(t / P.splits, P.tiles_m, P.tiles_n, tm, tn);
      const int n = tn * 2 * T2_BM + (int)rank * T2_BM + row;
      const int kstage0 = (t % P.splits) * kblocks;         // first 128-k stage of this work item
      uint32_t o[16];
      // ---------------- INT4 grouped metadata state ----------------
      // this warp's k offsets are kq*32 + 128*kb: a 32-wide span never straddles a group (group_size % 32 == 0)
      int g = 0, rem = 0;
      const uint16_t* sp = nullptr;
      const uint16_t* zp = nullptr;
      uint16_t sraw = 0, zraw = 0;                          // raw: touching the value would stall on the load
      uint32_t s2 = 0u, z2 = 0u;
      bool fresh = true;                                    // sraw / zraw belong to a group not yet expanded into s2 / z2
      // ---------------- MXFP4 metadata state: one e8m0 byte per stage, row stride N, requested one stage ahead ----------------
      const uint8_t* mxp = nullptr;
      uint8_t mxraw = 0;
      float w8_scale = 1.f;                                 // 8-bit weight-only: the feature's fp32 scale (pre-scaling only)
      if (W8) {
        if (P.mode == 2) w8_scale = reinterpret_cast<const float*>(P.scales)[n];
      } else if (WFMT == WFMT_INT4) {
        g = (kstage0 * T2_BK + kq * 32) / P.group_size; rem = (kstage0 * T2_BK + kq * 32) % P.group_size;
        sp = reinterpret_cast<const uint16_t*>(P.scales) + (size_t)g * P.N + n;
        zp = (P.mode != 2) ? reinterpret_cast<const uint16_t*>(P.zeros) + (size_t)g * P.N + n : sp;
        sraw = *sp; zraw = (P.mode != 2) ? *zp : (uint16_t)0;
      } else {
        mxp = reinterpret_cast<const uint8_t*>(P.scales) + (size_t)(kstage0 * 4 + kq) * P.N + n;
        mxraw = *mxp;
      }
      for (int kb = 0; kb < kblocks; ++kb) {
        float mx_sf = 0.f;
        uint32_t mx_s16 = 0u;
        if (W8) {
        } else if (WFMT == WFMT_INT4) {
          if (fresh) { s2 = dup(sraw); z2 = dup(zraw); fresh = false; }
          // where the next stage lands; if it is another group, request its metadata now and expand it next iteration
          rem += T2_BK;
          if (rem >= P.group_size && kb + 1 < kblocks) {
            int adv = 0;
            while (rem >= P.group_size) { rem -= P.group_size; ++adv; }
            sp += (size_t)adv * P.N; zp += (size_t)adv * P.N;
            sraw = *sp;
            if (P.mode != 2) zraw = *zp;
            fresh = true;
          }
        } else {
          const uint32_t e = mxraw;
          mx_sf = e8m0_to_float(e);
          const __half2 h = __float2half2_rn(mx_sf);
          mx_s16 = *reinterpret_cast<const uint32_t*>(&h);
          if (kb + 1 < kblocks) { mxp += (size_t)4 * P.N; mxraw = *mxp; }
        }
        // 1. packed words (landed long ago in steady state) -> dequantised pairs in registers; release the ring slot
        ptx::mbar_wait(&pfull[ps], pph);
        if (W8) {
          // row r of the tile is 128 bytes (k .. k+127) in 16-byte chunks, chunk c stored at c ^ (r & 7); this thread's
          // 32 k are chunks 2 kq and 2 kq + 1
          const unsigned char* rb = smem + T2_P_OFF + ps * PBYTES + row * 128;
          const uint4 w0 = *reinterpret_cast<const uint4*>(rb + (((2 * kq) ^ (row & 7)) << 4));
          const uint4 w1 = *reinterpret_cast<const uint4*>(rb + (((2 * kq + 1) ^ (row & 7)) << 4));
          const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) widen4_w8<BF16, WFMT>(w[cc], w8_scale, P.mode == 2, o[2 * cc], o[2 * cc + 1]);
        } else if (WFMT == WFMT_INT4) {
          const uint32_t* pk = pk0 + ps * (uint32_t)(PBYTES / 4);
          constexpr int PPW = 16 / NBITS;                       // pairs per word
          uint32_t w[NBITS];
#pragma unroll
          for (int cc = 0; cc < NBITS; ++cc) w[cc] = pk[cc * T2_BM];
#pragma unroll
          for (int cc = 0; cc < NBITS; ++cc) {
            uint32_t q[PPW];
            if (NBITS == 4) unpack8_nat<BF16>(w[cc], reinterpret_cast<uint32_t (&)[4]>(q));
            else unpack_word_nat<BF16, NBITS>(w[cc], q);
#pragma unroll
            for (int jj = 0; jj < PPW; ++jj) o[cc * PPW + jj] = dequant2_ref<BF16>(q[jj], s2, z2, P.mode);
          }
        } else {
          // row r of the tile is 64 bytes (128 codes); this thread's 32 codes are bytes kq*16 .. kq*16+15
          const uint4 wv = *reinterpret_cast<const uint4*>(smem + T2_P_OFF + ps * T2_P_BYTES + row * 64 + kq * 16);
          const uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              o[cc * 4 + jj] = mx_scale_pair<BF16>(e2m1x2_to_f16x2(w[cc] >> (8 * jj)), mx_sf, mx_s16);
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&pfree[ps]);          // every lane's LDS results have been consumed above
        if (++ps == PSTAGES) { ps = 0; pph ^= 1u; }
        // 2. A slot s is free once the MMAs of its previous use completed; store, publish
        ptx::mbar_wait(&empty[s], ph ^ 1u);
        ptx::tc_fence_after_sync();
        ptx::tmem_st_32x16(taddr0 + s * T2_ACOLS, o);
        ptx::tmem_st_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ready_leader + s * 8u);
        if (++s == T2_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp >= T2_EPI_WARP0 && warp < T2_EPI_WARP0 + T2_EPI_WARPS) {
    // ================= epilogue: TMEM -> registers -> global (each CTA drains its own 128 feature rows) =================
    const int q4 = warp & 3;                            // TMEM lane quarter == warp % 4
    const int chalf = (warp - T2_EPI_WARP0) >> 2;       // which half of the accumulator columns (two warps per lane quarter)
    const int row = q4 * 32 + lane;
    const uint32_t tempty_leader = ptx::mapa(ptx::smem_u32(tempty), 0u);
    const int ncols = P.n_mma;
    int it = 0;
    for (int t = pair; t < ntiles; t += npairs, ++it) {
      const int tile = t / P.splits, split = t % P.splits;
      int tm, tn;
      tile_coords(tile, P.tiles_m, P.tiles_n, tm, tn);
      const int n = tn * 2 * T2_BM + (int)rank * T2_BM + row;
      float bv = 0.f;
      const float es = (W8 && P.post_scale) ? reinterpret_cast<const float*>(P.scales)[n] : 1.f;   // channel_scale_mode 1
      if (P.bias) bv = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.bias)[n])
                            : __half2float(reinterpret_cast<const __half*>(P.bias)[n]);
      ptx::mbar_wait_sleep(tfull, (uint32_t)(it & 1));     // a whole main loop away: do not burn issue slots spinning
      ptx::tc_fence_after_sync();
      const int mrem = P.M - tm * T2_BT;                                  // valid tokens in this tile
      unsigned char* ybase = reinterpret_cast<unsigned char*>(P.y) + ((int64_t)(tm * T2_BT) * P.N + n) * 2;
      const int64_t ystride = (int64_t)P.N * 2;
      auto store_out = [&](int c, float v) {                              // one output element (token c of the tile, this row)
        if (W8) v *= es;
        if (BF16) {
          if (P.bias) v = __bfloat162float(__float2bfloat16_rn(v)) + bv;
          *reinterpret_cast<__nv_bfloat16*>(ybase + c * ystride) = __float2bfloat16_rn(v);
        } else {
          if (P.bias) v = __half2float(__float2half_rn(v)) + bv;
          *reinterpret_cast<__half*>(ybase + c * ystride) = __float2half_rn(v);
        }
      };
      if (P.splits == 1) {
        // the accumulator is single-buffered (tensor memory is full), so this drain is exposed: keep it lean --
        // one cvt + one address mad + one 2-byte store per element, no per-element predicates on whole tiles
#pragma unroll 1
        for (int c0 = chalf * (ncols >> 1); c0 < (chalf + 1) * (ncols >> 1); c0 += 32) {
          if (c0 >= mrem) break;
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)c0, r);
          ptx::tmem_ld_wait();
          const bool whole = (c0 + 32 <= mrem);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (whole || c0 + j < mrem) store_out(c0 + j, __uint_as_float(r[j]));
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader);
      } else if (chalf != 0) {
        // split-K slices are handled by the first warp of each lane quarter (one counter per slice); nothing to drain here
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader);
      } else {
        // split-K: park this slice's fp32 partial (row-major per feature row: 128 B per thread and chunk), release the
        // accumulator, then count in; the LAST of the `splits` warps serving this (tile, rank, quarter) slice reduces it
        float* mine = P.ws_partial + ((((size_t)tile * P.splits + split) * 2 + rank) * T2_BM + row) * (size_t)ncols;
#pragma unroll 1
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          if (c0 >= mrem) break;
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)c0, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(mine + c0 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader);
        __threadfence();                                                  // partial visible device-wide before the count
        __syncwarp();
        int* ctr = P.ws_counter + (tile * 2 + (int)rank) * 4 + q4;
        int old = 0;
        if (lane == 0) old = atomicAdd(ctr, 1);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old == P.splits - 1) {
          __threadfence();
          const float* base = P.ws_partial + (((size_t)tile * P.splits * 2 + rank) * T2_BM + row) * (size_t)ncols;
          const size_t sstride = (size_t)2 * T2_BM * ncols;              // between consecutive splits
#pragma unroll 1
          for (int c0 = 0; c0 < ncols; c0 += 4) {
            if (c0 >= mrem) break;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int sp = 0; sp < P.splits; ++sp) {                       // fixed order -> deterministic
              const float4 v = __ldcg(reinterpret_cast<const float4*>(base + sp * sstride + c0));
              acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            if (c0 + 0 < mrem) store_out(c0 + 0, acc.x);
            if (c0 + 1 < mrem) store_out(c0 + 1, acc.y);
            if (c0 + 2 < mrem) store_out(c0 + 2, acc.z);
            if (c0 + 3 < mrem) store_out(c0 + 3, acc.w);
          }
          if (lane == 0) *ctr = 0;                                        // self-cleaning: ready for the next launch
        }
      }
    }
  }

  // ---- teardown: nobody leaves (or frees tensor memory) while the pair may still touch this CTA ----
  ptx::tc_fence_before_sync();
  ptx::cluster_sync();
  ptx::tc_fence_after_sync();
  if (warp == T2_ALLOC_WARP) ptx::tmem_dealloc_2cta(tmem_base, 512);
}

// How a problem is cut into work items (shared by the launch and by glb_workspace_bytes).
constexpr size_t T2_COUNTER_REGION = 64 * 1024;      // 16384 counters = 2048 tiles (8 per tile)
struct Tc2Plan {
  int tiles_m, tiles_n, splits, kb_per_split, n_mma;
  size_t counter_bytes, partial_bytes;
};
Tc2Plan make_plan(const glb_linear& L, int64_t M) {
  Tc2Plan p;
  p.tiles_m = (int)((M + T2_BT - 1) / T2_BT);
  p.tiles_n = (int)(L.N / (2 * T2_BM));
  const int tiles = p.tiles_m * p.tiles_n, kblocks = (int)(L.K / T2_BK), pairs = num_sms() / 2;
  p.n_mma = (M > 128) ? 256 : ((M > 64) ? 128 : 64);      // tokens the MMA actually computes (TMA zero-fills beyond M)
  p.splits = 1;
  // split K only for M <= 128: measured r1 (profiles/configs_r1.txt) 45 -> 22 us at M=32, 51 -> 39 us at M=128 on a 4096^2
  // layer, but 39 -> 61 us at M=256, where parking and re-reading 256-column fp32 partials costs more than the idle SMs
  if (M <= 128 && (size_t)tiles * 8 * sizeof(int) <= T2_COUNTER_REGION)
    for (int s = std::min(16, pairs / std::max(tiles, 1)); s >= 2; --s)
      if (kblocks % s == 0 && kblocks / s >= T2_STAGES) { p.splits = s; break; }   // fill the SM pairs, keep >= one ring of stages
  p.kb_per_split = kblocks / p.splits;
  // The arrival counters sit in a FIXED region at the start of the workspace and the partial tiles behind it: one
  // workspace serves every layer on a stream, and a layout that depended on the layer's tile count would let one
  // layer's partial tiles land on the words another layer reads as (zero) counters.
  p.counter_bytes = (p.splits > 1) ? T2_COUNTER_REGION : 0;
  p.partial_bytes = (p.splits > 1) ? (size_t)tiles * p.splits * 2 * T2_BM * p.n_mma * sizeof(float) : 0;
  return p;
}

template <bool BF16, int WFMT, int NBITS = 4>
int launch_tc2(const glb_linear& L, const void* x, void* y, int64_t M, void* workspace, size_t ws_bytes, const void* bias,
               cudaStream_t stream) {
  CUtensorMap tx, tw;
  int rc = make_tmap_2d(&tx, BF16 ? GLB_BF16 : GLB_FP16, x, (uint64_t)L.K, (uint64_t)M, (uint64_t)L.K * 2, 64, T2_BT / 2, 1);
  if (rc != GLB_OK) return rc;
  if (WFMT == WFMT_INT4)
    rc = make_tmap_2d(&tw, GLB_INT32, L.W_q, (uint64_t)L.N, (uint64_t)(L.K * NBITS / 32), (uint64_t)L.N * 4, T2_BM, 4 * NBITS, 0);
  else if (WFMT == WFMT_W8I || WFMT == WFMT_W8F)   // (N, K) bytes, K contiguous: box = 128 bytes x 128 feature rows, 128B swizzle
    rc = make_tmap_2d(&tw, GLB_UINT8, L.W_q, (uint64_t)L.K, (uint64_t)L.N, (uint64_t)L.K, T2_BK, T2_BM, 1);
  else   // (N, K/2) bytes, K contiguous: box = 64 bytes of codes x 128 feature rows
    rc = make_tmap_2d(&tw, GLB_UINT8, L.W_q, (uint64_t)(L.K / 2), (uint64_t)L.N, (uint64_t)(L.K / 2), T2_BK / 2, T2_BM, 0);
  if (rc != GLB_OK) return rc;
  const Tc2Plan plan = make_plan(L, M);
  Tc2Params P;
  P.scales = L.scales; P.zeros = (L.W_group_mode == 2) ? nullptr : L.zeros;
  P.y = y; P.bias = bias; P.M = (int)M; P.N = (int)L.N; P.K = (int)L.K;
  P.group_size = L.group_size; P.mode = L.W_group_mode;
  P.post_scale = (L.channel_scale_mode == 1) ? 1 : 0;
  P.tiles_m = plan.tiles_m; P.tiles_n = plan.tiles_n;
  P.splits = plan.splits; P.kb_per_split = plan.kb_per_split; P.n_mma = plan.n_mma;
  P.ws_counter = nullptr; P.ws_partial = nullptr;
  if (plan.splits > 1) {
    GLB_REQUIRE(workspace != nullptr && ws_bytes >= plan.counter_bytes + plan.partial_bytes, GLB_ERR_WORKSPACE,
                "split-K needs %zu bytes of (zero-initialised) workspace, got %zu", plan.counter_bytes + plan.partial_bytes, ws_bytes);
    GLB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, GLB_ERR_INVALID, "workspace must be 16-byte aligned");
    P.ws_counter = reinterpret_cast<int*>(workspace);
    P.ws_partial = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + plan.counter_bytes);
  }
  {
    static int pf = -1;
    if (pf < 0) { const char* e = getenv("GLB_PREFILL_PREFETCH"); pf = e ? atoi(e) : 0; }
    P.prefetch = pf;
  }
  auto kern = gemm_a16w4_tc2_kernel<BF16, WFMT, NBITS>;
  GLB_ENSURE_SMEM(kern, T2_SMEM);
  const int pairs = std::min(num_sms() / 2, P.tiles_m * P.tiles_n * P.splits);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(T2_THREADS);
  cfg.dynamicSmemBytes = T2_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tx, tw, P));
  count_launch();
  return GLB_OK;
}

}  // namespace

// shapes the pair kernel takes on top of gemm_tc_supported(): whole 256-feature tiles
bool gemm_tc2_supported(const glb_linear& L, int64_t M) {
  if (M < 1) return false;
  if (L.N % (2 * T2_BM) != 0 || L.K % T2_BK != 0) return false;
  if (L.group_size % 32 != 0) return false;          // a dequant warp's 32-wide k span never straddles quantisation groups
  if (L.K % L.group_size != 0) return false;
  return true;
}

// MXFP4 weight-only layers (A16W4_MXFP: input_dtype MXFP16 / MXBF16, e2m1 codes two per byte, e8m0 scales per 32)
bool gemm_mx_tc2_supported(const glb_linear& L, int64_t M) {
  if (M < 1) return false;
  if (!(L.input_dtype == GLB_MXFP16 || L.input_dtype == GLB_MXBF16)) return false;
  if (L.W_nbits != 4 || L.elements_per_sample != 2 || L.wq_dtype != GLB_UINT8 || L.group_size != 32) return false;
  if (L.W_group_mode != 2 || L.channel_scale_mode != 0) return false;
  if (L.wq_stride0 != 1 || L.wq_stride1 != L.K / 2) return false;                 // (K/2, N) view of (N, K/2)
  if (!(L.scales_dtype == GLB_UINT8 || L.scales_dtype == GLB_E8M0)) return false;
  if (L.scales_numel != (L.K / 32) * L.N || L.s_stride0 != 1 || L.s_stride1 != L.N) return false;   // (N, K/32) view of (K/32, N)
  if (L.output_dtype != L.input_dtype) return false;
  if (L.N % (2 * T2_BM) != 0 || L.K % T2_BK != 0) return false;
  if (reinterpret_cast<uintptr_t>(L.W_q) & 15) return false;
  return true;
}

// 8-bit weight-only layers (A16W8 int8 / e4m3 with fp16 / bf16 activations; helper.py:88-170)
bool gemm_w8_tc2_supported(const glb_linear& L, int64_t M) {
  if (M < 1) return false;
  if (!(L.input_dtype == GLB_FP16 || L.input_dtype == GLB_BF16) || L.output_dtype != L.input_dtype) return false;
  if (L.W_nbits != 8 || L.elements_per_sample != 1) return false;
  if (!(L.wq_dtype == GLB_INT8 || L.wq_dtype == GLB_FP8E4)) return false;
  if (L.wq_stride0 != 1 || L.wq_stride1 != L.K) return false;                     // (K, N) view of (N, K)
  if (!((L.W_group_mode == 2 && L.channel_scale_mode == 0) || (L.W_group_mode == 0 && L.channel_scale_mode == 1))) return false;
  if (L.scales_dtype != GLB_FP32 || L.scales_numel != L.N || L.group_size != L.K) return false;
  if (L.N % (2 * T2_BM) != 0 || L.K % T2_BK != 0) return false;
  if (reinterpret_cast<uintptr_t>(L.W_q) & 15) return false;
  return true;
}

// bytes of zero-initialised workspace launch_gemm_tc2 needs for (L, M): per-slice counters + fp32 partials (0 without split-K)
size_t gemm_tc2_workspace(const glb_linear& L, int64_t M) {
  const Tc2Plan p = make_plan(L, M);
  return p.counter_bytes + p.partial_bytes;
}

int launch_gemm_tc2(const glb_linear& L, const void* x, void* y, int64_t M, void* workspace, size_t ws_bytes, const void* bias,
                    cudaStream_t stream) {
  if (gemm_mx_tc2_supported(L, M)) {
    set_kernel_name("gemm_a16w4_mxfp4_tcgen05_2cta");
    return (L.input_dtype == GLB_MXBF16) ? launch_tc2<true, WFMT_MXFP4>(L, x, y, M, workspace, ws_bytes, bias, stream)
                                         : launch_tc2<false, WFMT_MXFP4>(L, x, y, M, workspace, ws_bytes, bias, stream);
  }
  const bool bf = (L.input_dtype == GLB_BF16);
  if (gemm_w8_tc2_supported(L, M)) {
    const bool f8 = (L.wq_dtype == GLB_FP8E4);
    set_kernel_name(f8 ? "gemm_a16w8_fp8_tcgen05_2cta" : "gemm_a16w8_int8_tcgen05_2cta");
    if (f8) return bf ? launch_tc2<true, WFMT_W8F>(L, x, y, M, workspace, ws_bytes, bias, stream)
                      : launch_tc2<false, WFMT_W8F>(L, x, y, M, workspace, ws_bytes, bias, stream);
    return bf ? launch_tc2<true, WFMT_W8I>(L, x, y, M, workspace, ws_bytes, bias, stream)
              : launch_tc2<false, WFMT_W8I>(L, x, y, M, workspace, ws_bytes, bias, stream);
  }
#define GLB_TC2_INT(NB) (bf ? launch_tc2<true, WFMT_INT4, NB>(L, x, y, M, workspace, ws_bytes, bias, stream) \
                           : launch_tc2<false, WFMT_INT4, NB>(L, x, y, M, workspace, ws_bytes, bias, stream))
  switch (L.W_nbits) {
    case 8: set_kernel_name("gemm_a16w8_packed_tcgen05_2cta"); return GLB_TC2_INT(8);
    case 2: set_kernel_name("gemm_a16w2_tcgen05_2cta"); return GLB_TC2_INT(2);
    case 1: set_kernel_name("gemm_a16w1_tcgen05_2cta"); return GLB_TC2_INT(1);
    default: break;
  }
  set_kernel_name("gemm_a16w4_tcgen05_2cta");
  return GLB_TC2_INT(4);
#undef GLB_TC2_INT
}

}  // namespace glb
