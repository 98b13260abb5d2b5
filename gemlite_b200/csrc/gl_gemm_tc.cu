// tcgen05 tensor-core GEMM with fused unpack + dequant (A16W4 packed int32) -- the prefill path.
//
// Replaces gemlite/triton_kernels/gemm_kernels.py (gemm_INT_kernel, :248-413) and, for 16 < M <= 64,
// gemm_splitK_kernels.py.  Arithmetic contract: the dequantised weight tile is bit-identical to the
// reference's dequantize() (utils.py:57-89: fp16/bf16 fma / sub-mul in the scale dtype), products are
// accumulated in fp32 by the tensor core (tl.dot(..., out_dtype=fp32), gemm_kernels.py:384).
//
// Mapping (B200-first, not the reference's):
//   * the WEIGHTS are the MMA "A" operand: a CTA tile is 128 output features x BT tokens; the
//     accumulator D[feature, token] lives in TMEM (128 lanes x BT fp32 columns, double buffered).
//     Per FLOP this needs half the dequant work of the transposed mapping and lets BT shrink to 32/64
//     for small batches without wasting the dequant/HBM pipeline.
//   * warp roles (512 threads): w0 TMA producer, w1 MMA issuer (one elected lane), w2 TMEM allocator,
//     w4-11 dequant (two threads per feature row), w12-15 epilogue (TMEM -> registers -> global).
//   * per 64-wide k-block and stage s: TMA brings the token tile x[BT,64] (128B-swizzled, K-major)
//     and the packed weight tile Wq[8 words,128 features] (4 KB) into shared memory; dequant warps
//     turn the 4 KB into a 16 KB fp16/bf16 K-major SW128 tile (conflict-free 128-bit stores);
//     the MMA lane issues 4 x tcgen05.mma (K=16) and commits to the stage's "empty" barrier.
//   * persistent: grid = #SMs, tiles are walked token-block-fastest so CTAs running together share
//     the same weight tile in L2.
#include <stdlib.h>
#include <mutex>
#include <unordered_map>

#include "gl_common.cuh"
#include "gl_ptx.cuh"

namespace glb {
namespace {

constexpr int TC_BM = 128;            // features per tile (UMMA M)
constexpr int TC_BK = 64;             // k per stage (128 bytes of fp16 -> one swizzle atom row)
constexpr int TC_THREADS = 512;
constexpr int TC_DEQ_WARP0 = 4;       // warps 4..11 (two threads per feature row, 4 words each)
constexpr int TC_DEQ_WARPS = 8;
constexpr int TC_EPI_WARP0 = 12;      // warps 12..15 (warp % 4 == TMEM lane quarter)

struct TcParams {
  const void* scales;     // (G, N)
  const void* zeros;      // (G, N) or null
  void* y;                // (M, N)
  const void* bias;       // (N) or null
  int M, N, K;
  int group_size;
  int mode;               // 2, 3, 4
  int tiles_m;            // token blocks
  int tiles_n;            // feature blocks
};

template <int BT> struct TcSmem {
  static constexpr int X_BYTES = BT * TC_BK * 2;          // token tile
  static constexpr int P_BYTES = (TC_BK / 8) * TC_BM * 4;  // packed words
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;        // dequantised weights
  static constexpr int STAGE = X_BYTES + P_BYTES + A_BYTES;
  static constexpr int STAGES = (BT >= 256) ? 4 : ((BT >= 128) ? 5 : 6);
  static constexpr int BAR_BYTES = 1024;
  static constexpr int TOTAL = STAGES * STAGE + BAR_BYTES + 1024 /* alignment slack */;
};

// 8 nibbles -> 8 exact small integers in memory order k0..k7, as 4 packed 16-bit pairs
template <bool BF16> __device__ __forceinline__ void unpack8(uint32_t q, uint32_t (&o)[4]) {
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

// reference dequantize() on a pair, in the scale dtype (one rounding per op, fma fused for mode 4)
template <bool BF16> __device__ __forceinline__ uint32_t dequant2(uint32_t q2, uint32_t s2, uint32_t z2, int mode) {
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

template <bool BF16, int BT>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_a16w4_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const TcParams P) {
  using S = TcSmem<BT>;
  // 1024-byte alignment for the 128B swizzle atoms.  (No pointer round-trip through an integer here: that
  // would lose the shared address space and turn every LDS/STS below into generic LD/ST.)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0u) __trap();
  unsigned char* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE);
  uint64_t* full = bars;                       // [STAGES]  TMA landed (x + packed)
  uint64_t* aready = bars + S::STAGES;         // [STAGES]  dequant wrote the A tile
  uint64_t* empty = bars + 2 * S::STAGES;      // [STAGES]  MMAs reading the stage completed
  uint64_t* tfull = bars + 3 * S::STAGES;      // [2]       accumulator buffer complete
  uint64_t* tempty = bars + 3 * S::STAGES + 2; // [2]       accumulator buffer drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S::STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = P.K / TC_BK;
  const int ntiles = P.tiles_m * P.tiles_n;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_w);
    for (int s = 0; s < S::STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&aready[s], TC_DEQ_WARPS);     // one arrival per dequant warp
      ptx::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tfull[b], 1);
      ptx::mbar_init(&tempty[b], 4);     // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc(tmem_slot, 2 * BT);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tm = t % P.tiles_m, tn = t / P.tiles_m;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&empty[s], ph ^ 1);
          unsigned char* st = stage_base + s * S::STAGE;
          ptx::mbar_arrive_expect_tx(&full[s], S::X_BYTES + S::P_BYTES);
          ptx::tma_load_2d(st, &tmap_x, &full[s], kb * TC_BK, tm * BT);
          ptx::tma_load_2d(st + S::X_BYTES, &tmap_w, &full[s], tn * TC_BM, kb * (TC_BK / 8));
          if (++s == S::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = ptx::make_idesc(1u, BF16 ? 1u : 0u, BF16 ? 1u : 0u, TC_BM, BT);
    int s = 0; uint32_t ph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t bph = (it >> 1) & 1;
      ptx::mbar_wait(&tempty[buf], bph ^ 1);
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * BT;
      for (int kb = 0; kb < kblocks; ++kb) {
        ptx::mbar_wait(&full[s], ph);
        ptx::mbar_wait(&aready[s], ph);
        ptx::tc_fence_after_sync();
        if (ptx::elect_one()) {
          unsigned char* st = stage_base + s * S::STAGE;
          const uint64_t bdesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(st));
          const uint64_t adesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(st + S::X_BYTES + S::P_BYTES));
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)   // +32 bytes per K=16 step inside the swizzle atom
            ptx::mma_f16_ss(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          ptx::mma_commit(&empty[s]);
          if (kb == kblocks - 1) ptx::mma_commit(&tfull[buf]);
        }
        __syncwarp();
        if (++s == S::STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= TC_DEQ_WARP0 && warp < TC_DEQ_WARP0 + TC_DEQ_WARPS) {
    // ================= dequant: two threads per feature row, 4 packed words (32 k) each =================
    const int dt = threadIdx.x - TC_DEQ_WARP0 * 32;      // 0..255
    const int row = dt & (TC_BM - 1);                    // feature row inside the tile
    const int half = dt >> 7;                            // which 32-k half of the 64-k block
    const uint16_t* sc_g = reinterpret_cast<const uint16_t*>(P.scales);
    const uint16_t* zr_g = reinterpret_cast<const uint16_t*>(P.zeros);
    int s = 0; uint32_t ph = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int tn = t / P.tiles_m;
      const int n = tn * TC_BM + row;
      // group walk without divisions: g / rem track k0 = kb*64 + half*32 as it advances by 64 per block
      int g = (half * 32) / P.group_size, rem = (half * 32) % P.group_size;
      int cur_g = -1, pre_g = -1;
      uint32_t s2 = 0, z2 = 0;
      uint16_t psv = 0, pzv = 0;                           // raw prefetched scale / zero (consumed one block later)
      auto dup = [](uint16_t v) { return (uint32_t)v | ((uint32_t)v << 16); };
      for (int kb = 0; kb < kblocks; ++kb) {
        const int k0 = kb * TC_BK + half * 32;
        if (g != cur_g) {
          if (g == pre_g) { s2 = dup(psv); z2 = dup(pzv); }
          else {
            const int64_t mi = (int64_t)g * P.N + n;
            s2 = dup(sc_g[mi]);
            z2 = (P.mode != 2) ? dup(zr_g[mi]) : 0u;
          }
          cur_g = g;
        }
        // where the next block lands; issue its metadata loads now, touch the values only next iteration
        int gn = g, remn = rem + TC_BK;
        while (remn >= P.group_size) { remn -= P.group_size; ++gn; }
        if (kb + 1 < kblocks && gn != cur_g && gn != pre_g) {
          const int64_t mi = (int64_t)gn * P.N + n;
          psv = sc_g[mi];
          pzv = (P.mode != 2) ? zr_g[mi] : (uint16_t)0;
          pre_g = gn;
        }
        ptx::mbar_wait(&full[s], ph);
        unsigned char* st = stage_base + s * S::STAGE;
        const uint32_t* pk = reinterpret_cast<const uint32_t*>(st + S::X_BYTES);
        unsigned char* arow = st + S::X_BYTES + S::P_BYTES + (row >> 3) * 1024 + (row & 7) * 128;
        uint32_t w[4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) w[cc] = pk[(half * 4 + cc) * TC_BM + row];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = half * 4 + cc;
          uint32_t ls2 = s2, lz2 = z2;
          if (P.group_size < 32) {                         // a 32-k span may cross groups only for gs = 8 / 16 / 24
            const int gg = (k0 + cc * 8) / P.group_size;
            if (gg != cur_g) {
              const int64_t mi = (int64_t)gg * P.N + n;
              ls2 = dup(sc_g[mi]);
              lz2 = (P.mode != 2) ? dup(zr_g[mi]) : 0u;
            }
          }
          uint32_t q[4];
          unpack8<BF16>(w[cc], q);
          uint4 o;
          o.x = dequant2<BF16>(q[0], ls2, lz2, P.mode);
          o.y = dequant2<BF16>(q[1], ls2, lz2, P.mode);
          o.z = dequant2<BF16>(q[2], ls2, lz2, P.mode);
          o.w = dequant2<BF16>(q[3], ls2, lz2, P.mode);
          *reinterpret_cast<uint4*>(arow + ((c ^ (row & 7)) << 4)) = o;      // 128B swizzle: chunk ^= row % 8
        }
        ptx::fence_proxy_async_smem();    // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&aready[s]);
        if (++s == S::STAGES) { s = 0; ph ^= 1; }
        g = gn; rem = remn;
      }
    }
  } else if (warp >= TC_EPI_WARP0) {
    // ================= epilogue: TMEM -> registers -> global =================
    const int q4 = warp - TC_EPI_WARP0;                 // TMEM lane quarter (== warp % 4)
    const int row = q4 * 32 + lane;                     // feature row inside the tile
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int tm = t % P.tiles_m, tn = t / P.tiles_m;
      const int buf = it & 1;
      const uint32_t bph = (it >> 1) & 1;
      const int n = tn * TC_BM + row;
      float bv = 0.f;
      if (P.bias) bv = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.bias)[n])
                            : __half2float(reinterpret_cast<const __half*>(P.bias)[n]);
      ptx::mbar_wait(&tfull[buf], bph);
      ptx::tc_fence_after_sync();
#pragma unroll 1
      for (int c0 = 0; c0 < BT; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * BT + c0), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int m = tm * BT + c0 + j;
          if (m < P.M) {
            float v = __uint_as_float(r[j]);
            if (BF16) {
              if (P.bias) v = __bfloat162float(__float2bfloat16_rn(v)) + bv;
              reinterpret_cast<__nv_bfloat16*>(P.y)[(int64_t)m * P.N + n] = __float2bfloat16_rn(v);
            } else {
              if (P.bias) v = __half2float(__float2half_rn(v)) + bv;
              reinterpret_cast<__half*>(P.y)[(int64_t)m * P.N + n] = __float2half_rn(v);
            }
          }
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[buf]);
    }
  }

  // ---- teardown ----
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 2 * BT);
}


// =================================================================================================
// A8W8: 8-bit activations x 8-bit weights, no dequant (W_group_mode 0) -- rows a9/a10 for config C4.
// Both operands are K-major byte matrices and go straight from TMA (128B swizzle) to the tensor core:
//   int8 x int8 -> int32 (kind::i8, exact)   |   fp8 x fp8 -> fp32 (kind::f8f6f4)
// Epilogue = the reference's channel-scale modes (gemm_kernels.py:392-406) with fp32 metadata:
//   out = float(acc) * (scales_x[m] * scales_w[n])   (mode 3; modes 0/1/2 analogous), then bias.
// =================================================================================================
constexpr int T8_BK = 128;            // bytes == elements per stage row
constexpr int T8_THREADS = 256;       // w0 TMA, w1 MMA, w2 TMEM alloc, w4-7 epilogue

struct T8Params {
  const float* scales_w;   // (N) or null
  const float* scales_x;   // (M) or null
  void* y;
  const void* bias;
  int M, N, K;
  int csm;                 // channel_scale_mode 0..3
  int out_dtype;           // GLB_FP16 / BF16 / FP32 / INT32
  int tiles_m, tiles_n;
  // split-K (SPLIT instantiation only): K is cut into `splits` ranges of kb_per_split stages, one work item each;
  // raw accumulators are parked in ws_partial and summed in split order by the last warp to arrive (ws_counter)
  int splits, kb_per_split;
  int* ws_counter;         // [tile][lane quarter], zero on entry and on exit
  uint32_t* ws_partial;    // [tile][split][128 feature rows][BT] raw int32 / fp32 bits
};

template <int BT> struct T8Smem {
  static constexpr int X_BYTES = BT * T8_BK;
  static constexpr int W_BYTES = TC_BM * T8_BK;
  static constexpr int STAGE = X_BYTES + W_BYTES;
  static constexpr int STAGES = (BT >= 256) ? 4 : ((BT >= 128) ? 6 : 8);
  static constexpr int TOTAL = STAGES * STAGE + 1024 + 1024;
};

// SPLIT = true: split-K flavour for 16 < M <= 128 (not instantiated: validated in round 2, slower than unsplit; written without
// GPU time left to validate it).  The 128-feature tile gives only N / 128 CTAs (32 at N = 4096) for one token block, so a
// small batch streams its weights through a fraction of the SMs; with split-K every (tile, K range) is a work item.
// SPLIT = false compiles to exactly the kernel measured in round 1 (every difference is behind `if constexpr` / SPLIT ?:).
template <bool IS_INT, int BT, bool SPLIT = false>
__global__ void __launch_bounds__(T8_THREADS, 1)
gemm_a8w8_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const T8Params P,
                    const uint32_t idesc) {
  using S = T8Smem<BT>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + S::STAGES;
  uint64_t* tfull = bars + 2 * S::STAGES;
  uint64_t* tempty = bars + 2 * S::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = SPLIT ? P.kb_per_split : P.K / T8_BK;
  const int ntiles = P.tiles_m * P.tiles_n * (SPLIT ? P.splits : 1);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_w);
    for (int s = 0; s < S::STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tfull[b], 1); ptx::mbar_init(&tempty[b], 4); }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc(tmem_slot, 2 * BT);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tt = SPLIT ? t / P.splits : t;
        const int kb0 = SPLIT ? (t % P.splits) * P.kb_per_split : 0;
        const int tm = tt % P.tiles_m, tn = tt / P.tiles_m;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&empty[s], ph ^ 1);
          unsigned char* st = smem + s * S::STAGE;
          ptx::mbar_arrive_expect_tx(&full[s], S::X_BYTES + S::W_BYTES);
          ptx::tma_load_2d(st, &tmap_x, &full[s], (kb0 + kb) * T8_BK, tm * BT);
          ptx::tma_load_2d(st + S::X_BYTES, &tmap_w, &full[s], (kb0 + kb) * T8_BK, tn * TC_BM);
          if (++s == S::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    int s = 0; uint32_t ph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t bph = (it >> 1) & 1;
      ptx::mbar_wait(&tempty[buf], bph ^ 1);
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * BT;
      for (int kb = 0; kb < kblocks; ++kb) {
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after_sync();
        if (ptx::elect_one()) {
          unsigned char* st = smem + s * S::STAGE;
          const uint64_t bdesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(st));
          const uint64_t adesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(st + S::X_BYTES));
#pragma unroll
          for (int k = 0; k < T8_BK / 32; ++k) {     // K = 32 bytes per instruction
            if (IS_INT) ptx::mma_i8_ss(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            else ptx::mma_f8_ss(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          ptx::mma_commit(&empty[s]);
          if (kb == kblocks - 1) ptx::mma_commit(&tfull[buf]);
        }
        __syncwarp();
        if (++s == S::STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q4 = warp - 4;
    const int row = q4 * 32 + lane;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int tt = SPLIT ? t / P.splits : t;
      const int split = SPLIT ? t % P.splits : 0;
      const int tm = tt % P.tiles_m, tn = tt / P.tiles_m;
      const int buf = it & 1;
      const uint32_t bph = (it >> 1) & 1;
      const int n = tn * TC_BM + row;
      const float sb = (P.csm == 1 || P.csm == 3) ? P.scales_w[n] : 1.f;
      float bv = 0.f;
      if (P.bias) {
        if (P.out_dtype == GLB_FP16) bv = __half2float(reinterpret_cast<const __half*>(P.bias)[n]);
        else if (P.out_dtype == GLB_BF16) bv = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.bias)[n]);
        else bv = reinterpret_cast<const float*>(P.bias)[n];
      }
      ptx::mbar_wait(&tfull[buf], bph);
      ptx::tc_fence_after_sync();
      if constexpr (!SPLIT) {
#pragma unroll 1
        for (int c0 = 0; c0 < BT; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * BT + c0), r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int m = tm * BT + c0 + j;
            if (m < P.M) {
              const int64_t oi = (int64_t)m * P.N + n;
              if (IS_INT && P.out_dtype == GLB_INT32 && P.csm == 0 && !P.bias) {
                reinterpret_cast<int32_t*>(P.y)[oi] = (int32_t)r[j];
                continue;
              }
              float v = IS_INT ? (float)(int32_t)r[j] : __uint_as_float(r[j]);     // cvt.rn.f32.s32
              if (P.csm == 1) v = __fmul_rn(v, sb);
              else if (P.csm == 2) v = __fmul_rn(v, P.scales_x[m]);
              else if (P.csm == 3) v = __fmul_rn(v, __fmul_rn(P.scales_x[m], sb));
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
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty[buf]);
      } else {
        // one output element from its summed raw accumulator: the arithmetic of the branch above
        auto emit = [&](int m, uint32_t raw) {
          const int64_t oi = (int64_t)m * P.N + n;
          if (IS_INT && P.out_dtype == GLB_INT32 && P.csm == 0 && !P.bias) {
            reinterpret_cast<int32_t*>(P.y)[oi] = (int32_t)raw;
            return;
          }
          float v = IS_INT ? (float)(int32_t)raw : __uint_as_float(raw);
          if (P.csm == 1) v = __fmul_rn(v, sb);
          else if (P.csm == 2) v = __fmul_rn(v, P.scales_x[m]);
          else if (P.csm == 3) v = __fmul_rn(v, __fmul_rn(P.scales_x[m], sb));
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
        };
        // park this K range's raw accumulators (128 B per thread and 32-column chunk), release the accumulator buffer,
        // count in; the last of the `splits` warps serving this (tile, lane quarter) sums them in split order
        const int mrem = min(BT, P.M - tm * BT);                          // live token columns of this tile
        uint32_t* mine = P.ws_partial + (((size_t)tt * P.splits + split) * TC_BM + row) * (size_t)BT;
#pragma unroll 1
        for (int c0 = 0; c0 < BT; c0 += 32) {
          if (c0 >= mrem) break;
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * BT + c0), r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(mine + c0 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty[buf]);
        __threadfence();                                                  // partial visible device-wide before the count
        __syncwarp();
        int* ctr = P.ws_counter + tt * 4 + q4;
        int old = 0;
        if (lane == 0) old = atomicAdd(ctr, 1);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old == P.splits - 1) {
          __threadfence();
          const uint32_t* base = P.ws_partial + (((size_t)tt * P.splits) * TC_BM + row) * (size_t)BT;
          const size_t sstride = (size_t)TC_BM * BT;                      // between consecutive splits
#pragma unroll 1
          for (int c0 = 0; c0 < BT; c0 += 4) {
            if (c0 >= mrem) break;
            int32_t ai[4] = {0, 0, 0, 0};
            float af[4] = {0.f, 0.f, 0.f, 0.f};
            for (int sp = 0; sp < P.splits; ++sp) {                       // fixed order -> deterministic (int32: exact anyway)
              const uint4 v = __ldcg(reinterpret_cast<const uint4*>(base + sp * sstride + c0));
              const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (IS_INT) ai[j] += (int32_t)w[j]; else af[j] += __uint_as_float(w[j]);
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (c0 + j < mrem) emit(tm * BT + c0 + j, IS_INT ? (uint32_t)ai[j] : __float_as_uint(af[j]));
          }
          if (lane == 0) *ctr = 0;                                        // self-cleaning: ready for the next launch
        }
      }
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 2 * BT);
}

// ---------------------------------------------------------------------------------------------- host
template <bool BF16, int BT>
int launch_tc(const glb_linear& L, const void* x, void* y, int64_t M, const void* bias, cudaStream_t stream) {
  using S = TcSmem<BT>;
  CUtensorMap tx, tw;
  int rc = make_tmap_2d(&tx, BF16 ? GLB_BF16 : GLB_FP16, x, (uint64_t)L.K, (uint64_t)M, (uint64_t)L.K * 2, TC_BK, BT, 1);
  if (rc != GLB_OK) return rc;
  rc = make_tmap_2d(&tw, GLB_INT32, L.W_q, (uint64_t)L.N, (uint64_t)(L.K / 8), (uint64_t)L.N * 4, TC_BM, TC_BK / 8, 0);
  if (rc != GLB_OK) return rc;
  TcParams P;
  P.scales = L.scales; P.zeros = (L.W_group_mode == 2) ? nullptr : L.zeros;
  P.y = y; P.bias = bias; P.M = (int)M; P.N = (int)L.N; P.K = (int)L.K;
  P.group_size = L.group_size; P.mode = L.W_group_mode;
  P.tiles_m = (int)((M + BT - 1) / BT); P.tiles_n = (int)(L.N / TC_BM);
  auto kern = gemm_a16w4_tc_kernel<BF16, BT>;
  GLB_ENSURE_SMEM(kern, S::TOTAL);
  const int grid = std::min(num_sms(), P.tiles_m * P.tiles_n);
  kern<<<grid, TC_THREADS, S::TOTAL, stream>>>(tx, tw, P);
  count_launch();
  GLB_CHECK_CUDA(cudaGetLastError());
  return GLB_OK;
}


// Work decomposition of the single-CTA A8W8 kernel (shared by the launch and by glb_workspace_bytes).  Split-K is
// not used (see t8_plan) -- it would only apply to one token block (16 < M <= 128) when the feature tiles leave SMs idle.  The
// counters share the fixed region at the start of the workspace with the pair kernel's (gl_gemm_tc2.cu).
constexpr size_t T8_COUNTER_REGION = 64 * 1024;
struct T8Plan {
  int splits, kb_per_split;
  size_t counter_bytes, partial_bytes;
};
int t8_token_block(int64_t M) { return (M <= 32) ? 32 : ((M <= 64) ? 64 : ((M <= 128) ? 128 : 256)); }
T8Plan t8_plan(const glb_linear& L, int64_t M, int BT) {
  T8Plan p;
  const int kblocks = (int)(L.K / T8_BK), tiles = (int)(L.N / TC_BM) * (int)((M + BT - 1) / BT);
  p.splits = 1;
  // Split-K of this kernel was validated on hardware in round 2 (bit-exact) and measured SLOWER than the unsplit
  // launch (profiles/optin_validation_r2.txt: M = 64 37.9 vs 27.6 us, M = 128 58.4 vs 33.8 us), so it is never planned;
  // small batches go to the streaming kernel in gl_gemv_a8.cu instead.
  (void)tiles; (void)kblocks;
  p.kb_per_split = kblocks / p.splits;
  p.counter_bytes = (p.splits > 1) ? T8_COUNTER_REGION : 0;
  p.partial_bytes = (p.splits > 1) ? (size_t)tiles * p.splits * TC_BM * BT * sizeof(uint32_t) : 0;
  return p;
}

template <bool IS_INT, int BT>
int launch_t8(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M, void* workspace, size_t ws_bytes,
              const void* bias, cudaStream_t stream) {
  using S = T8Smem<BT>;
  CUtensorMap tx, tw;
  int rc = make_tmap_2d(&tx, GLB_UINT8, x, (uint64_t)L.K, (uint64_t)M, (uint64_t)L.K, T8_BK, BT, 1);
  if (rc != GLB_OK) return rc;
  // W_q is the (K, N) view with strides (1, K) of an (N, K) row-major byte matrix
  rc = make_tmap_2d(&tw, GLB_UINT8, L.W_q, (uint64_t)L.K, (uint64_t)L.N, (uint64_t)L.K, T8_BK, TC_BM, 1);
  if (rc != GLB_OK) return rc;
  T8Params P;
  P.scales_w = (L.channel_scale_mode == 1 || L.channel_scale_mode == 3) ? reinterpret_cast<const float*>(L.scales) : nullptr;
  P.scales_x = scales_x; P.y = y; P.bias = bias;
  P.M = (int)M; P.N = (int)L.N; P.K = (int)L.K; P.csm = L.channel_scale_mode; P.out_dtype = L.output_dtype;
  P.tiles_m = (int)((M + BT - 1) / BT); P.tiles_n = (int)(L.N / TC_BM);
  P.splits = 1; P.kb_per_split = (int)(L.K / T8_BK); P.ws_counter = nullptr; P.ws_partial = nullptr;
  uint32_t idesc;
  if (IS_INT) idesc = ptx::make_idesc(2u, 1u, 1u, TC_BM, BT);                         // s32 accumulate, signed 8-bit A/B
  else {
    const uint32_t wf = (L.wq_dtype == GLB_FP8E5) ? 1u : 0u, xf = (act_storage_dtype(L.input_dtype) == GLB_FP8E5) ? 1u : 0u;
    idesc = ptx::make_idesc(1u, wf, xf, TC_BM, BT);                                    // f32 accumulate, e4m3 = 0 / e5m2 = 1
  }
  const T8Plan plan = t8_plan(L, M, BT);
  if (plan.splits > 1) {
    GLB_REQUIRE(workspace != nullptr && ws_bytes >= plan.counter_bytes + plan.partial_bytes, GLB_ERR_WORKSPACE,
                "A8W8 split-K needs %zu bytes of (zero-initialised) workspace, got %zu", plan.counter_bytes + plan.partial_bytes, ws_bytes);
    GLB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, GLB_ERR_INVALID, "workspace must be 16-byte aligned");
    P.splits = plan.splits; P.kb_per_split = plan.kb_per_split;
    P.ws_counter = reinterpret_cast<int*>(workspace);
    P.ws_partial = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(workspace) + plan.counter_bytes);
    set_error("A8W8 split-K is not planned by t8_plan()");
    return GLB_ERR_UNSUPPORTED;
  }
  auto kern = gemm_a8w8_tc_kernel<IS_INT, BT>;
  GLB_ENSURE_SMEM(kern, S::TOTAL);
  const int grid = std::min(num_sms(), P.tiles_m * P.tiles_n);
  kern<<<grid, T8_THREADS, S::TOTAL, stream>>>(tx, tw, P, idesc);
  count_launch();
  GLB_CHECK_CUDA(cudaGetLastError());
  return GLB_OK;
}

bool a8w8_tc_supported(const glb_linear& L, int64_t M) {
  if (M < 1 || M > (1 << 24)) return false;
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
  if (L.N % TC_BM != 0 || L.K % T8_BK != 0) return false;
  if (reinterpret_cast<uintptr_t>(L.W_q) & 15) return false;
  return true;
}

}  // namespace

bool gemm_tc_supported(const glb_linear& L, int64_t M) {
  if (a8w8_tc_supported(L, M)) return true;
  if (gemm_mx_tc2_supported(L, M)) return true;
  if (gemm_w8_tc2_supported(L, M)) return true;
  if (M < 1 || M > (1 << 24)) return false;
  if (!(L.input_dtype == GLB_FP16 || L.input_dtype == GLB_BF16)) return false;
  if (L.output_dtype != L.input_dtype || L.meta_dtype != L.input_dtype) return false;
  // packed int32 words: 4-bit on both tcgen05 kernels; 8 / 2 / 1-bit (HQQ layers, packing_bitwidth 32) on the pair kernel only
  if (!(L.W_nbits == 4 || L.W_nbits == 8 || L.W_nbits == 2 || L.W_nbits == 1)) return false;
  if (L.elements_per_sample != 32 / L.W_nbits || L.wq_dtype != GLB_INT32) return false;
  if (L.W_nbits != 4 && !gemm_tc2_supported(L, M)) return false;
  if (L.wq_stride1 != 1 || L.wq_stride0 != L.N) return false;
  if (L.channel_scale_mode != 0) return false;
  if (!(L.W_group_mode == 2 || L.W_group_mode == 3 || L.W_group_mode == 4)) return false;
  if (L.W_group_mode != 2 && (L.zeros_numel != L.scales_numel || L.zeros_dtype != L.scales_dtype)) return false;
  if (L.group_size < 8 || L.group_size % 8 != 0 || L.K % L.group_size != 0) return false;
  if (L.scales_numel != (L.K / L.group_size) * L.N || L.s_stride1 != 1 || L.s_stride0 != L.N) return false;
  if (L.W_group_mode != 2 && (L.z_stride1 != 1 || L.z_stride0 != L.N)) return false;
  if (L.N % TC_BM != 0 || L.K % TC_BK != 0) return false;
  if (reinterpret_cast<uintptr_t>(L.W_q) & 15) return false;
  return true;
}

// which A16W4 / MXFP4 launches go to the CTA-pair kernel (gl_gemm_tc2.cu): MXFP4 always; INT4 above the decode range
// (GLB_PREFILL_VARIANT=1 keeps the single-CTA kernel for A/B comparisons)
static bool use_pair_kernel(const glb_linear& L, int64_t M) {
  if (gemm_mx_tc2_supported(L, M) || gemm_w8_tc2_supported(L, M)) return true;
  static int variant = -1;
  if (variant < 0) { const char* e = getenv("GLB_PREFILL_VARIANT"); variant = e ? atoi(e) : 2; }
  if (a8w8_tc_supported(L, M)) return false;
  if (L.W_nbits != 4) return true;                        // the single-CTA kernel unpacks 4-bit words only
  return variant != 1 && M > 16 && gemm_tc2_supported(L, M);
}

size_t gemm_tc_workspace(const glb_linear& L, int64_t M) {
  if (!gemm_tc_supported(L, M)) return 0;
  if (a8w8_tc_supported(L, M)) {
    if (M > 128) return 0;                                  // pair kernel or the 256-token block: no split-K
    const T8Plan p = t8_plan(L, M, t8_token_block(M));
    return p.counter_bytes + p.partial_bytes;
  }
  return use_pair_kernel(L, M) ? gemm_tc2_workspace(L, M) : 0;
}

int launch_gemm_tc(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M, void* workspace,
                   size_t ws_bytes, const void* bias, cudaStream_t stream) {
  GLB_REQUIRE(gemm_tc_supported(L, M), GLB_ERR_UNSUPPORTED, "tcgen05 GEMM: unsupported layer configuration");
  GLB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, GLB_ERR_INVALID, "tcgen05 GEMM: x must be 16-byte aligned");
  if (a8w8_tc_supported(L, M)) {
    {
      static int a8v = -1;        // GLB_A8W8_VARIANT=1 keeps the single-CTA kernel for A/B comparisons
      if (a8v < 0) { const char* e = getenv("GLB_A8W8_VARIANT"); a8v = e ? atoi(e) : 2; }
      if (a8v != 1 && gemm_a8_tc2_supported(L, M)) return launch_gemm_a8_tc2(L, x, scales_x, y, M, bias, stream);
    }
    const bool is_int = (act_storage_dtype(L.input_dtype) == GLB_INT8);
    set_kernel_name(is_int ? "gemm_a8w8_int8_tcgen05" : "gemm_a8w8_fp8_tcgen05");
#define GLB_T8(BT) (is_int ? launch_t8<true, BT>(L, x, scales_x, y, M, workspace, ws_bytes, bias, stream) \
                           : launch_t8<false, BT>(L, x, scales_x, y, M, workspace, ws_bytes, bias, stream))
    if (M <= 32) return GLB_T8(32);
    if (M <= 64) return GLB_T8(64);
    if (M <= 128) return GLB_T8(128);
    return GLB_T8(256);
#undef GLB_T8
  }
  if (use_pair_kernel(L, M)) return launch_gemm_tc2(L, x, y, M, workspace, ws_bytes, bias, stream);
  const bool bf = (L.input_dtype == GLB_BF16);
  set_kernel_name("gemm_a16w4_tcgen05");
  if (M <= 32) return bf ? launch_tc<true, 32>(L, x, y, M, bias, stream) : launch_tc<false, 32>(L, x, y, M, bias, stream);
  if (M <= 64) return bf ? launch_tc<true, 64>(L, x, y, M, bias, stream) : launch_tc<false, 64>(L, x, y, M, bias, stream);
  if (M <= 128) return bf ? launch_tc<true, 128>(L, x, y, M, bias, stream) : launch_tc<false, 128>(L, x, y, M, bias, stream);
  return bf ? launch_tc<true, 256>(L, x, y, M, bias, stream) : launch_tc<false, 256>(L, x, y, M, bias, stream);
}

}  // namespace glb
