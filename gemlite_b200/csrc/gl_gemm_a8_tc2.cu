// A8W8 (int8 x int8 -> int32, fp8 x fp8 -> fp32) on a CTA pair -- config C4 at large M (rows a9/a10).
//
// Replaces gemm_INT_kernel for W_group_mode 0 / 8-bit operands (gemm_kernels.py:248-413: tl.dot with an int32 / fp32
// accumulator, then the channel-scale epilogue :392-406).  No dequantisation: both operands are K-major byte matrices
// and go straight from TMA (128B swizzle) to the tensor core.
//
// Why a second kernel: the single-CTA A8W8 kernel (gl_gemm_tc.cu) stages 48 KB per 128-byte k-block, so only 4
// stages fit and the ring covers ~1500 tensor-core cycles -- less than a TMA round trip -- and its operand reads
// alone need 96 B/clk of shared-memory bandwidth (r1: 571 / 710 TFLOP/s at M = 4096, ~20 % of the 8-bit peak).
// Here (same structure as gl_gemm_tc2.cu, minus the dequant warps):
//   * tcgen05.mma.cta_group::2, M = 256 features (128 TMEM lanes per CTA) x N = 256 tokens; each CTA stages its 128
//     feature rows of W and its 128 tokens of x: 32 KB per 128-byte k-block -> 6 stages;
//   * the accumulator is double buffered (2 x 256 TMEM columns), so the epilogue of tile t overlaps the main loop of
//     tile t+1;
//   * MMA issuer on the highest warp id, one barrier wait + 4 MMAs (K = 32 bytes each) + one multicast commit per stage.
// Epilogue arithmetic is identical to the single-CTA kernel's (and the reference's): int32 -> cvt.rn.f32.s32, one
// fp32 multiply by (scales_x[m] * scales_w[n]) (rounded product), round to the output type, then bias.
// The first version of this kernel was EPILOGUE-bound (profiles/a8_int8_tc2_r1_ncu.txt: the MMA issuer never waited
// for data, only for the accumulator to be drained; 4 epilogue warps, one per scheduler, ~80 % busy in dependency
// stalls of a branchy per-element path).  Now: 8 epilogue warps (two per scheduler, 128 columns each) and a
// branch-free specialisation for the A8W8_dynamic case (channel_scale_mode 3, fp16 / bf16 output).
#include <stdlib.h>
#include <algorithm>

#include "gl_common.cuh"
#include "gl_ptx.cuh"

namespace glb {
namespace {

constexpr int A8_BM = 128;            // features per CTA (256 per pair)
constexpr int A8_BT = 256;            // tokens per pair tile (128 staged per CTA)
constexpr int A8_BK = 128;            // bytes (= elements) of K per stage
constexpr int A8_STAGES = 6;
constexpr int A8_THREADS = 384;       // w0 producer, w1 TMEM alloc, w2-9 epilogue (2 per lane quarter), w10 idle, w11 MMA issuer
constexpr int A8_EPI_WARPS = 8;
constexpr int A8_MMA_WARP = 11;
constexpr int A8_W_BYTES = A8_BM * A8_BK;          // 16 KB
constexpr int A8_X_BYTES = (A8_BT / 2) * A8_BK;    // 16 KB
constexpr int A8_STAGE = A8_W_BYTES + A8_X_BYTES;
constexpr int A8_BAR_OFF = A8_STAGES * A8_STAGE;
constexpr int A8_SMEM = A8_BAR_OFF + 1024;
constexpr int A8_GROUP_M = 4;

struct A8Params {
  const float* scales_w;   // (N) or null
  const float* scales_x;   // (M) or null
  void* y;
  const void* bias;
  int M, N, K;
  int csm;                 // channel_scale_mode 0..3
  int out_dtype;           // GLB_FP16 / BF16 / FP32 / INT32
  int tiles_m, tiles_n;
  uint32_t idesc;
};

__device__ __forceinline__ void a8_tile_coords(int t, int tiles_m, int tiles_n, int& tm, int& tn) {
  const int per_group = A8_GROUP_M * tiles_n;
  const int grp = t / per_group, r = t - grp * per_group;
  const int m0 = grp * A8_GROUP_M;
  const int gsz = (tiles_m - m0 < A8_GROUP_M) ? (tiles_m - m0) : A8_GROUP_M;
  tn = r / gsz;
  tm = m0 + (r - tn * gsz);
}

template <bool IS_INT>
__device__ __forceinline__ void mma_a8_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  if (IS_INT)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
                 : "memory");
}

template <bool IS_INT>
__global__ void __launch_bounds__(A8_THREADS, 1)
gemm_a8w8_tc2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const A8Params P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A8_BAR_OFF);
  uint64_t* full = bars;                          // [STAGES] leader: W + x bytes of both CTAs landed
  uint64_t* empty = bars + A8_STAGES;             // [STAGES] local : MMAs reading stage s completed (multicast commit)
  uint64_t* tfull = bars + 2 * A8_STAGES;         // [2]      local : accumulator buffer complete (multicast commit)
  uint64_t* tempty = tfull + 2;                   // [2]      leader: both CTAs' epilogue warps drained buffer b
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = (rank == 0);
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int kblocks = P.K / A8_BK;
  const int ntiles = P.tiles_m * P.tiles_n;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_w);
    for (int s = 0; s < A8_STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tfull[b], 1); ptx::mbar_init(&tempty[b], 2 * A8_EPI_WARPS); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc_2cta(tmem_slot, 512);
  ptx::tc_fence_before_sync();
  ptx::cluster_sync();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= producer (one lane in each CTA): this CTA's 128 feature rows of W and 128 tokens of x =================
    if (lane == 0) {
      const uint32_t full_leader = ptx::mapa(ptx::smem_u32(&full[0]), 0u);
      uint32_t s = 0, ph = 0;
      for (int t = pair; t < ntiles; t += npairs) {
        int tm, tn;
        a8_tile_coords(t, P.tiles_m, P.tiles_n, tm, tn);
        const int tok0 = tm * A8_BT + (int)rank * (A8_BT / 2), feat0 = tn * 2 * A8_BM + (int)rank * A8_BM;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&empty[s], ph ^ 1u);
          unsigned char* st = smem + s * A8_STAGE;
          if (leader) ptx::mbar_arrive_expect_tx(&full[s], 2 * A8_STAGE);
          ptx::tma_load_2d_2cta(st, &tmap_w, full_leader + s * 8u, kb * A8_BK, feat0);
          ptx::tma_load_2d_2cta(st + A8_W_BYTES, &tmap_x, full_leader + s * 8u, kb * A8_BK, tok0);
          if (++s == A8_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == A8_MMA_WARP) {
    // ================= MMA issuer (leader CTA) =================
    if (leader) {
      const uint64_t adesc0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem));
      const uint64_t bdesc0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + A8_W_BYTES));
      uint32_t s = 0, ph = 0;
      int it = 0;
      for (int t = pair; t < ntiles; t += npairs, ++it) {
        const int buf = it & 1;
        ptx::mbar_wait(&tempty[buf], (uint32_t)((it >> 1) & 1) ^ 1u);
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * A8_BT);
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after_sync();
          if (ptx::elect_one()) {
            const uint64_t so = (uint64_t)(s * (uint32_t)(A8_STAGE >> 4));
#pragma unroll
            for (int k = 0; k < A8_BK / 32; ++k)      // K = 32 bytes per instruction: 2 descriptor units
              mma_a8_ss_2cta<IS_INT>(d_tmem, adesc0 + so + (uint64_t)(2 * k), bdesc0 + so + (uint64_t)(2 * k), P.idesc, (kb | k) ? 1u : 0u);
            ptx::mma_commit_2cta(&empty[s], (uint16_t)3);
            if (kb == kblocks - 1) ptx::mma_commit_2cta(&tfull[buf], (uint16_t)3);
          }
          __syncwarp();
          if (++s == A8_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp >= 2 && warp < 2 + A8_EPI_WARPS) {
    // ================= epilogue: each CTA drains its own 128 feature rows; warp = (lane quarter, column half) =================
    const int q4 = warp & 3;                            // TMEM lane quarter == warp % 4
    const int chalf = (warp - 2) >> 2;                  // columns [chalf * 128, +128)
    const int row = q4 * 32 + lane;
    const uint32_t tempty_leader = ptx::mapa(ptx::smem_u32(&tempty[0]), 0u);
    const bool fast16 = (P.csm == 3) && (P.out_dtype == GLB_FP16 || P.out_dtype == GLB_BF16);
    int it = 0;
    for (int t = pair; t < ntiles; t += npairs, ++it) {
      int tm, tn;
      a8_tile_coords(t, P.tiles_m, P.tiles_n, tm, tn);
      const int buf = it & 1;
      const int n = tn * 2 * A8_BM + (int)rank * A8_BM + row;
      const float sb = (P.csm == 1 || P.csm == 3) ? P.scales_w[n] : 1.f;
      float bv = 0.f;
      if (P.bias) {
        if (P.out_dtype == GLB_FP16) bv = __half2float(reinterpret_cast<const __half*>(P.bias)[n]);
        else if (P.out_dtype == GLB_BF16) bv = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.bias)[n]);
        else bv = reinterpret_cast<const float*>(P.bias)[n];
      }
      const int mrem = P.M - tm * A8_BT;
      ptx::mbar_wait_sleep(&tfull[buf], (uint32_t)((it >> 1) & 1));
      ptx::tc_fence_after_sync();
#pragma unroll 1
      for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 32) {
        if (c0 >= mrem) break;
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * A8_BT + c0), r);
        // the 32 activation scales of this chunk: one coalesced load, then register shuffles
        float sxl = 1.f;
        if ((P.csm == 2 || P.csm == 3) && c0 + lane < mrem) sxl = P.scales_x[tm * A8_BT + c0 + lane];
        ptx::tmem_ld_wait();
        if (fast16) {
          // A8W8_dynamic: v = cvt(acc) * (sx * sb), rounded to fp16 / bf16 (+ bias): no per-element branches but the edge mask
          unsigned char* yp = reinterpret_cast<unsigned char*>(P.y) + ((int64_t)(tm * A8_BT + c0) * P.N + n) * 2;
          const int64_t ystride = (int64_t)P.N * 2;
          const bool f16 = (P.out_dtype == GLB_FP16), hb = (P.bias != nullptr);
          const int cnt = mrem - c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float sx = __shfl_sync(0xffffffffu, sxl, j);
            float v = IS_INT ? (float)(int32_t)r[j] : __uint_as_float(r[j]);
            v = __fmul_rn(v, __fmul_rn(sx, sb));
            unsigned short o;
            if (f16) {
              if (hb) v = __half2float(__float2half_rn(v)) + bv;
              o = __half_as_ushort(__float2half_rn(v));
            } else {
              if (hb) v = __bfloat162float(__float2bfloat16_rn(v)) + bv;
              o = __bfloat16_as_ushort(__float2bfloat16_rn(v));
            }
            if (j < cnt) *reinterpret_cast<unsigned short*>(yp + j * ystride) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float sx = __shfl_sync(0xffffffffu, sxl, j);
            if (c0 + j < mrem) {
              const int64_t oi = (int64_t)(tm * A8_BT + c0 + j) * P.N + n;
              if (IS_INT && P.out_dtype == GLB_INT32 && P.csm == 0 && !P.bias) {
                reinterpret_cast<int32_t*>(P.y)[oi] = (int32_t)r[j];
                continue;
              }
              float v = IS_INT ? (float)(int32_t)r[j] : __uint_as_float(r[j]);     // cvt.rn.f32.s32
              if (P.csm == 1) v = __fmul_rn(v, sb);
              else if (P.csm == 2) v = __fmul_rn(v, sx);
              else if (P.csm == 3) v = __fmul_rn(v, __fmul_rn(sx, sb));
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
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader + (uint32_t)buf * 8u);
    }
  }

  ptx::tc_fence_before_sync();
  ptx::cluster_sync();
  ptx::tc_fence_after_sync();
  if (warp == 1) ptx::tmem_dealloc_2cta(tmem_base, 512);
}

template <bool IS_INT>
int launch_a8(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M, const void* bias, cudaStream_t stream) {
  CUtensorMap tx, tw;
  int rc = make_tmap_2d(&tx, GLB_UINT8, x, (uint64_t)L.K, (uint64_t)M, (uint64_t)L.K, A8_BK, A8_BT / 2, 1);
  if (rc != GLB_OK) return rc;
  // W_q is the (K, N) view with strides (1, K) of an (N, K) row-major byte matrix
  rc = make_tmap_2d(&tw, GLB_UINT8, L.W_q, (uint64_t)L.K, (uint64_t)L.N, (uint64_t)L.K, A8_BK, A8_BM, 1);
  if (rc != GLB_OK) return rc;
  A8Params P;
  P.scales_w = (L.channel_scale_mode == 1 || L.channel_scale_mode == 3) ? reinterpret_cast<const float*>(L.scales) : nullptr;
  P.scales_x = scales_x; P.y = y; P.bias = bias;
  P.M = (int)M; P.N = (int)L.N; P.K = (int)L.K; P.csm = L.channel_scale_mode; P.out_dtype = L.output_dtype;
  P.tiles_m = (int)((M + A8_BT - 1) / A8_BT); P.tiles_n = (int)(L.N / (2 * A8_BM));
  if (IS_INT) P.idesc = ptx::make_idesc(2u, 1u, 1u, 2 * A8_BM, A8_BT);                  // s32 accumulate, signed 8-bit A/B
  else {
    const uint32_t wf = (L.wq_dtype == GLB_FP8E5) ? 1u : 0u, xf = (act_storage_dtype(L.input_dtype) == GLB_FP8E5) ? 1u : 0u;
    P.idesc = ptx::make_idesc(1u, wf, xf, 2 * A8_BM, A8_BT);                            // f32 accumulate, e4m3 = 0 / e5m2 = 1
  }
  auto kern = gemm_a8w8_tc2_kernel<IS_INT>;
  GLB_ENSURE_SMEM(kern, A8_SMEM);
  const int pairs = std::min(num_sms() / 2, P.tiles_m * P.tiles_n);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(A8_THREADS);
  cfg.dynamicSmemBytes = A8_SMEM;
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

// on top of a8w8_tc_supported(): whole 256-feature tiles and enough tokens to fill a 256-token tile reasonably
bool gemm_a8_tc2_supported(const glb_linear& L, int64_t M) {
  return M > 128 && L.N % (2 * A8_BM) == 0 && L.K % A8_BK == 0;
}

int launch_gemm_a8_tc2(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M, const void* bias,
                       cudaStream_t stream) {
  const bool is_int = (act_storage_dtype(L.input_dtype) == GLB_INT8);
  set_kernel_name(is_int ? "gemm_a8w8_int8_tcgen05_2cta" : "gemm_a8w8_fp8_tcgen05_2cta");
  return is_int ? launch_a8<true>(L, x, scales_x, y, M, bias, stream) : launch_a8<false>(L, x, scales_x, y, M, bias, stream);
}

}  // namespace glb
