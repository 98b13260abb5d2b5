// Block-scaled tcgen05 GEMM for the MX / NV layers with dynamically quantised activations (row f1, second stage):
//   A8W8_MXFP_dynamic   e4m3 x e4m3, one e8m0 scale per 32 elements on both operands   kind::mxf8f6f4.block_scale
//   A4W4_MXFP_dynamic   e2m1 x e2m1, e8m0 / 32                                          kind::mxf4.block_scale (scale_vec::2X)
//   A4W4_NVFP_dynamic   e2m1 x e2m1, e4m3 / 16, x 0.05^2                                kind::mxf4nvf4.block_scale (scale_vec::4X)
// -- what the reference reaches through tl.dot_scaled (gemm_kernels.py:423-547, :524).  The tensor core applies the
// group scales itself: no dequantisation, no bf16 expansion of the activations; products are exact and the accumulator
// is fp32, like tl.dot_scaled's.
//
// Structure = gl_gemm_a8_tc2.cu (CTA pair, tcgen05.mma.cta_group::2, M = 256 features x N = 256 tokens, each CTA stages
// its 128 feature rows of W and its 128 tokens of x by TMA with the 128-byte swizzle, 128 BYTES of K per stage = 4 MMAs, 5-stage ring)
// plus the scale plumbing, whose encodings were probed on hardware first (tools/bs_probe.cu,
// profiles/block_scaled_probe_r2.txt):
//   * scale factors live in TENSOR MEMORY next to the accumulator: for 128 rows one 32-bit word (4 consecutive group
//     scales) per row occupies 4 columns -- row r in lane r % 32, column r / 32, replicated in the four lane quarters;
//     tokens (N = 256) take 8 columns per word.  tcgen05.cp.32x128b.warpx4 copies one 512-byte block
//     (byte (r % 32) * 16 + (r / 32) * 4 + j) from shared memory into those 4 columns;
//   * four "scale" warps per CTA gather the words from the layouts the reference keeps -- weights (K / gs, N) bytes
//     N-contiguous, activations (M, K / gs) bytes row-major -- into that block layout, each warp a whole stage of every
//     fourth stage (they share the operand ring's barriers);
//   * per MMA the instruction descriptor selects the scale byte(s) inside the word: mxf8f6f4 (K = 32): byte k-step;
//     mxf4 (K = 64, 2 scales): byte 0 or 2 of word k-step / 2; nvf4 (K = 64, 4 scales): word k-step;
//   * tcgen05.cp and tcgen05.mma of one thread execute in issue order, so the scale columns are single buffered.
// Tensor memory: 256 accumulator columns + 48 scale columns (single-buffered accumulator: its drain is exposed, as in
// gl_gemm_tc2.cu).  Epilogue = the general kernel's (gl_generic.cu): NVFP4 x 0.0025, MXFP8 per-token post-scale
// (channel_scale_mode 2: round to the output type, x scales_x in fp32), round, bias.
#include <stdlib.h>
#include <algorithm>

#include "gl_common.cuh"
#include "gl_ptx.cuh"

namespace glb {
namespace {

constexpr int BS_MXFP8 = 0, BS_MXFP4 = 1, BS_NVFP4 = 2;
constexpr int BS_BM = 128;            // features per CTA (256 per pair)
constexpr int BS_BT = 256;            // tokens per pair tile (128 staged per CTA)
constexpr int BS_BK = 128;            // BYTES of K per row and stage
// ring geometry: a stage = 32 KB of operands + 6 KB of scale blocks (3 x 512 B per scale word, up to 4 words), 5 stages.
// (A 6-stage ring with the scale region trimmed per kind -- 34 / 35 KB stages for MXFP8 / MXFP4 -- measured 13-15 % SLOWER
// on two visits: MXFP8 8192^2 2743 -> 2377 TFLOP/s, MXFP4 16384^2 5622 -> 4787; not understood, reverted.)
constexpr int bs_sf_bytes(int) { return 3 * 4 * 512; }
constexpr int bs_stages(int) { return 5; }
constexpr int BS_THREADS = 512;       // w0 producer, w1 TMEM alloc, w2-9 epilogue, w11 MMA issuer, w10 + w12-15 scale warps
constexpr int BS_EPI_WARPS = 8, BS_MMA_WARP = 11, BS_SF_WARPS = 5;     // scale warps: 10, 12, 13, 14, 15
constexpr int BS_W_BYTES = BS_BM * BS_BK;            // 16 KB
constexpr int BS_X_BYTES = (BS_BT / 2) * BS_BK;      // 16 KB
constexpr int bs_stage_bytes(int nw) { return BS_W_BYTES + BS_X_BYTES + bs_sf_bytes(nw); }
constexpr int bs_smem_bytes(int nw) { return bs_stages(nw) * bs_stage_bytes(nw) + 1024; }
static_assert(bs_smem_bytes(1) <= 227 * 1024 && bs_smem_bytes(2) <= 227 * 1024 && bs_smem_bytes(4) <= 227 * 1024, "ring too large");
constexpr int BS_GROUP_M = 4;
constexpr uint32_t BS_SFA_COL = 256, BS_SFB_COL = 256 + 16;         // scale words behind the accumulator

struct BsParams {
  const uint8_t* scales_w;   // (K / gs, N) bytes, N contiguous
  const uint8_t* scales_x;   // (M, K / gs) bytes, or null (channel_scale_mode 2: no micro-scales on the activations)
  const float* post_x;       // (M) fp32 per-token scales (channel_scale_mode 2) or null
  void* y;
  const void* bias;
  int M, N, kbytes;          // kbytes = bytes of K per row
  int nsf;                   // scale bytes per row = K / gs
  int out_dtype;
  int tiles_m, tiles_n;
  int debug;                 // diagnostics (GLB_BS_DEBUG): 1 no scale copies, 2 scale warps arrive without writing, 4 scale warps do not take part
};

__device__ __forceinline__ void bs_tile_coords(int t, int tiles_m, int tiles_n, int& tm, int& tn) {
  const int per_group = BS_GROUP_M * tiles_n;
  const int grp = t / per_group, r = t - grp * per_group;
  const int m0 = grp * BS_GROUP_M;
  const int gsz = (tiles_m - m0 < BS_GROUP_M) ? (tiles_m - m0) : BS_GROUP_M;
  tn = r / gsz;
  tm = m0 + (r - tn * gsz);
}

// shared-memory matrix descriptor of one 512-byte scale block: 32 rows x 16 bytes, K-major, no swizzle, 8-row core
// matrices 128 bytes apart
__device__ __forceinline__ uint64_t make_sf_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void sf_copy_2cta(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// block-scaled instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptorBlockScaled): b_sf_id [4,6),
// a_format [7,10), b_format [10,13), N >> 3 at [17,23), scale format (1 = ue8m0) at 23, M >> 4 at [24,29), a_sf_id [29,31)
__device__ __forceinline__ uint32_t make_idesc_bs(uint32_t fmt, uint32_t m, uint32_t n, uint32_t scale_e8m0, uint32_t sf_id) {
  return (sf_id << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | (scale_e8m0 << 23) | ((m >> 4) << 24) | (sf_id << 29);
}
template <int KIND>
__device__ __forceinline__ void mma_bs_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t sfa, uint32_t sfb,
                                            uint32_t accumulate) {
  if (KIND == BS_MXFP8)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
                 : "memory");
  else if (KIND == BS_MXFP4)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::mxf4nvf4.block_scale.scale_vec::4X [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
                 : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(BS_THREADS, 1)
gemm_mx_bs_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const BsParams P) {
  constexpr int NW = (KIND == BS_MXFP8) ? 1 : ((KIND == BS_MXFP4) ? 2 : 4);     // 32-bit scale words per row and stage
  constexpr int BS_STAGES = bs_stages(NW), BS_STAGE = bs_stage_bytes(NW), BS_BAR_OFF = BS_STAGES * BS_STAGE;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BS_BAR_OFF);
  uint64_t* full = bars;                          // [STAGES] leader: operand bytes of both CTAs landed + both CTAs' scale warps wrote stage s
  uint64_t* empty = bars + BS_STAGES;             // [STAGES] local : MMAs (and scale copies) reading stage s completed (multicast commit)
  uint64_t* tfull = bars + 2 * BS_STAGES;         // [1]      local : accumulator complete (multicast commit)
  uint64_t* tempty = tfull + 1;                   // [1]      leader: both CTAs' epilogue warps drained it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = (rank == 0);
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int kblocks = P.kbytes / BS_BK;
  const int ntiles = P.tiles_m * P.tiles_n;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_w);
    for (int s = 0; s < BS_STAGES; ++s) { ptx::mbar_init(&full[s], (P.debug & 4) ? 1 : 1 + 2); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(tfull, 1);
    ptx::mbar_init(tempty, 2 * BS_EPI_WARPS);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc_2cta(tmem_slot, 512);
  ptx::tc_fence_before_sync();
  ptx::cluster_sync();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= operand producer (one lane in each CTA) =================
    if (lane == 0) {
      const uint32_t full_leader = ptx::mapa(ptx::smem_u32(&full[0]), 0u);
      uint32_t s = 0, ph = 0;
      for (int t = pair; t < ntiles; t += npairs) {
        int tm, tn;
        bs_tile_coords(t, P.tiles_m, P.tiles_n, tm, tn);
        const int tok0 = tm * BS_BT + (int)rank * (BS_BT / 2), feat0 = tn * 2 * BS_BM + (int)rank * BS_BM;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&empty[s], ph ^ 1u);
          unsigned char* st = smem + s * BS_STAGE;
          if (leader) ptx::mbar_arrive_expect_tx(&full[s], 2 * (BS_W_BYTES + BS_X_BYTES));
          ptx::tma_load_2d_2cta(st, &tmap_w, full_leader + s * 8u, kb * BS_BK, feat0);
          ptx::tma_load_2d_2cta(st + BS_W_BYTES, &tmap_x, full_leader + s * 8u, kb * BS_BK, tok0);
          if (++s == BS_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 10 || warp >= 12) {
    // ================= scale warps: warp sw owns every 5th stage of this pair's sequence and fills ALL its scale blocks =================
    // A stage's scale words are gathered by ONE warp: it issues its loads (weights: 32-bit words = 4 consecutive features of
    // one group, coalesced; activations: the NW words of a token, one vector load), waits for them, waits for the ring
    // slot, writes the blocks (weights: bytes of 4 groups re-packed into per-feature words; activations: the four tokens
    // lane, lane + 32, lane + 64, lane + 96 of a block are ONE 16-byte row) and arrives.  Five warps work on five
    // different stages, so the memory latency of one stage is hidden behind the other four.
    // (History, profiles/block_scaled_r2.txt: (1) all four warps on every stage, words assembled right after the byte
    // loads: every stage sat out the memory latency, 40 % of the unscaled fp8 kernel; (2) register prefetch 4 stages ahead:
    // fence.proxy.async waits for the OUTSTANDING prefetch loads, same stall; with the scale warps switched off the same
    // kernel ran 1.5x (MXFP8) / 2.5x (NVFP4) faster.)
    const int sw = (warp == 10) ? 0 : warp - 11;
    const uint32_t full_leader = ptx::mapa(ptx::smem_u32(&full[0]), 0u);
    const int my_tiles = (ntiles > pair) ? (ntiles - pair + npairs - 1) / npairs : 0;
    const int total = (P.debug & 4) ? 0 : my_tiles * kblocks;
    // this lane's 32-bit stores on the weight side: feature f = 4 lane + b -> byte offset (f % 32) * 16 + (f / 32) * 4
    const uint32_t foff0 = (uint32_t)((4 * (lane & 7)) * 16 + (lane >> 3) * 4);
    for (int i = sw; i < total; i += BS_SF_WARPS) {
      const int j = i / kblocks, kb = i - j * kblocks;
      const int s = i % BS_STAGES;
      const uint32_t ph = (uint32_t)((i / BS_STAGES) & 1);
      int tm, tn;
      bs_tile_coords(pair + j * npairs, P.tiles_m, P.tiles_n, tm, tn);
      // ---- loads ----
      uint32_t wfv[NW][4];                                              // [word][group inside the word]: features 4 lane .. 4 lane + 3
      const uint8_t* ws = P.scales_w + (size_t)(kb * 4 * NW) * P.N + (tn * 2 * BS_BM + (int)rank * BS_BM + 4 * lane);
#pragma unroll
      for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int g = 0; g < 4; ++g) wfv[w][g] = __ldg(reinterpret_cast<const uint32_t*>(ws + (size_t)(4 * w + g) * P.N));
      uint32_t wtv[8][NW];                                              // [token lane + 32 q][word]
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int tok = tm * BS_BT + lane + 32 * q;
#pragma unroll
        for (int w = 0; w < NW; ++w) wtv[q][w] = P.scales_x ? 0u : 0x7f7f7f7fu;       // e8m0 1.0 for post-scaled MXFP8 activations
        if (P.scales_x && tok < P.M) {
          const uint8_t* p = P.scales_x + (size_t)tok * P.nsf + kb * 4 * NW;
          if (NW == 1) wtv[q][0] = __ldg(reinterpret_cast<const uint32_t*>(p));
          else if (NW == 2) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); wtv[q][0] = v.x; wtv[q][NW - 1] = v.y; }
          else { const uint4 v = __ldg(reinterpret_cast<const uint4*>(p)); wtv[q][0] = v.x; wtv[q][1 % NW] = v.y; wtv[q][2 % NW] = v.z; wtv[q][3 % NW] = v.w; }
        }
      }
      // ---- the ring slot ----
      ptx::mbar_wait(&empty[s], ph ^ 1u);
      unsigned char* sf = smem + s * BS_STAGE + BS_W_BYTES + BS_X_BYTES;
      if (!(P.debug & 2)) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          unsigned char* fa = sf + (3 * w + 0) * 512 + foff0;
          // feature 4 lane + b: its word = byte b of the four group words
          const uint32_t lo01 = __byte_perm(wfv[w][0], wfv[w][1], 0x5140), hi01 = __byte_perm(wfv[w][0], wfv[w][1], 0x7362);   // (b0g0,b0g1,b1g0,b1g1) / (b2.., b3..)
          const uint32_t lo23 = __byte_perm(wfv[w][2], wfv[w][3], 0x5140), hi23 = __byte_perm(wfv[w][2], wfv[w][3], 0x7362);
          *reinterpret_cast<uint32_t*>(fa + 0 * 16) = __byte_perm(lo01, lo23, 0x5410);    // b = 0: g0 g1 g2 g3
          *reinterpret_cast<uint32_t*>(fa + 1 * 16) = __byte_perm(lo01, lo23, 0x7632);    // b = 1
          *reinterpret_cast<uint32_t*>(fa + 2 * 16) = __byte_perm(hi01, hi23, 0x5410);    // b = 2
          *reinterpret_cast<uint32_t*>(fa + 3 * 16) = __byte_perm(hi01, hi23, 0x7632);    // b = 3
          *reinterpret_cast<uint4*>(sf + (3 * w + 1) * 512 + lane * 16) = make_uint4(wtv[0][w], wtv[1][w], wtv[2][w], wtv[3][w]);
          *reinterpret_cast<uint4*>(sf + (3 * w + 2) * 512 + lane * 16) = make_uint4(wtv[4][w], wtv[5][w], wtv[6][w], wtv[7][w]);
        }
      }
      ptx::fence_proxy_async_smem();                 // generic-proxy stores before the tensor core's (async proxy) read
      __syncwarp();
      // plain arrive (.release.cta), as everywhere in these kernels: a cluster-scope release was tried for the peer's
      // arrivals and cost 25-30 % of the GEMM (MXFP8 8192^2 2743 -> 2082 TFLOP/s: the fence sits on every stage's path)
      if (lane == 0) ptx::mbar_arrive_cluster(full_leader + (uint32_t)s * 8u);
    }
  } else if (warp == BS_MMA_WARP) {
    // ================= MMA issuer (leader CTA): per stage NW x 3 scale copies, 4 MMAs, one multicast commit =================
    if (leader) {
      const uint64_t adesc0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem));
      const uint64_t bdesc0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + BS_W_BYTES));
      const uint32_t sf_smem0 = ptx::smem_u32(smem + BS_W_BYTES + BS_X_BYTES);
      const uint32_t fmt = (KIND == BS_MXFP8) ? 0u : 1u, e8 = (KIND == BS_NVFP4) ? 0u : 1u;
      uint32_t s = 0, ph = 0;
      int it = 0;
      for (int t = pair; t < ntiles; t += npairs, ++it) {
        ptx::mbar_wait(tempty, (uint32_t)(it & 1) ^ 1u);
        ptx::tc_fence_after_sync();
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after_sync();
          if (ptx::elect_one()) {
            const uint32_t sfs = sf_smem0 + s * (uint32_t)BS_STAGE;
#pragma unroll
            for (int w = 0; w < ((P.debug & 1) ? 0 : NW); ++w) {
              sf_copy_2cta(tmem_base + BS_SFA_COL + 4u * w, make_sf_desc(sfs + (3 * w + 0) * 512));
              sf_copy_2cta(tmem_base + BS_SFB_COL + 8u * w, make_sf_desc(sfs + (3 * w + 1) * 512));
              sf_copy_2cta(tmem_base + BS_SFB_COL + 8u * w + 4u, make_sf_desc(sfs + (3 * w + 2) * 512));
            }
            const uint64_t so = (uint64_t)(s * (uint32_t)(BS_STAGE >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k) {                 // 32 bytes of K per instruction: 2 descriptor units
              const uint32_t sel = (KIND == BS_MXFP8) ? (uint32_t)k : ((KIND == BS_MXFP4) ? (uint32_t)(2 * (k & 1)) : 0u);
              const uint32_t word = (KIND == BS_MXFP8) ? 0u : ((KIND == BS_MXFP4) ? (uint32_t)(k >> 1) : (uint32_t)k);
              mma_bs_2cta<KIND>(tmem_base, adesc0 + so + (uint64_t)(2 * k), bdesc0 + so + (uint64_t)(2 * k),
                                make_idesc_bs(fmt, 2 * BS_BM, BS_BT, e8, sel), tmem_base + BS_SFA_COL + 4u * word,
                                tmem_base + BS_SFB_COL + 8u * word, (kb | k) ? 1u : 0u);
            }
            ptx::mma_commit_2cta(&empty[s], (uint16_t)3);
            if (kb == kblocks - 1) ptx::mma_commit_2cta(tfull, (uint16_t)3);
          }
          __syncwarp();
          if (++s == BS_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp >= 2 && warp < 2 + BS_EPI_WARPS) {
    // ================= epilogue: each CTA drains its own 128 feature rows; warp = (lane quarter, column half) =================
    const int q4 = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t tempty_leader = ptx::mapa(ptx::smem_u32(tempty), 0u);
    const float alpha = (KIND == BS_NVFP4) ? 0.0025f : 1.f;      // NVFP4 meta scale 0.05^2 (gemm_kernels.py:461,530)
    int it = 0;
    for (int t = pair; t < ntiles; t += npairs, ++it) {
      int tm, tn;
      bs_tile_coords(t, P.tiles_m, P.tiles_n, tm, tn);
      const int n = tn * 2 * BS_BM + (int)rank * BS_BM + row;
      float bv = 0.f;
      if (P.bias) {
        if (P.out_dtype == GLB_FP16) bv = __half2float(reinterpret_cast<const __half*>(P.bias)[n]);
        else if (P.out_dtype == GLB_BF16) bv = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.bias)[n]);
        else bv = reinterpret_cast<const float*>(P.bias)[n];
      }
      const int mrem = P.M - tm * BS_BT;
      ptx::mbar_wait_sleep(tfull, (uint32_t)(it & 1));
      ptx::tc_fence_after_sync();
#pragma unroll 1
      for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 32) {
        if (c0 >= mrem) break;
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)c0, r);
        float sxl = 1.f;
        if (P.post_x && c0 + lane < mrem) sxl = P.post_x[tm * BS_BT + c0 + lane];
        ptx::tmem_ld_wait();
        if (!P.post_x && P.out_dtype != GLB_FP32) {
          // micro-scaled activations, 16-bit output: the accumulator is single buffered, so this drain is exposed -- one
          // cvt, one address add and one 2-byte store per element, no per-element branches but the edge mask
          unsigned char* yp = reinterpret_cast<unsigned char*>(P.y) + ((int64_t)(tm * BS_BT + c0) * P.N + n) * 2;
          const int64_t ystride = (int64_t)P.N * 2;
          const bool f16 = (P.out_dtype == GLB_FP16), hb = (P.bias != nullptr);
          const int cnt = mrem - c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = __uint_as_float(r[j]);
            if (KIND == BS_NVFP4) v *= alpha;
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
          continue;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float sx = __shfl_sync(0xffffffffu, sxl, j);
          if (c0 + j < mrem) {
            const int64_t oi = (int64_t)(tm * BS_BT + c0 + j) * P.N + n;
            float v = __uint_as_float(r[j]);
            if (KIND == BS_NVFP4) v *= alpha;
            if (P.out_dtype == GLB_FP16) {
              if (P.post_x) v = __fmul_rn(__half2float(__float2half_rn(v)), sx);
              if (P.bias) v = __half2float(__float2half_rn(v)) + bv;
              reinterpret_cast<__half*>(P.y)[oi] = __float2half_rn(v);
            } else if (P.out_dtype == GLB_BF16) {
              if (P.post_x) v = __fmul_rn(__bfloat162float(__float2bfloat16_rn(v)), sx);
              if (P.bias) v = __bfloat162float(__float2bfloat16_rn(v)) + bv;
              reinterpret_cast<__nv_bfloat16*>(P.y)[oi] = __float2bfloat16_rn(v);
            } else {
              if (P.post_x) v = __fmul_rn(v, sx);
              reinterpret_cast<float*>(P.y)[oi] = v + bv;
            }
          }
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader);
    }
  }

  ptx::tc_fence_before_sync();
  ptx::cluster_sync();
  ptx::tc_fence_after_sync();
  if (warp == 1) ptx::tmem_dealloc_2cta(tmem_base, 512);
}

int bs_kind(const glb_linear& L) {
  if (L.input_dtype == GLB_MXFP8) return BS_MXFP8;
  if (L.input_dtype == GLB_MXFP4) return BS_MXFP4;
  return BS_NVFP4;
}

template <int KIND>
int launch_bs(const glb_linear& L, const void* x, const void* scales_x, void* y, int64_t M, const void* bias, cudaStream_t stream) {
  const uint64_t kbytes = (KIND == BS_MXFP8) ? (uint64_t)L.K : (uint64_t)(L.K / 2);
  CUtensorMap tx, tw;
  int rc = make_tmap_2d(&tx, GLB_UINT8, x, kbytes, (uint64_t)M, kbytes, BS_BK, BS_BT / 2, 1);
  if (rc != GLB_OK) return rc;
  rc = make_tmap_2d(&tw, GLB_UINT8, L.W_q, kbytes, (uint64_t)L.N, kbytes, BS_BK, BS_BM, 1);
  if (rc != GLB_OK) return rc;
  BsParams P;
  P.scales_w = reinterpret_cast<const uint8_t*>(L.scales);
  P.scales_x = (L.channel_scale_mode == 4) ? reinterpret_cast<const uint8_t*>(scales_x) : nullptr;
  P.post_x = (L.channel_scale_mode == 2) ? reinterpret_cast<const float*>(scales_x) : nullptr;
  P.y = y; P.bias = bias; P.M = (int)M; P.N = (int)L.N; P.kbytes = (int)kbytes; P.nsf = (int)(L.K / L.group_size);
  P.out_dtype = act_storage_dtype(L.output_dtype);
  P.tiles_m = (int)((M + BS_BT - 1) / BS_BT); P.tiles_n = (int)(L.N / (2 * BS_BM));
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("GLB_BS_DEBUG"); dbg = e ? atoi(e) : 0; }
    P.debug = dbg;
  }
  auto kern = gemm_mx_bs_kernel<KIND>;
  constexpr int NW = (KIND == BS_MXFP8) ? 1 : ((KIND == BS_MXFP4) ? 2 : 4);
  constexpr int BS_SMEM = bs_smem_bytes(NW);
  GLB_ENSURE_SMEM(kern, BS_SMEM);
  const int pairs = std::min(num_sms() / 2, P.tiles_m * P.tiles_n);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(BS_THREADS);
  cfg.dynamicSmemBytes = BS_SMEM;
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

// MX / NV layers with dynamically quantised activations whose operands the tensor core takes as they are: both operands
// the same element type (e4m3 / e2m1), K-contiguous, group scales in the reference's layouts, whole 256-feature tiles
bool gemm_bs_supported(const glb_linear& L, int64_t M) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("GLB_BLOCK_SCALED"); enabled = e ? atoi(e) : 1; }
  if (!enabled || M < 1) return false;
  const bool f8 = (L.input_dtype == GLB_MXFP8), f4 = (L.input_dtype == GLB_MXFP4), nv = (L.input_dtype == GLB_NVFP4);
  if (!f8 && !f4 && !nv) return false;
  if (!(L.channel_scale_mode == 4 || (f8 && L.channel_scale_mode == 2))) return false;
  if (f8) {
    if (L.W_nbits != 8 || L.elements_per_sample != 1 || L.wq_dtype != GLB_FP8E4) return false;
    if (L.wq_stride0 != 1 || L.wq_stride1 != L.K) return false;
    if (L.K % 128 != 0) return false;
  } else {
    if (L.W_nbits != 4 || L.elements_per_sample != 2 || L.wq_dtype != GLB_UINT8) return false;
    if (L.wq_stride0 != 1 || L.wq_stride1 != L.K / 2) return false;
    if (L.K % 256 != 0) return false;
  }
  if (L.group_size != (nv ? 16 : 32)) return false;
  if (nv ? (L.scales_dtype != GLB_FP8E4) : !(L.scales_dtype == GLB_UINT8 || L.scales_dtype == GLB_E8M0)) return false;
  if (L.scales_numel != (L.K / L.group_size) * L.N || L.s_stride0 != 1 || L.s_stride1 != L.N) return false;
  const int od = act_storage_dtype(L.output_dtype);
  if (!(od == GLB_FP16 || od == GLB_BF16 || od == GLB_FP32)) return false;
  if (L.N % (2 * BS_BM) != 0) return false;
  if (reinterpret_cast<uintptr_t>(L.W_q) & 15) return false;
  return true;
}

int launch_gemm_bs(const glb_linear& L, const void* x, const void* scales_x, void* y, int64_t M, const void* bias, cudaStream_t stream) {
  GLB_REQUIRE(gemm_bs_supported(L, M), GLB_ERR_UNSUPPORTED, "block-scaled GEMM: unsupported layer configuration");
  GLB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, GLB_ERR_INVALID, "block-scaled GEMM: x must be 16-byte aligned");
  if (L.channel_scale_mode == 4)
    GLB_REQUIRE((reinterpret_cast<uintptr_t>(scales_x) & 3) == 0, GLB_ERR_INVALID, "block-scaled GEMM: activation scales must be 4-byte aligned");
  switch (bs_kind(L)) {
    case BS_MXFP8: set_kernel_name("gemm_mxfp8_block_scaled_tcgen05_2cta"); return launch_bs<BS_MXFP8>(L, x, scales_x, y, M, bias, stream);
    case BS_MXFP4: set_kernel_name("gemm_mxfp4_block_scaled_tcgen05_2cta"); return launch_bs<BS_MXFP4>(L, x, scales_x, y, M, bias, stream);
    default: set_kernel_name("gemm_nvfp4_block_scaled_tcgen05_2cta"); return launch_bs<BS_NVFP4>(L, x, scales_x, y, M, bias, stream);
  }
}

}  // namespace glb
