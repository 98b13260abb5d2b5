// Decode-regime kernel (M <= 8) for grouped layers whose codes are packed 8 / 2 / 1 bits wide in int32 words
// (A16W8 / A16W2 / A16W1_HQQ_INT and friends, packing_bitwidth 32: helper.py:344-400; reference kernels
// gemv_revsplitK / gemv_splitK, gemv_kernels.py:244-330).  Until this kernel those layers ran the CTA-pair tcgen05
// kernel with split-K at every M: 15 us per call at M = 1 on a 4096^2 2-bit layer whose 4 MB stream in 0.7 us
// (the reference's Triton GEMV: 8.1 us, profiles/vs_reference_timing_r2_visit12.txt).
//
// HBM-bound byte work, no tensor cores: W_q (K / eps, N) int32 N-contiguous, eps = 32 / NBITS codes per word.
//   * grid = (N / 128 column blocks) x KS k-slices, the KS CTAs of a column block form a CLUSTER (KS <= 8);
//   * thread = one column: its words of the slice are 32-bit loads, coalesced across the 128 threads of the CTA (512 B per
//     packed row), issued BEFORE griddepcontrol.wait (weights do not depend on the previous kernel), 32 rows in flight;
//   * x slice (<= 8 tokens) and the per-group sums of x in shared memory as fp32;
//   * per group: acc = sum_k code_k x_k in fp32 (integer codes are exact), then the group's scale / zero applied once:
//     y += s (acc - z sx)  (W_group_mode 3),  s acc + z' sx  (4),  s acc  (2)  -- the factorisation of the A16W4 decode
//     kernel (gl_gemv.cu): closer to the exact product than rounding every dequantised weight to 16 bits first, within the
//     decode tolerance of the oracle (tests: REL_DECODE);
//   * k-slices meet through distributed shared memory in rank order (deterministic), rank 0 adds the bias and stores.
// Algorithmic bytes per launch: K N NBITS / 8 + 2 (2) G N (scales, zeros) + 2 M K + 2 M N.
#include <stdlib.h>

#include "gl_common.cuh"
#include "gl_ptx.cuh"

namespace glb {
namespace {

constexpr int PV_COLS = 128;          // columns (threads) per CTA
constexpr int PV_CHUNK = 32;          // packed rows a thread keeps in flight

struct PvParams {
  const uint32_t* W;        // (K / eps, N)
  const void* scales;       // (G, N) 16-bit
  const void* zeros;        // (G, N) 16-bit or null (mode 2)
  const void* x;            // (M, K) 16-bit
  void* y;                  // (M, N) 16-bit
  const void* bias;
  int M, N, K, group_size, mode, ks;
};

template <bool BF16> __device__ __forceinline__ float pv_to_float(uint16_t raw) {
  return BF16 ? __uint_as_float((uint32_t)raw << 16) : __half2float(__ushort_as_half(raw));
}

template <bool BF16, int NBITS, int MT>
__global__ void __launch_bounds__(PV_COLS, 4) gemv_packed_kernel(const PvParams P) {
  constexpr int EPS = 32 / NBITS;
  constexpr uint32_t MASK = (NBITS == 32) ? 0xffffffffu : ((1u << NBITS) - 1u);
  extern __shared__ __align__(16) unsigned char smem[];
  const int ks = P.ks;
  const uint32_t rank = (ks > 1) ? ptx::cluster_ctarank() : 0u;
  const int col = (blockIdx.x / ks) * PV_COLS + threadIdx.x;
  const int kslice = P.K / ks;                       // k per CTA (whole groups)
  const int k0 = (int)rank * kslice;
  const int rows = kslice / EPS, row0 = k0 / EPS;    // packed rows of this CTA
  const int gslice = kslice / P.group_size, g0 = k0 / P.group_size;
  const int wpg = P.group_size / EPS;                // packed rows per group
  float* xs = reinterpret_cast<float*>(smem);                        // [MT][kslice]
  float* sx = xs + (size_t)MT * kslice;                              // [MT][gslice]
  float* red = sx + (size_t)MT * gslice;                             // [MT][128] partial outputs of this CTA
  float* meta = red + (size_t)MT * PV_COLS;                          // [gslice][2][128] scale, zero of (group, column) as fp32

  // ---- weights of the first chunk: independent of the previous kernel ----
  const uint32_t* wp = P.W + (size_t)row0 * P.N + col;
  uint32_t w[PV_CHUNK];
#pragma unroll
  for (int i = 0; i < PV_CHUNK; ++i) w[i] = (i < rows) ? __ldg(wp + (size_t)i * P.N) : 0u;
  // ---- this column's scales / zeros of the slice: also before the hand-off, parked in the thread's own shared-memory
  //      slots (loading them at each group boundary put two dependent global loads per group on the critical path:
  //      12 us per call for a 4 MB layer) ----
  {
    const uint16_t* sp0 = reinterpret_cast<const uint16_t*>(P.scales) + (size_t)g0 * P.N + col;
    const uint16_t* zp0 = P.zeros ? reinterpret_cast<const uint16_t*>(P.zeros) + (size_t)g0 * P.N + col : nullptr;
    for (int gg = 0; gg < gslice; ++gg) {
      meta[(2 * gg) * PV_COLS + threadIdx.x] = pv_to_float<BF16>(__ldg(sp0 + (size_t)gg * P.N));
      meta[(2 * gg + 1) * PV_COLS + threadIdx.x] = zp0 ? pv_to_float<BF16>(__ldg(zp0 + (size_t)gg * P.N)) : 0.f;
    }
  }
  pdl_wait_prior_grids();
  pdl_launch_dependents();

  // ---- x slice -> fp32 in shared memory, group sums ----
  const uint16_t* xg = reinterpret_cast<const uint16_t*>(P.x);
  for (int i = threadIdx.x; i < MT * kslice; i += PV_COLS) {
    const int m = i / kslice, k = i - m * kslice;
    xs[i] = (m < P.M) ? pv_to_float<BF16>(xg[(size_t)m * P.K + k0 + k]) : 0.f;
  }
  __syncthreads();
  {
    // one warp per (token, group): lanes stride over the group, shuffle tree (a serial loop per thread was 64 dependent
    // shared-memory loads = ~1 us on the critical path)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < MT * gslice; i += PV_COLS / 32) {
      const int m = i / gslice, g = i - m * gslice;
      const float* xr = xs + (size_t)m * kslice + (size_t)g * P.group_size;
      float s = 0.f;
      for (int k = lane; k < P.group_size; k += 32) s += xr[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) sx[i] = s;
    }
  }
  __syncthreads();

  // ---- main loop: chunks of PV_CHUNK packed rows, groups closed as they complete ----
  float out[MT], acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) { out[m] = 0.f; acc[m] = 0.f; }
  int in_group = 0, g = 0;
  for (int r0 = 0; r0 < rows; r0 += PV_CHUNK) {
    if (r0 > 0) {
#pragma unroll
      for (int i = 0; i < PV_CHUNK; ++i) w[i] = (r0 + i < rows) ? __ldg(wp + (size_t)(r0 + i) * P.N) : 0u;
    }
#pragma unroll
    for (int i = 0; i < PV_CHUNK; ++i) {
      if (r0 + i < rows) {
        const float* xr = xs + (size_t)(r0 + i) * EPS;
        const uint32_t q = w[i];
#pragma unroll
        for (int j = 0; j < EPS; j += 4) {
          // code -> fp32 through the 2^23 magic (exact; cvt.f32.u32 runs at a quarter of the FMA rate and was the limiter)
          float c[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) c[u] = __uint_as_float(((q >> (NBITS * (j + u))) & MASK) | 0x4B000000u) - 8388608.f;
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            const float4 xv = *reinterpret_cast<const float4*>(xr + (size_t)m * kslice + j);     // broadcast, 16-byte aligned
            acc[m] = fmaf(c[0], xv.x, acc[m]); acc[m] = fmaf(c[1], xv.y, acc[m]);
            acc[m] = fmaf(c[2], xv.z, acc[m]); acc[m] = fmaf(c[3], xv.w, acc[m]);
          }
        }
        if (++in_group == wpg) {                       // group complete: apply its scale / zero once
          const float s = meta[(2 * g) * PV_COLS + threadIdx.x], z = meta[(2 * g + 1) * PV_COLS + threadIdx.x];
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            const float sxm = sx[(size_t)m * gslice + g];
            if (P.mode == 3) out[m] = fmaf(s, acc[m] - z * sxm, out[m]);
            else if (P.mode == 4) out[m] = fmaf(s, acc[m], fmaf(z, sxm, out[m]));
            else out[m] = fmaf(s, acc[m], out[m]);
            acc[m] = 0.f;
          }
          in_group = 0; ++g;
        }
      }
    }
  }

  // ---- k-slices: through distributed shared memory, rank order ----
#pragma unroll
  for (int m = 0; m < MT; ++m) red[m * PV_COLS + threadIdx.x] = out[m];
  if (ks > 1) ptx::cluster_sync(); else __syncthreads();
  if (rank == 0) {
    const uint32_t red_addr = ptx::smem_u32(red);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      float v = out[m];
      for (int r = 1; r < ks; ++r) {
        float u;
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(u) : "r"(ptx::mapa(red_addr + (uint32_t)(m * PV_COLS + threadIdx.x) * 4u, (uint32_t)r)) : "memory");
        v += u;
      }
      if (m < P.M) {
        const int64_t oi = (int64_t)m * P.N + col;
        if (BF16) {
          if (P.bias) v = __bfloat162float(__float2bfloat16_rn(v)) + __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(P.bias)[col]);
          reinterpret_cast<__nv_bfloat16*>(P.y)[oi] = __float2bfloat16_rn(v);
        } else {
          if (P.bias) v = __half2float(__float2half_rn(v)) + __half2float(reinterpret_cast<const __half*>(P.bias)[col]);
          reinterpret_cast<__half*>(P.y)[oi] = __float2half_rn(v);
        }
      }
    }
  }
  if (ks > 1) ptx::cluster_sync();                     // peers keep their partials alive until rank 0 has read them
}

int pv_ks(const glb_linear& L) {
  const int G = (int)(L.K / L.group_size);
  for (int ks = 8; ks > 1; ks >>= 1)
    if (G % ks == 0) return ks;
  return 1;
}
size_t pv_smem(const glb_linear& L, int mt) {
  const int ks = pv_ks(L), kslice = (int)(L.K / ks), gslice = kslice / L.group_size;
  return ((size_t)mt * (kslice + gslice + PV_COLS) + (size_t)2 * gslice * PV_COLS) * sizeof(float);
}

template <bool BF16, int NBITS, int MT>
int launch_pv(const glb_linear& L, const PvParams& P, cudaStream_t stream) {
  auto kern = gemv_packed_kernel<BF16, NBITS, MT>;
  const size_t smem = pv_smem(L, MT);
  if (smem > 48 * 1024) GLB_ENSURE_SMEM(kern, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(L.N / PV_COLS) * (unsigned)P.ks);
  cfg.blockDim = dim3(PV_COLS);
  cfg.dynamicSmemBytes = smem;
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
  GLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  count_launch();
  return GLB_OK;
}

template <bool BF16, int NBITS>
int dispatch_pv(const glb_linear& L, const PvParams& P, cudaStream_t stream) {
  if (P.M == 1) return launch_pv<BF16, NBITS, 1>(L, P, stream);
  if (P.M == 2) return launch_pv<BF16, NBITS, 2>(L, P, stream);
  if (P.M <= 4) return launch_pv<BF16, NBITS, 4>(L, P, stream);
  return launch_pv<BF16, NBITS, 8>(L, P, stream);
}

}  // namespace

bool gemv_packed_supported(const glb_linear& L, int64_t M) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("GLB_GEMV_PACKED"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return false;
  if (M < 1 || M > 8) return false;
  if (!(L.input_dtype == GLB_FP16 || L.input_dtype == GLB_BF16)) return false;
  if (L.output_dtype != L.input_dtype || L.meta_dtype != L.input_dtype) return false;
  if (!(L.W_nbits == 8 || L.W_nbits == 2 || L.W_nbits == 1)) return false;
  if (L.elements_per_sample != 32 / L.W_nbits || L.wq_dtype != GLB_INT32) return false;
  if (L.wq_stride1 != 1 || L.wq_stride0 != L.N) return false;
  if (L.channel_scale_mode != 0) return false;
  if (!(L.W_group_mode == 2 || L.W_group_mode == 3 || L.W_group_mode == 4)) return false;
  if (L.W_group_mode != 2 && (L.zeros_numel != L.scales_numel || L.zeros_dtype != L.scales_dtype)) return false;
  const int gs = L.group_size, eps = 32 / L.W_nbits;
  if (gs < eps || gs % eps != 0 || L.K % gs != 0) return false;
  if (L.scales_numel != (L.K / gs) * L.N || L.s_stride1 != 1 || L.s_stride0 != L.N) return false;
  if (L.W_group_mode != 2 && (L.z_stride1 != 1 || L.z_stride0 != L.N)) return false;
  if (L.N % PV_COLS != 0) return false;
  if (pv_smem(L, M <= 1 ? 1 : (M <= 2 ? 2 : (M <= 4 ? 4 : 8))) > 160 * 1024) return false;
  return true;
}

int launch_gemv_packed(const glb_linear& L, const void* x, void* y, int64_t M, const void* bias, cudaStream_t stream) {
  GLB_REQUIRE(gemv_packed_supported(L, M), GLB_ERR_UNSUPPORTED, "packed-width decode kernel: unsupported layer configuration");
  PvParams P;
  P.W = reinterpret_cast<const uint32_t*>(L.W_q); P.scales = L.scales; P.zeros = (L.W_group_mode == 2) ? nullptr : L.zeros;
  P.x = x; P.y = y; P.bias = bias; P.M = (int)M; P.N = (int)L.N; P.K = (int)L.K; P.group_size = L.group_size; P.mode = L.W_group_mode;
  P.ks = pv_ks(L);
  const bool bf = (L.input_dtype == GLB_BF16);
  switch (L.W_nbits) {
    case 8: set_kernel_name("gemv_w8_packed"); return bf ? dispatch_pv<true, 8>(L, P, stream) : dispatch_pv<false, 8>(L, P, stream);
    case 2: set_kernel_name("gemv_w2_packed"); return bf ? dispatch_pv<true, 2>(L, P, stream) : dispatch_pv<false, 2>(L, P, stream);
    default: set_kernel_name("gemv_w1_packed"); return bf ? dispatch_pv<true, 1>(L, P, stream) : dispatch_pv<false, 1>(L, P, stream);
  }
}

}  // namespace glb
