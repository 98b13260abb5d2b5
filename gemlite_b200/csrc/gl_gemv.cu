// Decode-regime kernels (small M, HBM-bound): fused unpack + dequant + dot.
//
// Replaces gemlite/triton_kernels/gemv_revsplitK_kernels.py (packed W_nbits < 8, the reference's
// default at M == 1, core.py:105) and serves forward_manual("GEMV"/"GEMV_REVSPLITK").
//
// Layout consumed (unchanged reference layout, core.py:387-392): W_q int32 (K/8, N), N contiguous,
// word (r, n) holds codes k = 8r .. 8r+7 of column n in nibbles 0..7; scales / zeros (K/gs, N).
//
// Design for B200 (148 SMs, ~6.5 TB/s measured => ~23 B/clk/SM => ~46 four-bit weights/clk/SM):
//  * one CTA per 32-column slab, full K, 8 warps; a warp-load is 4 packed rows x 128 contiguous
//    bytes (uint4 per lane, ld.global.nc.L1::no_allocate); every lane issues ALL of its weight
//    loads (16 x 128-bit for K=4096) before it consumes anything, so the whole slab (64 KB) is in
//    flight at once -- the bandwidth-delay product of the chip is about the size of the matrix;
//  * weights do not depend on the previous layer: they are requested BEFORE griddepcontrol.wait,
//    x (which does) after it => with programmatic dependent launch the next layer's HBM traffic
//    overlaps this layer's tail;
//  * the multiply-accumulate is issue-bound on CUDA cores at this bandwidth (>= 1.9 issue slots per
//    weight for lop3 + sub + fma + flush), so the dot products run on the tensor pipe instead:
//    nibbles -> fp16 integers by the 0x6400 magic (1 shift + 4 lop3 per 8 weights), exact
//    (h - 1024) / (h/16 - 64), then mma.sync.m16n8k16 with fp32 accumulation; the k order inside a
//    word is permuted consistently on the x side, so no shuffles are needed;
//  * per group g:  y += s_g * sum_k x_k q_k  +  z'_g * sum_k x_k   (mode 4; modes 2/3 analogous) --
//    the group-wise factorisation of sum_k x_k * fma(q_k, s_g, z'_g), applied in fp32;
//  * deterministic: no atomics, no memset; cross-warp reduction through shared memory; bias fused.
// The same kernel handles 1 <= M <= 16 rows of x (rows of the MMA A operand).
#include "gl_common.cuh"

namespace glb {
namespace {

constexpr int GV_THREADS = 256;
constexpr int GV_WARPS = 8;
constexpr int GV_COLS = 32;        // columns per CTA
constexpr int GV_MAXR = 16;        // packed rows (uint4 loads) a lane keeps in flight per batch

struct GemvParams {
  const uint32_t* W;     // (K/8, N) int32
  const void* scales;    // (G, N)
  const void* zeros;     // (G, N) or null
  const void* x;         // (M, K)
  void* y;               // (M, N)
  const void* bias;      // (N) or null
  int K, N, M;
  int group_size;        // multiple of 32
  int mode;              // 2, 3, 4
  int rows_total;        // K / 8
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

// 8 nibbles of one word -> 4 registers of exact small integers as fp16x2 / bf16x2.
// Pairing: h[0] = (k0,k4), h[1] = (k1,k5), h[2] = (k2,k6), h[3] = (k3,k7).
template <bool BF16> __device__ __forceinline__ void nibbles_to_f16x2(uint32_t q, uint32_t (&h)[4]) {
  if (!BF16) {
    const uint32_t t = q >> 8;
    const uint32_t h0 = lop3_and_or(q, 0x000F000Fu, 0x64006400u);   // 1024 + n
    const uint32_t h1 = lop3_and_or(q, 0x00F000F0u, 0x64006400u);   // 1024 + 16 n
    const uint32_t h2 = lop3_and_or(t, 0x000F000Fu, 0x64006400u);
    const uint32_t h3 = lop3_and_or(t, 0x00F000F0u, 0x64006400u);
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

template <bool BF16> __device__ __forceinline__ float ld_meta(const void* p, int64_t i) {
  if (BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}
template <bool BF16> __device__ __forceinline__ void st_out(void* p, int64_t i, float v) {
  if (BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}
template <bool BF16> __device__ __forceinline__ float rnd_out(float v) {
  if (BF16) return __bfloat162float(__float2bfloat16_rn(v));
  return __half2float(__float2half_rn(v));
}

// Shared memory: xs[MT][K] permuted 16-bit activations | sx[MT][G] fp32 group sums | red[8][MT][32] fp32
template <bool BF16, int MT>
__global__ void __launch_bounds__(GV_THREADS) gemv_w4_mma_kernel(const GemvParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = P.K, N = P.N, G = K / P.group_size;
  uint16_t* xs = reinterpret_cast<uint16_t*>(smem_raw);
  float* sx = reinterpret_cast<float*>(smem_raw + (size_t)MT * K * 2);
  float* red = sx + (size_t)MT * G;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g8 = lane >> 2, t4 = lane & 3;        // mma lane coordinates
  const int col0 = blockIdx.x * GV_COLS;
  const int rpg = P.group_size >> 3;               // packed rows per group (multiple of 4)
  const int runs_per_group = rpg >> 2;             // 4-row runs per group

  // ---- work split: warp w owns groups w, w+8, ...; inside a group the warp walks runs of 4 rows;
  //      lane (g8, t4) loads row (run*4 + t4), columns col0 + 4*g8 .. +3
  // Issue every weight load of the first batch before anything else.
  const uint32_t* wbase = P.W + (size_t)col0 + 4 * g8;
  uint4 q[GV_MAXR];
  int nq = 0;                       // loads held
  int grp = warp, run = 0;          // cursor of the next (group, run) to load
  auto issue_batch = [&]() {
    nq = 0;
#pragma unroll
    for (int i = 0; i < GV_MAXR; ++i) {
      if (grp < G) {
        const int row = grp * rpg + run * 4 + t4;
        q[i] = ldg_nc_v4(wbase + (size_t)row * N);
        ++nq;
        if (++run == runs_per_group) { run = 0; grp += GV_WARPS; }
      }
    }
  };
  issue_batch();

  // ---- x depends on the previous kernel: wait for it only now (PDL), then stage it permuted ----
  pdl_wait_prior_grids();
  {
    // word r of row m: source k = 8r + j -> slot order [0,4,1,5,2,6,3,7] so that an LDS.128 yields
    // {(k0,k4), (k1,k5), (k2,k6), (k3,k7)} = the A-fragment pairs matching nibbles_to_f16x2
    const uint16_t* xg = reinterpret_cast<const uint16_t*>(P.x);
    const int words = K >> 3;
    for (int i = threadIdx.x; i < MT * words; i += GV_THREADS) {
      const int m = i / words, r = i - m * words;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (m < P.M) {
        const uint4 v = *reinterpret_cast<const uint4*>(xg + (size_t)m * K + (size_t)r * 8);
        // v.x = (k0,k1) v.y = (k2,k3) v.z = (k4,k5) v.w = (k6,k7)
        o.x = __byte_perm(v.x, v.z, 0x5410);   // (k0,k4)
        o.y = __byte_perm(v.x, v.z, 0x7632);   // (k1,k5)
        o.z = __byte_perm(v.y, v.w, 0x5410);   // (k2,k6)
        o.w = __byte_perm(v.y, v.w, 0x7632);   // (k3,k7)
      }
      *reinterpret_cast<uint4*>(xs + (size_t)m * K + (size_t)r * 8) = o;
    }
  }
  __syncthreads();
  if (P.mode != 2) {   // group sums of x (fp32) for the zero-point term
    for (int i = warp; i < MT * G; i += GV_WARPS) {
      const int m = i / G, g = i - m * G;
      float s = 0.f;
      for (int k = lane; k < P.group_size; k += 32) {
        const uint16_t raw = xs[(size_t)m * K + (size_t)g * P.group_size + k];
        s += BF16 ? __bfloat162float(__ushort_as_bfloat16(raw)) : __half2float(__ushort_as_half(raw));
      }
      s = warp_sum(s);
      if (lane == 0) sx[i] = s;
    }
    __syncthreads();
  }
  pdl_launch_dependents();

  // ---- main loop ----
  // out[c][0..3]: D fragment of column set c (columns col0 + 4*n' + c, n' = mma n index)
  float out[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { out[c][0] = out[c][1] = out[c][2] = out[c][3] = 0.f; }
  float dacc[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { dacc[c][0] = dacc[c][1] = dacc[c][2] = dacc[c][3] = 0.f; }

  int cgrp = warp, crun = 0;        // cursor of the (group, run) being consumed
  const bool a_lo = (g8 < MT), a_hi = (g8 + 8 < MT);
  while (nq > 0) {
#pragma unroll
    for (int i = 0; i < GV_MAXR; ++i) {
      if (i < nq) {
        const int row = cgrp * rpg + crun * 4 + t4;
        // A fragments: rows g8 / g8+8 of x, k-slots of packed row `row`
        uint32_t a1[4] = {0u, 0u, 0u, 0u}, a2[4] = {0u, 0u, 0u, 0u};
        if (a_lo) {
          const uint4 v = *reinterpret_cast<const uint4*>(xs + (size_t)g8 * K + (size_t)row * 8);
          a1[0] = v.x; a1[2] = v.y; a2[0] = v.z; a2[2] = v.w;
        }
        if (MT > 8 && a_hi) {
          const uint4 v = *reinterpret_cast<const uint4*>(xs + (size_t)(g8 + 8) * K + (size_t)row * 8);
          a1[1] = v.x; a1[3] = v.y; a2[1] = v.z; a2[3] = v.w;
        }
        const uint32_t w4[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t h[4];
          nibbles_to_f16x2<BF16>(w4[c], h);
          mma16816<BF16>(dacc[c], a1, h[0], h[1]);
          mma16816<BF16>(dacc[c], a2, h[2], h[3]);
        }
        if (++crun == runs_per_group) {
          // ---- end of group: apply scale / zero in fp32 ----
          // this lane holds rows (g8, g8+8) x n' = (2*t4, 2*t4+1) -> columns col0 + 8*t4 + {0..7}
          const int64_t mo = (int64_t)cgrp * N + col0 + 8 * t4;
          float sc[8], zc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            sc[j] = ld_meta<BF16>(P.scales, mo + j);
            zc[j] = (P.mode != 2) ? ld_meta<BF16>(P.zeros, mo + j) : 0.f;
          }
          const float sx_lo = (P.mode != 2 && a_lo) ? sx[g8 * G + cgrp] : 0.f;
          const float sx_hi = (P.mode != 2 && MT > 8 && a_hi) ? sx[(g8 + 8) * G + cgrp] : 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            // dacc[c][0]: (row g8, n'=2t4) -> column 8*t4 + c ; [1]: n'=2t4+1 -> column 8*t4 + 4 + c
            const float s0 = sc[c], s1 = sc[4 + c], z0 = zc[c], z1 = zc[4 + c];
            if (P.mode == 4) {
              out[c][0] += s0 * dacc[c][0] + z0 * sx_lo;  out[c][1] += s1 * dacc[c][1] + z1 * sx_lo;
              out[c][2] += s0 * dacc[c][2] + z0 * sx_hi;  out[c][3] += s1 * dacc[c][3] + z1 * sx_hi;
            } else if (P.mode == 3) {
              out[c][0] += s0 * (dacc[c][0] - z0 * sx_lo);  out[c][1] += s1 * (dacc[c][1] - z1 * sx_lo);
              out[c][2] += s0 * (dacc[c][2] - z0 * sx_hi);  out[c][3] += s1 * (dacc[c][3] - z1 * sx_hi);
            } else {
              out[c][0] += s0 * dacc[c][0];  out[c][1] += s1 * dacc[c][1];
              out[c][2] += s0 * dacc[c][2];  out[c][3] += s1 * dacc[c][3];
            }
            dacc[c][0] = dacc[c][1] = dacc[c][2] = dacc[c][3] = 0.f;
          }
          crun = 0; cgrp += GV_WARPS;
        }
      }
    }
    issue_batch();   // K larger than one batch: next set of loads
  }

  // ---- cross-warp reduction + epilogue ----
  // red[warp][m][col]: lane writes rows g8 (+8), columns 8*t4 + c and 8*t4 + 4 + c
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (g8 < MT) {
      red[(warp * MT + g8) * GV_COLS + 8 * t4 + c] = out[c][0];
      red[(warp * MT + g8) * GV_COLS + 8 * t4 + 4 + c] = out[c][1];
    }
    if (MT > 8 && g8 + 8 < MT) {
      red[(warp * MT + g8 + 8) * GV_COLS + 8 * t4 + c] = out[c][2];
      red[(warp * MT + g8 + 8) * GV_COLS + 8 * t4 + 4 + c] = out[c][3];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MT * GV_COLS; i += GV_THREADS) {
    const int m = i / GV_COLS, c = i - m * GV_COLS;
    if (m < P.M && col0 + c < N) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < GV_WARPS; ++w) v += red[(w * MT + m) * GV_COLS + c];
      if (P.bias) v = rnd_out<BF16>(v) + ld_meta<BF16>(P.bias, col0 + c);
      st_out<BF16>(P.y, (int64_t)m * N + col0 + c, v);
    }
  }
}

template <bool BF16, int MT>
int launch_w4(const GemvParams& P, cudaStream_t stream) {
  const int G = P.K / P.group_size;
  const size_t smem = (size_t)MT * P.K * 2 + (size_t)MT * G * 4 + (size_t)GV_WARPS * MT * GV_COLS * 4;
  auto kern = gemv_w4_mma_kernel<BF16, MT>;
  if (smem > 48 * 1024) {
    static thread_local size_t configured = 0;
    if (configured < smem) {
      GLB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(P.N / GV_COLS));
  cfg.blockDim = dim3(GV_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_enabled.load() ? 1 : 0;
  GLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  count_launch();
  return GLB_OK;
}

}  // namespace

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
  if (L.group_size < 32 || L.group_size % 32 != 0 || L.K % L.group_size != 0) return false;
  if (L.scales_numel != (L.K / L.group_size) * L.N || L.s_stride1 != 1 || L.s_stride0 != L.N) return false;
  if (L.W_group_mode != 2 && (L.z_stride1 != 1 || L.z_stride0 != L.N)) return false;
  if (L.N % GV_COLS != 0 || L.K % 8 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(L.W_q) & 15) || (reinterpret_cast<uintptr_t>(L.scales) & 1)) return false;
  const int mt = (M == 1) ? 1 : (M <= 8 ? 8 : 16);
  const size_t smem = (size_t)mt * L.K * 2 + (size_t)mt * (L.K / L.group_size) * 4 + (size_t)GV_WARPS * mt * GV_COLS * 4;
  if (smem > 200 * 1024) return false;
  return true;
}

int launch_gemv(const glb_linear& L, const void* x, const float* scales_x, void* y, int64_t M, const void* bias,
                cudaStream_t stream) {
  (void)scales_x;
  GLB_REQUIRE(gemv_supported(L, M), GLB_ERR_UNSUPPORTED, "decode kernel: unsupported layer configuration");
  GLB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, GLB_ERR_INVALID, "decode kernel: x must be 16-byte aligned");
  GemvParams P;
  P.W = reinterpret_cast<const uint32_t*>(L.W_q);
  P.scales = L.scales; P.zeros = (L.W_group_mode == 2) ? nullptr : L.zeros;
  P.x = x; P.y = y; P.bias = bias;
  P.K = (int)L.K; P.N = (int)L.N; P.M = (int)M;
  P.group_size = L.group_size; P.mode = L.W_group_mode; P.rows_total = (int)(L.K / 8);
  const bool bf = (L.input_dtype == GLB_BF16);
  set_kernel_name("gemv_w4_mma");
  if (M == 1) return bf ? launch_w4<true, 1>(P, stream) : launch_w4<false, 1>(P, stream);
  if (M <= 8) return bf ? launch_w4<true, 8>(P, stream) : launch_w4<false, 8>(P, stream);
  return bf ? launch_w4<true, 16>(P, stream) : launch_w4<false, 16>(P, stream);
}

}  // namespace glb
