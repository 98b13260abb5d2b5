// Probe for the block-scaled tcgen05.mma encodings (kind::mxf8f6f4 / mxf4 / mxf4nvf4 .block_scale) used by gl_gemm_bs.cu:
// one CTA, one 128-feature x N-token tile, K swept stage by stage without pipelining, every open question about the
// scale-factor plumbing (byte select in the instruction descriptor, tensor-memory column per MMA) run as a variant and
// compared with a double-precision product of the decoded operands on the host.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/_build/bs_probe tools/bs_probe.cu
//   tools/_build/bs_probe            (prints max abs / rel error per kind and variant)
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../gemlite_b200/csrc/gl_ptx.cuh"

using namespace glb;

enum { KIND_MXFP8 = 0, KIND_MXFP4 = 1, KIND_NVFP4 = 2 };

struct ProbeParams {
  const uint8_t* sfa;      // [128][K / vs] scale bytes of the A rows (features)
  const uint8_t* sfb;      // [N][K / vs] scale bytes of the B rows (tokens)
  float* d;                // [128][N]
  int stages;              // K bytes per row / 128
  int n;                   // 128 or 256
  int variant;             // bit 0: also put the byte select into the top two address bits; bit 1: alternative pair numbering
};

__device__ __forceinline__ uint64_t make_nosw_kmajor_desc(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(sbo_bytes >> 4) << 16;          // leading-dimension offset (not used by a 16-byte-wide source)
  d |= (uint64_t)(sbo_bytes >> 4) << 32;          // stride between 8-row core matrices
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void utccp_32x128b_warpx4(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// block-scaled instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptorBlockScaled)
__device__ __forceinline__ uint32_t make_idesc_bs(uint32_t a_fmt, uint32_t b_fmt, uint32_t m, uint32_t n, uint32_t scale_e8m0,
                                                  uint32_t a_sf_id, uint32_t b_sf_id) {
  return (b_sf_id << 4) | (a_fmt << 7) | (b_fmt << 10) | ((n >> 3) << 17) | (scale_e8m0 << 23) | ((m >> 4) << 24) | (a_sf_id << 29);
}
template <int KIND>
__device__ __forceinline__ void mma_bs(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t sfa, uint32_t sfb,
                                       uint32_t accumulate) {
  if (KIND == KIND_MXFP8) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
        : "memory");
  } else if (KIND == KIND_MXFP4) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4nvf4.block_scale.scale_vec::4X [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
        : "memory");
  }
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) bs_probe_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                          const __grid_constant__ CUtensorMap tmap_b, const ProbeParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int SF_PER_STAGE = (KIND == KIND_MXFP8) ? 4 : ((KIND == KIND_MXFP4) ? 8 : 16);   // scale bytes per row and stage
  constexpr int NBLK = SF_PER_STAGE / 4;                                                      // 32-bit words per row and stage
  unsigned char* sA = smem;                       // 128 x 128 B
  unsigned char* sB = smem + 16384;               // N x 128 B
  unsigned char* sSFA = smem + 16384 + 32768;     // NBLK x 512 B
  unsigned char* sSFB = sSFA + 4 * 512;           // (N / 128) x NBLK x 512 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSFB + 8 * 512);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nsf = P.stages * SF_PER_STAGE;        // scale bytes per row

  if (tid == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_sfa = tmem + 256, t_sfb = tmem + 256 + 16;     // accumulator: columns [0, N); scale words behind it

  for (int st = 0; st < P.stages; ++st) {
    if (tid == 0) {
      ptx::mbar_arrive_expect_tx(&bars[0], 16384u + (uint32_t)P.n * 128u);
      ptx::tma_load_2d(sA, &tmap_a, &bars[0], st * 128, 0);
      ptx::tma_load_2d(sB, &tmap_b, &bars[0], st * 128, 0);
      if (P.n == 256) ptx::tma_load_2d(sB + 16384, &tmap_b, &bars[0], st * 128, 128);
    }
    // scale words of this stage in the 128 x 4 block layout: byte (r % 32) * 16 + (r / 32) * 4 + j
    for (int b = 0; b < NBLK; ++b) {
      uint32_t w = 0;
      for (int j = 0; j < 4; ++j) w |= (uint32_t)P.sfa[(size_t)tid * nsf + st * SF_PER_STAGE + 4 * b + j] << (8 * j);
      *reinterpret_cast<uint32_t*>(sSFA + b * 512 + (tid & 31) * 16 + (tid >> 5) * 4) = w;
      for (int h = 0; h < P.n / 128; ++h) {
        uint32_t v = 0;
        for (int j = 0; j < 4; ++j) v |= (uint32_t)P.sfb[(size_t)(h * 128 + tid) * nsf + st * SF_PER_STAGE + 4 * b + j] << (8 * j);
        *reinterpret_cast<uint32_t*>(sSFB + (h * NBLK + b) * 512 + (tid & 31) * 16 + (tid >> 5) * 4) = v;
      }
    }
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      ptx::mbar_wait(&bars[0], (uint32_t)(st & 1));
      ptx::tc_fence_after_sync();
      for (int b = 0; b < NBLK; ++b) {
        utccp_32x128b_warpx4(t_sfa + 4 * b, make_nosw_kmajor_desc(ptx::smem_u32(sSFA + b * 512), 128));
        for (int h = 0; h < P.n / 128; ++h)     // word b of all tokens: columns [(n/32) b, (n/32)(b+1)), token t in column t / 32
          utccp_32x128b_warpx4(t_sfb + (uint32_t)((P.n / 128) * 4 * b + 4 * h), make_nosw_kmajor_desc(ptx::smem_u32(sSFB + (h * NBLK + b) * 512), 128));
      }
      const uint64_t ad = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sA)), bd = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sB));
      for (int i = 0; i < 4; ++i) {
        uint32_t sel, word;
        if (KIND == KIND_MXFP8) { sel = (uint32_t)i; word = 0; }
        else if (KIND == KIND_MXFP4) { sel = (P.variant & 2) ? (uint32_t)(i & 1) : (uint32_t)(2 * (i & 1)); word = (uint32_t)(i >> 1); }
        else { sel = 0; word = (uint32_t)i; }
        const uint32_t fmt = (KIND == KIND_MXFP8) ? 0u : 1u;                       // E4M3 / (MXF4) E2M1
        const uint32_t idesc = make_idesc_bs(fmt, fmt, 128, (uint32_t)P.n, (KIND == KIND_NVFP4) ? 0u : 1u, sel, sel);
        uint32_t a_sf = t_sfa + 4 * word, b_sf = t_sfb + (uint32_t)(P.n / 128) * 4 * word;
        if (P.variant & 1) { a_sf |= sel << 30; b_sf |= sel << 30; }
        mma_bs<KIND>(tmem, ad + (uint64_t)(2 * i), bd + (uint64_t)(2 * i), idesc, a_sf, b_sf, (st | i) ? 1u : 0u);
      }
      ptx::mma_commit(&bars[1]);
    }
    __syncthreads();
    ptx::mbar_wait(&bars[1], (uint32_t)(st & 1));      // MMAs (and the scale copies before them) done: buffers reusable
    ptx::tc_fence_after_sync();
    __syncthreads();
  }
  for (int c0 = 0; c0 < P.n; c0 += 32) {
    uint32_t r[32];
    ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) P.d[(size_t)tid * P.n + c0 + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------- host
static float e4m3_to_float(uint8_t v) {
  const int s = v >> 7, e = (v >> 3) & 15, m = v & 7;
  float f = (e == 0) ? ldexpf((float)m, -9) : ldexpf(1.f + m / 8.f, e - 7);
  return s ? -f : f;
}
static float e2m1_to_float(uint8_t c) {
  static const float t[8] = {0.f, 0.5f, 1.f, 1.5f, 2.f, 3.f, 4.f, 6.f};
  return (c & 8) ? -t[c & 7] : t[c & 7];
}
static float e8m0_to_float_h(uint8_t e) { return ldexpf(1.f, (int)e - 127); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static uint32_t rng_state = 12345u;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

template <int KIND>
static void run(EncodeTiledFn enc, int n, int kbytes, int variant, const char* label) {
  const int vs = (KIND == KIND_NVFP4) ? 16 : 32;
  const int epb = (KIND == KIND_MXFP8) ? 1 : 2;             // elements per byte
  const int K = kbytes * epb, nsf = K / vs;
  std::vector<uint8_t> A((size_t)128 * kbytes), B((size_t)n * kbytes), SA((size_t)128 * nsf), SB((size_t)n * nsf);
  auto fill = [&](std::vector<uint8_t>& v) {
    for (auto& x : v) {
      x = (uint8_t)rnd();
      if (KIND == KIND_MXFP8 && (x & 0x7f) == 0x7f) x &= 0xf7;      // no NaN codes
    }
  };
  fill(A); fill(B);
  for (auto& x : SA) x = (KIND == KIND_NVFP4) ? (uint8_t)(0x30 + rnd() % 0x18) : (uint8_t)(124 + rnd() % 7);
  for (auto& x : SB) x = (KIND == KIND_NVFP4) ? (uint8_t)(0x30 + rnd() % 0x18) : (uint8_t)(124 + rnd() % 7);
  auto val = [&](const std::vector<uint8_t>& m, int r, int k) -> double {
    if (KIND == KIND_MXFP8) return e4m3_to_float(m[(size_t)r * kbytes + k]);
    const uint8_t b = m[(size_t)r * kbytes + k / 2];
    return e2m1_to_float((k & 1) ? (b >> 4) : (b & 15));
  };
  auto sc = [&](const std::vector<uint8_t>& s, int r, int k) -> double {
    const uint8_t b = s[(size_t)r * nsf + k / vs];
    return (KIND == KIND_NVFP4) ? e4m3_to_float(b) : e8m0_to_float_h(b);
  };
  std::vector<double> ref((size_t)128 * n);
  double refmax = 0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < n; ++c) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += val(A, r, k) * sc(SA, r, k) * val(B, c, k) * sc(SB, c, k);
      ref[(size_t)r * n + c] = acc;
      refmax = fmax(refmax, fabs(acc));
    }
  uint8_t *dA, *dB, *dSA, *dSB;
  float* dD;
  cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dSA, SA.size()); cudaMalloc(&dSB, SB.size());
  cudaMalloc(&dD, (size_t)128 * n * 4);
  cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dSA, SA.data(), SA.size(), cudaMemcpyHostToDevice); cudaMemcpy(dSB, SB.data(), SB.size(), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, (size_t)128 * n * 4);
  CUtensorMap ta, tb;
  auto mk = [&](CUtensorMap* t, void* base, int rows) {
    cuuint64_t dims[2] = {(cuuint64_t)kbytes, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kbytes};
    cuuint32_t box[2] = {128, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(t, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  };
  mk(&ta, dA, 128); mk(&tb, dB, n);
  ProbeParams P;
  P.sfa = dSA; P.sfb = dSB; P.d = dD; P.stages = kbytes / 128; P.n = n; P.variant = variant;
  const int smem_bytes = 16384 + 32768 + 12 * 512 + 256 + 1024;
  cudaFuncSetAttribute(bs_probe_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  bs_probe_kernel<KIND><<<1, 128, smem_bytes>>>(ta, tb, P);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-28s variant %d: CUDA error %s\n", label, variant, cudaGetErrorString(e)); exit(2); }
  std::vector<float> D((size_t)128 * n);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double emax = 0;
  int bad = 0;
  for (size_t i = 0; i < D.size(); ++i) {
    const double d = fabs((double)D[i] - ref[i]);
    if (!(d <= 1e-3 * refmax)) ++bad;
    if (d > emax || d != d) emax = d;
  }
  printf("%-28s n=%3d K=%4d variant %d: max|err| %.4g of max|ref| %.4g  (%d / %zu beyond 1e-3)  %s\n", label, n, K, variant, emax,
         refmax, bad, D.size(), bad == 0 ? "MATCH" : "differs");
  cudaFree(dA); cudaFree(dB); cudaFree(dSA); cudaFree(dSB); cudaFree(dD);
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no driver entry\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)p;
  for (int v = 0; v < 2; ++v) {
    run<KIND_MXFP8>(enc, 128, 512, v, "mxf8f6f4 e4m3 x e4m3 /32");
    run<KIND_MXFP8>(enc, 256, 512, v, "mxf8f6f4 e4m3 x e4m3 /32");
  }
  for (int v = 0; v < 2; ++v) {      // (variant bit 1 -- pair index instead of byte index as the scale id -- is an illegal instruction)
    run<KIND_MXFP4>(enc, 128, 256, v, "mxf4 e2m1 x e2m1 /32");
    run<KIND_MXFP4>(enc, 256, 256, v, "mxf4 e2m1 x e2m1 /32");
  }
  for (int v = 0; v < 2; ++v) {
    run<KIND_NVFP4>(enc, 128, 256, v, "mxf4nvf4 e2m1 x e2m1 /16 ue4m3");
    run<KIND_NVFP4>(enc, 256, 256, v, "mxf4nvf4 e2m1 x e2m1 /16 ue4m3");
  }
  return 0;
}
