// Ceiling of the tcgen05.mma issue forms used by the prefill kernels: back-to-back kind::f16 MMAs on garbage operands,
// one CTA (pair) per SM (pair), cycles per instruction measured with clock64 around issue + commit + wait.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gemlite_b200/csrc -o tools/mma_bench tools/mma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "gl_ptx.cuh"
using namespace glb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void mma_f16_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
               "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void mma_i8_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
               "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f8_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
               "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

// MODE 0: 2-CTA TS (A in TMEM)  1: 2-CTA SS  2: 1-CTA SS  3: 1-CTA TS  4: 2-CTA SS kind::i8  5: 2-CTA SS kind::f8f6f4 (e4m3)
// N = accumulator width; modes 4 / 5 use K = 32 bytes per instruction (same 128 ideal cycles at N = 256)
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  constexpr bool TWO = (MODE == 0 || MODE == 1 || MODE == 4 || MODE == 5);
  const uint32_t rank = TWO ? ptx::cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) { if (TWO) ptx::tmem_alloc_2cta(&slot, 512); else ptx::tmem_alloc(&slot, 512); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  if (TWO) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = (MODE == 4) ? ptx::make_idesc(2u, 1u, 1u, 256, N) : ptx::make_idesc(1u, 0u, 0u, TWO ? 256 : 128, N);
    const uint64_t ad = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem));
    const uint64_t bd = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + 32 * 1024));
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (MODE == 0) ptx::mma_f16_ts_2cta(tm, tm + 256 + 8 * k, bd + 2 * k, idesc, 1u);
        else if (MODE == 1) mma_f16_ss_2cta(tm, ad + 2 * k, bd + 2 * k, idesc, 1u);
        else if (MODE == 2) ptx::mma_f16_ss(tm, ad + 2 * k, bd + 2 * k, idesc, 1u);
        else if (MODE == 4) mma_i8_ss_2cta(tm, ad + 2 * k, bd + 2 * k, idesc, 1u);
        else if (MODE == 5) mma_f8_ss_2cta(tm, ad + 2 * k, bd + 2 * k, idesc, 1u);
        else ptx::mma_f16_ts(tm, tm + 256 + 8 * k, bd + 2 * k, idesc, 1u);
      }
    }
    if (TWO) ptx::mma_commit_2cta(&bar, (uint16_t)1); else ptx::mma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before_sync();
  if (TWO) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after_sync();
  if (threadIdx.x < 32) { if (TWO) ptx::tmem_dealloc_2cta(tm, 512); else ptx::tmem_dealloc(tm, 512); }
}

template <int MODE, int N> int run(const char* name, long long* out) {
  constexpr bool TWO = (MODE == 0 || MODE == 1 || MODE == 4 || MODE == 5);
  auto k = bench<MODE, N>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 96 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = TWO ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, k, out, iters));
    CK(cudaDeviceSynchronize());
  }
  long long c[148];
  CK(cudaMemcpy(c, out, sizeof(c), cudaMemcpyDeviceToHost));
  double avg = 0; int n = 0;
  for (int i = 0; i < 148; i += (TWO ? 2 : 1)) { avg += (double)c[i]; ++n; }
  avg /= n;
  const double per = avg / (iters * 4.0);
  const double ideal = (double)N / 2.0;   // 128 rows per CTA x N x 16 MACs / 4096 MACs per clk per SM
  printf("%-28s N=%3d  cycles/MMA = %7.1f  (ideal %5.1f -> %5.1f %% of the tensor peak)\n", name, N, per, ideal, 100.0 * ideal / per);
  return 0;
}

int main() {
  long long* out;
  CK(cudaMalloc(&out, 148 * 8));
  CK(cudaMemset(out, 0, 148 * 8));
  if (run<0, 256>("2-CTA TS (A in TMEM)", out)) return 1;
  if (run<1, 256>("2-CTA SS", out)) return 1;
  if (run<2, 256>("1-CTA SS", out)) return 1;
  if (run<3, 256>("1-CTA TS", out)) return 1;
  if (run<4, 256>("2-CTA SS kind::i8 (K=32)", out)) return 1;
  if (run<5, 256>("2-CTA SS kind::f8f6f4 (K=32)", out)) return 1;
  if (run<0, 128>("2-CTA TS (A in TMEM)", out)) return 1;
  if (run<1, 128>("2-CTA SS", out)) return 1;
  if (run<2, 128>("1-CTA SS", out)) return 1;
  printf("done\n");
  return 0;
}
