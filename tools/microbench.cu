// Micro-benchmarks that decide the decode kernel's inner loop (run under gpurun, results -> profiles/):
//   * issue rate of FHFMA (fma.rn.f32.f16), HFMA2, FFMA, LOP3 and legacy HMMA.16816.F32 per SM
//   * the per-launch floor of a chain of programmatic-dependent launches replayed from a CUDA graph
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITERS = 2048;

__device__ __forceinline__ float fhfma(uint32_t a, uint32_t b, float c, bool hi) {
  unsigned short ha = hi ? (unsigned short)(a >> 16) : (unsigned short)(a & 0xffff);
  unsigned short hb = hi ? (unsigned short)(b >> 16) : (unsigned short)(b & 0xffff);
  asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(c) : "h"(ha), "h"(hb));
  return c;
}

// mode 0: FHFMA x8 chains, 1: HFMA2 x8, 2: FFMA x8, 3: LOP3 x8, 4: decode-like mix (4 lop3 + 8 fhfma), 5: HMMA (CH chains)
template <int MODE, int CH>
__global__ void rate_kernel(uint32_t* out, long long* cycles, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + threadIdx.x;
  float f[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = (float)i; h[i] = 0x3c003c00u + i; }
  float d[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = 0.f;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fhfma(a, b, f[i], i & 1);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __half2 r = __hfma2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b), *reinterpret_cast<__half2*>(&h[i]));
        h[i] = *reinterpret_cast<uint32_t*>(&r);
      }
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], __uint_as_float(a), __uint_as_float(b));
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(h[i]) : "r"(a), "r"(b));
    } else if (MODE == 4) {
      // one packed word: shift + 4 lop3 + 8 fhfma (what the decode loop does per 8 weights)
      uint32_t q = a ^ (uint32_t)it, t = q >> 8, m0, m1, m2, m3;
      asm volatile("lop3.b32 %0, %1, 0x000F000F, 0x64006400, 0xEA;" : "=r"(m0) : "r"(q));
      asm volatile("lop3.b32 %0, %1, 0x00F000F0, 0x64006400, 0xEA;" : "=r"(m1) : "r"(q));
      asm volatile("lop3.b32 %0, %1, 0x000F000F, 0x64006400, 0xEA;" : "=r"(m2) : "r"(t));
      asm volatile("lop3.b32 %0, %1, 0x00F000F0, 0x64006400, 0xEA;" : "=r"(m3) : "r"(t));
      f[0] = fhfma(m0, b, f[0], false); f[1] = fhfma(m0, b, f[1], true);
      f[2] = fhfma(m1, b, f[2], false); f[3] = fhfma(m1, b, f[3], true);
      f[4] = fhfma(m2, b, f[4], false); f[5] = fhfma(m2, b, f[5], true);
      f[6] = fhfma(m3, b, f[6], false); f[7] = fhfma(m3, b, f[7], true);
    } else {
#pragma unroll
      for (int c = 0; c < CH; ++c)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                     : "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(a), "r"(b));
    }
  }
  long long t1 = clock64();
  __syncthreads();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= __float_as_uint(f[i]) ^ h[i];
#pragma unroll
  for (int c = 0; c < 4; ++c) acc ^= __float_as_uint(d[c][0] + d[c][1] + d[c][2] + d[c][3]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int CH>
int run_rate(const char* name, int warps, int per_iter_ops, uint32_t* out, long long* cyc) {
  rate_kernel<MODE, CH><<<148, warps * 32>>>(out, cyc, 12345u);
  CK(cudaDeviceSynchronize());
  rate_kernel<MODE, CH><<<148, warps * 32>>>(out, cyc, 12345u);
  CK(cudaDeviceSynchronize());
  long long c[148];
  CK(cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += (double)c[i];
  avg /= 148;
  const double warp_instrs = (double)warps * ITERS * per_iter_ops;
  printf("%-28s warps=%2d  cycles=%9.0f  warp-instr/clk/SM=%6.3f  (clk per warp-instr per SMSP=%6.2f)\n", name, warps, avg,
         warp_instrs / avg, avg / (warp_instrs / 4.0));
  return 0;
}

// ---------------------------------------------------------------- PDL chain floor
__global__ void chain_kernel(const float* in, float* outp, int trigger_early, int work) {
  extern __shared__ float sm[];
  if (trigger_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float v = in[threadIdx.x & 31];
  for (int i = 0; i < work; ++i) v = v * 1.0001f + 0.5f;
  if (!trigger_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x < 32) outp[blockIdx.x * 32 + threadIdx.x] = v + (float)blockIdx.x;
  if (threadIdx.x == 1000000) sm[0] = v;
}

int run_chain(int grid, int threads, int smem, int pdl, int trigger_early, int layers, float* buf) {
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  CK(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  auto enqueue = [&]() -> int {
    for (int l = 0; l < layers; ++l) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
      const float* in = buf + (size_t)(l & 1) * 8192;
      float* o = buf + (size_t)((l + 1) & 1) * 8192;
      CK(cudaLaunchKernelEx(&cfg, chain_kernel, in, o, trigger_early, 8));
    }
    return 0;
  };
  if (enqueue()) return 1;
  CK(cudaStreamSynchronize(st));
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  if (enqueue()) return 1;
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  for (int i = 0; i < 5; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaStreamSynchronize(st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int reps = 50;
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("chain grid=%3d threads=%4d smem=%6d pdl=%d early=%d : %.3f us/launch\n", grid, threads, smem, pdl, trigger_early,
         ms * 1e3 / (reps * layers));
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(st);
  return 0;
}

int main() {
  uint32_t* out; long long* cyc; float* buf;
  CK(cudaMalloc(&out, 148 * 1024 * 4));
  CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaMalloc(&buf, 2 * 8192 * 4));
  CK(cudaMemset(buf, 0, 2 * 8192 * 4));
  for (int w : {4, 8, 16}) {
    if (run_rate<0, 1>("FHFMA (fma.rn.f32.f16) x8", w, 8, out, cyc)) return 1;
    if (run_rate<1, 1>("HFMA2 x8", w, 8, out, cyc)) return 1;
    if (run_rate<2, 1>("FFMA x8", w, 8, out, cyc)) return 1;
    if (run_rate<3, 1>("LOP3 x8", w, 8, out, cyc)) return 1;
    if (run_rate<4, 1>("decode mix 5 alu + 8 fhfma", w, 13, out, cyc)) return 1;
    if (run_rate<5, 1>("HMMA.16816.F32 1 chain", w, 1, out, cyc)) return 1;
    if (run_rate<5, 2>("HMMA.16816.F32 2 chains", w, 2, out, cyc)) return 1;
    if (run_rate<5, 4>("HMMA.16816.F32 4 chains", w, 4, out, cyc)) return 1;
  }
  for (int pdl : {1, 0})
    for (int early : {1, 0}) {
      if (run_chain(128, 512, 80 * 1024, pdl, early, 64, buf)) return 1;
      if (run_chain(128, 128, 0, pdl, early, 64, buf)) return 1;
      if (run_chain(148, 256, 100 * 1024, pdl, early, 64, buf)) return 1;
    }
  printf("done\n");
  return 0;
}
