/* TEST / BASELINE INFRASTRUCTURE -- not part of the product (only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build or call this file).
 *
 * CPU restatement of the reference's A16W4 (W_group_mode 4) decode path for bench.py's CPU baseline:
 *     y[m, n] = sum_k x[m, k] * fma(q[k, n], s[k / gs, n], z'[k / gs, n])
 * reference: gemlite/triton_kernels/utils.py:57-89 (dequantize, mode 4 = fma), bitpack.py:42-54 (word (r, n) of the
 * (K/8, N) int32 matrix holds the codes k = 8r .. 8r+7 of column n in nibbles 0..7), core.py:433-436 (z' = -z*s).
 * Arithmetic: fp32; the group-wise factorisation  s * sum(x q) + z' * sum(x)  (same as oracle.cpu_dequant_matmul_fast,
 * which the tests compare it with).  OpenMP over column blocks: each thread streams its columns of every packed row,
 * so the weights are read exactly once per call.
 *
 * build: gcc -O3 -march=native -fopenmp -shared -fPIC cpu_a16w4.c -o _build/libcpu_a16w4.so   (oracle/build_cpu.py)
 */
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CB 64 /* columns per task: 64 words = one 256-byte run of every packed row */

int a16w4_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Wp (Kp, N) int32, s / zf (G, N) fp32 (converted once at load time), x (M, K) fp32, y (M, N) fp32; K = 8 Kp, N % CB == 0 */
void a16w4_forward_f32(const int32_t* Wp, const float* s, const float* zf, const float* x, float* y, int M, int Kp, int N, int G) {
  const int rpg = Kp / G; /* packed rows per group */
#pragma omp parallel for schedule(static)
  for (int n0 = 0; n0 < N; n0 += CB) {
    for (int m = 0; m < M; ++m) {
      const float* xm = x + (size_t)m * Kp * 8;
      float out[CB];
      memset(out, 0, sizeof(out));
      for (int g = 0; g < G; ++g) {
        float acc[CB];
        memset(acc, 0, sizeof(acc));
        float sx = 0.f;
        for (int r = g * rpg; r < (g + 1) * rpg; ++r) {
          const uint32_t* w = (const uint32_t*)(Wp + (size_t)r * N + n0);
          const float* xr = xm + (size_t)r * 8;
          const float x0 = xr[0], x1 = xr[1], x2 = xr[2], x3 = xr[3], x4 = xr[4], x5 = xr[5], x6 = xr[6], x7 = xr[7];
          sx += ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
          for (int c = 0; c < CB; ++c) {
            const uint32_t q = w[c];
            acc[c] += x0 * (float)(q & 15u) + x1 * (float)((q >> 4) & 15u) + x2 * (float)((q >> 8) & 15u) +
                      x3 * (float)((q >> 12) & 15u) + x4 * (float)((q >> 16) & 15u) + x5 * (float)((q >> 20) & 15u) +
                      x6 * (float)((q >> 24) & 15u) + x7 * (float)(q >> 28);
          }
        }
        const float* sg = s + (size_t)g * N + n0;
        const float* zg = zf + (size_t)g * N + n0;
        for (int c = 0; c < CB; ++c) out[c] += sg[c] * acc[c] + zg[c] * sx;
      }
      memcpy(y + (size_t)m * N + n0, out, sizeof(out));
    }
  }
}
