// Dependent-issue latencies that bound the QP kernel's sweeps on sm_100a (one warp, clock64): DFMA chain, DFMA with 2 / 4
// independent chains, LDS.64 pointer chase, SHFL chain, __syncwarp, DADD chain, MUFU.RCP64H + Newton.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/micro/lat scripts/micro/lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void k(long long* out, double* sink, double a, double b) {
  __shared__ double sm[256];
  __shared__ int nxt[256];
  const int l = threadIdx.x;
  for (int i = l; i < 256; i += 32) { sm[i] = 1.0 + i * 1e-9; nxt[i] = (i * 8 + 8) & 2047 ? ((i + 1) & 255) : 0; }
  __syncwarp();
  double x = a + l, y = a + 2 * l, z = a - l, w = a * l;
  long long t0, t1;
  // 1 chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = fma(x, a, b);
  t1 = clock64(); if (l == 0) out[0] = t1 - t0;
  // 2 chains
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; i++) { x = fma(x, a, b); y = fma(y, a, b); }
  t1 = clock64(); if (l == 0) out[1] = t1 - t0;
  // 4 chains
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) { x = fma(x, a, b); y = fma(y, a, b); z = fma(z, a, b); w = fma(w, a, b); }
  t1 = clock64(); if (l == 0) out[2] = t1 - t0;
  // DADD chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = x + b;
  t1 = clock64(); if (l == 0) out[3] = t1 - t0;
  // LDS pointer chase (int index -> next)
  int p = l;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) p = nxt[p];
  t1 = clock64(); if (l == 0) out[4] = t1 - t0;
  // LDS.64 feeding a DFMA chain (load then fma dependent on the load address through the value)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) { const double v = sm[(int)(x) & 255]; x = fma(v, a, b); }
  t1 = clock64(); if (l == 0) out[5] = t1 - t0;
  // SHFL chain (64-bit = 2 SHFL)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) y = __shfl_xor_sync(0xffffffffu, y, 1) + 1.0;
  t1 = clock64(); if (l == 0) out[6] = t1 - t0;
  // STS + syncwarp + LDS round trip
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) { sm[l] = z; __syncwarp(); z = sm[(l + 1) & 31] + 1.0; __syncwarp(); }
  t1 = clock64(); if (l == 0) out[7] = t1 - t0;
  // reciprocal: rcp.approx + 2 Newton
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; i++) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w)); r = fma(fma(-w, r, 1.0), r, r); r = fma(fma(-w, r, 1.0), r, r); w = r + 1.5; }
  t1 = clock64(); if (l == 0) out[8] = t1 - t0;
  // FSEL/IMAD chain (fp32/int ALU dependent latency)
  int q = l;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) q = q * 3 + i;
  t1 = clock64(); if (l == 0) out[9] = t1 - t0;
  sink[l] = x + y + z + w + p + q;
}
int main() {
  long long* d; double* s; cudaMalloc(&d, 16 * 8); cudaMalloc(&s, 32 * 8);
  k<<<1, 32>>>(d, s, 0.999999, 1e-9); k<<<1, 32>>>(d, s, 0.999999, 1e-9);
  long long h[16]; cudaMemcpy(h, d, 16 * 8, cudaMemcpyDeviceToHost);
  const char* nm[] = {"DFMA chain", "DFMA x2 chains (per pair)", "DFMA x4 chains (per quad)", "DADD chain", "LDS.32 pointer chase", "LDS.64 -> DFMA -> cvt -> address", "SHFL.64 + DADD", "STS + syncwarp + LDS + DADD + syncwarp", "RCP64H + 2 Newton + DADD", "IMAD chain"};
  for (int i = 0; i < 10; i++) printf("%-42s %.2f cycles/iter\n", nm[i], (double)h[i] / N);
  return 0;
}
