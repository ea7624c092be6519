// dmma_probe.cu -- checks the fragment layout of mma.sync.m8n8k4.f64 on this device: C (8x8) = A (8xK) * B (Kx8), K = 96,
// one warp, against a host loop.  nvcc -gencode arch=compute_100a,code=sm_100a -o dmma_probe scripts/micro/dmma_probe.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const double* A, const double* B, double* C, int K, long long* clk) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double c0 = 0.0, c1 = 0.0;
  const long long t0 = clock64();
  for (int j = 0; j < K; j += 4) {
    const double a = A[g * K + j + t];        // A[row g][col j + t]
    const double b = B[(j + t) * 8 + g];      // B[row j + t][col g]
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  C[g * 8 + 2 * t] = c0; C[g * 8 + 2 * t + 1] = c1;
  if (lane == 0) *clk = t1 - t0;
}
// dependent and independent DMMA chains on registers only (latency / issue interval of DMMA.8x8x4)
__global__ void lat(double* out, long long* clk, int n) {
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x, c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0, g0 = 0.0, g1 = 0.0;
  long long t0 = clock64();
  for (int i = 0; i < n; i++)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  long long t1 = clock64();
  for (int i = 0; i < n; i++) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e0), "+d"(e1) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(f0), "+d"(f1) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(g0), "+d"(g1) : "d"(a), "d"(b));
  }
  long long t2 = clock64();
  out[threadIdx.x] = c0 + c1 + e0 + e1 + f0 + f1 + g0 + g1;
  if (threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; }
}
int main() {
  const int K = 96;
  double hA[8 * K], hB[K * 8], hC[64], ref[64];
  for (int i = 0; i < 8 * K; i++) { hA[i] = sin(0.37 * i) + 0.1; hB[i] = cos(0.11 * i) - 0.2; }
  for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) { double s = 0; for (int q = 0; q < K; q++) s += hA[r * K + q] * hB[q * 8 + c]; ref[r * 8 + c] = s; }
  double *dA, *dB, *dC; long long* dclk; long long hclk;
  cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dC, sizeof hC); cudaMalloc(&dclk, 8);
  cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
  k<<<1, 32>>>(dA, dB, dC, K, dclk); k<<<1, 32>>>(dA, dB, dC, K, dclk);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost); cudaMemcpy(&hclk, dclk, 8, cudaMemcpyDeviceToHost);
  double worst = 0; for (int i = 0; i < 64; i++) worst = fmax(worst, fabs(hC[i] - ref[i]));
  printf("cuda: %s; max |C - ref| = %.3e (|ref| up to %.2f); 24 dependent DMMA + loads: %lld cycles\n", cudaGetErrorString(e), worst, fabs(ref[0]), hclk);
  double* dout; long long* dc2; long long hc2[2]; cudaMalloc(&dout, 32 * 8); cudaMalloc(&dc2, 16);
  const int n = 256;
  lat<<<1, 32>>>(dout, dc2, n); lat<<<1, 32>>>(dout, dc2, n); cudaDeviceSynchronize();
  cudaMemcpy(hc2, dc2, 16, cudaMemcpyDeviceToHost);
  printf("DMMA.8x8x4: dependent chain %.1f cycles each; four independent chains %.1f cycles per DMMA\n", (double)hc2[0] / n, (double)hc2[1] / (4.0 * n));
  return 0;
}
