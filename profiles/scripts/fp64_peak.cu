// fp64_peak.cu — FP64 FMA throughput of the GPU (the denominator of the FP64 roofline fractions in DESIGN.md / bench.py).
// Each thread runs 8 independent DFMA chains; 148 x 8 CTAs of 256 threads keep every SM's FP64 pipe saturated.
// build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cuda_runtime.h>
#include <cstdio>

__global__ void __launch_bounds__(256) fma_chains(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  fma_chains<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 10; ++rep) {
    cudaEventRecord(e0);
    fma_chains<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double fmas = (double)blocks * threads * iters * 16.0 * 8.0;
  const double tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"fp64_tflops\": %.3f, \"fp64_fma_per_clk_per_sm\": %.2f, \"ms\": %.4f, \"sms\": %d, \"sm_clock_mhz_max\": %.0f, "
         "\"gpu\": \"%s\", \"how\": \"8 independent DFMA chains per thread, %d CTAs x 256 threads, best of 10, CUDA events\"}\n",
         tflops, fmas / (best * 1e-3) / (clk * 1e3) / p.multiProcessorCount, best, p.multiProcessorCount, clk / 1e3, p.name, blocks);
  return 0;
}
