// FP64 DFMA peak and HBM copy bandwidth probe.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double *out, double a, double b, int iters)
{
   double x[8];
   for (int k = 0; k < 8; k++) { x[k] = threadIdx.x * 1e-3 + k; }
   for (int i = 0; i < iters; i++)
   {
#pragma unroll
      for (int k = 0; k < 8; k++) { x[k] = fma(x[k], a, b); }
   }
   double s = 0;
   for (int k = 0; k < 8; k++) { s += x[k]; }
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void copy(const double4 *in, double4 *out, size_t n)
{
   for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { out[i] = in[i]; }
}
int main()
{
   cudaDeviceProp p;
   cudaGetDeviceProperties(&p, 0);
   printf("device %s SMs %d\n", p.name, p.multiProcessorCount);
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   double *out;
   const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
   cudaMalloc(&out, sizeof(double) * blocks * threads);
   float best = 1e30f;
   for (int r = 0; r < 6; r++)
   {
      cudaEventRecord(e0);
      dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r > 0 && ms < best) { best = ms; }
   }
   const double flops = 2.0 * 8 * (double)iters * blocks * threads;
   printf("FP64 DFMA: %.3f ms  %.2f TFLOP/s\n", best, flops / best / 1e9);
   const size_t n = (size_t)1 << 27; // 128 Mi double4 = 4 GiB
   double4 *a, *b;
   cudaMalloc(&a, n * sizeof(double4) / 4); cudaMalloc(&b, n * sizeof(double4) / 4);
   cudaMemset(a, 0, n * sizeof(double4) / 4);
   best = 1e30f;
   for (int r = 0; r < 6; r++)
   {
      cudaEventRecord(e0);
      copy<<<p.multiProcessorCount * 16, 512>>>(a, b, n / 4);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r > 0 && ms < best) { best = ms; }
   }
   printf("HBM copy: %.3f ms  %.1f GB/s (read+write)\n", best, 2.0 * n / 4 * sizeof(double4) / best / 1e6);
   return 0;
}
