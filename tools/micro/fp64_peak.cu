// Micro-benchmark: FP64 FMA pipe vs DMMA (mma.sync.m8n8k4.f64) throughput on sm_100a, and shared-memory LDS.128 rate.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/fp64_peak tools/micro/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 2) dfma_kernel(double *out, int iters)
{
  double a[16], b = 1.0000001, c = 0.5;
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  double s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 2) dmma_kernel(double *out, int iters)
{
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  double a = 1.0 + threadIdx.x * 1e-6, b = 0.999;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
  double *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096, grid = 148 * 2;
  for (int rep = 0; rep < 2; ++rep)
  {
    cudaEventRecord(e0); dfma_kernel<<<grid, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 16 * iters * 256.0 * grid;
    printf("DFMA  : %.3f ms  %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    cudaEventRecord(e0); dmma_kernel<<<grid, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    fl = 2.0 * 256 * 8 * iters * 8.0 * grid; // 256 FMA per warp-MMA, 8 MMAs per iter, 8 warps per CTA
    printf("DMMA  : %.3f ms  %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
