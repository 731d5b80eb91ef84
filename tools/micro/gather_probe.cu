// Probe: achievable SM <- L2 / HBM bandwidth for gathers of whole 8*m-byte rows (what the SpMM X gather does).
//   gather_probe <table MB> <row bytes> : every 16-byte lane group of (row bytes / 16) lanes loads pseudo-random rows
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); std::exit(1);} } while (0)

template <int UN>
__global__ void __launch_bounds__(256) gather(const double2 *__restrict__ T, unsigned nrows, int tpr, int iters, int window, double2 *out)
{
  const int t = threadIdx.x % tpr, grp = (blockIdx.x * 256 + threadIdx.x) / tpr;
  unsigned s = grp * 2654435761u + 12345u;
  double2 acc = make_double2(0, 0);
  for (int it = 0; it < iters; ++it)
  {
    double2 v[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u)
    {
      s = s * 1664525u + 1013904223u;
      // window == 0: uniformly random rows; else rows near a slowly moving base (stencil-like locality)
      unsigned r = window ? ((unsigned)(((unsigned long long)blockIdx.x * iters + it) * 16u) + (s >> 8) % window) % nrows : (s >> 4) % nrows;
      v[u] = __ldg(T + (size_t)r * tpr + t);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
  }
  if (acc.x == 123.456) out[0] = acc;
}

int main(int argc, char **argv)
{
  const double mb = argc > 1 ? atof(argv[1]) : 64;
  const int rowbytes = argc > 2 ? atoi(argv[2]) : 256;
  const int window = argc > 3 ? atoi(argv[3]) : 0;
  const int tpr = rowbytes / 16;
  const unsigned nrows = (unsigned)(mb * 1e6 / rowbytes);
  double2 *T, *out;
  CK(cudaMalloc(&T, (size_t)nrows * rowbytes));
  CK(cudaMalloc(&out, 64));
  CK(cudaMemset(T, 0, (size_t)nrows * rowbytes));
  for (int ctas = 2; ctas <= 8; ctas *= 2)
  {
    const int grid = 148 * ctas, iters = 400;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<8><<<grid, 256>>>(T, nrows, tpr, iters, window, out);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    gather<8><<<grid, 256>>>(T, nrows, tpr, iters, window, out);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)grid * 256 * 16.0 * iters * 8;
    std::printf("table %.0f MB rows %d B window %d  %d CTAs/SM x 8 loads in flight: %.3f ms  %.1f GB/s\n", mb, rowbytes, window, ctas, ms, bytes / ms / 1e6);
  }
  return 0;
}
