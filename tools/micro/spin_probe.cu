// Which host-side CUDA calls block while a kernel of ANOTHER stream of the same process spins on a flag?
// (single-process multi-rank mode, de_multi.cu: a rank's reduction tail spins until its peers' contributions arrive;
// a peer whose host thread is stuck in such a call never delivers -> deadlock.) Every call is probed on its own: a
// spinner is started, the call is made, a watchdog releases the spinner after 0.5 s. A call that only returns after the
// release was blocked by the resident kernel.
#include <chrono>
#include <cstdio>
#include <functional>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

__global__ void spinner(volatile int *flag)
{
  while (*flag == 0)
    __nanosleep(200);
}
__global__ void fresh_kernel_a(int *p) { if (p) *p = 1; }
__global__ void fresh_kernel_b(int *p) { if (p) *p = 2; }
__global__ void __launch_bounds__(256) big_smem_kernel(int *p)
{
  extern __shared__ int sm[];
  sm[threadIdx.x] = 1;
  if (p) *p = sm[0];
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main()
{
  int *hflag;
  cudaHostAlloc((void **)&hflag, sizeof(int), cudaHostAllocMapped);
  int *dflag;
  cudaHostGetDevicePointer((void **)&dflag, hflag, 0);
  cudaStream_t s1, s2;
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  int *dummy;
  cudaMalloc(&dummy, 4);
  char *dbuf, *dbuf2;
  cudaMalloc(&dbuf, 8 << 20);
  cudaMalloc(&dbuf2, 8 << 20);
  void *pinned;
  cudaMallocHost(&pinned, 4 << 20);
  std::vector<char> pageable(4 << 20, 1);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, fresh_kernel_b); // the library's remedy: force the load BEFORE anything can spin
  cudaFuncGetAttributes(&fa, big_smem_kernel);
  cudaDeviceSynchronize();
  const double release_after = 0.5;
  auto probe = [&](const char *name, const std::function<void()> &f) {
    *hflag = 0;
    spinner<<<1, 32, 0, s1>>>(dflag);
    const double a = now();
    std::thread rel([&] {
      std::this_thread::sleep_for(std::chrono::milliseconds((int)(release_after * 1e3)));
      *hflag = 1;
    });
    f();
    const double b = now();
    rel.join();
    cudaDeviceSynchronize();
    std::printf("%-58s %9.3f ms%s\n", name, (b - a) * 1e3, (b - a > 0.8 * release_after) ? "   <-- BLOCKED until release" : "");
  };
  void *p1 = nullptr, *p2 = nullptr, *h1 = nullptr;
  cudaStream_t s3;
  cudaEvent_t ev;
  probe("cudaMalloc 1 MB", [&] { cudaMalloc(&p1, 1 << 20); });
  probe("cudaMalloc 1 GB", [&] { cudaMalloc(&p2, (size_t)1 << 30); });
  probe("cudaMallocHost 8 MB", [&] { cudaMallocHost(&h1, 8 << 20); });
  probe("cudaStreamCreate", [&] { cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking); });
  probe("cudaEventCreate", [&] { cudaEventCreate(&ev); });
  probe("launch of a PRELOADED kernel + sync", [&] { fresh_kernel_b<<<1, 1, 0, s2>>>(dummy); cudaStreamSynchronize(s2); });
  probe("cudaFuncSetAttribute(max dyn smem) on a preloaded kernel", [&] { cudaFuncSetAttribute(big_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
  probe("launch big-smem kernel (148 CTAs x 200 KB) + sync", [&] { big_smem_kernel<<<148, 256, 200 * 1024, s2>>>(dummy); cudaStreamSynchronize(s2); });
  probe("cudaMemsetAsync 4 B + sync", [&] { cudaMemsetAsync(dbuf, 0, 4, s2); cudaStreamSynchronize(s2); });
  probe("cudaMemsetAsync 32 B + sync", [&] { cudaMemsetAsync(dbuf, 0, 32, s2); cudaStreamSynchronize(s2); });
  probe("cudaMemsetAsync 20 B at odd offset + sync", [&] { cudaMemsetAsync(dbuf + 3, 0, 20, s2); cudaStreamSynchronize(s2); });
  probe("cudaMemsetAsync 1 MB + sync", [&] { cudaMemsetAsync(dbuf, 0, 1 << 20, s2); cudaStreamSynchronize(s2); });
  probe("cudaMemcpyAsync D2D 256 B + sync", [&] { cudaMemcpyAsync(dbuf, dbuf2, 256, cudaMemcpyDeviceToDevice, s2); cudaStreamSynchronize(s2); });
  probe("cudaMemcpyAsync D2D 4 MB + sync", [&] { cudaMemcpyAsync(dbuf, dbuf2, 4 << 20, cudaMemcpyDeviceToDevice, s2); cudaStreamSynchronize(s2); });
  probe("cudaMemcpyAsync H2D pageable 256 B + sync", [&] { cudaMemcpyAsync(dbuf, pageable.data(), 256, cudaMemcpyHostToDevice, s2); cudaStreamSynchronize(s2); });
  probe("cudaMemcpyAsync H2D pageable 4 MB + sync", [&] { cudaMemcpyAsync(dbuf, pageable.data(), 4 << 20, cudaMemcpyHostToDevice, s2); cudaStreamSynchronize(s2); });
  probe("cudaMemcpyAsync D2H pageable 256 B + sync", [&] { cudaMemcpyAsync(pageable.data(), dbuf, 256, cudaMemcpyDeviceToHost, s2); cudaStreamSynchronize(s2); });
  probe("cudaMemcpyAsync D2H pageable 4 MB + sync", [&] { cudaMemcpyAsync(pageable.data(), dbuf, 4 << 20, cudaMemcpyDeviceToHost, s2); cudaStreamSynchronize(s2); });
  probe("cudaMemcpyAsync H2D pinned 4 MB + sync", [&] { cudaMemcpyAsync(dbuf, pinned, 4 << 20, cudaMemcpyHostToDevice, s2); cudaStreamSynchronize(s2); });
  probe("cudaMemcpyAsync D2H pinned 16 B + sync", [&] { cudaMemcpyAsync(pinned, dbuf, 16, cudaMemcpyDeviceToHost, s2); cudaStreamSynchronize(s2); });
  probe("cudaMemcpy (sync API, legacy stream) D2H 4 B", [&] { int v; cudaMemcpy(&v, dummy, 4, cudaMemcpyDeviceToHost); });
  probe("cudaOccupancyMaxActiveBlocksPerMultiprocessor", [&] { int o; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fresh_kernel_b, 256, 0); });
  probe("cudaEventRecord + cudaEventSynchronize on idle stream", [&] { cudaEventRecord(ev, s2); cudaEventSynchronize(ev); });
  probe("stream capture + cudaGraphInstantiate + cudaGraphLaunch", [&] {
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s2, cudaStreamCaptureModeThreadLocal);
    fresh_kernel_b<<<1, 1, 0, s2>>>(dummy);
    cudaStreamEndCapture(s2, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, s2);
    cudaStreamSynchronize(s2);
  });
  probe("cudaFree 1 MB", [&] { cudaFree(p1); });
  probe("cudaFreeHost", [&] { cudaFreeHost(h1); });
  probe("cudaStreamDestroy / cudaEventDestroy", [&] { cudaStreamDestroy(s3); cudaEventDestroy(ev); });
  probe("first launch of a fresh kernel (lazy load) + sync", [&] { fresh_kernel_a<<<1, 1, 0, s2>>>(dummy); cudaStreamSynchronize(s2); });
  std::printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
