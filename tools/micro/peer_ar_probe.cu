// Stand-alone reproduction of the one-shot peer all-reduce (csrc/kernels_peer.cuh) with R ranks on ONE device:
// R streams, R windows, one host thread per rank, N back-to-back all-reduces of a short vector.
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "../../dune_eigensolver_b200/csrc/kernels_peer.cuh"

int main(int argc, char **argv)
{
  const int R = argc > 1 ? std::atoi(argv[1]) : 4, N = argc > 2 ? std::atoi(argv[2]) : 1000, len = 64;
  std::vector<unsigned char *> win(R);
  std::vector<cudaStream_t> st(R);
  std::vector<double *> buf(R);
  std::vector<int *> err(R);
  const size_t bytes = de::kPeerHaloOff + 1024;
  for (int r = 0; r < R; ++r)
  {
    cudaMalloc(&win[r], bytes);
    cudaMemset(win[r], 0, bytes);
    cudaStreamCreateWithFlags(&st[r], cudaStreamNonBlocking);
    cudaMalloc(&buf[r], len * sizeof(double));
    cudaMalloc(&err[r], sizeof(int));
    cudaMemset(err[r], 0, sizeof(int));
  }
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, de::peer_allreduce_kernel);
  cudaDeviceSynchronize();
  std::vector<int> bad(R, 0);
  auto work = [&](int r) {
    std::vector<double> h(len);
    for (int it = 1; it <= N; ++it)
    {
      for (int i = 0; i < len; ++i)
        h[i] = r + 1.0;
      cudaMemcpyAsync(buf[r], h.data(), len * sizeof(double), cudaMemcpyHostToDevice, st[r]);
      de::PeerArgs pa{};
      pa.rank = r;
      pa.nranks = R;
      for (int q = 0; q < R; ++q)
        pa.base[q] = win[q];
      pa.epoch = (unsigned long long)it;
      pa.done = nullptr;
      pa.err = err[r];
      pa.timeout = 4000000000LL;
      de::peer_allreduce_kernel<<<1, 1024, 0, st[r]>>>(pa, buf[r], len);
      cudaMemcpyAsync(h.data(), buf[r], len * sizeof(double), cudaMemcpyDeviceToHost, st[r]);
      cudaStreamSynchronize(st[r]);
      if (h[0] != R * (R + 1) / 2.0)
      {
        int e = 0;
        cudaMemcpy(&e, err[r], sizeof(int), cudaMemcpyDeviceToHost);
        std::printf("rank %d iteration %d: got %g, err word 0x%x (what %d from %d epoch %d flag %d)\n", r, it, h[0], e, e & 15, (e >> 4) & 15,
                    (e >> 8) & 0xfff, (e >> 20) & 0x7ff);
        bad[r] = 1;
        return;
      }
    }
  };
  std::vector<std::thread> th;
  for (int r = 0; r < R; ++r)
    th.emplace_back(work, r);
  for (auto &t : th)
    t.join();
  int nb = 0;
  for (int b : bad)
    nb += b;
  std::printf("R=%d N=%d: %s (%s)\n", R, N, nb ? "FAILED" : "ok", cudaGetErrorString(cudaGetLastError()));
  return nb;
}
