// Development bench for SpMM kernel designs (not part of the library): builds a 3D stencil matrix on the host,
// runs the candidate kernels on one GPU, checks them against a plain one-thread-per-entry kernel and prints
// time, algorithmic GB/s (12*nnz + 4*(n+1) + 16*n*m bytes) and the fraction of the measured HBM peak.
//
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -o tools/micro/spmm_lab tools/micro/spmm_lab.cu
//   tools/micro/spmm_lab <grid> <fd|q1> <m> [reps]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../dune_eigensolver_b200/csrc/kernels_sparse.cuh"
#include "../../dune_eigensolver_b200/csrc/kernels_spmm_blocked.cuh"
#include "../../dune_eigensolver_b200/csrc/brb_format.hpp"
#include <chrono>

#define CK(x)                                                                                     \
  do                                                                                              \
  {                                                                                               \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess)                                                                        \
    {                                                                                             \
      std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      std::exit(1);                                                                               \
    }                                                                                             \
  } while (0)

static void build_stencil(int N, bool q1, std::vector<int> &rp, std::vector<int> &ci, std::vector<double> &v)
{
  const long long n = (long long)N * N * N;
  rp.assign(n + 1, 0);
  ci.clear();
  v.clear();
  ci.reserve((size_t)n * (q1 ? 27 : 7));
  v.reserve((size_t)n * (q1 ? 27 : 7));
  for (int z = 0; z < N; ++z)
    for (int y = 0; y < N; ++y)
      for (int x = 0; x < N; ++x)
      {
        const long long r = ((long long)z * N + y) * N + x;
        for (int dz = -1; dz <= 1; ++dz)
          for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx)
            {
              const int s = std::abs(dx) + std::abs(dy) + std::abs(dz);
              if (!q1 && s > 1)
                continue;
              const int xx = x + dx, yy = y + dy, zz = z + dz;
              if (xx < 0 || yy < 0 || zz < 0 || xx >= N || yy >= N || zz >= N)
                continue;
              ci.push_back((int)(((long long)zz * N + yy) * N + xx));
              v.push_back(s == 0 ? (q1 ? 8.0 / 3.0 : 6.0) : (q1 ? -1.0 / (3.0 * s) - 0.01 * dx : -1.0));
            }
        rp[r + 1] = (int)ci.size();
      }
}

__global__ void ref_spmm(long long n, int m, const int *rp, const int *ci, const double *v, const double *X, double *Y)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * m)
    return;
  const long long r = e / m;
  const int c = (int)(e % m);
  double acc = 0.0;
  for (int k = rp[r]; k < rp[r + 1]; ++k)
    acc = fma(v[k], X[(size_t)ci[k] * m + c], acc);
  Y[e] = acc;
}

__global__ void maxdiff(long long cnt, const double *a, const double *b, double *out)
{
  double mx = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += (long long)gridDim.x * blockDim.x)
    mx = fmax(mx, fabs(a[e] - b[e]));
  atomicMax(reinterpret_cast<unsigned long long *>(out), (unsigned long long)__double_as_longlong(mx));
}

template <class F>
static float time_it(F f, int reps)
{
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i)
    f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i)
    f();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main(int argc, char **argv)
{
  const int N = argc > 1 ? std::atoi(argv[1]) : 100;
  const bool q1 = argc > 2 ? std::string(argv[2]) == "q1" : true;
  const int m = argc > 3 ? std::atoi(argv[3]) : 32;
  const int reps = argc > 4 ? std::atoi(argv[4]) : 20;
  const double peak = 6529.1;

  std::vector<int> rp, ci;
  std::vector<double> v;
  build_stencil(N, q1, rp, ci, v);
  const long long n = (long long)rp.size() - 1, nnz = (long long)ci.size();
  const double bytes = 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n * m;
  std::printf("grid %d^3 %s n=%lld nnz=%lld m=%d  algorithmic bytes %.3f GB  (%.3f ms at %.0f GB/s)\n", N, q1 ? "q1" : "fd",
              n, nnz, m, bytes / 1e9, bytes / peak / 1e6, peak);

  int *d_rp, *d_ci;
  double *d_v, *d_X, *d_Y, *d_Yref, *d_part, *d_diff;
  CK(cudaMalloc(&d_rp, (n + 1 + 8) * sizeof(int)));
  CK(cudaMalloc(&d_ci, (nnz + 8) * sizeof(int)));
  CK(cudaMalloc(&d_v, (nnz + 8) * sizeof(double)));
  CK(cudaMalloc(&d_X, n * m * sizeof(double)));
  CK(cudaMalloc(&d_Y, n * m * sizeof(double)));
  CK(cudaMalloc(&d_Yref, n * m * sizeof(double)));
  CK(cudaMalloc(&d_part, 4096 * 64 * sizeof(double)));
  CK(cudaMalloc(&d_diff, sizeof(double)));
  CK(cudaMemset(d_ci + nnz, 0, 8 * sizeof(int)));
  CK(cudaMemset(d_v + nnz, 0, 8 * sizeof(double)));
  CK(cudaMemcpy(d_rp, rp.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ci, ci.data(), nnz * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_v, v.data(), nnz * sizeof(double), cudaMemcpyHostToDevice));
  {
    std::vector<double> X((size_t)n * m);
    unsigned long long s = 88172645463325252ULL;
    for (auto &x : X)
    {
      s ^= s << 13;
      s ^= s >> 7;
      s ^= s << 17;
      x = (double)(s >> 11) / 9007199254740992.0 - 0.5;
    }
    CK(cudaMemcpy(d_X, X.data(), X.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  ref_spmm<<<(unsigned)((n * m + 255) / 256), 256>>>(n, m, d_rp, d_ci, d_v, d_X, d_Yref);
  CK(cudaDeviceSynchronize());

  auto check = [&](const char *name, float ms)
  {
    CK(cudaMemset(d_diff, 0, sizeof(double)));
    maxdiff<<<592, 256>>>(n * m, d_Y, d_Yref, d_diff);
    double diff;
    CK(cudaMemcpy(&diff, d_diff, sizeof(double), cudaMemcpyDeviceToHost));
    std::printf("%-28s %8.4f ms  %8.1f GB/s  %.3f of peak   max|dY| %.2e\n", name, ms, bytes / ms / 1e6, bytes / ms / 1e6 / peak, diff);
    CK(cudaMemset(d_Y, 0, n * m * sizeof(double)));
  };

  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

  // ---- baseline: the library's staged kernel -----------------------------------------------------------
  {
    std::vector<int4> meta;
    long long r0 = 0;
    while (r0 < n)
    {
      long long r1 = r0;
      while (r1 < n && r1 - r0 < de::kStageMaxRows && (long long)(rp[r1 + 1] - rp[r0]) <= de::kStageCapNnz)
        ++r1;
      if (r1 == r0)
        r1 = r0 + 1;
      meta.push_back(make_int4((int)r0, (int)r1, rp[r0], rp[r1]));
      r0 = r1;
    }
    int4 *d_meta;
    CK(cudaMalloc(&d_meta, meta.size() * sizeof(int4)));
    CK(cudaMemcpy(d_meta, meta.data(), meta.size() * sizeof(int4), cudaMemcpyHostToDevice));
    de::StagedArgs a{};
    a.nblocks = (int)meta.size();
    a.blk_meta = d_meta;
    a.rowmap = nullptr;
    a.rowptr = d_rp;
    a.col = d_ci;
    a.val = d_v;
    a.X = d_X;
    a.H = nullptr;
    a.n_owned = (int)n;
    a.m = m;
    a.Y = d_Y;
    a.partials = d_part;
    const size_t smem = de::spmm_staged_smem_bytes();
    const int grid = std::min((int)meta.size(), sms * 3);
    auto run = [&](auto kern, const char *name)
    {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      float ms = time_it([&] { kern<<<grid, 256, smem>>>(a); }, reps);
      check(name, ms);
    };
    if (m == 8)
      run(de::spmm_staged_kernel<4, false, false>, "staged(v3)");
    if (m == 16)
      run(de::spmm_staged_kernel<8, false, false>, "staged(v3)");
    if (m == 32)
      run(de::spmm_staged_kernel<16, false, false>, "staged(v3)");
    if (m == 64)
      run(de::spmm_staged_kernel<32, false, false>, "staged(v3)");

  }

  // ---- production path: BRB format (host builder) + warp-specialised tensor-core kernel ----------------------
  {
    de::brb::Format F;
    const auto t0 = std::chrono::steady_clock::now();
    const bool ok = de::brb::build(n, n, rp.data(), ci.data(), v.data(), n, F);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("brb: ok=%d grid=%d S=(%lld,%lld) tile %dx%dx%d, %d tiles, %.2f steps/block, fill %.3f, union/rows %.2f, max_u %d, blob %.3f GB, host build %.3f s\n",
                (int)ok, (int)F.grid, F.S1, F.S2, F.tw, F.th, F.td, F.ntiles, (double)F.nsteps / std::max(1LL, F.nblocks),
                (double)nnz / (32.0 * std::max(1LL, F.nsteps)), (double)F.ucol.size() / n, F.max_u, F.blob.size() * 4e-9, dt);
    if (ok)
    {
      int4 *d_tile, *d_blob;
      int *d_uc;
      CK(cudaMalloc(&d_tile, F.tile.size() * sizeof(int4)));
      CK(cudaMalloc(&d_blob, F.blob.size() * 4 + 64));
      CK(cudaMalloc(&d_uc, F.ucol.size() * 4 + 64));
      CK(cudaMemcpy(d_tile, F.tile.data(), F.tile.size() * sizeof(int4), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_blob, F.blob.data(), F.blob.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_uc, F.ucol.data(), F.ucol.size() * 4, cudaMemcpyHostToDevice));
      const int np = std::min(m / 8, 4), passes = m / (8 * np);
      int stages = de::kBrbMaxStages;
      while (stages > 2 && de::spmm_brb_smem_bytes(np, F.max_len16, F.max_u, stages) > 227 * 1024)
        --stages;
      const size_t smem = de::spmm_brb_smem_bytes(np, F.max_len16, F.max_u, stages);
      const int grid = std::min(F.ntiles, sms);
      auto launch = [&](auto kern, bool dot)
      {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int ps = 0; ps < passes; ++ps)
        {
          de::BrbArgs b{};
          b.ntiles = F.ntiles;
          b.n = n;
          b.tile = d_tile;
          b.blob = d_blob;
          b.ucol = d_uc;
          b.X = d_X + 8 * np * ps;
          b.H = nullptr;
          b.n_owned = (int)n;
          b.ldx = m;
          b.Y = d_Y + 8 * np * ps;
          b.partials = d_part + 8 * np * ps;
          b.pstride = m;
          b.gram_off = m;
          b.done = nullptr;
          b.blob_cap16 = F.max_len16;
          b.xs_cap = F.max_u;
          b.stages = stages;
          kern<<<grid, de::brb_threads(false), smem>>>(b);
        }
        (void)dot;
      };
      char nm[64];
      std::snprintf(nm, sizeof nm, "brb %d stage(s) x%d pass", stages, passes);
      float ms = 0;
      if (np == 1)
        ms = time_it([&] { launch(de::spmm_brb_kernel<1, false, false, false>, false); }, reps);
      if (np == 2)
        ms = time_it([&] { launch(de::spmm_brb_kernel<2, false, false, false>, false); }, reps);
      if (np == 4)
        ms = time_it([&] { launch(de::spmm_brb_kernel<4, false, false, false>, false); }, reps);
      check(nm, ms);
      if (np == 4)
      {
        ms = time_it([&] { launch(de::spmm_brb_kernel<4, false, true, false>, false); }, reps);
        check("brb HALO variant (no halo cols)", ms);
      }
      if (np == 1)
        ms = time_it([&] { launch(de::spmm_brb_kernel<1, true, false, false>, true); }, reps);
      if (np == 2)
        ms = time_it([&] { launch(de::spmm_brb_kernel<2, true, false, false>, true); }, reps);
      if (np == 4)
        ms = time_it([&] { launch(de::spmm_brb_kernel<4, true, false, false>, true); }, reps);
      // dot check on the host: sum of partials vs sum_i X(i,j) Yref(i,j)
      {
        std::vector<double> part((size_t)grid * m), X((size_t)n * m), Yr((size_t)n * m);
        CK(cudaMemcpy(part.data(), d_part, part.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(X.data(), d_X, X.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(Yr.data(), d_Yref, Yr.size() * 8, cudaMemcpyDeviceToHost));
        double worst = 0;
        for (int j = 0; j < m; ++j)
        {
          double sref = 0, sg = 0;
          for (long long i = 0; i < n; ++i)
            sref += X[(size_t)i * m + j] * Yr[(size_t)i * m + j];
          for (int c = 0; c < grid; ++c)
            sg += part[(size_t)c * m + j];
          worst = std::max(worst, std::fabs(sref - sg) / std::max(1.0, std::fabs(sref)));
        }
        std::printf("  dot: max relative error %.2e\n", worst);
      }
      std::snprintf(nm, sizeof nm, "brb+dot %d stage(s) x%d pass", stages, passes);
      check(nm, ms);
    }
  }
  return 0;
}
