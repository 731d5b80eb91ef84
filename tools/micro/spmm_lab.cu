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

    // ---- VA: staged CSR with vectorised metadata reads and predicated gathers ----
    const size_t smem4 = de::spmm_csr4_smem_bytes();
    auto run4 = [&](auto kern, const char *name, int ctas)
    {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
      const int g = std::min((int)meta.size(), sms * ctas);
      float ms = time_it([&] { kern<<<g, 256, smem4>>>(a); }, reps);
      check(name, ms);
    };
    if (m == 8)
      run4(de::spmm_csr4_kernel<4, false, false>, "csr4 (VA)", 3);
    if (m == 16)
      run4(de::spmm_csr4_kernel<8, false, false>, "csr4 (VA)", 3);
    if (m == 32)
      run4(de::spmm_csr4_kernel<16, false, false>, "csr4 (VA)", 3);
    if (m == 64)
      run4(de::spmm_csr4_kernel<32, false, false>, "csr4 (VA)", 3);
  }

  // ---- VB: 8-row block-sparse DMMA kernel ------------------------------------------------------------------
  {
    de::Brb8Host H;
    de::brb8_build_host(n, rp.data(), ci.data(), v.data(), nullptr, H);
    std::printf("brb8: %lld row blocks, %lld steps (%.2f per block), fill %.3f, matrix bytes %.3f GB (CSR %.3f GB)\n",
                (long long)H.nblocks, (long long)H.stepmask.size(), (double)H.stepmask.size() / H.nblocks,
                (double)nnz / (32.0 * H.stepmask.size()),
                (20.0 * H.stepmask.size() + 8.0 * nnz + 12.0 * H.nblocks) / 1e9, (12.0 * nnz + 4.0 * n) / 1e9);
    de::Brb8Args b{};
    int *d_bs, *d_bv, *d_sc, *d_rows = nullptr;
    unsigned *d_sm;
    CK(cudaMalloc(&d_bs, H.blkstep.size() * sizeof(int)));
    CK(cudaMalloc(&d_bv, H.blkval.size() * sizeof(int)));
    CK(cudaMalloc(&d_sc, (H.stepcol.size() + 64) * sizeof(int)));
    CK(cudaMalloc(&d_sm, (H.stepmask.size() + 16) * sizeof(unsigned)));
    CK(cudaMemset(d_sc, 0, (H.stepcol.size() + 64) * sizeof(int)));
    CK(cudaMemset(d_sm, 0, (H.stepmask.size() + 16) * sizeof(unsigned)));
    CK(cudaMemcpy(d_bs, H.blkstep.data(), H.blkstep.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_bv, H.blkval.data(), H.blkval.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sc, H.stepcol.data(), H.stepcol.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sm, H.stepmask.data(), H.stepmask.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
    double *d_bval;
    CK(cudaMalloc(&d_bval, (H.val.size() + 8) * sizeof(double)));
    CK(cudaMemcpy(d_bval, H.val.data(), H.val.size() * sizeof(double), cudaMemcpyHostToDevice));
    b.nblocks = H.nblocks;
    b.n = n;
    b.blkstep = d_bs;
    b.blkval = d_bv;
    b.stepcol = d_sc;
    b.stepmask = d_sm;
    b.val = d_bval;
    b.blkrows = d_rows;
    b.X = d_X;
    b.H = nullptr;
    b.n_owned = (int)n;
    b.Y = d_Y;
    b.partials = d_part;
    auto runb = [&](auto kern, const char *name, int threads, int ctas)
    {
      const int wpb = threads / 32;
      const int g = (int)std::min<long long>((H.nblocks + wpb - 1) / wpb, (long long)sms * ctas);
      float ms = time_it([&] { kern<<<g, threads>>>(b); }, reps);
      check(name, ms);
    };
    if (m == 8)
      runb(de::spmm_brb8_kernel<1, false, false>, "brb8 dmma (VB)", 256, 3);
    if (m == 16)
      runb(de::spmm_brb8_kernel<2, false, false>, "brb8 dmma (VB)", 256, 3);
    if (m == 32)
    {
      runb(de::spmm_brb8_kernel<4, false, false>, "brb8 dmma (VB) 3cta", 256, 3);
      runb(de::spmm_brb8_kernel<4, false, false>, "brb8 dmma (VB) 2cta", 256, 2);
      runb(de::spmm_brb8_kernel<4, false, false>, "brb8 dmma (VB) 4cta", 256, 4);
    }
    if (m == 64)
      runb(de::spmm_brb8_kernel<8, false, false>, "brb8 dmma (VB)", 256, 2);
  }
  // ---- VC: tiled BRB8 with the X rows of a tile staged in shared memory --------------------------------------
  {
    const char *env = std::getenv("LAB_TILES");
    std::string spec = env ? env : "8,6,6,8,1,1;8,6,6,2,2,2;8,5,5,2,2,2;4,8,8,2,2,2;8,4,4,2,2,2";
    size_t pos = 0;
    while (pos < spec.size())
    {
      size_t end = spec.find(';', pos);
      if (end == std::string::npos)
        end = spec.size();
      int tw, th, td, bw, bh, bd;
      if (std::sscanf(spec.substr(pos, end - pos).c_str(), "%d,%d,%d,%d,%d,%d", &tw, &th, &td, &bw, &bh, &bd) != 6)
        break;
      pos = end + 1;
      std::vector<int> rows, tilecut;
      de::brb8t_grid_order(n, N, (long long)N * N, tw, th, td, bw, bh, bd, rows, tilecut);
      de::Brb8THost H;
      de::brb8t_build_host(rp.data(), ci.data(), v.data(), rows, tilecut, H);
      const size_t smem = (size_t)H.max_u * (m + 4) * sizeof(double);
      const double mbytes = 12.0 * H.stepmask.size() + 8.0 * H.val.size() + 40.0 * H.nblocks + 4.0 * H.ucol.size();
      std::printf("brb8t tile %dx%dx%d block %dx%dx%d: %d tiles, %d blocks, %.2f steps/block, fill %.3f, union/rows %.2f, max_u %d, smem %.1f KB, matrix bytes %.3f GB\n",
                  tw, th, td, bw, bh, bd, H.ntiles, H.nblocks, (double)H.stepmask.size() / H.nblocks,
                  (double)nnz / (32.0 * H.stepmask.size()), (double)H.ucol.size() / n, H.max_u, smem / 1024.0, mbytes / 1e9);
      if (smem > 227 * 1024)
      {
        std::printf("  (tile does not fit in shared memory)\n");
        continue;
      }
      de::Brb8TArgs b{};
      int4 *d_tile;
      int *d_uc, *d_bs, *d_bv, *d_rows;
      unsigned short *d_lc;
      unsigned *d_sm;
      double *d_bval;
      CK(cudaMalloc(&d_tile, H.tile.size() * sizeof(int4)));
      CK(cudaMalloc(&d_uc, H.ucol.size() * sizeof(int)));
      CK(cudaMalloc(&d_bs, H.blkstep.size() * sizeof(int)));
      CK(cudaMalloc(&d_bv, H.blkval.size() * sizeof(int)));
      CK(cudaMalloc(&d_rows, H.blkrows.size() * sizeof(int)));
      CK(cudaMalloc(&d_lc, (H.steplc.size() + 64) * sizeof(unsigned short)));
      CK(cudaMalloc(&d_sm, (H.stepmask.size() + 16) * sizeof(unsigned)));
      CK(cudaMalloc(&d_bval, (H.val.size() + 8) * sizeof(double)));
      CK(cudaMemset(d_lc, 0, (H.steplc.size() + 64) * sizeof(unsigned short)));
      CK(cudaMemset(d_sm, 0, (H.stepmask.size() + 16) * sizeof(unsigned)));
      CK(cudaMemcpy(d_tile, H.tile.data(), H.tile.size() * sizeof(int4), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_uc, H.ucol.data(), H.ucol.size() * sizeof(int), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_bs, H.blkstep.data(), H.blkstep.size() * sizeof(int), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_bv, H.blkval.data(), H.blkval.size() * sizeof(int), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_rows, H.blkrows.data(), H.blkrows.size() * sizeof(int), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_lc, H.steplc.data(), H.steplc.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_sm, H.stepmask.data(), H.stepmask.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_bval, H.val.data(), H.val.size() * sizeof(double), cudaMemcpyHostToDevice));
      b.ntiles = H.ntiles;
      b.n = n;
      b.tile = d_tile;
      b.ucol = d_uc;
      b.blkstep = d_bs;
      b.blkval = d_bv;
      b.steplc = d_lc;
      b.stepmask = d_sm;
      b.val = d_bval;
      b.blkrows = d_rows;
      b.X = d_X;
      b.H = nullptr;
      b.n_owned = (int)n;
      b.Y = d_Y;
      b.partials = d_part;
      auto runt = [&](auto kern, const char *name)
      {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
        const int gridt = std::min(H.ntiles, sms * std::max(occ, 1));
        float ms = time_it([&] { kern<<<gridt, 256, smem>>>(b); }, reps);
        char nm[64];
        std::snprintf(nm, sizeof nm, "%s occ%d", name, occ);
        check(nm, ms);
      };
      if (m == 8)
        runt(de::spmm_brb8t_kernel<1, false, false>, "brb8t (VC)");
      if (m == 16)
        runt(de::spmm_brb8t_kernel<2, false, false>, "brb8t (VC)");
      if (m == 32)
        runt(de::spmm_brb8t_kernel<4, false, false>, "brb8t (VC)");
      if (m == 64)
        runt(de::spmm_brb8t_kernel<8, false, false>, "brb8t (VC)");

      // ---- VD: pipelined tiles ----
      {
        de::Brb8PHost P;
        de::brb8p_pack_host(H, P);
        const size_t smemp = 2 * ((size_t)P.max_len16 * 16 + (size_t)P.max_u * (m + 4) * sizeof(double));
        std::printf("  brb8p: blob %.3f GB, max blob %.1f KB, smem %.1f KB (2 buffers)\n", P.blob.size() * 4.0 / 1e9,
                    P.max_len16 * 16.0 / 1024, smemp / 1024.0);
        if (smemp <= 227 * 1024 && (size_t)P.max_u * (m / 2) <= 16 * 512)
        {
          de::TileDesc *d_td;
          int4 *d_blob;
          CK(cudaMalloc(&d_td, P.tile.size() * sizeof(de::TileDesc)));
          CK(cudaMalloc(&d_blob, P.blob.size() * 4 + 64));
          CK(cudaMemcpy(d_td, P.tile.data(), P.tile.size() * sizeof(de::TileDesc), cudaMemcpyHostToDevice));
          CK(cudaMemcpy(d_blob, P.blob.data(), P.blob.size() * 4, cudaMemcpyHostToDevice));
          de::Brb8PArgs pa{};
          pa.ntiles = H.ntiles;
          pa.n = n;
          pa.tile = d_td;
          pa.blob = d_blob;
          pa.ucol = d_uc;
          pa.X = d_X;
          pa.H = nullptr;
          pa.n_owned = (int)n;
          pa.ldx = m;
          pa.Y = d_Y;
          pa.partials = d_part;
          pa.blob_cap16 = P.max_len16;
          pa.xs_cap = P.max_u;
          auto runp = [&](auto kern, const char *name)
          {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemp));
            const int gridp = std::min(H.ntiles, sms);
            float ms = time_it([&] { kern<<<gridp, de::kBrbThreads, smemp>>>(pa); }, reps);
            check(name, ms);
          };
          if (m == 8)
            runp(de::spmm_brb8p_kernel<1, false, false>, "  brb8p (VD)");
          if (m == 16)
            runp(de::spmm_brb8p_kernel<2, false, false>, "  brb8p (VD)");
          if (m == 32)
            runp(de::spmm_brb8p_kernel<4, false, false>, "  brb8p (VD)");
          cudaFree(d_td);
          cudaFree(d_blob);
        }
        else
          std::printf("  (does not fit)\n");
      }

      // ---- VE: warp-specialised TMA pipeline ----
      {
        de::Brb8PHost P;
        de::brb8q_pack_host(H, P);
        const size_t per = (size_t)P.max_len16 * 16 + (size_t)P.max_u * (m + 4) * sizeof(double);
        std::printf("  brb8q: blob %.3f GB, max blob %.1f KB, per-stage %.1f KB\n", P.blob.size() * 4.0 / 1e9, P.max_len16 * 16.0 / 1024, per / 1024.0);
        de::TileDesc *d_td;
        int4 *d_blob;
        CK(cudaMalloc(&d_td, P.tile.size() * sizeof(de::TileDesc)));
        CK(cudaMalloc(&d_blob, P.blob.size() * 4 + 64));
        CK(cudaMemcpy(d_td, P.tile.data(), P.tile.size() * sizeof(de::TileDesc), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_blob, P.blob.data(), P.blob.size() * 4, cudaMemcpyHostToDevice));
        de::Brb8PArgs pa{};
        pa.ntiles = H.ntiles;
        pa.n = n;
        pa.tile = d_td;
        pa.blob = d_blob;
        pa.ucol = d_uc;
        pa.X = d_X;
        pa.H = nullptr;
        pa.n_owned = (int)n;
        pa.ldx = m;
        pa.Y = d_Y;
        pa.partials = d_part;
        pa.blob_cap16 = P.max_len16;
        pa.xs_cap = P.max_u;
        auto runq = [&](auto kern, const char *name, int threads, int stages)
        {
          const size_t smemq = 128 + stages * per;
          if (smemq > 227 * 1024)
          {
            std::printf("  %s: does not fit (%.1f KB)\n", name, smemq / 1024.0);
            return;
          }
          CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemq));
          const int gridq = std::min(H.ntiles, sms);
          float ms = time_it([&] { kern<<<gridq, threads, smemq>>>(pa); }, reps);
          check(name, ms);
        };
        if (m == 16)
        {
          runq(de::spmm_brb8q_kernel<2, 16, 2, false, false>, "  brb8q (VE) 16w 2st", 32 * 18, 2);
          runq(de::spmm_brb8q_kernel<2, 16, 3, false, false>, "  brb8q (VE) 16w 3st", 32 * 18, 3);
        }
        if (m == 32)
        {
          runq(de::spmm_brb8q_kernel<4, 12, 2, false, false>, "  brb8q (VE) 12w 2st", 32 * 14, 2);
          runq(de::spmm_brb8q_kernel<4, 16, 2, false, false>, "  brb8q (VE) 16w 2st", 32 * 18, 2);
          runq(de::spmm_brb8q_kernel<4, 8, 2, false, false>, "  brb8q (VE) 8w 2st", 32 * 10, 2);
          runq(de::spmm_brb8q_kernel<4, 12, 3, false, false>, "  brb8q (VE) 12w 3st", 32 * 14, 3);
          runq(de::spmm_brb8q_kernel<4, 16, 3, false, false>, "  brb8q (VE) 16w 3st", 32 * 18, 3);
        }
        cudaFree(d_td);
        cudaFree(d_blob);
      }
      cudaFree(d_tile); cudaFree(d_uc); cudaFree(d_bs); cudaFree(d_bv); cudaFree(d_rows); cudaFree(d_lc); cudaFree(d_sm); cudaFree(d_bval);
    }
  }
  return 0;
}
