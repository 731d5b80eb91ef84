// dune-eigensolver driver program on the B200 drop-in headers.
//
// Re-creates the reference's executable (reference src/dune-eigensolver.cc:448-787) on top of
// include/dune/eigensolver/eigensolver.hh: same parameter file (sections / keys of src/dune-eigensolver.ini, `-ev.N 100`
// style command line overrides as dune-common's ParameterTreeParser::readOptions accepts), same matrix generators
// (:98-156), same driver calls, same machine-greppable output tables, and the thread-replica harness behind a barrier
// (:42-89, :756-773; `parallel.numthreads` replicas, each host thread with its own GPU context).
//
//   dune_eigensolver [-ini file] [-test largest|eigenvalues|smallest|mgs] [-section.key value ...]
//
//   largest      largest_eigenvalues_convergence_test (:620-730, what the shipped main() runs)
//   eigenvalues  eigenvalues_test (:448-525); ev.method = raes (GeneralizedInverse) | lobpcg (NEW: GeneralizedLOBPCG)
//   smallest     smallest_eigenvalues_convergence_test (:528-617)
//   mgs          mgs_performance_test (:164-311)
//
// New keys only: parallel.numgpus (row-partitioned multi-GPU run of the non-factored drivers), ev.method = lobpcg.
// ARPACK++ -- the reference's ground truth, not available -- is replaced by the analytic spectrum where one exists
// (:437-446) and by the driver's own run at tol 1e-13 otherwise; the ARPACK columns of the tables then hold that.
//
// The matrix type is a ~100-line BCRSMatrix stand-in with the iterator surface the headers use (at a DUNE site: the
// real dune-istl type, nothing else changes).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "simple_bcrs.hh"

#include <dune/eigensolver/eigensolver.hh>

using Matrix = Dune::BCRSMatrix<Dune::FieldMatrix<double, 1, 1>>;
using Vector = std::vector<double>;

// ---- parameter tree (sections, `key = value`, `#` comments; `-section.key value` overrides) -------------------------
struct ParameterTree
{
  std::map<std::string, std::string> kv;
  static std::string trim(const std::string &s)
  {
    const auto a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
  }
  void read_ini(const std::string &path)
  {
    std::ifstream in(path);
    if (!in)
      throw std::runtime_error("cannot open parameter file " + path);
    std::string line, section;
    while (std::getline(in, line))
    {
      line = trim(line.substr(0, line.find('#')));
      if (line.empty())
        continue;
      if (line.front() == '[' && line.back() == ']')
      {
        section = trim(line.substr(1, line.size() - 2));
        continue;
      }
      const auto eq = line.find('=');
      if (eq == std::string::npos)
        continue;
      kv[(section.empty() ? "" : section + ".") + trim(line.substr(0, eq))] = trim(line.substr(eq + 1));
    }
  }
  void read_options(int argc, char **argv)
  {
    for (int i = 1; i + 1 < argc; i += 2)
      if (argv[i][0] == '-')
        kv[argv[i] + 1] = argv[i + 1];
  }
  template <class T>
  T get(const std::string &key, T fallback) const
  {
    auto it = kv.find(key);
    if (it == kv.end())
      return fallback;
    std::istringstream is(it->second);
    T v;
    is >> v;
    return is.fail() ? fallback : v;
  }
  std::string get(const std::string &key, const char *fallback) const
  {
    auto it = kv.find(key);
    return it == kv.end() ? std::string(fallback) : it->second;
  }
};

// ---- the reference's matrices (src/dune-eigensolver.cc:98-156): 2D 5-point, N x N nodes, lexicographic ---------------
static Matrix laplacian(int N, const char *kind, int overlap)
{
  const std::string what(kind);
  std::vector<long> ptr(1, 0), col;
  std::vector<double> val;
  auto pu = [&](int x, int y) { return (x < overlap || x > N - 1 - overlap || y < overlap || y > N - 1 - overlap) ? 0.0 : 1.0; };
  for (int y = 0; y < N; ++y)
    for (int x = 0; x < N; ++x)
    {
      const int nb = (y > 0) + (x > 0) + (x < N - 1) + (y < N - 1);
      auto put = [&](int xx, int yy, double v) {
        if (what == "B") // partition-of-unity masked Laplacian (:124-143)
          v *= pu(x, y) * pu(xx, yy);
        if (what == "identity") // identity on the Laplacian pattern (:145-156)
          v = (xx == x && yy == y) ? 1.0 : 0.0;
        col.push_back(yy * N + xx);
        val.push_back(v);
      };
      const double diag = what == "neumann" ? (double)nb : 4.0; // Neumann: |sum of the off-diagonals| (:105-121)
      if (y > 0) put(x, y - 1, -1.0);
      if (x > 0) put(x - 1, y, -1.0);
      put(x, y, diag);
      if (x < N - 1) put(x + 1, y, -1.0);
      if (y < N - 1) put(x, y + 1, -1.0);
      ptr.push_back((long)col.size());
    }
  return Matrix((std::size_t)N * N, (std::size_t)N * N, ptr.data(), col.data(), val.data());
}

static std::vector<double> eigenvalues_laplace_dirichlet_2d(std::size_t N) // (:437-446), ascending
{
  std::vector<double> ev(N * N);
  const double h = 1.0 / (N + 1.0);
  for (std::size_t i = 0; i < N; ++i)
    for (std::size_t j = 0; j < N; ++j)
    {
      const double a = std::sin(0.5 * h * (i + 1) * M_PI), b = std::sin(0.5 * h * (j + 1) * M_PI);
      ev[j * N + i] = 4.0 * (a * a + b * b);
    }
  std::sort(ev.begin(), ev.end());
  return ev;
}

// ---- replica harness (:42-89): P threads run the same test behind a barrier -----------------------------------------
class Barrier
{
  std::mutex mu_;
  std::condition_variable cv_;
  int n_, waiting_ = 0;
  unsigned long generation_ = 0;

public:
  explicit Barrier(int n) : n_(n) {}
  int nthreads() const { return n_; }
  void wait()
  {
    std::unique_lock<std::mutex> lock(mu_);
    const unsigned long g = generation_;
    if (++waiting_ == n_)
    {
      waiting_ = 0;
      ++generation_;
      cv_.notify_all();
      return;
    }
    cv_.wait(lock, [&] { return generation_ != g; });
  }
};

struct Timer
{
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void reset() { t0 = std::chrono::steady_clock::now(); }
  double elapsed() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

static std::mutex g_print;

// ---- largest_eigenvalues_convergence_test (:620-730) ----------------------------------------------------------------
static int largest_eigenvalues_convergence_test(const ParameterTree &ptree, int rank, Barrier *barrier)
{
  const int N = ptree.get("ev.N", 200), m = ptree.get("ev.m", 4), maxiter = ptree.get("ev.maxiter", 4000);
  const double tol = ptree.get("ev.tol", 2e-3);
  const int verbose = ptree.get("ev.verbose", 0);
  const unsigned seed = ptree.get("ev.seed", 123u);
  const double shift = 0.0; // the reference forces it (:643)
  Matrix A = laplacian(N, "dirichlet", 0);
  const std::size_t n = A.N();
  std::vector<double> eval(m);
  std::vector<Vector> evec(m, Vector(n));
  // ground truth in the ARPACK columns: the analytic spectrum (largest m, descending)
  std::vector<double> analytic = eigenvalues_laplace_dirichlet_2d(N);
  std::reverse(analytic.begin(), analytic.end());
  barrier->wait();
  Timer timer;
  StandardLargest(A, shift, tol, maxiter, m, eval, evec, verbose, seed);
  const double time_eigensolver = timer.elapsed();
  barrier->wait();
  if (rank != 0)
    return 0;
  std::lock_guard<std::mutex> lock(g_print);
  std::vector<double> sorted(eval);
  std::sort(sorted.begin(), sorted.end(), std::greater<double>());
  std::cout << std::scientific << std::setprecision(6);
  for (int i = 0; i < m; ++i)
    std::cout << "EV " << i << "  " << std::setprecision(12) << sorted[i] << "  " << analytic[i] << "  " << std::setprecision(6)
              << std::abs(sorted[i] - analytic[i]) << std::endl;
  std::cout << ": eigensolver elapsed time " << time_eigensolver << std::endl;
  double maxerror = 0.0;
  for (int i = 0; i < m; ++i)
    maxerror = std::max(maxerror, std::abs(sorted[i] - analytic[i]));
  std::cout << "N_M_TOL_ESARERROR_ARPERROR_ESANERROR_TIMERATIO_ARPACKITER " << std::endl;
  std::cout << n << " & " << m << " & " << tol << " & " << maxerror << " & " << 0.0 << " & " << maxerror << " & "
            << "nan" << " & " << 0 << " \\\\" << std::endl;
  return 0;
}

// ---- eigenvalues_test (:448-525) and smallest_eigenvalues_convergence_test (:528-617) -------------------------------
static int eigenvalues_test(const ParameterTree &ptree, int rank, Barrier *barrier, bool convergence_table)
{
  const int N = ptree.get("ev.N", 200), overlap = ptree.get("ev.overlap", 3), m = ptree.get("ev.m", 4);
  const int maxiter = ptree.get("ev.maxiter", 4000), verbose = ptree.get("ev.verbose", 0);
  const double shift = ptree.get("ev.shift", 1e-3), reg = ptree.get("ev.regularization", 0.0), tol = ptree.get("ev.tol", 2e-3);
  const unsigned seed = ptree.get("ev.seed", 123u);
  const std::string method = ptree.get("ev.method", "raes");
  Matrix A = laplacian(N, "neumann", 0), B = laplacian(N, "B", overlap);
  const std::size_t n = A.N();
  std::vector<double> eval;
  std::vector<Vector> evec;
  barrier->wait();
  Timer timer;
  if (method == "raes")
    GeneralizedInverse(A, B, shift, reg, tol, maxiter, m, eval, evec, verbose, seed);
  else if (method == "lobpcg")
  {
    // NEW driver (the reference has no LOBPCG). It needs a positive definite B, and the partition-of-unity-masked
    // Laplacian of this test is only semi-definite: the smallest eigenpairs of the STANDARD problem for A + shift I are
    // computed instead (A = Neumann Laplacian is singular without the shift) and the shift is subtracted again.
    Matrix As(A);
    de_b200::add_to_diagonal(As, shift);
    eval.assign(m, 0.0);
    evec.assign(m, Vector(n));
    StandardLOBPCG(As, tol, maxiter, m, eval, evec, verbose, seed);
    for (double &e : eval)
      e -= shift;
  }
  else
    throw std::invalid_argument("ev.method must be raes or lobpcg (arpack is not available)");
  const double time_eigensolver = timer.elapsed();
  barrier->wait();
  if (rank != 0)
    return 0;
  std::lock_guard<std::mutex> lock(g_print);
  std::cout << std::scientific;
  for (int i = 0; i < m; ++i)
    std::cout << "EV " << i << "  " << std::setprecision(12) << eval[i] << std::endl;
  std::cout << std::setprecision(6) << ": eigensolver elapsed time " << time_eigensolver << std::endl;
  if (convergence_table && method == "raes")
  {
    // truth: the same driver at tol 1e-13 (stands in for ARPACK++ at 1e-14, :564-571)
    std::vector<double> truth;
    std::vector<Vector> tv;
    Timer t2;
    GeneralizedInverse(A, B, shift, reg, 1e-13, maxiter, m, truth, tv, 0, seed);
    const double time_truth = t2.elapsed();
    double maxerror = 0.0;
    for (int i = 0; i < m; ++i)
      maxerror = std::max(maxerror, std::abs(eval[i] - truth[i]));
    std::cout << "N_M_TOL_RASERROR_ARPERROR_TIMERATIO_ARPACKITER " << std::endl;
    std::cout << n << " & " << m << " & " << tol << " & " << maxerror << " & " << 0.0 << " & " << time_eigensolver / time_truth
              << " & " << 0 << " \\\\" << std::endl;
  }
  return 0;
}

// ---- mgs_performance_test (:164-311): GFLOP/s of orthonormalize_blocked with the reference's analytic models --------
static int mgs_performance_test(const ParameterTree &ptree, int rank, Barrier *barrier)
{
  const int n = ptree.get("mgs.n", 20), m = ptree.get("mgs.m", 16), n_iter = ptree.get("mgs.n_iter", 15);
  const int b = 8;
  MultiVector<double, 8> Q = de_b200::random_start_block(n, m, 123), W{(std::size_t)n, (std::size_t)m};
  auto &ctx = de_b200::Context::thread_default();
  de_b200::DeviceMV dQ(ctx, Q), dW(ctx, (std::size_t)n, (std::size_t)m);
  de_b200::check(de_mv_copy(dW.get(), dQ.get()), ctx.get());
  de_b200::check(de_orthonormalize(dW.get()), ctx.get()); // warm-up
  barrier->wait();
  Timer timer;
  for (int it = 0; it < n_iter; ++it)
  {
    de_b200::check(de_mv_copy(dW.get(), dQ.get()), ctx.get()); // every iteration orthonormalises the same random block
    de_b200::check(de_orthonormalize(dW.get()), ctx.get());
  }
  de_b200::check(de_context_synchronize(ctx.get()), ctx.get());
  const double time = timer.elapsed();
  barrier->wait();
  if (rank != 0)
    return 0;
  std::lock_guard<std::mutex> lock(g_print);
  const double P = barrier->nthreads();
  const double flops = P * n_iter * flops_orthonormalize(n, m);
  const double bytes = P * n_iter * bytes_orthonormalize_blocked(n, m, b, sizeof(double));
  const double bytesn = P * n_iter * bytes_orthonormalize_naive(n, m, sizeof(double));
  // reference columns: P n m AI_naive AI_blocked GF_naive GF_blocked [GF_simd]; the naive kernel does not exist here
  std::cout << "P_n_m_i_iblocked_perfn_perfb_perfv " << barrier->nthreads() << " " << n << " " << m << " " << flops / bytesn << " "
            << flops / bytes << " " << 0.0 << " " << flops / time * 1e-9 << " " << flops / time * 1e-9 << std::endl;
  return 0;
}

int main(int argc, char **argv)
{
  try
  {
    ParameterTree ptree;
    std::string ini = "dune-eigensolver.ini", test = "largest";
    for (int i = 1; i + 1 < argc; i += 2)
    {
      if (std::string(argv[i]) == "-ini")
        ini = argv[i + 1];
      if (std::string(argv[i]) == "-test")
        test = argv[i + 1];
    }
    ptree.read_ini(ini);
    ptree.read_options(argc, argv);
    const int numthreads = std::max(1, ptree.get("parallel.numthreads", 1));
    const int numgpus = std::max(1, ptree.get("parallel.numgpus", 1));
    if (numgpus > 1)
    {
      if (numthreads > 1)
        throw std::invalid_argument("parallel.numgpus > 1 and parallel.numthreads > 1 exclude each other");
      de_b200::Parallel::instance().set_num_gpus(numgpus);
      de_b200::Parallel::instance().set_row_align(ptree.get("ev.N", 200)); // one grid line of the 2D grid
    }
    Barrier barrier(numthreads);
    std::vector<int> rc(numthreads, 0);
    std::vector<std::string> err(numthreads);
    auto body = [&](int rank) {
      try
      {
        if (test == "largest")
          rc[rank] = largest_eigenvalues_convergence_test(ptree, rank, &barrier);
        else if (test == "eigenvalues")
          rc[rank] = eigenvalues_test(ptree, rank, &barrier, false);
        else if (test == "smallest")
          rc[rank] = eigenvalues_test(ptree, rank, &barrier, true);
        else if (test == "mgs")
          rc[rank] = mgs_performance_test(ptree, rank, &barrier);
        else
          throw std::invalid_argument("unknown -test " + test);
      }
      catch (const std::exception &e)
      {
        err[rank] = e.what();
        rc[rank] = 3;
      }
    };
    std::vector<std::thread> threads; // the replica harness (:756-773)
    for (int r = 1; r < numthreads; ++r)
      threads.emplace_back(body, r);
    body(0);
    for (auto &t : threads)
      t.join();
    for (int r = 0; r < numthreads; ++r)
      if (rc[r] != 0)
      {
        std::cerr << "replica " << r << ": " << err[r] << std::endl;
        return rc[r];
      }
    return 0;
  }
  catch (const std::exception &e)
  {
    std::cerr << "dune_eigensolver: " << e.what() << std::endl;
    return 3;
  }
}
