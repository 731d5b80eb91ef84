// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for Dune::Timer (reference eigensolver.hh:221,255-257,343).
#ifndef DE_ORACLE_SHIM_TIMER_HH
#define DE_ORACLE_SHIM_TIMER_HH

#include <chrono>

namespace Dune
{
  class Timer
  {
    std::chrono::steady_clock::time_point t0_;

  public:
    Timer() { reset(); }
    void reset() { t0_ = std::chrono::steady_clock::now(); }
    double elapsed() const
    {
      return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count();
    }
  };
} // namespace Dune

#endif
