// Stand-in for the NLopt C++ API surface BsplineOptimizer::optimize() touches (bspline_optimizer.cpp:165-227).
// NLopt 2.7.1 is a third-party dependency that is absent from the image; this is NOT an optimiser.  optimize() records
// what the reference hands to NLopt (start point after its clamp, bounds, stopping parameters) and evaluates the
// registered objective -- BsplineOptimizer::costFunction -> combineCost, the hot-path function -- at the start point
// and at every probe point the test supplies, keeping each (f, gradient).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <vector>

namespace nlopt {
typedef int algorithm;
typedef int result;
typedef double (*vfunc)(const std::vector<double>& x, std::vector<double>& grad, void* data);

struct Recorder {
  std::vector<double> x0, lb, ub;
  int maxeval = 0;
  double maxtime = 0, xtol_rel = 0;
  int alg = 0;
  std::vector<std::vector<double>> probes;   // in: extra points to evaluate
  std::vector<double> f;                     // out: f at x0, then at each probe
  std::vector<std::vector<double>> grad;     // out
};
inline Recorder& recorder() {
  static thread_local Recorder r;  // one per thread: the batch driver runs one optimiser per host thread
  return r;
}

class opt {
public:
  opt(algorithm a, unsigned n) : n_(n) { recorder().alg = a; }
  void set_min_objective(vfunc f, void* data) {
    f_ = f;
    data_ = data;
  }
  void set_maxeval(int n) { recorder().maxeval = n; }
  void set_maxtime(double t) { recorder().maxtime = t; }
  void set_xtol_rel(double t) { recorder().xtol_rel = t; }
  void set_lower_bounds(const std::vector<double>& v) { recorder().lb = v; }
  void set_upper_bounds(const std::vector<double>& v) { recorder().ub = v; }
  result optimize(std::vector<double>& x, double& opt_f) {
    Recorder& r = recorder();
    r.x0 = x;
    r.f.clear();
    r.grad.clear();
    std::vector<double> g(n_);
    opt_f = f_(x, g, data_);
    r.f.push_back(opt_f);
    r.grad.push_back(g);
    for (auto& p : r.probes) {
      std::vector<double> gp(n_);
      r.f.push_back(f_(p, gp, data_));
      r.grad.push_back(gp);
    }
    return 1;
  }

private:
  unsigned n_;
  vfunc f_ = nullptr;
  void* data_ = nullptr;
};
}  // namespace nlopt
