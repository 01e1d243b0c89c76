// CPU check of the resumable Hager-Zhang line search (csrc/lbfgs_ctl.h) against a direct, recursive transcription
// of the same algorithm on synthetic 1-D functions: both must ask for exactly the same sequence of steps and
// end at the same point.  Built and run by tests/test_host_logic.py (g++, no GPU).
#include <cstdio>
#include <functional>
#include <vector>

#include "../../pinn_based_online_pde_calculator_b200/csrc/lbfgs_ctl.h"

struct Phi { double a, f, d; };
typedef std::function<void(double, double&, double&)> Fn;

// direct transcription (nested lambdas, the shape of the round-1 host code)
static bool direct(const Fn& fn, double fcur, double dphi0, std::vector<double>& steps, Phi& c, int max_evals = 50) {
  int evals = 0;
  const double f_lim = fcur + 1e-6 * fabs(fcur), phi0 = fcur;
  auto wolfe = [&](const Phi& p) {
    const double delta = 0.1, sigma = 0.9;
    if (!(std::isfinite(p.f) && std::isfinite(p.d))) return false;
    const bool exact = (p.f <= phi0 + delta * p.a * dphi0) && (p.d >= sigma * dphi0);
    const bool approx = (p.f <= f_lim) && ((2 * delta - 1) * dphi0 >= p.d) && (p.d >= sigma * dphi0);
    return exact || approx;
  };
  auto ev = [&](double a, Phi& p) {
    steps.push_back(a);
    ++evals;
    p.a = a;
    fn(a, p.f, p.d);
    if (!std::isfinite(p.f)) { p.f = INFINITY; p.d = -1.0; }
    return wolfe(p);
  };
  Phi lo{0.0, fcur, dphi0}, hi{};
  bool found = false, ok = true;
  auto bisect = [&](Phi& A, Phi& B) -> bool {
    while (ok && evals < max_evals) {
      Phi d;
      if (ev(0.5 * (A.a + B.a), d)) { c = d; return true; }
      if (d.d >= 0) { B = d; return false; }
      if (d.f <= f_lim) A = d; else B = d;
      if (B.a - A.a <= 1e-16 * fmax(1.0, fabs(B.a))) break;
    }
    return false;
  };
  auto update = [&](Phi& A, Phi& B, const Phi& p) -> bool {
    if (!(p.a > A.a && p.a < B.a)) return false;
    if (p.d >= 0) { B = p; return false; }
    if (p.f <= f_lim) { A = p; return false; }
    Phi Bb = p;
    const bool f = bisect(A, Bb);
    B = Bb;
    return f;
  };
  {
    Phi prev = lo;
    double a = 1.0;
    bool bracketed = false;
    while (ok && evals < max_evals) {
      if (ev(a, c)) { found = true; break; }
      if (c.d >= 0) { lo = prev; hi = c; bracketed = true; break; }
      if (c.f > f_lim) {
        lo = Phi{0.0, fcur, dphi0}; hi = c;
        if (bisect(lo, hi)) found = true;
        bracketed = true;
        break;
      }
      prev = c;
      a *= 5.0;
    }
    if (!found && !bracketed) ok = false;
  }
  auto secant = [](const Phi& A, const Phi& B) { return (A.a * B.d - B.a * A.d) / (B.d - A.d); };
  while (ok && !found && evals < max_evals) {
    const Phi a0 = lo, b0 = hi;
    Phi p;
    double cs = secant(lo, hi);
    if (!std::isfinite(cs) || !(cs > lo.a && cs < hi.a)) cs = 0.5 * (lo.a + hi.a);
    if (ev(cs, p)) { c = p; found = true; break; }
    if (update(lo, hi, p)) { found = true; break; }
    double c2 = NAN;
    if (p.a == hi.a) c2 = secant(b0, hi);
    else if (p.a == lo.a) c2 = secant(a0, lo);
    if (std::isfinite(c2) && c2 > lo.a && c2 < hi.a && evals < max_evals) {
      Phi p2;
      if (ev(c2, p2)) { c = p2; found = true; break; }
      if (update(lo, hi, p2)) { found = true; break; }
    }
    if (hi.a - lo.a > 0.66 * (b0.a - a0.a) && evals < max_evals) {
      Phi pm;
      if (ev(0.5 * (lo.a + hi.a), pm)) { c = pm; found = true; break; }
      if (update(lo, hi, pm)) { found = true; break; }
    }
    if (hi.a - lo.a <= 1e-16 * fmax(1.0, hi.a)) break;
  }
  return found;
}

static bool resumable(const Fn& fn, double fcur, double dphi0, std::vector<double>& steps, Phi& c) {
  LbfgsCtl s{};
  s.ls_max_evals = 50;
  s.value_unnorm = 1;
  s.lref = 1.0;
  ls_begin(s, fcur, dphi0);
  int r = ls_resume(s, LsPhi{0, 0, 0});
  while (r == LS_REQUEST) {
    steps.push_back(s.a_next);
    double f, d;
    fn(s.a_next, f, d);
    s.evals += 1;
    r = ls_resume(s, ls_result(s, s.a_next, f, d));
  }
  c = Phi{s.c.a, s.c.f, s.c.d};
  return r == LS_FOUND;
}

int main() {
  std::vector<std::pair<const char*, Fn>> fns = {
      {"quadratic, minimum at 0.3", [](double a, double& f, double& d) { f = (a - 0.3) * (a - 0.3); d = 2 * (a - 0.3); }},
      {"quadratic, minimum at 40", [](double a, double& f, double& d) { f = 1e-3 * (a - 40) * (a - 40); d = 2e-3 * (a - 40); }},
      {"quartic with a flat start", [](double a, double& f, double& d) { const double t = a - 2.5; f = t * t * t * t - 3 * t * t; d = 4 * t * t * t - 6 * t; }},
      {"steep wall after 0.01", [](double a, double& f, double& d) { f = -a + 1e4 * a * a * a; d = -1 + 3e4 * a * a; }},
      {"oscillating", [](double a, double& f, double& d) { f = -0.2 * a + sin(7 * a) * 0.05 + 0.01 * a * a; d = -0.2 + 0.35 * cos(7 * a) + 0.02 * a; }},
      {"blows up (inf) beyond 0.5", [](double a, double& f, double& d) { if (a > 0.5) { f = INFINITY; d = NAN; } else { f = -a + a * a; d = -1 + 2 * a; } }},
      {"never satisfiable (monotone decrease, no curvature)", [](double a, double& f, double& d) { f = -a; d = -1; }},
      {"tiny minimum step", [](double a, double& f, double& d) { f = -1e-9 * a + a * a; d = -1e-9 + 2 * a; }},
  };
  int bad = 0;
  for (auto& kv : fns) {
    double f0, d0;
    kv.second(0.0, f0, d0);
    std::vector<double> s1, s2;
    Phi c1{}, c2{};
    const bool r1 = direct(kv.second, f0, d0, s1, c1);
    const bool r2 = resumable(kv.second, f0, d0, s2, c2);
    const bool same = r1 == r2 && s1 == s2 && (!r1 || (c1.a == c2.a && c1.f == c2.f && c1.d == c2.d));
    printf("%-55s found=%d evals=%zu step=%.17g  %s\n", kv.first, (int)r1, s1.size(), r1 ? c1.a : 0.0, same ? "same" : "DIFFERENT");
    if (!same) ++bad;
  }
  return bad;
}
