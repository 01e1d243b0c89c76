// L-BFGS controller shared by the host-driven and the device-resident optimiser loops
// (pinn_app/software.py:499-514: tfp.optimizer.lbfgs_minimize, m = 10, Hager-Zhang line search).
//
// The Hager & Zhang (2006, "Algorithm 851: CG_DESCENT") line search -- bracketing from step 1 with
// expansion factor 5, U3 bisection, secant^2 updates, (approximate) Wolfe tests with delta = 0.1,
// sigma = 0.9, eps = 1e-6, gamma = 0.66, at most 50 evaluations -- is written as a RESUMABLE state machine:
// ls_resume() consumes the result of the evaluation it asked for and runs until it needs the next one
// (LS_REQUEST, step in a_next), has found its point (LS_FOUND, point in c) or gives up (LS_FAIL).
// The host loop calls it between evaluations; the device loop calls the very same code from a one-thread
// kernel inside a CUDA-graph WHILE node, so both produce bit-identical iterates (the translation unit of
// the device kernels is compiled with -fmad=false: no contraction the host compiler would not do).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LB_HD __host__ __device__
#else
#define LB_HD
#endif

struct LsPhi { double a, f, d; };  // step, value, directional derivative

enum { LS_REQUEST = 0, LS_FOUND = 1, LS_FAIL = 2 };

struct LbfgsCtl {
  // ---- configuration
  int max_iter, m, value_unnorm, ls_max_evals, ring_cap, trace_cap;
  double tol, lref;
  // ---- optimiser state
  int iter, cnt, head, converged, failed, stop, total_evals, rows, started;
  double fcur, ginf;
  // ---- what the kernels of the current loop trip have to do
  int do_push;      // the line search accepted its point: push (s, y), move the iterate
  int init_eval;    // the evaluation in flight is the one at the initial point
  int need_dir;     // compute a new search direction and start the next line search
  double a_next;    // step of the evaluation to enqueue
  // ---- line-search state (survives between evaluations)
  int pc, evals, bracketed;
  double f_lim, phi0, dphi0, a, cs, c2;
  LsPhi lo, hi, c, prev, a0, b0, p, dd, bA, bB;
};

LB_HD inline bool ls_finite(double x) { return x == x && x - x == 0.0; }

LB_HD inline bool ls_wolfe(const LbfgsCtl& s, const LsPhi& p) {
  const double delta = 0.1, sigma = 0.9;
  if (!(ls_finite(p.f) && ls_finite(p.d))) return false;
  const bool exact = (p.f <= s.phi0 + delta * p.a * s.dphi0) && (p.d >= sigma * s.dphi0);
  const bool approx = (p.f <= s.f_lim) && ((2 * delta - 1) * s.dphi0 >= p.d) && (p.d >= sigma * s.dphi0);
  return exact || approx;
}
LB_HD inline double ls_secant(const LsPhi& A, const LsPhi& B) { return (A.a * B.d - B.a * A.d) / (B.d - A.d); }

// start the line search of one iteration: phi(0) = fcur, phi'(0) = dphi0; first trial step 1 (tfp initial_step_size)
LB_HD inline void ls_begin(LbfgsCtl& s, double fcur, double dphi0) {
  s.evals = 0;
  s.phi0 = fcur;
  s.dphi0 = dphi0;
  s.f_lim = fcur + 1e-6 * fabs(fcur);
  s.lo = LsPhi{0.0, fcur, dphi0};
  s.prev = s.lo;
  s.a = 1.0;
  s.bracketed = 0;
  s.pc = 0;
}

// U3 bisection on [bA, bB] (phi'(bA) < 0, phi(bA) <= f_lim, phi'(bB) < 0, phi(bB) > f_lim) as a coroutine fragment:
// leaves with `found` set when a Wolfe point turned up (in s.c), else with the shrunk interval in bA, bB
#define LS_BISECT(ID)                                                                         \
  found = false;                                                                              \
  while (s.evals < s.ls_max_evals) {                                                          \
    s.a_next = 0.5 * (s.bA.a + s.bB.a);                                                       \
    s.pc = ID;                                                                                \
    return LS_REQUEST;                                                                        \
    case ID:                                                                                  \
    s.dd = res;                                                                               \
    if (ls_wolfe(s, s.dd)) { s.c = s.dd; found = true; break; }                               \
    if (s.dd.d >= 0) { s.bB = s.dd; break; }                                                  \
    if (s.dd.f <= s.f_lim) s.bA = s.dd; else s.bB = s.dd;                                     \
    if (s.bB.a - s.bA.a <= 1e-16 * fmax(1.0, fabs(s.bB.a))) break;                            \
  }

// update(lo, hi, p) of Hager-Zhang (U0-U3); a nested bisection may find the Wolfe point
#define LS_UPDATE(ID)                                                                         \
  found = false;                                                                              \
  if (s.p.a > s.lo.a && s.p.a < s.hi.a) {                                                     \
    if (s.p.d >= 0) s.hi = s.p;                                                               \
    else if (s.p.f <= s.f_lim) s.lo = s.p;                                                    \
    else {                                                                                    \
      s.bA = s.lo; s.bB = s.p;                                                                \
      LS_BISECT(ID)                                                                           \
      s.lo = s.bA; s.hi = s.bB;                                                               \
    }                                                                                         \
  }

// res = result of the evaluation requested by the previous call (ignored on the first call after ls_begin)
LB_HD inline int ls_resume(LbfgsCtl& s, LsPhi res) {
  bool found = false;
  switch (s.pc) {
    case 0:
      // ---- bracket, starting from step 1
      while (s.evals < s.ls_max_evals) {
        s.a_next = s.a;
        s.pc = 1;
        return LS_REQUEST;
        case 1:
        s.c = res;
        if (ls_wolfe(s, s.c)) return LS_FOUND;
        if (s.c.d >= 0) { s.lo = s.prev; s.hi = s.c; s.bracketed = 1; break; }
        if (s.c.f > s.f_lim) {
          s.bA = LsPhi{0.0, s.phi0, s.dphi0}; s.bB = s.c;
          LS_BISECT(2)
          s.lo = s.bA; s.hi = s.bB;
          if (found) return LS_FOUND;
          s.bracketed = 1;
          break;
        }
        s.prev = s.c;
        s.a *= 5.0;
      }
      if (!s.bracketed) return LS_FAIL;
      // ---- secant^2 iterations
      while (s.evals < s.ls_max_evals) {
        s.a0 = s.lo; s.b0 = s.hi;
        s.cs = ls_secant(s.lo, s.hi);
        if (!ls_finite(s.cs) || !(s.cs > s.lo.a && s.cs < s.hi.a)) s.cs = 0.5 * (s.lo.a + s.hi.a);
        s.a_next = s.cs;
        s.pc = 3;
        return LS_REQUEST;
        case 3:
        s.p = res;
        if (ls_wolfe(s, s.p)) { s.c = s.p; return LS_FOUND; }
        LS_UPDATE(4)
        if (found) return LS_FOUND;
        s.c2 = NAN;
        if (s.p.a == s.hi.a) s.c2 = ls_secant(s.b0, s.hi);
        else if (s.p.a == s.lo.a) s.c2 = ls_secant(s.a0, s.lo);
        if (ls_finite(s.c2) && s.c2 > s.lo.a && s.c2 < s.hi.a && s.evals < s.ls_max_evals) {
          s.a_next = s.c2;
          s.pc = 5;
          return LS_REQUEST;
          case 5:
          s.p = res;
          if (ls_wolfe(s, s.p)) { s.c = s.p; return LS_FOUND; }
          LS_UPDATE(6)
          if (found) return LS_FOUND;
        }
        if (s.hi.a - s.lo.a > 0.66 * (s.b0.a - s.a0.a) && s.evals < s.ls_max_evals) {
          s.a_next = 0.5 * (s.lo.a + s.hi.a);
          s.pc = 7;
          return LS_REQUEST;
          case 7:
          s.p = res;
          if (ls_wolfe(s, s.p)) { s.c = s.p; return LS_FOUND; }
          LS_UPDATE(8)
          if (found) return LS_FOUND;
        }
        if (s.hi.a - s.lo.a <= 1e-16 * fmax(1.0, s.hi.a)) break;
      }
      return LS_FAIL;
  }
  return LS_FAIL;
}

// result of an evaluation as the line search sees it (loss_info[0] is the UN-normalised loss; the reference
// pairs it with the gradient of the normalised loss, software.py:479-490)
LB_HD inline LsPhi ls_result(const LbfgsCtl& s, double a, double loss0, double gdotd) {
  LsPhi r;
  r.a = a;
  r.f = s.value_unnorm ? loss0 : loss0 / s.lref;
  r.d = gdotd;
  if (!ls_finite(r.f)) { r.f = INFINITY; r.d = -1.0; }
  return r;
}

// after the accepted point was pushed (sy = s.y, ginf = |g_new|_inf): iteration bookkeeping and the convergence
// tests of tfp (tolerance on |g|_inf; x_tolerance = f_relative_tolerance = 0)
LB_HD inline void lb_after_push(LbfgsCtl& s, double sy, double ginf) {
  if (sy > 0.0 && ls_finite(sy)) {
    s.head = (s.head + 1) % s.m;
    s.cnt = s.cnt + 1 < s.m ? s.cnt + 1 : s.m;
  }
  s.ginf = ginf;
  const double fprev = s.fcur;
  s.fcur = s.c.f;
  s.iter += 1;
  if (ginf <= s.tol) s.converged = 1;
  if (s.fcur == fprev && s.c.a == 0.0) s.converged = 1;
}
