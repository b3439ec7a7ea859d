// The Levenberg-Marquardt state machine of the solver kernel (ceres TrustRegionMinimizer + LevenbergMarquardtStrategy + the projected Armijo line
// search, restated; see solve.cu) as plain C++: compiled for the device by solve.cu and, unchanged, for the HOST by tests/cpp/lm_host.cpp, where a CPU test
// drives it with the oracle's evaluations (tests/test_lm_core.py) -- the control flow of the kernel's speculative step (both ComputeStep hypotheses,
// the deferred gradient test, the pending hand-over) is checked against the oracle's solver without a GPU.
#pragma once
#include <cfloat>
#include <cmath>
#ifdef __CUDACC__
#define LL_HD __device__
#define LL_INLINE __forceinline__
#define LL_NOINLINE __noinline__
#else
#define LL_HD
#define LL_INLINE inline
#define LL_NOINLINE
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::isfinite;
#endif

#include "lm_state.h"

// ------------------------------------------------------------------------------------------------ small algebra (one thread)
LL_HD void d_qmul(const double a[4], const double b[4], double o[4]) {  // (w,x,y,z)
  o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  o[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
  o[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}
LL_HD void d_qrot(const double q[4], const double v[3], double o[3]) {
  double ux = q[1], uy = q[2], uz = q[3], w = q[0];
  double cx = uy * v[2] - uz * v[1], cy = uz * v[0] - ux * v[2], cz = ux * v[1] - uy * v[0];
  cx += cx; cy += cy; cz += cz;
  o[0] = v[0] + w * cx + (uy * cz - uz * cy); o[1] = v[1] + w * cy + (uz * cx - ux * cz); o[2] = v[2] + w * cz + (ux * cy - uy * cx);
}
LL_HD double d_angdist(const double a[4], const double b[4]) {
  double bc[4] = {b[0], -b[1], -b[2], -b[3]}, d[4]; d_qmul(a, bc, d);
  return 2.0 * atan2(sqrt(d[1] * d[1] + d[2] * d[2] + d[3] * d[3]), fabs(d[0]));
}
// EigenQuaternionParameterization::Plus (x: q as x,y,z,w then t) + box projection of the t block
LL_HD void d_plus(const double x[7], const double delta[6], double bound, double out[7]) {
  double nd = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (nd > 0.0) {
    double sn, cs; sincos(nd, &sn, &cs);
    const double sbd = sn / nd;
    double dq[4] = {cs, sbd * delta[0], sbd * delta[1], sbd * delta[2]}, q[4] = {x[3], x[0], x[1], x[2]}, r[4];
    d_qmul(dq, q, r); out[0] = r[1]; out[1] = r[2]; out[2] = r[3]; out[3] = r[0];
  } else { out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3]; }
  for (int k = 0; k < 3; k++) { double v = x[4 + k] + delta[3 + k]; out[4 + k] = fmin(fmax(v, -bound), bound); }
}
LL_HD LL_INLINE bool d_chol6(const double A[6][6], const double b[6], double x[6]) {
  double L[6][6], inv[6];
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j < 6; j++) L[i][j] = 0;
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 6; i++) {
#pragma unroll
    for (int j = 0; j <= i; j++) {
      double s = A[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
      if (i == j) { if (!(s > 0.0)) ok = false; inv[i] = rsqrt(s); L[i][i] = s * inv[i]; } else L[i][j] = s * inv[j];   // one rsqrt instead of sqrt + reciprocal: the 6 pivots are the dependent chain of the solve
    }
  }
  if (!ok) return false;
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; i++) { double s = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) s -= L[i][k] * y[k]; y[i] = s * inv[i]; }
#pragma unroll
  for (int i = 5; i >= 0; i--) { double s = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) s -= L[k][i] * x[k]; x[i] = s * inv[i]; }
#pragma unroll
  for (int i = 0; i < 6; i++) if (!isfinite(x[i])) ok = false;
  return ok;
}
LL_HD LL_INLINE int hidx(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }  // i <= j, upper-triangular row-major

// ---- ceres line-search polynomial helpers (polynomial.cc), highest degree first ----
LL_HD double d_poly_eval(const double* p, int n, double x) { double v = 0; for (int i = 0; i < n; i++) v = v * x + p[i]; return v; }
LL_HD int d_fit_poly(const FnSample* s, int ns, double* coef) {  // returns number of coefficients
  int nc = 0; for (int i = 0; i < ns; i++) { if (s[i].value_valid) nc++; if (s[i].gradient_valid) nc++; }
  const int deg = nc - 1; double A[6][6], rhs[6]; int perm[6];
  int row = 0;
  for (int i = 0; i < ns; i++) {
    if (s[i].value_valid) { for (int j = 0; j <= deg; j++) A[row][j] = pow(s[i].x, (double)(deg - j)); rhs[row] = s[i].value; row++; }
    if (s[i].gradient_valid) { for (int j = 0; j <= deg; j++) A[row][j] = j < deg ? (deg - j) * pow(s[i].x, (double)(deg - j - 1)) : 0.0; rhs[row] = s[i].gradient; row++; }
  }
  for (int i = 0; i < nc; i++) perm[i] = i;
  for (int c = 0; c < nc; c++) {  // full pivoting
    int pr = c, pc = c; double best = 0;
    for (int i = c; i < nc; i++) for (int j = c; j < nc; j++) if (fabs(A[i][j]) > best) { best = fabs(A[i][j]); pr = i; pc = j; }
    if (best == 0) break;
    for (int j = 0; j < nc; j++) { double t = A[c][j]; A[c][j] = A[pr][j]; A[pr][j] = t; }
    { double t = rhs[c]; rhs[c] = rhs[pr]; rhs[pr] = t; }
    for (int i = 0; i < nc; i++) { double t = A[i][c]; A[i][c] = A[i][pc]; A[i][pc] = t; }
    { int t = perm[c]; perm[c] = perm[pc]; perm[pc] = t; }
    for (int i = c + 1; i < nc; i++) { double f = A[i][c] / A[c][c]; for (int j = c; j < nc; j++) A[i][j] -= f * A[c][j]; rhs[i] -= f * rhs[c]; }
  }
  double y[6];
  for (int i = nc - 1; i >= 0; i--) { double sacc = rhs[i]; for (int j = i + 1; j < nc; j++) sacc -= A[i][j] * y[j]; y[i] = sacc / A[i][i]; }
  for (int i = 0; i < nc; i++) coef[perm[i]] = y[i];
  return nc;
}
// real parts of all roots of p (degree n-1), Durand-Kerner for degree > 2
LL_HD int d_root_real_parts(const double* pin, int n, double* out) {
  double p[6]; int m = 0; bool lead = true;
  for (int i = 0; i < n; i++) { if (lead && pin[i] == 0.0) continue; lead = false; p[m++] = pin[i]; }
  int deg = m - 1; if (deg < 1) return 0;
  if (deg == 1) { out[0] = -p[1] / p[0]; return 1; }
  if (deg == 2) {
    double a = p[0], b = p[1], c = p[2], D = b * b - 4 * a * c, sD = sqrt(fabs(D));
    if (D >= 0) { if (b >= 0) { out[0] = (-b - sD) / (2.0 * a); out[1] = (2.0 * c) / (-b - sD); } else { out[0] = (2.0 * c) / (-b + sD); out[1] = (-b + sD) / (2.0 * a); } }
    else { out[0] = -b / (2.0 * a); out[1] = out[0]; }
    return 2;
  }
  double zr[5], zi[5], cr[6]; for (int i = 0; i <= deg; i++) cr[i] = p[i] / p[0];
  double rad = 0; for (int i = 1; i <= deg; i++) rad = fmax(rad, fabs(cr[i])); rad = 1.0 + rad;
  for (int i = 0; i < deg; i++) { double ang = 2.0 * 3.14159265358979323846 * i / deg + 0.4; zr[i] = rad * 0.5 * cos(ang); zi[i] = rad * 0.5 * sin(ang); }
  for (int it = 0; it < 500; it++) {
    double change = 0;
    for (int i = 0; i < deg; i++) {
      double nr = 0, ni = 0; for (int k = 0; k <= deg; k++) { double tr = nr * zr[i] - ni * zi[i] + cr[k], ti = nr * zi[i] + ni * zr[i]; nr = tr; ni = ti; }
      double dr = 1, di = 0; for (int j = 0; j < deg; j++) if (j != i) { double ar = zr[i] - zr[j], ai = zi[i] - zi[j]; double tr = dr * ar - di * ai, ti = dr * ai + di * ar; dr = tr; di = ti; }
      double den = dr * dr + di * di; if (den == 0) { dr = 1e-300; di = 0; den = 1e-600 > 0 ? 1e-300 * 1e-300 : DBL_MIN; }
      double qr = (nr * dr + ni * di) / den, qi = (ni * dr - nr * di) / den;
      zr[i] -= qr; zi[i] -= qi; change = fmax(change, sqrt(qr * qr + qi * qi));
    }
    if (change < 1e-15 * rad) break;
  }
  for (int i = 0; i < deg; i++) out[i] = zr[i];
  return deg;
}
LL_HD double d_minimize_interp(const FnSample* s, int ns, double x_min, double x_max) {
  double coef[6]; int n = d_fit_poly(s, ns, coef);
  double ox = (x_min + x_max) / 2.0, ov = d_poly_eval(coef, n, ox);
  double v = d_poly_eval(coef, n, x_min); if (v < ov) { ov = v; ox = x_min; }
  v = d_poly_eval(coef, n, x_max); if (v < ov) { ov = v; ox = x_max; }
  if (n <= 2) return ox;
  double der[5]; int deg = n - 1; for (int i = 0; i < deg; i++) der[i] = (deg - i) * coef[i];
  double roots[5]; int nr = d_root_real_parts(der, deg, roots);
  for (int i = 0; i < nr; i++) { double r = roots[i]; if (!(r >= x_min && r <= x_max)) continue; v = d_poly_eval(coef, n, r); if (v < ov) { ov = v; ox = r; } }
  return ox;
}

// ------------------------------------------------------------------------------------------------ LM state machine (one thread)
#define LM_FTOL 1e-6
#define LM_GTOL 1e-10
#define LM_PTOL 1e-8

LL_HD double lm_gmax(const LmState& L, double bound) {
  double ng[6], pg[7]; for (int c = 0; c < 6; c++) ng[c] = -L.g[c];
  d_plus(L.x, ng, bound, pg); double m = 0; for (int k = 0; k < 7; k++) m = fmax(m, fabs(L.x[k] - pg[k])); return m;
}
LL_HD void lm_finish(LmState& L, int termination) {
  L.done = 1; L.termination = termination; L.final_cost = fmin(L.initial_cost, L.min_iter_cost);
}
// LevenbergMarquardtStrategy::ComputeStep + the trial point of one LM iteration as a PURE function of its inputs, so that it can be evaluated
// ahead of the accept / reject decision for both outcomes (lm_hypothesis below).  One copy of the code (noinline): the speculative evaluation and
// the in-line one round identically.
struct StepIn { double H[21], g[6], x[7], scaling[6], diagonal[6], radius; int reuse_diagonal; };
struct StepOut { double delta[6], trial[7], diagonal[6], model_cost_change, gd, dmax; int valid; };
LL_HD LL_NOINLINE void compute_step(const StepIn& I, double bound, StepOut& O) {
  for (int c = 0; c < 6; c++) { if (I.reuse_diagonal) O.diagonal[c] = I.diagonal[c]; else { double d = I.H[hidx(c, c)] * I.scaling[c] * I.scaling[c]; O.diagonal[c] = fmin(fmax(d, 1e-6), 1e32); } }
  double A[6][6], rhs[6], Hs[6][6];
  for (int i = 0; i < 6; i++) { rhs[i] = I.g[i] * I.scaling[i]; for (int j = i; j < 6; j++) { double v = I.H[hidx(i, j)] * I.scaling[i] * I.scaling[j]; Hs[i][j] = v; Hs[j][i] = v; } }
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) A[i][j] = Hs[i][j];
  // Ceres forms lm_diagonal = sqrt(diagonal / radius) and the linear solver adds its square; diagonal / radius is the same number to an ulp and
  // saves 6 square roots and 5 divisions on the serial path of every LM iteration
  { const double inv_radius = 1.0 / I.radius; for (int c = 0; c < 6; c++) A[c][c] += O.diagonal[c] * inv_radius; }
  double step[6]; const bool solved = d_chol6(A, rhs, step);
  O.valid = 0; O.model_cost_change = 0; O.gd = 0; O.dmax = 0;
  if (solved) {
    for (int c = 0; c < 6; c++) step[c] = -step[c];
    // model_cost_change = -(J s)'(f + J s / 2) = -(s' J'f) - s' J'J s / 2
    double sg = 0, shs = 0; for (int i = 0; i < 6; i++) { sg += step[i] * rhs[i]; double r = 0; for (int j = 0; j < 6; j++) r += Hs[i][j] * step[j]; shs += step[i] * r; }
    O.model_cost_change = -sg - 0.5 * shs; O.valid = O.model_cost_change > 0.0 ? 1 : 0;
  }
  if (!O.valid) return;
  for (int c = 0; c < 6; c++) { O.delta[c] = step[c] * I.scaling[c]; O.gd += I.g[c] * O.delta[c]; O.dmax = fmax(O.dmax, fabs(O.delta[c])); }
  d_plus(I.x, O.delta, bound, O.trial);
}
LL_HD LL_INLINE double radius_after_success(double radius, double rel) {   // HandleSuccessfulStep
  const double c1 = 2.0 * rel - 1.0; radius = radius / fmax(1.0 / 3.0, 1.0 - c1 * c1 * c1); return fmin(1e16, radius);
}
// Inputs of the NEXT iteration's ComputeStep under hypothesis h, built from the state BEFORE the evaluation `sums` is digested:
// h = 0: the candidate is accepted (or this is iteration zero): x <- trial, H, g <- sums, radius grows, diagonal recomputed;
// h = 1: the candidate is rejected: x, H, g stay, radius shrinks, diagonal reused.
LL_HD void lm_hypothesis(const LmState& L, const double* sums, int h, StepIn& I) {
  if (L.phase == 0 || h == 0) {
    for (int i = 0; i < 21; i++) I.H[i] = sums[i]; for (int i = 0; i < 6; i++) I.g[i] = sums[21 + i]; for (int k = 0; k < 7; k++) I.x[k] = L.trial[k];
    if (L.phase == 0) { for (int c = 0; c < 6; c++) I.scaling[c] = 1.0 / (1.0 + sqrt(I.H[hidx(c, c)])); I.radius = 1e4; }
    else { for (int c = 0; c < 6; c++) I.scaling[c] = L.scaling[c]; I.radius = radius_after_success(L.radius, (L.x_cost - sums[27]) / L.model_cost_change); }
    for (int c = 0; c < 6; c++) I.diagonal[c] = 0.0; I.reuse_diagonal = 0;
  } else {
    for (int i = 0; i < 21; i++) I.H[i] = L.H[i]; for (int i = 0; i < 6; i++) I.g[i] = L.g[i]; for (int k = 0; k < 7; k++) I.x[k] = L.x[k];
    for (int c = 0; c < 6; c++) { I.scaling[c] = L.scaling[c]; I.diagonal[c] = L.diagonal[c]; }
    I.radius = L.radius / L.decrease_factor; I.reuse_diagonal = 1;
  }
}
// Starts LM iterations until one needs an evaluation (sets L.trial, phase = 1) or the solve terminates.  `pre`: the ComputeStep of the first
// iteration started here, already evaluated for exactly the state L is in (see lm_hypothesis), or null.
LL_HD void lm_next_iteration(LmState& L, double bound, const StepOut* pre) {
  for (;;) {
    if (L.iteration >= L.max_iterations) { lm_finish(L, 0); return; }
    if (L.last_successful && L.last_gmax <= LM_GTOL) { lm_finish(L, 1); return; }
    if (L.radius <= 1e-32) { lm_finish(L, 5); return; }
    L.iteration++; L.total_iterations++;
    StepOut tmp; const StepOut* so = pre; pre = nullptr;
    if (!so) {
      StepIn in; for (int i = 0; i < 21; i++) in.H[i] = L.H[i]; for (int i = 0; i < 6; i++) { in.g[i] = L.g[i]; in.scaling[i] = L.scaling[i]; in.diagonal[i] = L.diagonal[i]; }
      for (int k = 0; k < 7; k++) in.x[k] = L.x[k]; in.radius = L.radius; in.reuse_diagonal = L.reuse_diagonal;
      compute_step(in, bound, tmp); so = &tmp;
    }
    for (int c = 0; c < 6; c++) L.diagonal[c] = so->diagonal[c];
    L.reuse_diagonal = 1;
    if (!so->valid) {  // HandleInvalidStep
      if (++L.num_invalid >= 5) { lm_finish(L, 4); return; }
      L.radius = L.radius / L.decrease_factor; L.decrease_factor *= 2.0; L.reuse_diagonal = 1; L.last_successful = 0; continue;
    }
    L.num_invalid = 0;
    L.model_cost_change = so->model_cost_change; L.gd = so->gd; L.dmax = so->dmax;
    for (int c = 0; c < 6; c++) L.delta[c] = so->delta[c];
    for (int k = 0; k < 7; k++) L.trial[k] = so->trial[k];
    L.ls_iters = 0; L.prev.value_valid = 0; L.prev.gradient_valid = 0; L.ls_alpha = 1.0;
    L.phase = 1; return;
  }
}
// Candidate point L.trial evaluated: cost + sums (normal equations at the candidate).
LL_HD void lm_accept_test(LmState& L, const double* sums, double bound, bool defer_gmax) {
  const double cand_cost = sums[27];
  double step_norm = 0; for (int k = 0; k < 7; k++) step_norm += (L.x[k] - L.trial[k]) * (L.x[k] - L.trial[k]); step_norm = sqrt(step_norm);
  if (step_norm <= LM_PTOL * (L.x_norm + LM_PTOL)) { lm_finish(L, 2); return; }
  const double cost_change = L.x_cost - cand_cost;
  if (fabs(cost_change) <= LM_FTOL * L.x_cost) { lm_finish(L, 3); return; }
  const double rel = (L.x_cost - cand_cost) / L.model_cost_change;
  if (rel > 1e-3) {  // HandleSuccessfulStep
    for (int k = 0; k < 7; k++) L.x[k] = L.trial[k];
    double n = 0; for (int k = 0; k < 7; k++) n += L.x[k] * L.x[k]; L.x_norm = sqrt(n);
    L.x_cost = cand_cost; for (int i = 0; i < 21; i++) L.H[i] = sums[i]; for (int i = 0; i < 6; i++) L.g[i] = sums[21 + i];
    if (!defer_gmax) L.last_gmax = lm_gmax(L, bound);
    L.last_successful = 1;
    L.radius = radius_after_success(L.radius, rel);
    L.decrease_factor = 2.0; L.reuse_diagonal = 0;
    L.min_iter_cost = fmin(L.min_iter_cost, L.x_cost);
    if (L.x_cost < L.minimum_cost) { L.minimum_cost = L.x_cost; for (int k = 0; k < 7; k++) L.x_best[k] = L.x[k]; }
    L.pending = 0;
  } else {  // HandleUnsuccessfulStep
    L.radius = L.radius / L.decrease_factor; L.decrease_factor *= 2.0; L.reuse_diagonal = 1; L.last_successful = 0;
    L.min_iter_cost = fmin(L.min_iter_cost, cand_cost);
    L.pending = 1;
  }
}
// One evaluation finished; sums = normal equations at L.trial.  Digests it up to the point where the next LM iteration would start: L.pending = 0 / 1
// (next iteration from the accepted / the old point: the caller finishes with lm_next_iteration and the matching pre-computed step) or -1 (the solve
// ended, or the line search goes on and L.trial is its next sample).
// defer_gmax: the caller supplies L.last_gmax of the accepted point itself (it is evaluated on another warp meanwhile), whenever pending == 0.
LL_HD LL_NOINLINE void lm_step(LmState& L, const double* sums, double bound, bool defer_gmax) {
  L.total_evaluations++; L.pending = -1;
  if (L.phase == 0) {  // IterationZero
    for (int k = 0; k < 7; k++) { L.x[k] = L.trial[k]; L.x_best[k] = L.trial[k]; }
    double n = 0; for (int k = 0; k < 7; k++) n += L.x[k] * L.x[k]; L.x_norm = sqrt(n);
    L.x_cost = sums[27]; for (int i = 0; i < 21; i++) L.H[i] = sums[i]; for (int i = 0; i < 6; i++) L.g[i] = sums[21 + i];
    L.n_valid = (int)(sums[28] + 0.5);
    L.initial_cost = L.x_cost; L.min_iter_cost = L.x_cost; L.minimum_cost = L.x_cost; L.final_cost = L.x_cost;
    for (int c = 0; c < 6; c++) L.scaling[c] = 1.0 / (1.0 + sqrt(L.H[hidx(c, c)]));
    if (!defer_gmax) L.last_gmax = lm_gmax(L, bound);
    L.last_successful = 1; L.iteration = 0; L.radius = 1e4; L.decrease_factor = 2.0; L.reuse_diagonal = 0; L.num_invalid = 0;
    if (L.n_valid == 0) { lm_finish(L, -1); return; }
    if (!isfinite(L.x_cost)) { lm_finish(L, 4); return; }
    L.pending = 0; return;
  }
  if (L.phase == 1) {  // projected Armijo line search sample at ls_alpha (ArmijoLineSearch::DoSearch, CUBIC interpolation)
    double gt = 0; for (int c = 0; c < 6; c++) gt += sums[21 + c] * L.delta[c];
    L.cur.x = L.ls_alpha; L.cur.value = sums[27]; L.cur.gradient = gt; L.cur.value_valid = isfinite(sums[27]) ? 1 : 0; L.cur.gradient_valid = (L.cur.value_valid && isfinite(gt)) ? 1 : 0;
    if (L.cur.value_valid && !(L.cur.value > L.x_cost + 1e-4 * L.gd * L.cur.x)) {
      for (int c = 0; c < 6; c++) L.delta[c] *= L.cur.x;   // success: trial == Plus(x, alpha*delta) is the candidate
      lm_accept_test(L, sums, bound, defer_gmax); return;
    }
    bool failed = false; double step_size = 0;
    if (++L.ls_iters >= 20) failed = true;
    else {
      const double mn = 1e-3 * L.cur.x, mx = 0.6 * L.cur.x;
      if (!L.cur.value_valid) step_size = fmin(fmax(L.cur.x * 0.5, mn), mx);
      else {
        FnSample s[3]; int ns = 0;
        s[ns].x = 0; s[ns].value = L.x_cost; s[ns].gradient = L.gd; s[ns].value_valid = 1; s[ns].gradient_valid = 1; ns++;
        s[ns++] = L.cur; if (L.prev.value_valid) s[ns++] = L.prev;
        step_size = d_minimize_interp(s, ns, mn, mx);
      }
      if (step_size * L.dmax < 1e-9) failed = true;
    }
    if (!failed) { L.prev = L.cur; L.ls_alpha = step_size; double sd[6]; for (int c = 0; c < 6; c++) sd[c] = step_size * L.delta[c]; d_plus(L.x, sd, bound, L.trial); return; }
    // line search failed: delta stays; the candidate is Plus(x, delta)
    if (L.cur.x == 1.0) { lm_accept_test(L, sums, bound, defer_gmax); return; }
    d_plus(L.x, L.delta, bound, L.trial); L.phase = 2; return;
  }
  lm_accept_test(L, sums, bound, defer_gmax);  // phase 2
}

