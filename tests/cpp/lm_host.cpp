// The solver kernel's LM state machine (loam_livox_b200/csrc/lm_core.cuh) compiled for the host, with the kernel's orchestration of one LM step
// restated around it: both ComputeStep hypotheses evaluated ahead of the accept test, the gradient test of the candidate evaluated aside, the
// pending hand-over.  tests/test_lm_core.py drives it with the oracle's evaluations and compares the trajectory with the oracle's solver.
#include "../../loam_livox_b200/csrc/lm_core.cuh"

extern "C" {

struct LmHost { LmState L; double bound; };

// x0: start point (q x,y,z,w ; t).  Returns the first trial point (Plus(x0, 0)) in trial_out.
void lmh_init(LmHost* h, const double x0[7], double bound, int max_iterations, double trial_out[7]) {
  LmState& L = h->L; h->bound = bound;
  double z[6] = {0, 0, 0, 0, 0, 0}, tr[7];
  d_plus(x0, z, bound, tr);
  L = LmState();
  L.phase = 0; L.iteration = 0; L.max_iterations = max_iterations; L.num_invalid = 0; L.done = 0; L.termination = 0; L.last_successful = 1; L.reuse_diagonal = 0;
  L.ls_iters = 0; L.n_valid = 0; L.total_iterations = 0; L.total_evaluations = 0; L.pending = -1;
  for (int k = 0; k < 7; k++) { L.trial[k] = tr[k]; L.x_best[k] = tr[k]; trial_out[k] = tr[k]; }
}

// One evaluation arrived: sums = 21 JtJ (upper, row-major), 6 Jtr, cost, block count.  Exactly what lm_solve_kernel does between two evaluations.
// Returns 1 when the solve has ended (x_best in x_out), else 0 (next trial point in x_out).  speculative = 0 runs the plain serial path instead
// (no hypotheses, gradient test in line) so that the test can also check that both paths agree bit for bit.
int lmh_step(LmHost* h, const double sums[29], int speculative, double x_out[7]) {
  LmState& L = h->L; const double bound = h->bound;
  if (speculative) {
    StepIn in[2]; StepOut pre[2]; double gx[7], gg[6];
    for (int hyp = 0; hyp < 2; hyp++) lm_hypothesis(L, sums, hyp, in[hyp]);            // a private copy of the state BEFORE lm_step touches it
    for (int k = 0; k < 7; k++) gx[k] = L.trial[k]; for (int c = 0; c < 6; c++) gg[c] = -sums[21 + c];
    for (int hyp = 0; hyp < 2; hyp++) compute_step(in[hyp], bound, pre[hyp]);
    double pg[7]; d_plus(gx, gg, bound, pg); double gmax = 0; for (int k = 0; k < 7; k++) gmax = fmax(gmax, fabs(gx[k] - pg[k]));
    lm_step(L, sums, bound, true);
    if (L.pending == 0) L.last_gmax = gmax;
    if (L.pending >= 0) lm_next_iteration(L, bound, &pre[L.pending]);
  } else {
    lm_step(L, sums, bound, false);
    if (L.pending >= 0) lm_next_iteration(L, bound, nullptr);
  }
  for (int k = 0; k < 7; k++) x_out[k] = L.done ? L.x_best[k] : L.trial[k];
  return L.done;
}
void lmh_summary(const LmHost* h, double out[6]) { out[0] = h->L.initial_cost; out[1] = h->L.final_cost; out[2] = h->L.iteration; out[3] = h->L.termination; out[4] = h->L.total_evaluations; out[5] = h->L.n_valid; }
int lmh_sizeof() { return (int)sizeof(LmHost); }

}  // extern "C"
