// State of the Levenberg-Marquardt machine (lm_core.cuh); plain structs, shared by the device code, the C-ABI diagnostics and the host test.
#pragma once
struct FnSample { double x, value, gradient; int value_valid, gradient_valid; };

// Levenberg-Marquardt state machine (ceres TrustRegionMinimizer restated), advanced after every grid-wide evaluation (identically by every CTA).
struct LmState {
  int phase, iteration, max_iterations, num_invalid, done, termination, last_successful, reuse_diagonal;
  int ls_iters, n_valid, total_iterations, total_evaluations;
  int pending, _pad;   // after lm_step: -1 nothing to start / 0 next iteration from the accepted point / 1 from the old point
  double x[7], x_norm, x_cost, g[6], H[21];
  double trial[7];
  double scaling[6], diagonal[6], radius, decrease_factor;
  double delta[6], model_cost_change, gd, dmax, ls_alpha;
  FnSample prev, cur;
  double x_best[7], minimum_cost, min_iter_cost, initial_cost, final_cost, last_gmax;
};

