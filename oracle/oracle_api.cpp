// ORACLE — TEST INFRASTRUCTURE ONLY. C entry points (ctypes) over the restatement headers in this directory.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
// PARITY UNPINNED (no golden vectors exist in the reference; PCL/Ceres/Eigen are not installed here).
#include <chrono>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "orc_cellmap.hpp"
#include "orc_cloud.hpp"
#include "orc_extract.hpp"
#include "orc_registration.hpp"

using namespace orc;

extern "C" {

int orc_hw_threads() {
#ifdef _OPENMP
  return omp_get_num_procs();   // not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 and the caller passes its team size explicitly
#else
  return 1;
#endif
}
float orc_cap_uniform(int seed, int icp_iter, int stream, int index) { return cap_uniform(seed, icp_iter, stream, index); }

// ---------------------------------------------------------------- VoxelGrid / kNN
int orc_voxel_grid(const float* in, int n, float leaf, float* out) { return voxel_grid(in, n, leaf, out); }

void* orc_kdtree_build(const float* pts4, int n) { KdTree* t = new KdTree(); t->build(pts4, n); return t; }
void orc_kdtree_free(void* t) { delete (KdTree*)t; }
// queries: nq x 4 floats. idx/d2: nq x k. found: nq.
void orc_kdtree_knn(void* tree, const float* q4, int nq, int k, int* idx, float* d2, int* found, int threads) {
  const KdTree* t = (const KdTree*)tree;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 256) if (threads > 1)
  for (int i = 0; i < nq; i++) {
    for (int j = 0; j < k; j++) { idx[(size_t)i * k + j] = -1; d2[(size_t)i * k + j] = std::numeric_limits<float>::infinity(); }
    found[i] = t->knn(q4 + (size_t)i * 4, k, idx + (size_t)i * k, d2 + (size_t)i * k);
  }
}
void orc_knn_brute(const float* map4, int n, const float* q4, int nq, int k, int* idx, float* d2, int* found) {
  for (int i = 0; i < nq; i++) {
    for (int j = 0; j < k; j++) { idx[(size_t)i * k + j] = -1; d2[(size_t)i * k + j] = std::numeric_limits<float>::infinity(); }
    found[i] = knn_brute(map4, n, q4 + (size_t)i * 4, k, idx + (size_t)i * k, d2 + (size_t)i * k);
  }
}

// ---------------------------------------------------------------- extractor
void* orc_extractor_create(float corner_curvature, float surface_curvature, float minimum_view_angle, float min_dis, float min_sigma) {
  Extractor* e = new Extractor();
  e->cfg.thr_corner_curvature = corner_curvature; e->cfg.thr_surface_curvature = surface_curvature; e->cfg.minimum_view_angle = minimum_view_angle;
  e->cfg.livox_min_allow_dis = min_dis; e->cfg.livox_min_sigma = min_sigma;
  return e;
}
void orc_extractor_free(void* e) { delete (Extractor*)e; }
int orc_extractor_extract(void* e, const float* raw4, int n, double stamp) { return ((Extractor*)e)->extract(raw4, n, stamp); }
// per-point arrays of length n (any may be null)
void orc_extractor_point_info(void* ev, int* pt_type, int* pt_label, float* curvature, float* view_angle, float* depth_sq2, float* time_stamp, float* polar_dis_sq2, int* polar_direction) {
  Extractor* e = (Extractor*)ev;
  for (size_t i = 0; i < e->info.size(); i++) {
    const PtInfo& p = e->info[i];
    if (pt_type) pt_type[i] = p.pt_type; if (pt_label) pt_label[i] = p.pt_label; if (curvature) curvature[i] = p.curvature; if (view_angle) view_angle[i] = p.view_angle;
    if (depth_sq2) depth_sq2[i] = p.depth_sq2; if (time_stamp) time_stamp[i] = p.time_stamp; if (polar_dis_sq2) polar_dis_sq2[i] = p.polar_dis_sq2; if (polar_direction) polar_direction[i] = p.polar_direction;
  }
}
int orc_extractor_split_idx(void* ev, int* out, int cap) { Extractor* e = (Extractor*)ev; int m = (int)e->split_idx.size(); for (int i = 0; i < m && i < cap; i++) out[i] = e->split_idx[i]; return m; }
int orc_extractor_scans(void* ev, int* first, int* last, int cap) { Extractor* e = (Extractor*)ev; int m = (int)e->scan_first_idx.size(); for (int i = 0; i < m && i < cap; i++) { first[i] = e->scan_first_idx[i]; last[i] = e->scan_last_idx[i]; } return m; }
void orc_extractor_piece_bounds(void* ev, int pieces, float* start, float* end) { ((Extractor*)ev)->piece_bounds(pieces, start, end); }
void orc_extractor_get_features(void* ev, float min_blur, float max_blur, float* corners, int* nc, float* surf, int* ns, float* full, int* nf) {
  ((Extractor*)ev)->get_features(min_blur, max_blur, corners, nc, surf, ns, full, nf);
}
double orc_extractor_current_time(void* ev) { return ((Extractor*)ev)->current_time; }

// ---------------------------------------------------------------- registration
// C mirror of RegParams/RegResult (plain doubles/ints so ctypes stays trivial)
struct orc_reg_params {
  int if_motion_deblur, current_frame_index, mapping_init_accumulate_frames, icp_max_iterations, cere_max_iterations, cere_prerun_times;
  int icp_plane, icp_line, maximum_allow_residual_block, num_threads, rng_seed, _pad;
  double para_max_angular_rate, para_max_speed, max_final_cost, minimum_pt_time_stamp, maximum_pt_time_stamp;
  double minimum_icp_R_diff, minimum_icp_T_diff, inliner_dis, inlier_ratio, maximum_dis_plane_for_match, maximum_dis_line_for_match, huber_a;
  double q_w_last[4], t_w_last[3], q_w_curr[4], t_w_curr[3], para_buffer_incremental[7];
};
struct orc_reg_result {
  int status, registered, num_residual_blocks, icp_iterations, corner_used, surf_used, total_lm_iterations, total_cost_evals, total_jac_evals, total_line_search_steps;
  double q_w_curr[4], t_w_curr[3], q_w_incre[4], t_w_incre[3], inlier_threshold, final_cost, initial_cost, angular_diff, t_diff;
  double seconds_knn_build, seconds_total;
};
static RegParams to_params(const orc_reg_params* p) {
  RegParams P; P.if_motion_deblur = p->if_motion_deblur; P.current_frame_index = p->current_frame_index; P.mapping_init_accumulate_frames = p->mapping_init_accumulate_frames;
  P.icp_max_iterations = p->icp_max_iterations; P.cere_max_iterations = p->cere_max_iterations; P.cere_prerun_times = p->cere_prerun_times; P.icp_plane = p->icp_plane; P.icp_line = p->icp_line;
  P.maximum_allow_residual_block = p->maximum_allow_residual_block; P.num_threads = p->num_threads < 1 ? 1 : p->num_threads; P.rng_seed = p->rng_seed;
  P.para_max_angular_rate = (float)p->para_max_angular_rate; P.para_max_speed = (float)p->para_max_speed; P.max_final_cost = (float)p->max_final_cost;
  P.minimum_pt_time_stamp = (float)p->minimum_pt_time_stamp; P.maximum_pt_time_stamp = (float)p->maximum_pt_time_stamp;
  P.minimum_icp_R_diff = p->minimum_icp_R_diff; P.minimum_icp_T_diff = p->minimum_icp_T_diff; P.inliner_dis = p->inliner_dis; P.inlier_ratio = p->inlier_ratio;
  P.maximum_dis_plane_for_match = p->maximum_dis_plane_for_match; P.maximum_dis_line_for_match = p->maximum_dis_line_for_match; P.huber_a = p->huber_a;
  for (int k = 0; k < 4; k++) { P.q_w_last[k] = p->q_w_last[k]; P.q_w_curr[k] = p->q_w_curr[k]; }
  for (int k = 0; k < 3; k++) { P.t_w_last[k] = p->t_w_last[k]; P.t_w_curr[k] = p->t_w_curr[k]; }
  for (int k = 0; k < 7; k++) P.para_buffer_incremental[k] = p->para_buffer_incremental[k];
  return P;
}
static void from_result(const RegResult& R, orc_reg_result* o) {
  o->status = R.status; o->registered = R.registered; o->num_residual_blocks = R.num_residual_blocks; o->icp_iterations = R.icp_iterations; o->corner_used = R.corner_used; o->surf_used = R.surf_used;
  o->total_lm_iterations = R.total_lm_iterations; o->total_cost_evals = R.total_cost_evals; o->total_jac_evals = R.total_jac_evals; o->total_line_search_steps = R.total_line_search_steps;
  for (int k = 0; k < 4; k++) { o->q_w_curr[k] = R.q_w_curr[k]; o->q_w_incre[k] = R.q_w_incre[k]; }
  for (int k = 0; k < 3; k++) { o->t_w_curr[k] = R.t_w_curr[k]; o->t_w_incre[k] = R.t_w_incre[k]; }
  o->inlier_threshold = R.inlier_threshold; o->final_cost = R.final_cost; o->initial_cost = R.initial_cost; o->angular_diff = R.angular_diff; o->t_diff = R.t_diff;
}

struct TraceC { int corner_avail, surf_avail, blocks_before_select, blocks_after_select, lm_iters1, lm_iters2; double inlier_threshold, x1[7], x2[7], cost1_initial, cost1_final, cost2_initial, cost2_final; };

// Registration against prebuilt trees (the 6-argument reference overload). trace may be null (cap entries).
int orc_register(const float* map_c, int nmc, void* tree_c, const float* map_s, int nms, void* tree_s,
                 const float* scan_c, int nc, const float* scan_s, int ns, const orc_reg_params* p, orc_reg_result* out, TraceC* trace, int trace_cap, int* trace_n) {
  auto t0 = std::chrono::steady_clock::now();
  Registration reg; reg.init(to_params(p)); RegResult R;
  int st = reg.find_out_incremental_transfrom(map_c, nmc, *(KdTree*)tree_c, map_s, nms, *(KdTree*)tree_s, scan_c, nc, scan_s, ns, &R);
  if (st < 0) { out->status = st; return st; }
  from_result(R, out); out->seconds_knn_build = 0;
  out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (trace_n) *trace_n = (int)reg.trace.size();
  if (trace) for (size_t i = 0; i < reg.trace.size() && (int)i < trace_cap; i++) {
    const IcpIterTrace& t = reg.trace[i]; TraceC& c = trace[i];
    c.corner_avail = t.corner_avail; c.surf_avail = t.surf_avail; c.blocks_before_select = t.blocks_before_select; c.blocks_after_select = t.blocks_after_select; c.lm_iters1 = t.lm_iters1; c.lm_iters2 = t.lm_iters2;
    c.inlier_threshold = t.inlier_threshold; for (int k = 0; k < 7; k++) { c.x1[k] = t.x_after_solve1[k]; c.x2[k] = t.x_after_solve2[k]; }
    c.cost1_initial = t.cost1_initial; c.cost1_final = t.cost1_final; c.cost2_initial = t.cost2_initial; c.cost2_final = t.cost2_final;
  }
  return st;
}

// Transform a cloud with pointAssociateToMap (non-deblur path): q (w,x,y,z), t.
void orc_transform(const float* in4, int n, const double q[4], const double t[3], float* out4) {
  Registration reg; RegParams P; for (int k = 0; k < 4; k++) P.q_w_curr[k] = q[k]; for (int k = 0; k < 3; k++) P.t_w_curr[k] = t[k]; reg.init(P);
  for (int i = 0; i < n; i++) reg.pointAssociateToMap(in4 + (size_t)i * 4, out4 + (size_t)i * 4);
}

// ---- pieces for step-by-step parity tests --------------------------------------------------------------
// Build the residual blocks of one ICP iteration at the pose in `p`; returns M. blocks_out: M x 11 doubles
// (type, p[3], a[3], v[3], s); src_out: M x 2 ints (0 corner / 1 surf, feature index).
int orc_build_blocks(const float* map_c, int nmc, void* tree_c, const float* map_s, int nms, void* tree_s, const float* scan_c, int nc, const float* scan_s, int ns,
                     const orc_reg_params* p, double* blocks_out, int* src_out, int cap, int* corner_avail, int* surf_avail) {
  Registration reg; reg.init(to_params(p)); std::vector<ResidualBlock> blocks;
  reg.build_blocks(map_c, *(KdTree*)tree_c, map_s, *(KdTree*)tree_s, scan_c, nc, scan_s, ns, blocks, corner_avail, surf_avail);
  for (size_t i = 0; i < blocks.size() && (int)i < cap; i++) {
    double* o = blocks_out + i * 11; const ResidualBlock& b = blocks[i];
    o[0] = b.type; for (int k = 0; k < 3; k++) { o[1 + k] = b.p[k]; o[4 + k] = b.a[k]; o[7 + k] = b.v[k]; } o[10] = b.motion_blur ? b.s : std::nan("");   // NaN = not a *_mb block (s itself may be negative: refine_blur does not clamp below 0)
    src_out[2 * i] = b.src; src_out[2 * i + 1] = b.src_index;
  }
  return (int)blocks.size();
}
static void fill_problem(Problem& prob, const double* blocks, int M, const double q_last[4], const double t_last[3], double huber_a, double bound) {
  prob.blocks.resize(M);
  for (int i = 0; i < M; i++) { const double* o = blocks + (size_t)i * 11; ResidualBlock& b = prob.blocks[i]; b.type = (int)o[0]; for (int k = 0; k < 3; k++) { b.p[k] = o[1 + k]; b.a[k] = o[4 + k]; b.v[k] = o[7 + k]; } b.motion_blur = !std::isnan(o[10]); b.s = b.motion_blur ? o[10] : 1.0; b.src = 0; b.src_index = i; }
  prob.q_last = {q_last[0], q_last[1], q_last[2], q_last[3]}; prob.t_last = {t_last[0], t_last[1], t_last[2]}; prob.huber_a = huber_a; prob.t_bound = bound;
}
// Evaluate at x: cost, gradient(6), JtJ (36, row-major, loss-corrected, unscaled), residuals (3M, optional), jac (18M, optional)
void orc_evaluate(const double* blocks, int M, const double q_last[4], const double t_last[3], double huber_a, double bound, const double x[7],
                  double* cost, double* gradient, double* jtj, double* residuals, double* jac) {
  Problem prob; fill_problem(prob, blocks, M, q_last, t_last, huber_a, bound);
  std::vector<double> r, J; prob.evaluate(x, cost, &r, gradient, &J);
  if (jtj) { for (int i = 0; i < 36; i++) jtj[i] = 0; for (size_t rr = 0; rr < r.size(); rr++) for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) jtj[i * 6 + j] += J[6 * rr + i] * J[6 * rr + j]; }
  if (residuals) std::memcpy(residuals, r.data(), r.size() * sizeof(double));
  if (jac) std::memcpy(jac, J.data(), J.size() * sizeof(double));
}
// ceres::Solve restatement on given blocks; x in/out. summary: [initial_cost, final_cost, iterations, successful, unsuccessful, line_search_steps, termination, cost_evals, jac_evals]
void orc_solve(const double* blocks, int M, const double q_last[4], const double t_last[3], double huber_a, double bound, int max_iter, double x[7], double* summary) {
  Problem prob; fill_problem(prob, blocks, M, q_last, t_last, huber_a, bound);
  SolveOptions so; so.max_num_iterations = max_iter; SolveSummary s; solve(prob, so, x, &s);
  summary[0] = s.initial_cost; summary[1] = s.final_cost; summary[2] = s.iterations; summary[3] = s.num_successful_steps; summary[4] = s.num_unsuccessful_steps; summary[5] = s.num_line_search_steps; summary[6] = s.termination; summary[7] = s.num_cost_evals; summary[8] = s.num_jac_evals;
}
double orc_inlier_threshold(const double* residuals, int M, double ratio) { std::vector<double> r(residuals, residuals + (size_t)3 * M); return Registration::inlier_residual_threshold(r, ratio); }
void orc_plus(const double x[7], const double delta[6], double bound, double out[7]) { Problem p; p.t_bound = bound; p.plus(x, delta, out); }

// ---------------------------------------------------------------- cell map (matching_mode 1)
void* orc_cellmap_create(float resolution, int revisit_threshold) { CellMap* m = new CellMap(); m->set_resolution(resolution); m->revisit_threshold = revisit_threshold; return m; }
void orc_cellmap_free(void* m) { delete (CellMap*)m; }
void orc_cellmap_append(void* m, const float* pts4, int n) { ((CellMap*)m)->append_cloud(pts4, n); }
int orc_cellmap_cells(void* m) { return (int)((CellMap*)m)->cells.size(); }
int orc_cellmap_points(void* m) { return ((CellMap*)m)->total_points(); }
int orc_cellmap_frame_idx(void* m) { return ((CellMap*)m)->current_frame_idx; }
// returns the number of points written (<= cap); *n_total = the number assemble produced
int orc_cellmap_assemble(void* m, const double q[4], const double t[3], float search_range, float fov_angle, float leaf, int replace, float* out4, int cap, int* n_total, int* cells_in_fov) {
  std::vector<float> out; Qd qq{q[0], q[1], q[2], q[3]}; V3d tt{t[0], t[1], t[2]};
  int n = ((CellMap*)m)->assemble(qq, tt, search_range, fov_angle, leaf, replace != 0, out, cells_in_fov);
  if (n_total) *n_total = n;
  int w = n < cap ? n : cap; if (w > 0) std::memcpy(out4, out.data(), (size_t)w * 16);
  return w;
}

}  // extern "C"
