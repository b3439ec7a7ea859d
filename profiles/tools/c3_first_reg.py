"""Diagnostic: where the first registered scan of a C3 stream spends its time (the call that showed 0.4 - 1.6 s in matching_mode 1).
usage: python profiles/tools/c3_first_reg.py <matching_mode>   (58 scans, init gate 50, same settings as bench.py --workload c3)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from loam_livox_b200 import capi, synthetic as S
from loam_livox_b200.registration import Context, Laser_mapping

mode = int(sys.argv[1]); n_total, init, LINE, PLANE, N_SCAN = 58, 50, 0.05, 0.1, 100_000
poses = S.trajectory(n_scans=1000, n_static=init + 1, speed=1.0, zero_mean_yaw=True, y0=-1.6)
ctx = Context(0, max_scan_points=N_SCAN, max_features=N_SCAN)
pipe = capi.PipelineCfg(pieces=3, use_piece=0, extractor_leaf_corner=LINE, extractor_leaf_surf=PLANE / 2, mapping_leaf_corner=LINE, mapping_leaf_surf=PLANE, whole_frame=1)
t0 = time.perf_counter()
gm = Laser_mapping(ctx, reg=capi.default_reg_state(mapping_init_accumulate_frames=init), pipeline=pipe, line_resolution=LINE, plane_resolution=PLANE,
                   matching_mode=mode, maximum_history_size=400, reserve_map_points=1 << 22, reserve_store_points=1 << 23)
print(f"mode {mode}: mapper created in {1e3 * (time.perf_counter() - t0):.1f} ms")
for k in range(n_total):
    raw = S.make_scan(N_SCAN, poses[k], seed=S.SEED + k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res, st = gm.process_new_scan(raw, 100.0 + 0.1 * k)
    dt = 1e3 * (time.perf_counter() - t0)
    if k >= 47 or dt > 10:
        print(f"scan {k}: {dt:8.2f} ms | front {st.ms_front_end:.2f} refresh {st.ms_refresh:.2f} register {st.ms_register:.2f} append {st.ms_append:.2f} | registered {res.registered} icp {res.icp_iterations} "
              f"blocks {res.num_residual_blocks} evals {res.total_evaluations} lm {res.total_lm_iterations} | gpu ms total {res.gpu_ms_total:.3f} knn {res.gpu_ms_knn_all:.3f} solve {res.gpu_ms_solve_all:.3f} sort {res.gpu_ms_sort:.3f} "
              f"| map {st.map_corner}+{st.map_surf} feats {st.n_corner}+{st.n_surf}", flush=True)
