"""The oracle's streaming mapper over the WHOLE C3 sequence of bench.py --workload c3 (same trajectory, seeds, leaves, init gate), on the CPU:
prints the position error at the scans bench.py checkpoints, and the final pose, so the GPU run's numbers (profiles/r2/bench_c3_mode*.json) can be
compared at scan 1000 and not only at scan 80.  usage: python profiles/tools/c3_oracle_full.py <matching_mode> [n_total]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from loam_livox_b200 import synthetic as S
from oracle import oracle

mode = int(sys.argv[1]); n_total = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
N_SCAN, init, LINE, PLANE = 100_000, 50, 0.05, 0.1
poses = S.trajectory(n_scans=n_total, n_static=init + 1, speed=1.0, zero_mean_yaw=True, y0=-1.6)
om = oracle.Mapper(oracle.default_params(mapping_init_accumulate_frames=init, num_threads=8), threads=8, line_resolution=LINE, plane_resolution=PLANE,
                   extractor_leaf_corner=LINE, extractor_leaf_surf=PLANE / 2, matching_mode=mode, maximum_history_size=400)
R0, t0w = poses[0].R(), poses[0].t
out, t_start, track = {}, time.time(), []
for k in range(n_total):
    st, q, t = om.process_scan(S.make_scan(N_SCAN, poses[k], seed=S.SEED + k), 100.0 + 0.1 * k)
    track.append([float(x) for x in q] + [float(x) for x in t])
    if k + 1 in (80, 200, 500, n_total):
        out[k + 1] = float(np.linalg.norm(t - R0.T @ (poses[k].t - t0w)))
        print(k + 1, out[k + 1], f"{time.time() - t_start:.0f} s", flush=True)
np.save(f"/tmp/c3_oracle_track_mode{mode}.npy", np.array(track))
print(json.dumps({"matching_mode": mode, "oracle_position_error_m_at_scan": out, "final_pose": {"scan": n_total, "q_wxyz": [float(x) for x in q], "t": [float(x) for x in t]}}))
