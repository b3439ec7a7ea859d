"""Regenerates tests/golden/golden_small.npz from the oracle (oracle/ = CPU restatement of the reference; the reference itself
ships no fixtures and cannot be built here — PARITY UNPINNED, see oracle/orc_math.hpp).  Run: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from loam_livox_b200 import synthetic as S  # noqa: E402
from oracle import oracle as O  # noqa: E402


def build():
    pose = S.default_pose()
    raw = S.make_scan(3000, pose, seed=7)
    ex = O.Extractor()
    n_scans = ex.extract(raw, 100.0)
    info = ex.point_info()
    c, s, f = ex.get_features(0.0, 1.0)
    ps, pe = ex.piece_bounds(3)
    mc, ms = S.make_map(400, 3600, seed=11)
    vg = O.voxel_grid(ms, 0.4)
    fc, fs = S.make_features(60, 540, pose, seed=13)
    q = O.transform(fs[:200], pose.q, pose.t)
    tc, ts = O.KdTree(mc), O.KdTree(ms)
    ki, kd, _ = ts.knn(q)
    guess = S.perturb_pose(pose, np.random.default_rng(5))
    p = O.default_params(q_w_last=guess.q, t_w_last=guess.t, q_w_curr=guess.q, t_w_curr=guess.t)
    st, res = O.register(mc, tc, ms, ts, fc, fs, p)
    blocks, src, ca, sa = O.build_blocks(mc, tc, ms, ts, fc, fs, p)
    x = O.plus([0, 0, 0, 1, 0, 0, 0], [0.01, -0.02, 0.015, 0.05, -0.04, 0.03])
    cost, g, H = O.evaluate(blocks, guess.q, guess.t, x)
    return dict(raw=raw, n_scans=n_scans, pt_type=info["pt_type"], pt_label=info["pt_label"], curvature=info["curvature"], view_angle=info["view_angle"],
                split_idx=ex.split_idx(), corners=c, surface=s, n_full=f.shape[0], piece_start=ps, piece_end=pe,
                map_corner=mc, map_surf=ms, voxel=vg, feat_corner=fc, feat_surf=fs, knn_q=q, knn_idx=ki, knn_d2=kd,
                guess_q=guess.q, guess_t=guess.t, reg_status=st, reg_q=np.array(res.q_w_curr), reg_t=np.array(res.t_w_curr), reg_iters=res.icp_iterations,
                reg_blocks=res.num_residual_blocks, reg_final_cost=res.final_cost, reg_inlier_thr=res.inlier_threshold, n_blocks=blocks.shape[0], corner_avail=ca,
                surf_avail=sa, eval_x=x, eval_cost=cost, eval_g=g, eval_H=H)


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz"), **build())
    print("written")
