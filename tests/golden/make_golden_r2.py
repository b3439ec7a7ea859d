"""Regenerates tests/golden/golden_r2.npz.  Two kinds of vectors:
  * flann_*: exact 5-NN answers of REAL FLANN (the KDTreeSingleIndex that pcl::KdTreeFLANN wraps, here through OpenCV's bundled copy, cv2.flann,
    leaf_max_size 15, exact, sorted) on the map / queries of golden_small.npz -- third-party golden vectors, not oracle output;
  * cap_*: the residual-block cap (point_cloud_registration.hpp:232-238,339-345,434-458) under the counter-based generator, from the oracle.
Run: python tests/golden/make_golden_r2.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from loam_livox_b200 import synthetic as S  # noqa: E402
from oracle import oracle as O  # noqa: E402


def build():
    import cv2
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz"))
    out = {}
    for name, cloud in (("surf", g["map_surf"]), ("corner", g["map_corner"])):
        index = cv2.flann_Index(np.ascontiguousarray(cloud[:, :3]), dict(algorithm=4, leaf_max_size=15))
        fi, fd = index.knnSearch(np.ascontiguousarray(g["knn_q"][:, :3]), 5, params=dict(checks=-1, eps=0.0, sorted=True))
        assert np.all(np.diff(fd, axis=1) > 0), "a tie: FLANN's order would be unspecified"
        out[f"flann_idx_{name}"], out[f"flann_d2_{name}"] = fi.astype(np.int32), fd.astype(np.float32)
    pose = S.default_pose()
    mc, ms = S.make_map(2000, 18000, seed=21)
    fc, fs = S.make_features(300, 2700, pose, seed=23)
    guess = S.perturb_pose(pose, np.random.default_rng(6))
    cap, seed = 100, 3
    p = O.default_params(q_w_last=guess.q, t_w_last=guess.t, q_w_curr=guess.q, t_w_curr=guess.t, maximum_allow_residual_block=cap, rng_seed=seed)
    tc, ts = O.KdTree(mc), O.KdTree(ms)
    blocks, src, ca, sa = O.build_blocks(mc, tc, ms, ts, fc, fs, p)
    slot = src[:, 1] + np.where(src[:, 0] == 1, fc.shape[0], 0)
    st, res, tr = O.register(mc, tc, ms, ts, fc, fs, p, want_trace=True)
    u = np.array([[O.lib().orc_cap_uniform(s, it, stream, i) for i in (0, 1, 7, 1000, 399999)] for s in (0, 3, -1) for it in (0, 5) for stream in (0, 1, 2)], np.float32)
    out.update(cap_map_corner=mc, cap_map_surf=ms, cap_feat_corner=fc, cap_feat_surf=fs, cap_guess_q=guess.q, cap_guess_t=guess.t, cap=cap, cap_seed=seed,
               cap_slots=slot.astype(np.int32), cap_corner_avail=ca, cap_surf_avail=sa, cap_status=st, cap_q=np.array(res.q_w_curr), cap_t=np.array(res.t_w_curr),
               cap_iters=res.icp_iterations, cap_blocks=res.num_residual_blocks, cap_blocks_per_iter=np.array([t.blocks_before_select for t in tr], np.int32), cap_uniform=u)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_r2.npz"), **build())
    print("written")
