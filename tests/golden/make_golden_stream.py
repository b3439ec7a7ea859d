"""Regenerates tests/golden/golden_stream.npz: the oracle's streaming mapper (Laser_mapping::process_new_scan restated, oracle.Mapper) on a 14-scan
synthetic sequence (24k-point scans, mapping_init_accumulate_frames 3), in matching_mode 0 with a window of 400 and of 3 clouds (the latter pops from
scan 4 on) and in matching_mode 1: pose, frame index, feature / map / append counts and ICP iteration count after every scan, plus a checksum of the
generated scans so that a change of the generator is told apart from a change of the mapper.  Run: python tests/golden/make_golden_stream.py"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from loam_livox_b200 import synthetic as S  # noqa: E402
from oracle import oracle as O  # noqa: E402

N_SCANS, N_PTS, INIT = 14, 24000, 3
CASES = [(0, 400), (0, 3), (1, 400)]


def scans():
    poses = S.trajectory(n_scans=N_SCANS, n_static=INIT + 1, speed=1.0)
    return [S.make_scan(N_PTS, p, seed=S.SEED + k) for k, p in enumerate(poses)]


def run(mode, window, raws, threads=4):
    om = O.Mapper(O.default_params(mapping_init_accumulate_frames=INIT, num_threads=threads), threads=threads, matching_mode=mode, maximum_history_size=window)
    rows = []
    for k, raw in enumerate(raws):
        st, q, t = om.process_scan(raw, 100.0 + 0.1 * k)
        L = om.last
        icp = L["res"].icp_iterations if L["res"] is not None and L["res"].registered else 0
        rows.append([st, om.frame_index, L["n_corner"], L["n_surf"], L["map_corner"], L["map_surf"], L["appended_corner"], L["appended_surf"], icp] + [float(x) for x in q] + [float(x) for x in t])
    return np.array(rows, dtype=np.float64)


def build():
    raws = scans()
    out = {"scan_crc": np.array([zlib.crc32(np.ascontiguousarray(r).tobytes()) for r in raws], dtype=np.uint32)}
    for mode, window in CASES:
        out[f"track_m{mode}_w{window}"] = run(mode, window, raws)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_stream.npz"), **build())
    print("written")
