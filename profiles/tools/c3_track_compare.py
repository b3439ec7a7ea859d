"""GPU vs oracle pose track of the 1000-scan C3 stream, matching_mode 0 (profiles/r2/c3_track_{gpu,oracle}_mode0.npy: q_wxyz, t after every scan; the GPU
file comes from `bench.py --workload c3 --matching-mode 0 --dump-poses`, the oracle file from profiles/tools/c3_oracle_full.py 0).  Prints the first scan at
which the translations differ by more than each threshold, and the difference at a few scans."""
import os
import numpy as np

d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "r2")
g, o = np.load(os.path.join(d, "c3_track_gpu_mode0.npy")), np.load(os.path.join(d, "c3_track_oracle_mode0.npy"))
dt = np.linalg.norm(g[:, 4:] - o[:, 4:], axis=1)
for th in (1e-12, 1e-9, 1e-6, 1e-3):
    idx = np.nonzero(dt > th)[0]
    print(f"first scan with |t_gpu - t_oracle| > {th:g} m: {int(idx[0]) if len(idx) else None}")
print("max over scans 0..636:", dt[:637].max())
for k in (80, 400, 636, 637, 638, 700, 999):
    print(f"scan {k}: {dt[k]:.3e} m")
