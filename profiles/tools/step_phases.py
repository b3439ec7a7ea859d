"""Per-phase GPU time of the per-scan step (CUDA events inside the library, ll_reg_result.gpu_ms_*): python profiles/tools/step_phases.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, time
import torch
import bench as B
from loam_livox_b200 import capi
from loam_livox_b200.registration import Context, Map, scan_to_pose
mc, ms, scans, guesses, truths = B.make_inputs(0)
ctx = Context(0, max_scan_points=B.N_SCAN, max_features=B.N_SCAN)
m = Map(ctx, mc, ms)
pc = capi.PipelineCfg(**B.PIPE)
dev = [torch.from_numpy(s).cuda() for s in scans]
stream = torch.cuda.ExternalStream(ctx.stream(), device=0)
for rep in range(3):
    for k in range(len(scans)):
        st = capi.default_reg_state(q_w_last=guesses[k].q, t_w_last=guesses[k].t, q_w_curr=guesses[k].q, t_w_curr=guesses[k].t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0.record(stream)
        res, nc, ns = scan_to_pose(ctx, m, dev[k].data_ptr(), 100.0, pc, st, where=capi.LL_DEVICE, n=B.N_SCAN, fmt=capi.LL_FMT_XYZI16)
        e1.record(stream); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
        if rep == 2:
            tot = e0.elapsed_time(e1)
            print(f"scan {k}: feats {nc}+{ns} icp {res.icp_iterations} step {tot:.3f} ms (wall {wall:.3f}) | front-end {tot - res.gpu_ms_total:.3f} | register {res.gpu_ms_total:.3f}: sort {res.gpu_ms_sort:.3f} knn {res.gpu_ms_knn_all:.3f} (first {res.gpu_ms_knn:.3f}) solve {res.gpu_ms_solve_all:.3f} select {res.gpu_ms_select_all:.3f} other {res.gpu_ms_total - res.gpu_ms_sort - res.gpu_ms_knn_all - res.gpu_ms_solve_all - res.gpu_ms_select_all:.3f} | evals {res.total_evaluations}")
