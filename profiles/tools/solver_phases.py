import sys, numpy as np, ctypes as C, time
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import bench
from loam_livox_b200 import capi
from loam_livox_b200.registration import Context, Map, scan_to_pose
import torch
mc, ms, scans, guesses, truths = bench.make_inputs(0)
ctx = Context(0, max_scan_points=bench.N_SCAN, max_features=bench.N_SCAN)
m = Map(ctx, mc, ms)
pc = capi.PipelineCfg(**bench.PIPE)
dev = [torch.from_numpy(s).cuda() for s in scans]
for rep in range(3):
    for k in range(len(scans)):
        st = capi.default_reg_state(q_w_last=guesses[k].q, t_w_last=guesses[k].t, q_w_curr=guesses[k].q, t_w_curr=guesses[k].t)
        res, nc, ns = scan_to_pose(ctx, m, dev[k].data_ptr(), 100.0, pc, st, where=capi.LL_DEVICE, n=bench.N_SCAN, fmt=capi.LL_FMT_XYZI16)
        out = (C.c_longlong * 16)()
        ctx.check(ctx._lib.ll_debug_solver_cycles(ctx.h, out))
        o = np.array(list(out), dtype=np.float64)
        if rep == 2:
            n = o[5]
            us = o / 1965.0
            print(f"scan {k}: icp {res.icp_iterations} evals {int(n)} solve_ms {res.gpu_ms_solve_all:.3f} | per eval us: eval {us[0]/n:.2f} wait {us[1]/n:.2f} reduce {us[2]/n:.2f} lm {us[3]/n:.2f} publish {us[4]/n:.2f} | per launch us: staging {us[6]/res.icp_iterations:.2f} L1+K10+tail {us[7]/res.icp_iterations:.2f} | K10 per launch us: l1+insert {us[8]/res.icp_iterations:.2f} bar {us[9]/res.icp_iterations:.2f} select {us[10]/res.icp_iterations:.2f} bar {us[11]/res.icp_iterations:.2f} drop {us[12]/res.icp_iterations:.2f} | LM step per eval us: compute_step (warp 1) {us[13]/n:.2f} lm_step (warp 0) {us[14]/n:.2f} both+sync {us[15]/n:.2f} | total accounted {us[:5].sum()/1e3 + (us[6]+us[7])/1e3:.3f} ms")
