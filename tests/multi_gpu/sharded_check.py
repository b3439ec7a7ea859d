"""Run under torchrun with >= 2 GPUs: sharded registration (every rank indexes only the cells it owns + the 1.42 m / 7.07 m halos of the match gates
and processes the features that fall into its cells; 29 sums all-reduced inside the solver kernel over peer memory, L1 norms and -- when the residual
cap binds -- block counts exchanged over peer memory) against the same registration on one GPU with the whole map.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu/sharded_check.py

Prints one line 'SHARDED_OK ...' from rank 0 on success; exits non-zero otherwise.  SURVEY.md 8(e): results may differ from one GPU only by
the fp64 re-association of <= 8 partial sums (the angle test is 1e-6 because 2*acos(|q1.q2|) resolves only ~sqrt(eps) = 1.5e-8 near identity)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from loam_livox_b200 import capi, synthetic as S
    from loam_livox_b200.distributed import connect
    from loam_livox_b200.registration import Context, Map, Point_cloud_registration
    pose = S.default_pose()
    mc, ms = S.make_map(40000, 400000)
    ok = True
    report = []
    # single-GPU results first (context not connected)
    solo = Context(local)
    m1 = Map(solo, mc, ms)
    cases = []
    for k in range(3):
        fc, fs = S.make_features(3000, 27000, pose, seed=S.SEED + 17 * k)
        guess = S.perturb_pose(pose, np.random.default_rng(40 + k), dt=0.08, dang_deg=1.5)
        reg = Point_cloud_registration(solo)
        reg.set_pose(guess.q, guess.t)
        st = reg.find_out_incremental_transfrom(m1, fc, fs)
        cases.append((fc, fs, guess, st, reg.result))
    # the residual-block cap of the shipped YAML on top (pre-skip + drop: the drop probability needs the block count of ALL ranks)
    for k in range(2):
        fc, fs = S.make_features(3000, 27000, pose, seed=S.SEED + 17 * k)
        guess = S.perturb_pose(pose, np.random.default_rng(50 + k), dt=0.08, dang_deg=1.5)
        reg = Point_cloud_registration(solo, maximum_allow_residual_block=200, rng_seed=3 + k)
        reg.set_pose(guess.q, guess.t)
        st = reg.find_out_incremental_transfrom(m1, fc, fs)
        cases.append((fc, fs, guess, st, reg.result, dict(maximum_allow_residual_block=200, rng_seed=3 + k)))
    # sharded: every rank indexes its cells + halo only, and owns the queries that fall into its cells
    ctx = Context(local)
    connect(ctx, rank, world, dist)
    m = Map(ctx, mc, ms, rank=rank, world=world, cell_size=2.0)
    info, owner = m.shard_info()
    sizes = torch.tensor([info.kept_corner, info.kept_surf], dtype=torch.int64, device="cuda")
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    ok = ok and (info.kept_surf < info.total_surf and info.kept_corner < info.total_corner)   # a real subset on every rank
    for rep in range(2):            # twice: the cross-GPU generation counters must survive a new registration
        for k, case in enumerate(cases):
            fc, fs, guess, st1, r1 = case[:5]
            reg = Point_cloud_registration(ctx, **(case[5] if len(case) > 5 else {}))
            reg.set_pose(guess.q, guess.t)
            st = reg.find_out_incremental_transfrom(m, fc, fs)
            r = reg.result
            dt = float(np.linalg.norm(np.array(r.t_w_curr) - np.array(r1.t_w_curr)))
            dq = float(S.quat_angle(np.array(r.q_w_curr), np.array(r1.q_w_curr)))
            same = st == st1 and r.icp_iterations == r1.icp_iterations and r.num_residual_blocks == r1.num_residual_blocks and dt < 1e-9 and dq < 1e-6 \
                and abs(r.inlier_threshold - r1.inlier_threshold) <= 1e-12 * max(1.0, abs(r1.inlier_threshold))
            ok = ok and same
            report.append((rep, k, st, r.icp_iterations, r.num_residual_blocks, r.corner_used + r.surf_used, dt, dq))
            # all ranks must hold the same pose bit for bit (fixed rank order in the one-shot reduce)
            mine = torch.tensor(list(r.q_w_curr) + list(r.t_w_curr), dtype=torch.float64, device="cuda")
            allp = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allp, mine)
            ok = ok and all(bool(torch.equal(allp[0], p)) for p in allp)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        for r in report:
            print("case rep=%d k=%d status=%d icp=%d blocks=%d owned_here=%d dt=%.3e dq=%.3e" % r)
        print("shard sizes (corner, surf) per rank of (%d, %d): %s" % (info.total_corner, info.total_surf, [tuple(int(v) for v in t.tolist()) for t in all_sizes]))
        print(("SHARDED_OK" if int(flag.item()) else "SHARDED_MISMATCH") + f" world={world}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
