"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: the one-time IPC-handle all-gather, the cell ownership that
splits the queries, and the fact that per-rank normal equations summed in rank order equal the single-rank ones."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from loam_livox_b200 import synthetic as S
    from loam_livox_b200.distributed import all_gather_handles, cell_owner, plan_shards
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # 1. handle exchange
    mine = bytes([rank + 1] * 64)
    blob = all_gather_handles(mine, world, dist)
    ok_handles = blob == b"".join(bytes([r + 1] * 64) for r in range(world))
    # 2. ownership partitions the queries, 3. the reduced normal equations equal the unsharded ones
    mc, ms = S.make_map(1000, 9000)
    pose = S.default_pose()
    fc, fs = S.make_features(100, 900, pose)
    guess = S.perturb_pose(pose, np.random.default_rng(0))
    p = O.default_params(q_w_last=guess.q, t_w_last=guess.t, q_w_curr=guess.q, t_w_curr=guess.t)
    blocks, src, _, _ = O.build_blocks(mc, O.KdTree(mc), ms, O.KdTree(ms), fc, fs, p)
    feats = np.concatenate([fc, fs])
    slot = src[:, 1] + np.where(src[:, 0] == 1, fc.shape[0], 0)
    world_pts = O.transform(feats, guess.q, guess.t)
    origin, dims, table = plan_shards(np.concatenate([mc, ms]), world, 2.0)   # the partition ll_map_build_sharded computes
    own = cell_owner(world_pts, origin, 2.0, dims, table)
    x = O.plus([0, 0, 0, 1, 0, 0, 0], [0.01, -0.02, 0.015, 0.05, -0.04, 0.03])
    mine_blocks = blocks[own[slot] == rank]
    c, g, H = O.evaluate(mine_blocks, guess.q, guess.t, x) if len(mine_blocks) else (0.0, np.zeros(6), np.zeros((6, 6)))
    part = torch.tensor(np.concatenate([[c], g, H.ravel(), [len(mine_blocks)]]))
    parts = [torch.zeros_like(part) for _ in range(world)]
    dist.all_gather(parts, part)
    total = sum(parts[r] for r in range(world)).numpy()      # fixed rank order, like the in-kernel reduction
    cf, gf, Hf = O.evaluate(blocks, guess.q, guess.t, x)
    ok_sum = abs(total[0] - cf) <= 1e-12 * cf and np.allclose(total[1:7], gf, rtol=1e-10, atol=1e-12) and np.allclose(total[7:43].reshape(6, 6), Hf, rtol=1e-10)
    ok_part = int(total[43]) == blocks.shape[0] and own.min() >= 0 and own.max() < world and len(set(own.tolist())) == world
    # 4. K10 in sharded mode: every rank fills a slot-indexed exchange array with the loss-corrected L1 norms of the blocks it owns (NaN elsewhere),
    #    the arrays are merged slot by slot (what l1_exchange_kernel does with NVLink stores), and the order statistic over the merged array equals
    #    the single-rank one (point_cloud_registration.hpp:153-161)
    def l1_of(b):
        r = O.evaluate(b, guess.q, guess.t, x, want_full=True)[3].reshape(-1, 3)
        return np.abs(r).sum(1)
    xl1 = np.full(feats.shape[0], np.nan)
    sel = own[slot] == rank
    if sel.any():
        xl1[slot[sel]] = l1_of(blocks[sel])
    gathered = [torch.zeros(feats.shape[0], dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor(xl1))
    stack = np.stack([t.numpy() for t in gathered])
    owners_per_slot = (~np.isnan(stack)).sum(0)
    merged = np.nanmin(np.where(np.isnan(stack), np.inf, stack), axis=0)
    full = np.full(feats.shape[0], np.inf); full[slot] = l1_of(blocks)
    thr_merged = O.inlier_threshold(np.repeat(merged[np.isfinite(merged)] / 3.0, 3), 0.8)    # three equal residuals whose L1 norm is the value
    thr_full = O.inlier_threshold(np.repeat(full[np.isfinite(full)] / 3.0, 3), 0.8)
    ok_k10 = owners_per_slot.max() == 1 and np.array_equal(merged, full) and thr_merged == thr_full
    q.put((rank, ok_handles, ok_sum, ok_part and ok_k10))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_host_protocol():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res


def _numpy_shard(points, origin, cell, dims, owner, rank, halo):
    """Reference restatement of sh_keep_kernel: the points within `halo` of a cell owned by `rank` (border cells reach to infinity)."""
    p = points[:, :3].astype(np.float64)
    keep = np.zeros(p.shape[0], bool)
    d = [int(v) for v in dims]
    for c in np.nonzero(owner == rank)[0]:
        ci = (c % d[0], (c // d[0]) % d[1], c // (d[0] * d[1]))
        d2 = np.zeros(p.shape[0])
        for a in range(3):
            lo = float(origin[a]) + ci[a] * cell
            hi = lo + cell
            e = np.zeros(p.shape[0])
            if ci[a] > 0:
                e = np.maximum(e, lo - p[:, a])
            if ci[a] < d[a] - 1:
                e = np.maximum(e, p[:, a] - hi)
            d2 += e * e
        keep |= d2 <= halo * halo
    return keep


def test_owner_plus_halo_shards_give_the_full_map_blocks():
    """with a halo of sqrt(gate) around the cells a rank owns, the residual blocks of the features that rank owns are exactly the
    blocks a search of the whole map gives (gates :254 / :353 compare SQUARED distances with 2.0 / 50.0).  Checked with the oracle on numpy shards,
    for the partition rule the library exports (ll_shard_plan): every cell has one owner, ranges are balanced, all features are covered once."""
    sys.path.insert(0, ROOT)
    from loam_livox_b200 import synthetic as S
    from loam_livox_b200.distributed import cell_owner, plan_shards
    from oracle import oracle as O
    mc, ms = S.make_map(3000, 30000)
    pose = S.default_pose()
    fc, fs = S.make_features(300, 2700, pose)
    guess = S.perturb_pose(pose, np.random.default_rng(3))
    p = O.default_params(q_w_last=guess.q, t_w_last=guess.t, q_w_curr=guess.q, t_w_curr=guess.t)
    full, src, ca, sa = O.build_blocks(mc, O.KdTree(mc), ms, O.KdTree(ms), fc, fs, p)
    key = {(int(s0), int(s1)): b for (s0, s1), b in zip(src, full)}
    for world, cell in ((2, 2.0), (4, 2.0), (8, 3.0)):
        origin, dims, owner = plan_shards(np.concatenate([mc, ms]), world, cell)
        counts = np.bincount(owner, minlength=world)
        assert owner.min() >= 0 and owner.max() == world - 1 and counts.min() > 0
        n_all = mc.shape[0] + ms.shape[0]
        per_rank = np.bincount(cell_owner(np.concatenate([mc, ms]), origin, cell, dims, owner), minlength=world)
        assert per_rank.sum() == n_all and per_rank.max() < 1.5 * n_all / world + 0.05 * n_all     # balanced by point count (up to one cell)
        own_c = cell_owner(O.transform(fc, guess.q, guess.t), origin, cell, dims, owner)
        own_s = cell_owner(O.transform(fs, guess.q, guess.t), origin, cell, dims, owner)
        seen = 0
        for r in range(world):
            kc = _numpy_shard(mc, origin, cell, dims, owner, r, 2.0 ** 0.5)
            ks = _numpy_shard(ms, origin, cell, dims, owner, r, 50.0 ** 0.5)
            assert ks.sum() < ms.shape[0] or world == 2            # the shard really is a subset (the 7.07 m halo covers a lot of a 42 m room)
            ic, is_ = np.nonzero(own_c == r)[0], np.nonzero(own_s == r)[0]
            shc, shs = mc[kc], ms[ks]
            if shc.shape[0] == 0 or shs.shape[0] < 5:
                continue
            blk, s2, _, _ = O.build_blocks(shc, O.KdTree(shc), shs, O.KdTree(shs), fc[ic], fs[is_], p)
            for (s0, s1), b in zip(s2, blk):
                g = (int(s0), int(ic[s1] if s0 == 0 else is_[s1]))
                assert g in key and np.array_equal(key[g], b, equal_nan=True), (world, r, g)
            seen += blk.shape[0]
        assert seen == full.shape[0]
