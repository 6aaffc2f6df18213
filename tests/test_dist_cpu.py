"""CPU tests of the N > 1 host logic (vapc_b200/dist.py) with world_size-2 gloo process groups: splitter
planning, destination partitioning and the all-to-all(v) row exchange, followed by a numpy Chan merge (the
checker) and comparison with the single-process oracle.  No CUDA kernels run here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vapc_oracle as O
from vapc_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def partial_table(x, y, z, vs, vmin, bits):
    """numpy stand-in of the local GPU stage: key-sorted partial rows (key, n, mean - c, M2)."""
    vx, vy, vz = O.voxelize(x, y, z, [0, 0, 0], vs)
    g = O.group_by_voxel(vx, vy, vz)
    key = ((g.vx - vmin[0]) << (bits[1] + bits[2])) | ((g.vy - vmin[1]) << bits[2]) | (g.vz - vmin[2])
    pts = np.stack([x, y, z], 1)[g.order]
    c = (np.stack([g.vx, g.vy, g.vz], 1) + 0.5) * vs
    n = g.counts.astype(np.float64)
    mean = np.add.reduceat(pts, g.starts, axis=0) / n[:, None]
    dev = pts - np.repeat(mean, g.counts, axis=0)
    prod = np.stack([dev[:, a] * dev[:, b] for a, b in ((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))], 1)
    m2 = np.add.reduceat(prod, g.starts, axis=0)
    assert np.all(np.diff(key) > 0)
    return key.astype(np.int64), g.counts.astype(np.int32), np.concatenate([mean - c, m2], axis=1)


def chan_merge(keys, count, mom):
    """Reference merge of partial rows with equal keys (source order), numpy."""
    order = np.argsort(keys, kind="stable")
    keys, count, mom = keys[order], count[order].astype(np.float64), mom[order]
    out_k, out_n, out_m = [], [], []
    i = 0
    idx6 = ((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))
    while i < len(keys):
        na, ma, Ma = count[i], mom[i, :3].copy(), mom[i, 3:].copy()
        j = i + 1
        while j < len(keys) and keys[j] == keys[i]:
            nb, mb, Mb = count[j], mom[j, :3], mom[j, 3:]
            n = na + nb
            d = mb - ma
            ma = ma + d * nb / n
            Ma = Ma + Mb + np.array([d[a] * d[b] for a, b in idx6]) * na * nb / n
            na = n
            j += 1
        out_k.append(keys[i]); out_n.append(na); out_m.append(np.concatenate([ma, Ma]))
        i = j
    return np.array(out_k), np.array(out_n), np.array(out_m)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        n, vs = 40000, 0.5
        pts = np.round(rng.uniform([0, 0, 0], [30, 20, 4], (n, 3)), 3)
        # disjoint point sets whose x-ranges overlap by 20 %, like bench.py's slabs: the middle fifth of the points
        # (sorted by x) is dealt alternately to the two ranks
        order = np.argsort(pts[:, 0], kind="stable")
        a, b = int(0.4 * n), int(0.6 * n)
        mid = order[a:b]
        mine = pts[np.concatenate([order[:a], mid[0::2]])] if rank == 0 else pts[np.concatenate([mid[1::2], order[b:]])]
        vmin = np.floor(pts.min(0) / vs).astype(np.int64)
        vmax = np.floor(pts.max(0) / vs).astype(np.int64)
        bits = [int(int(b - a).bit_length()) for a, b in zip(vmin, vmax)]
        total_bits = sum(bits)
        key, cnt, mom = partial_table(mine[:, 0], mine[:, 1], mine[:, 2], vs, vmin, bits)
        hb = min(D.HIST_BITS, total_bits)
        hist = torch.from_numpy(np.bincount(key >> (total_bits - hb), minlength=1 << hb).astype(np.int64))
        dist.all_reduce(hist)
        thresholds = D.plan_splitters(hist.numpy(), world, total_bits, hb)
        send = D.partition_counts(torch.from_numpy(key), thresholds)
        assert sum(send) == len(key)
        recv, recv_counts = D.exchange_rows({"keys": torch.from_numpy(key), "count": torch.from_numpy(cnt), "moments": torch.from_numpy(mom)},
                                            send)
        rk = recv["keys"].numpy()
        # every received key is in this rank's range
        t = [0] + thresholds + [1 << 62]
        assert np.all((rk >= t[rank]) & (rk < t[rank + 1]))
        mk, mn, mm = chan_merge(rk, recv["count"].numpy(), recv["moments"].numpy())
        # compare with the single-process oracle restricted to this rank's key range
        gk, gc, gm = partial_table(pts[:, 0], pts[:, 1], pts[:, 2], vs, vmin, bits)
        sel = (gk >= t[rank]) & (gk < t[rank + 1])
        assert np.array_equal(mk, gk[sel])
        assert np.array_equal(mn.astype(np.int64), gc[sel].astype(np.int64))
        np.testing.assert_allclose(mm[:, :3], gm[sel][:, :3], rtol=0, atol=1e-12)
        np.testing.assert_allclose(mm[:, 3:], gm[sel][:, 3:], rtol=1e-9, atol=1e-15)
        # reverse exchange (D.rows_from_owner): a column of the merged table comes back to every local partial row — the rows a
        # rank owns itself locally, the rows it sent away answered by their owners in one all-to-all
        off = np.concatenate([[0], np.cumsum(recv_counts)])
        remote = np.concatenate([rk[off[s]:off[s + 1]] for s in range(world) if s != rank]) if world > 1 else rk[:0]
        table = D.ShardedVoxelTable(grid=None, nvoxels=len(mk), voxel_keys=torch.from_numpy(mk.astype(np.int64)), count=None, mean3=None, m2=None,
                                    thresholds=thresholds, partial_rows_sent=0, partial_rows_received=len(remote))
        shard = D.LocalShard(cloud=None, gkeys=torch.from_numpy(key), lo=int(sum(send[:rank])), n_own=int(send[rank]),
                             send_remote=[0 if d == rank else int(c) for d, c in enumerate(send)],
                             recv_remote=[0 if d == rank else int(c) for d, c in enumerate(recv_counts)],
                             recv_keys=torch.from_numpy(remote.astype(np.int64)), group=None)
        cols = torch.from_numpy(np.stack([mn.astype(np.float64), mm[:, 0]]))  # merged count and mean_x of every owned voxel
        per_row = D.rows_from_owner(table, shard, cols).numpy()
        where = np.searchsorted(gk, key)
        assert np.array_equal(gk[where], key)
        assert np.array_equal(per_row[0].astype(np.int64), gc[where].astype(np.int64)), "every partial row sees its voxel's global count"
        np.testing.assert_allclose(per_row[1], gm[where, 0], rtol=0, atol=1e-12)
        # second epoch over the SAME key ranges (D.two_epoch_change): destination counts from the given thresholds, the
        # send-count matrix by one all-gather; afterwards the voxel join is local to the owner
        pts2 = np.round(np.concatenate([pts[: n // 2] + rng.normal(0, 0.05, (n // 2, 3)), rng.uniform([10, 0, 0], [30, 20, 4], (n // 8, 3))]), 3)
        pts2 = np.clip(pts2, pts.min(0), pts.max(0))  # same grid as epoch 1
        mine2 = pts2[rank::world]
        key2, cnt2, mom2 = partial_table(mine2[:, 0], mine2[:, 1], mine2[:, 2], vs, vmin, bits)
        counts = D.destination_counts(torch.from_numpy(key2), torch.tensor(thresholds, dtype=torch.int64))
        matrix = torch.empty(world * world, dtype=torch.int64)
        dist.all_gather_into_tensor(matrix, counts)
        matrix = matrix.view(world, world)
        assert matrix[rank].tolist() == D.partition_counts(torch.from_numpy(key2), thresholds)
        recv2, rc2 = D.exchange_rows({"keys": torch.from_numpy(key2), "count": torch.from_numpy(cnt2), "moments": torch.from_numpy(mom2)},
                                     matrix[rank].tolist())
        assert rc2 == matrix[:, rank].tolist()
        k2, _, _ = chan_merge(recv2["keys"].numpy(), recv2["count"].numpy(), recv2["moments"].numpy())
        assert np.all((k2 >= t[rank]) & (k2 < t[rank + 1]))
        g2, _, _ = partial_table(pts2[:, 0], pts2[:, 1], pts2[:, 2], vs, vmin, bits)
        joined_here = np.intersect1d(mk, k2)
        joined_global = np.intersect1d(gk, g2)
        assert np.array_equal(joined_here, joined_global[(joined_global >= t[rank]) & (joined_global < t[rank + 1])])
        ret[rank] = (len(mk), sum(send), sum(recv_counts), int(sel.sum()))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_and_merge_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    owned = [ret[r][0] for r in range(world)]
    assert all(o > 0 for o in owned)
    assert abs(owned[0] - owned[1]) < 0.35 * sum(owned), f"unbalanced key ranges {owned}"
    assert sum(ret[r][1] for r in range(world)) == sum(ret[r][2] for r in range(world))  # rows sent == rows received


def test_plan_splitters_properties():
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        for total_bits in (5, 16, 24, 40):
            hb = min(D.HIST_BITS, total_bits)
            hist = rng.integers(0, 50, 1 << hb)
            hist[rng.random(1 << hb) < 0.5] = 0
            t = D.plan_splitters(hist, world, total_bits, hb)
            assert len(t) == world - 1 and all(a <= b for a, b in zip(t, t[1:]))
            edges = [0] + [x >> (total_bits - hb) for x in t] + [1 << hb]
            loads = [hist[a:b].sum() for a, b in zip(edges[:-1], edges[1:])]
            assert sum(loads) == hist.sum()
            if world > 1 and hist.sum() > 0:
                assert max(loads) <= hist.sum() / world + hist.max() + 1
    assert D.plan_splitters(np.zeros(16, np.int64), 4, 4, 4) == sorted(D.plan_splitters(np.zeros(16, np.int64), 4, 4, 4))


def test_partition_counts_unsigned_keys():
    keys = torch.tensor([1, 5, 5, 9, 0x7FFFFFFF, 0x80000000, 0xFFFFFFF0], dtype=torch.int64).to(torch.int32)  # uint32 storage
    assert D.partition_counts(keys, [5, 0x80000000]) == [1, 4, 2]
    assert D.partition_counts(keys, []) == [7]
    assert D.bench_slab_shift(0, 50.0) == 0.0 and D.bench_slab_shift(2, 50.0) == 90.0


def test_device_side_planning_matches_host_planning():
    """voxel_statistics plans splitters and destinations with tensor ops on the device; on CPU tensors they must give
    exactly what the (numpy) host planner gives."""
    rng = np.random.default_rng(1)
    for world in (1, 2, 3, 4, 8):
        for total_bits in (5, 16, 24, 40, 63):
            hb = min(D.HIST_BITS, total_bits)
            hist = rng.integers(0, 50, 1 << hb)
            hist[rng.random(1 << hb) < 0.5] = 0
            t_host = D.plan_splitters(hist, world, total_bits, hb)
            t_dev = D.plan_splitters_tensor(torch.from_numpy(hist), world, total_bits, hb)
            assert t_dev.tolist() == t_host
            keys = np.sort(rng.integers(0, 1 << min(total_bits, 62), 5000))
            assert D.destination_counts(torch.from_numpy(keys), t_dev).tolist() == D.partition_counts(torch.from_numpy(keys), t_host)
    assert D.plan_splitters_tensor(torch.zeros(16, dtype=torch.int64), 4, 4, 4).tolist() == D.plan_splitters(np.zeros(16, np.int64), 4, 4, 4)


def test_plan_exchange_arithmetic_matches_host_planner():
    """The planning kernel's arithmetic (host mirror in the library) == plan_splitters + partition of every rank's rows."""
    from vapc_b200 import _lib as L

    rng = np.random.default_rng(4)
    for world in (1, 2, 3, 8, 16):
        for hb, total_bits in ((4, 4), (10, 24), (16, 27), (16, 40)):
            nb = 1 << hb
            hist_all = rng.integers(0, 40, (world, nb)).astype(np.uint32)
            hist_all[rng.random((world, nb)) < 0.6] = 0
            if world > 2:
                hist_all[1] = 0  # a rank without rows
            shift = max(total_bits - hb, 0)
            out = np.zeros(world * world + world - 1, dtype=np.int64)
            assert L.raw("vapc_selftest_plan_exchange_host")(hist_all.ctypes.data, world, nb, shift, out.ctypes.data) == 0
            thresholds = D.plan_splitters(hist_all.sum(0), world, total_bits, hb)
            assert out[world * world:].tolist() == thresholds
            edges = [0] + [t >> shift for t in thresholds] + [nb]
            for r in range(world):
                assert out[r * world:(r + 1) * world].tolist() == [int(hist_all[r, a:b].sum()) for a, b in zip(edges[:-1], edges[1:])]


# ---- the independent check bench.py --gpus N runs on its sharded table (tools/dist_check.py), world_size 2 over gloo ----------------
def _check_worker(rank, world, port, ret):
    from types import SimpleNamespace

    from tools.dist_check import verify_sharded_table

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        n, vs = 30000, 0.5
        pts = np.round(rng.uniform([0, 0, 0], [20, 12, 3], (n, 3)), 3)
        mine = pts[rank::world]  # every rank holds points of (almost) every voxel
        vmin = np.floor(pts.min(0) / vs).astype(np.int64)
        vmax = np.floor(pts.max(0) / vs).astype(np.int64)
        bits = [int(int(b - a).bit_length()) for a, b in zip(vmin, vmax)]
        grid = SimpleNamespace(voxel_size=vs, origin=[0.0, 0.0, 0.0], vmin=vmin.tolist(), bits=bits, total_bits=sum(bits))
        v = np.floor(pts / vs).astype(np.int64) - vmin
        key = (v[:, 0] << (bits[1] + bits[2])) | (v[:, 1] << bits[2]) | v[:, 2]
        uk, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
        inv = inv.ravel()
        threshold = int(uk[len(uk) // 2])
        sel = (uk < threshold) if rank == 0 else (uk >= threshold)
        mean = np.stack([np.bincount(inv, weights=pts[:, k], minlength=len(uk)) / cnt for k in range(3)], 1)
        d = pts - mean[inv]
        prod = [d[:, a] * d[:, b] for a, b in ((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))]
        with np.errstate(invalid="ignore", divide="ignore"):
            cov = np.stack([np.bincount(inv, weights=p, minlength=len(uk)) / (cnt - 1) for p in prod], 1)

        def table_and_stats(count):
            t = SimpleNamespace(grid=grid, voxel_keys=torch.from_numpy(uk[sel]), count=torch.from_numpy(count.astype(np.int32)), thresholds=[threshold])
            s = SimpleNamespace(cog=torch.from_numpy(np.ascontiguousarray(mean[sel].T)), cov=torch.from_numpy(np.ascontiguousarray(cov[sel].T)))
            return t, s

        x, y, z = (torch.from_numpy(np.ascontiguousarray(mine[:, k])) for k in range(3))
        good = verify_sharded_table(*table_and_stats(cnt[sel]), x, y, z, sample_rows=500)
        assert good["ok"], good
        assert good["voxels"] == len(uk) == good["voxels_torch_unique"] and good["points"] == n
        wrong = cnt[sel].copy()
        wrong[len(wrong) // 2] += 1
        bad = verify_sharded_table(*table_and_stats(wrong), x, y, z, sample_rows=500)
        assert not bad["ok"] and not bad["points_ok"]
        t, s = table_and_stats(cnt[sel])
        t.voxel_keys = t.voxel_keys[:-1]  # a voxel is missing
        t.count = t.count[:-1]
        s.cog, s.cov = s.cog[:, :-1], s.cov[:, :-1]
        bad = verify_sharded_table(t, s, x, y, z, sample_rows=500)
        assert not bad["ok"] and not bad["key_set_ok"]
        ret[rank] = good["voxels"]
    finally:
        dist.destroy_process_group()


def test_sharded_table_check_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_check_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world and len(set(ret.values())) == 1
