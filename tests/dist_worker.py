"""Worker for tests/test_gpu_dist.py: run under torch.distributed.run with one rank per GPU.  Every rank holds a
chunk of one seeded cloud; the sharded multi-GPU voxel table must equal the single-GPU table of the whole cloud
(integers exactly, fp64 within 1e-9), independent of the number of ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gather_cat(a):
    """Concatenate a per-rank numpy array over the ranks, in rank order (= global key order)."""
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, a)
    return np.concatenate(parts, axis=-1)


def two_epoch(rank, world, dev, pts, rng, vs, origin):
    """Multi-GPU change detection (shared grid + shared key ranges, owner-local join) == the single-GPU flow."""
    from vapc_b200 import dist as D
    from vapc_b200 import engine as E

    n = len(pts)
    e2 = np.round(np.concatenate([pts[: n * 3 // 4] + rng.normal(0, 0.02, (n * 3 // 4, 3)), rng.uniform([40, 30, 0], [75, 50, 6], (n // 8, 3))]), 3)
    chunks = []
    for e, order in ((pts, np.argsort(pts[:, 0], kind="stable")), (e2, rng.permutation(len(e2)))):  # epoch 1 coherent, epoch 2 shuffled
        c = e[order[rank * len(e) // world:(rank + 1) * len(e) // world]]
        chunks.append([torch.from_numpy(np.ascontiguousarray(c[:, k])).to(dev) for k in range(3)])
    res = D.two_epoch_change(chunks[0], chunks[1], vs, origin, alpha=0.005)
    # single-GPU flow on every rank
    only = dict(cog=True, cov=True)
    refs = [E.VoxelCloud.build(e[:, 0].copy(), e[:, 1].copy(), e[:, 2].copy(), vs, origin) for e in (pts, e2)]
    i1, i2, only1, only2 = E.join_voxel_tables(refs[0].voxel_coords(), refs[1].voxel_coords())
    s1, s2 = refs[0].stats(**only), refs[1].stats(**only)
    ref = E.mahalanobis(s1.cog, s1.cov, s2.cog, s2.cov, i1, i2, alpha=0.005)
    vox = lambda cloud, rows: torch.stack(cloud.voxel_coords())[:, rows].cpu().numpy()  # noqa: E731
    got_pairs = gather_cat(torch.stack(res["table1"].voxel_coords())[:, res["i1"]].cpu().numpy())
    assert np.array_equal(got_pairs, vox(refs[0], i1)), "joined voxels"
    assert np.array_equal(gather_cat(torch.stack(res["table2"].voxel_coords())[:, res["i2"]].cpu().numpy()), vox(refs[1], i2)), "joined voxels (epoch 2)"
    assert np.array_equal(gather_cat(torch.stack(res["table1"].voxel_coords())[:, res["only1"]].cpu().numpy()), vox(refs[0], only1)), "disappeared"
    assert np.array_equal(gather_cat(torch.stack(res["table2"].voxel_coords())[:, res["only2"]].cpu().numpy()), vox(refs[1], only2)), "appeared"
    empty = torch.empty(0, dtype=torch.float64, device=dev)
    dd = gather_cat(res.get("distance", empty).cpu().numpy())
    rd = ref["distance"].cpu().numpy()
    assert np.all(np.abs(dd - rd) <= 1e-9 * np.abs(rd) + 1e-12), "distance"
    p, pr = gather_cat(res.get("p_value", empty).cpu().numpy()), ref["p_value"].cpu().numpy()
    assert np.array_equal(np.isnan(p), np.isnan(pr))
    ct = gather_cat(res.get("change_type", torch.empty(0, dtype=torch.int32, device=dev)).cpu().numpy())
    rct = ref["change_type"].cpu().numpy()
    # classes are exact away from the alpha threshold and from p == 0 (covariances agree to 1e-9 only)
    d1, r1 = gather_cat(res.get("d1", empty).cpu().numpy()), ref["d1"].cpu().numpy()
    d2, r2 = gather_cat(res.get("d2", empty).cpu().numpy()), ref["d2"].cpu().numpy()
    with np.errstate(invalid="ignore"):
        stable = (~np.isnan(pr) & (np.abs(pr - 0.005) > 1e-4) & (pr > 1e-300) & (np.abs(d1 - r1) <= 1e-6 * np.abs(r1))
                  & (np.abs(d2 - r2) <= 1e-6 * np.abs(r2)))
    assert stable.sum() > 1000, int(stable.sum())
    assert np.array_equal(ct[stable], rct[stable]), "change_type"
    assert np.all(np.abs(p[stable] - pr[stable]) <= 1e-5 * pr[stable] + 1e-12), "p_value"
    if rank == 0:
        print(f"[two-epoch] world={world}: {len(rd)} joined voxels, {len(only1)} disappeared, {len(only2)} appeared, "
              f"{int((rct == 1).sum() + (rct == 2).sum())} significant", flush=True)


def per_point_outputs(rank, world, dev, pts, rng, vs, origin):
    """Reverse exchange: per-point distance to the GLOBAL centroid, per-point broadcast of the global point count, and the
    closest-to-centroid subsample over all ranks == the single-GPU results."""
    from vapc_b200 import dist as D
    from vapc_b200 import engine as E

    n = len(pts)
    cloud = pts[rng.permutation(n)]  # chunks of a shuffled cloud: every voxel is shared by several ranks
    a, b = rank * n // world, (rank + 1) * n // world
    x, y, z = (torch.from_numpy(np.ascontiguousarray(cloud[a:b, k])).to(dev) for k in range(3))
    table, st, shard = D.voxel_statistics(x, y, z, vs, origin, keep_local=True)
    ref = E.VoxelCloud.build(cloud[:, 0].copy(), cloud[:, 1].copy(), cloud[:, 2].copy(), vs, origin)
    # per-point columns
    dist_pt = D.point_distance(table, shard, st.cog).cpu().numpy()
    ref_pt = ref.point_distance().cpu().numpy()[a:b]
    assert np.all(np.abs(dist_pt - ref_pt) <= 1e-9 * ref_pt + 1e-12), "distance to the global centroid"
    cnt_pt = D.broadcast_to_points(table, shard, table.count.to(torch.float64)[None, :])[0].cpu().numpy()
    ref_cnt = ref.broadcast(ref.count).cpu().numpy()[a:b]
    assert np.array_equal(cnt_pt.astype(np.int64), ref_cnt.astype(np.int64)), "global point_count per point"
    # the same per-point outputs with another axis leading the key (z, x, y): results do not depend on the internal order
    t2, s2, sh2 = D.voxel_statistics(x, y, z, vs, origin, keep_local=True, axis_order=(2, 0, 1))
    d2 = D.point_distance(t2, sh2, s2.cog).cpu().numpy()
    assert np.all(np.abs(d2 - ref_pt) <= 1e-9 * ref_pt + 1e-12), "distance to the global centroid, z-leading key"
    md2, gi2 = D.closest_to_cog(t2, sh2, s2.cog)
    v2 = gather_cat(torch.stack(t2.voxel_coords()).cpu().numpy())
    o2 = np.lexsort((v2[2], v2[1], v2[0]))
    rmd2 = ref.closest_to_cog()[0].cpu().numpy()
    md2 = gather_cat(md2.cpu().numpy())[o2]
    assert np.all(np.abs(md2 - rmd2) <= 1e-9 * rmd2 + 1e-12), "min_distance, z-leading key"
    del t2, s2, sh2
    # closest-to-centroid subsample
    md, gi = D.closest_to_cog(table, shard, st.cog)
    md, gi = gather_cat(md.cpu().numpy()), gather_cat(gi.cpu().numpy())
    rmd, ram = ref.closest_to_cog()
    rmd, ram = rmd.cpu().numpy(), (ram.to(torch.int64) & 0xFFFFFFFF).cpu().numpy()
    assert len(md) == ref.nvoxels
    assert np.all(np.abs(md - rmd) <= 1e-9 * rmd + 1e-12), "min_distance"
    same = gi == ram
    # H5: the winner must lie in the single-GPU run's tied set of its voxel (distance <= min * (1 + 1e-12) + the
    # rounding scale of a last-bit centroid difference); where the runner-up is clearly farther it must be the same point
    p2v = (ref.point2voxel.to(torch.int64) & 0xFFFFFFFF).cpu().numpy()
    full_d = ref.point_distance().cpu().numpy()
    scale = 8 * np.finfo(np.float64).eps * max(vs, float(np.abs(cloud).max()))
    rows = np.arange(ref.nvoxels)
    assert np.array_equal(p2v[gi], rows), "the winner belongs to its voxel"
    assert np.all(full_d[gi] <= rmd * (1 + 1e-12) + scale), "the winner is in the tied set"
    far = np.where(full_d <= rmd[p2v] * (1 + 1e-12) + scale, np.inf, full_d)
    second = np.full(ref.nvoxels, np.inf)
    np.minimum.at(second, p2v, far)
    near_count = np.bincount(p2v, weights=(far == np.inf).astype(np.float64), minlength=ref.nvoxels)
    clear = (near_count == 1) & (second > rmd * (1 + 1e-9) + 1e-12)
    assert np.array_equal(gi[clear], ram[clear]), "clear winners are identical"
    near_ties = int((~clear).sum())
    # mode / mode_count of an integer attribute: (voxel, value, count) runs merged at the owners == one GPU, exactly
    cls_all = rng.integers(0, 9, n).astype(np.float64)
    cls_all[cloud[:, 0] > 20] = np.minimum(cls_all[cloud[:, 0] > 20], 3)  # skewed classes in half of the scene
    mode, counts = D.mode_count(table, shard, torch.from_numpy(cls_all[a:b]).to(dev), thresholds=[0.1, 0.3], want_mode=True)
    rmode, rcounts = ref.mode_count(cls_all, [0.1, 0.3], want_mode=True)
    assert np.array_equal(gather_cat(mode.cpu().numpy()), rmode.cpu().numpy()), "mode"
    for p in (0.1, 0.3):
        assert np.array_equal(gather_cat(counts[p].cpu().numpy()), rcounts[p].cpu().numpy()), f"mode_count,{p}"
    if rank == 0:
        print(f"[per-point] world={world}: {len(md)} voxels, subsample identical for {int(same.sum())} ({near_ties} near-tie voxels), "
              f"partial rows answered by owners: {sum(shard.send_remote)}", flush=True)


def main():
    from vapc_b200 import dist as D
    from vapc_b200 import engine as E

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(17)
    n, vs, origin = 600_000, 0.2, [0.5, -1.0, 0.25]
    pts = np.round(np.concatenate([rng.uniform([-20, -5, 0], [60, 45, 6], (n // 2, 3)),
                                   rng.normal([20, 20, 3], [3, 3, 0.5], (n - n // 2, 3))]), 3)
    ref = E.VoxelCloud.build(pts[:, 0].copy(), pts[:, 1].copy(), pts[:, 2].copy(), vs, origin)  # single-GPU table of the whole cloud
    rs = ref.stats(density=True, cog=True, std=True, cov=True, eig=True, eig_n=True, feat=True)
    ref_vox = torch.stack(ref.voxel_coords()).cpu().numpy()

    def close(name, a, b, mode):
        assert np.array_equal(np.isnan(a), np.isnan(b)), name
        ok = ~np.isnan(b)
        scale = np.abs(b[ok]) if name in ("density", "cog") else np.maximum(np.abs(b[ok]), 1e-12)
        err = np.abs(a[ok] - b[ok])
        if name in ("cov", "eig", "std"):  # norm-wise per voxel (H1 / H7)
            big = np.nanmax(np.abs(b), axis=0, keepdims=True) * np.ones_like(b)
            scale = np.maximum(big[ok], 1e-300)
        assert np.all(err <= 1e-9 * scale + 1e-18), (mode, name, float(np.max(err / scale)))

    for mode, exchange in (("coherent", "peer"), ("shuffled", "peer"), ("coherent", "nccl"), ("coherent_y", "peer")):
        if mode == "shuffled":
            order = rng.permutation(n)
        else:  # contiguous chunks of the cloud sorted along x — or along y: the key then leads with y (choose_axis_order)
            order = np.argsort(pts[:, 1 if mode == "coherent_y" else 0], kind="stable")
        chunk = pts[order[rank * n // world:(rank + 1) * n // world]]
        x, y, z = (torch.from_numpy(np.ascontiguousarray(chunk[:, k])).to(dev) for k in range(3))
        table, st = D.voxel_statistics(x, y, z, vs, origin, exchange=exchange)
        assert table.exchange == exchange, f"asked for the {exchange} exchange, got {table.exchange}"
        assert table.axis_order == ((1, 0, 2) if mode == "coherent_y" and world > 1 else (0, 1, 2)), table.axis_order
        counts = torch.tensor([table.nvoxels], device=dev)
        allc = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(allc, counts)
        allc = [int(c.item()) for c in allc]
        assert sum(allc) == ref.nvoxels, (allc, ref.nvoxels)
        # order-free comparison: every rank's rows by voxel triple against the single-GPU table (lexicographic x, y, z)
        vox = gather_cat(torch.stack(table.voxel_coords()).cpu().numpy())
        by_voxel = np.lexsort((vox[2], vox[1], vox[0]))
        assert np.array_equal(vox[:, by_voxel], ref_vox), "the voxel set"
        assert np.array_equal(gather_cat(table.count.cpu().numpy())[by_voxel], ref.count.cpu().numpy()), "counts are exact"
        for name in ("density", "cog", "std", "cov", "eig"):
            close(name, gather_cat(getattr(st, name).cpu().numpy())[..., by_voxel], getattr(rs, name).cpu().numpy(), mode)
        if table.axis_order == (0, 1, 2):
            assert table.grid.total_bits == ref.grid.total_bits and list(table.grid.vmin) == list(ref.grid.vmin)
            lo = sum(allc[:rank])
            sl = slice(lo, lo + table.nvoxels)
            assert torch.equal(table.voxel_keys, ref.voxel_keys[sl]), "keys: ranks in rank order = global key order"
            assert torch.equal(table.count, ref.count[sl]), "counts are exact"
        else:
            # chunks separated along y: a y-leading key keeps most rows at home, the x-leading key of the same chunks sends most away
            forced, _ = D.voxel_statistics(x, y, z, vs, origin, exchange=exchange, axis_order=(0, 1, 2))
            sent = torch.tensor([table.partial_rows_sent, forced.partial_rows_sent], device=dev)
            dist.all_reduce(sent)
            assert int(sent[0]) < 0.6 * int(sent[1]), ("rows sent with the y-leading / x-leading key", sent.tolist())
        # the chunks' bounding boxes as an input (LAS headers: no min/max pass, no collective for the bounds): the same table bit for bit
        hdr = D.gather_bounds(x, y, z)
        t_h, s_h = D.voxel_statistics(x, y, z, vs, origin, exchange=exchange, all_bounds=hdr)
        assert t_h.axis_order == table.axis_order and t_h.nvoxels == table.nvoxels
        assert torch.equal(t_h.voxel_keys, table.voxel_keys) and torch.equal(t_h.count, table.count), "all_bounds= changes the table"
        for name in ("cog", "cov", "eig"):
            assert torch.equal(torch.nan_to_num(getattr(s_h, name)), torch.nan_to_num(getattr(st, name))), ("all_bounds=", name)
        if rank == 0:
            print(f"[{mode}, {exchange}] world={world}: axis order {table.axis_order}, owned voxels {allc}, partial rows sent by rank 0: "
                  f"{table.partial_rows_sent}", flush=True)
    two_epoch(rank, world, dev, pts, rng, vs, origin)
    per_point_outputs(rank, world, dev, pts, rng, vs, origin)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_OK", flush=True)


if __name__ == "__main__":
    main()
