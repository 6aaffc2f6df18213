"""Independent check of a sharded voxel table (bench.py --gpus N runs it after the timed region and prints the outcome as
"check" in its JSON line, so the driver's SCALE record carries the parity of the N-GPU result).

Everything here is plain torch + torch.distributed on the ranks' own points — none of the library's kernels:

  points     the counts of all rows of all ranks add up to the number of points of all ranks
  order      keys strictly ascending inside every rank and across the rank boundaries (rank order = key order)
  key set    every rank voxelises its points with torch (floor((c - o) / vs), fp64), packs the global key, de-duplicates
             (torch.unique) and sends the keys to the owners of their key ranges; the union must equal the owner's
             voxel_keys exactly — hence also the global number of voxels
  sample     for a sampled key range per owner every rank sends the points falling into it; the owner recomputes count,
             centroid and covariance with torch (two passes, fp64) and compares: counts exact, centroid 1e-9 relative,
             covariance 1e-9 of its largest entry
"""
import torch
import torch.distributed as dist


def _i64(keys):
    return (keys.to(torch.int64) & 0xFFFFFFFF) if keys.dtype == torch.int32 else keys


def _global_keys(x, y, z, grid):
    vs = float(grid.voxel_size)
    out = None
    shifts = (grid.bits[1] + grid.bits[2], grid.bits[2], 0)
    for c, k, sh in zip((x, y, z), range(3), shifts):
        v = torch.floor((c - float(grid.origin[k])) / vs).to(torch.int64) - int(grid.vmin[k])
        if bool(((v < 0) | (v >= (1 << grid.bits[k]))).any()):
            raise AssertionError("a point lies outside the global grid")
        out = (v << sh) if out is None else out | (v << sh)
    return out


def _exchange(parts, group):
    """all-to-all of a list of world tensors (same dtype / trailing shape); returns the received list."""
    world = dist.get_world_size(group)
    dev = parts[0].device
    sc = torch.tensor([p.shape[0] for p in parts], dtype=torch.int64, device=dev)
    rc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    rc = rc.cpu().tolist()
    src = torch.cat(parts).contiguous()
    dst = torch.empty((int(sum(rc)),) + tuple(src.shape[1:]), dtype=src.dtype, device=dev)
    dist.all_to_all_single(dst, src, output_split_sizes=rc, input_split_sizes=sc.cpu().tolist(), group=group)
    return dst


def verify_sharded_table(table, stats, x, y, z, group=None, sample_rows=100_000):
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = x.device
    g = table.grid
    out = {}
    tk = _i64(table.voxel_keys).contiguous()
    v = tk.numel()
    cnt = table.count.to(torch.int64) & 0xFFFFFFFF

    # ---- points ------------------------------------------------------------------------------------------------------
    tot = torch.stack([cnt.sum(), torch.tensor(x.numel(), dtype=torch.int64, device=dev), torch.tensor(v, dtype=torch.int64, device=dev)])
    dist.all_reduce(tot, group=group)
    out["points"] = int(tot[1])
    out["points_ok"] = bool(tot[0] == tot[1])
    out["voxels"] = int(tot[2])

    # ---- order -------------------------------------------------------------------------------------------------------
    asc = bool((tk[1:] > tk[:-1]).all()) if v > 1 else True
    edge = torch.tensor([int(tk[0]) if v else -1, int(tk[-1]) if v else -1], dtype=torch.int64, device=dev)
    edges = torch.empty(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(edges, edge, group=group)
    last = -1
    for r, (lo, hi) in enumerate(edges.view(world, 2).cpu().tolist()):
        if lo < 0:
            continue
        asc = asc and lo > last
        last = hi
    th = torch.tensor(list(table.thresholds), dtype=torch.int64, device=dev)
    if v and world > 1:
        owner = torch.searchsorted(th, tk, right=True)
        asc = asc and bool((owner == rank).all())
    flag = torch.tensor([int(asc)], dtype=torch.int64, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    out["order_ok"] = bool(flag.item())

    # ---- key set -----------------------------------------------------------------------------------------------------
    order = tuple(getattr(table, "axis_order", (0, 1, 2)))  # the table's grid is expressed in its internal axis order
    pk = _global_keys(*[(x, y, z)[a] for a in order], g)
    uk = torch.unique(pk)
    dest = torch.searchsorted(th, uk, right=True) if world > 1 else torch.zeros_like(uk)
    cuts = torch.searchsorted(dest, torch.arange(world + 1, device=dev)).cpu().tolist()
    got = _exchange([uk[cuts[r]:cuts[r + 1]] for r in range(world)], group)
    mine = torch.unique(got)
    same = mine.numel() == v and bool(torch.equal(mine, tk))
    flag = torch.tensor([int(same), mine.numel()], dtype=torch.int64, device=dev)
    tot2 = flag.clone()
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(tot2, group=group)
    out["key_set_ok"] = bool(flag[0].item())
    out["voxels_torch_unique"] = int(tot2[1])
    del uk, dest, got, mine

    # ---- sample ------------------------------------------------------------------------------------------------------
    # owner r samples the rows [a, a + m) in the middle of its range; everybody sends it the points with keys in that range
    m = min(v, sample_rows)
    a = (v - m) // 2
    rng_mine = torch.tensor([int(tk[a]) if m else 1, int(tk[a + m - 1]) if m else 0], dtype=torch.int64, device=dev)
    rngs = torch.empty(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(rngs, rng_mine, group=group)
    parts = []
    for lo, hi in rngs.view(world, 2).cpu().tolist():
        sel = (pk >= lo) & (pk <= hi)
        parts.append(torch.stack([x[sel], y[sel], z[sel], pk[sel].view(torch.float64)], dim=1))
    pts = _exchange(parts, group)
    del parts, pk
    ok_cnt = ok_cog = ok_cov = True
    worst_cog = worst_cov = 0.0
    if m:
        keys = pts[:, 3].contiguous().view(torch.int64)
        row = torch.searchsorted(tk[a:a + m].contiguous(), keys)
        ok_cnt = bool((tk[a:a + m][row.clamp(max=m - 1)] == keys).all())
        n_ref = torch.bincount(row, minlength=m)
        ok_cnt = ok_cnt and bool(torch.equal(n_ref, cnt[a:a + m]))
        xyz = pts[:, :3]
        mean = torch.zeros((m, 3), dtype=torch.float64, device=dev).index_add_(0, row, xyz) / n_ref[:, None]
        d = xyz - mean[row]
        prod = torch.stack([d[:, 0] * d[:, 0], d[:, 0] * d[:, 1], d[:, 0] * d[:, 2], d[:, 1] * d[:, 1], d[:, 1] * d[:, 2], d[:, 2] * d[:, 2]], dim=1)
        cov = torch.zeros((m, 6), dtype=torch.float64, device=dev).index_add_(0, row, prod) / (n_ref[:, None] - 1).clamp(min=1)
        cog = stats.cog[:, a:a + m].t()
        # 1e-9 relative, norm-wise per voxel (a coordinate that happens to be ~0 has no relative scale of its own)
        err = ((cog - mean).abs().max(dim=1).values / mean.abs().max(dim=1).values.clamp(min=1e-3 * float(g.voxel_size))).max()
        worst_cog = float(err)
        ok_cog = worst_cog <= 1e-9
        multi = n_ref > 1
        if bool(multi.any()):
            gc = stats.cov[:, a:a + m].t()[multi]
            rc = cov[multi]
            scale = rc.abs().max(dim=1, keepdim=True).values + 1e-3 * float(g.voxel_size) ** 2  # degenerate voxels: relative to the voxel
            worst_cov = float(((gc - rc).abs() / scale).max())
            ok_cov = worst_cov <= 1e-9
    flag = torch.tensor([int(ok_cnt), int(ok_cog), int(ok_cov)], dtype=torch.int64, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    worst = torch.tensor([worst_cog, worst_cov], dtype=torch.float64, device=dev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=group)
    out.update(sample_rows_per_rank=m, sample_count_ok=bool(flag[0].item()), sample_centroid_ok=bool(flag[1].item()),
               sample_covariance_ok=bool(flag[2].item()), sample_centroid_worst_rel=float(worst[0]), sample_covariance_worst_rel=float(worst[1]))
    out["ok"] = all(out[k] for k in ("points_ok", "order_ok", "key_set_ok", "sample_count_ok", "sample_centroid_ok", "sample_covariance_ok"))
    return out
