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
    for mode, exchange in (("coherent", "peer"), ("shuffled", "peer"), ("coherent", "nccl")):
        order = np.argsort(pts[:, 0], kind="stable") if mode == "coherent" else rng.permutation(n)
        chunk = pts[order[rank * n // world:(rank + 1) * n // world]]
        x, y, z = (torch.from_numpy(np.ascontiguousarray(chunk[:, k])).to(dev) for k in range(3))
        table, st = D.voxel_statistics(x, y, z, vs, origin, exchange=exchange)
        assert table.exchange == exchange, f"asked for the {exchange} exchange, got {table.exchange}"
        # single-GPU reference of the whole cloud on every rank
        ref = E.VoxelCloud.build(pts[:, 0].copy(), pts[:, 1].copy(), pts[:, 2].copy(), vs, origin)
        rs = ref.stats(density=True, cog=True, std=True, cov=True, eig=True, eig_n=True, feat=True)
        assert table.grid.total_bits == ref.grid.total_bits and list(table.grid.vmin) == list(ref.grid.vmin)
        counts = torch.tensor([table.nvoxels], device=dev)
        allc = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(allc, counts)
        allc = [int(c.item()) for c in allc]
        assert sum(allc) == ref.nvoxels, (allc, ref.nvoxels)
        lo = sum(allc[:rank])
        sl = slice(lo, lo + table.nvoxels)
        assert torch.equal(table.voxel_keys, ref.voxel_keys[sl]), "keys: ranks in rank order = global key order"
        assert torch.equal(table.count, ref.count[sl]), "counts are exact"
        for name in ("density", "cog", "std", "cov", "eig", "eig_n", "feat"):
            a, b = getattr(st, name).cpu().numpy(), getattr(rs, name).cpu().numpy()[..., sl]
            assert np.array_equal(np.isnan(a), np.isnan(b)), name
            ok = ~np.isnan(b)
            scale = np.abs(b[ok]) if name in ("density", "cog") else np.maximum(np.abs(b[ok]), 1e-12)
            err = np.abs(a[ok] - b[ok])
            if name in ("cov", "eig", "std"):  # norm-wise per voxel (H1 / H7)
                big = np.nanmax(np.abs(b), axis=0, keepdims=True) * np.ones_like(b)
                scale = np.maximum(big[ok], 1e-300)
            if name in ("eig_n", "feat"):
                continue  # ratios of noise on rank-deficient voxels; covered through eig
            assert np.all(err <= 1e-9 * scale + 1e-18), (mode, name, float(np.max(err / scale)))
        if rank == 0:
            print(f"[{mode}, {exchange}] world={world}: owned voxels {allc}, partial rows sent by rank 0: {table.partial_rows_sent}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_OK", flush=True)


if __name__ == "__main__":
    main()
