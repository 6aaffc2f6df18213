"""Multi-GPU voxel statistics: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

Points are sharded in contiguous chunks (each rank holds its own slice, as it would come out of a
file).  The path has exactly one real exchange step, as BASELINE.json's north_star prescribes:

  1. bounding boxes: local min/max kernel; the all-reduce(min/max) of 6 doubles that gives the GLOBAL key layout is
     started asynchronously — nothing local depends on it;
  2. every rank builds its local PARTIAL voxel table (key, count, mean - c(key), M2) with its OWN grid (keys
     relative to its own lowest voxel: fewer key bits and radix passes than the global layout): keys + records in
     bucket order -> segmented radix sort -> segmentation -> moments, all local, no communication;
  3. the sorted local voxel keys are re-expressed in the global layout (same lexicographic order) and the key space
     is cut into `world` ranges balancing the number of partial rows: ONE all-gather of the ranks' 2^16-bin key
     histograms, from which a single kernel derives the splitters and the complete send-count matrix;
  4. ONE all-to-all(v) moves the rows that leave their rank, packed row-major, to the owner of their key range (the
     local table is sorted, so each destination's rows are a contiguous slice); the rows a rank owns itself are
     never copied;
  5. the owner merges its own rows with the received ones (stable radix sort by key keeps source order; equal keys
     are combined with Chan's update in that fixed order) and finalises statistics.

Because c(key) is an exact function of the key, partial (count, mean, M2) triples combine without
loss; integer outputs do not depend on the number of ranks, floating point ones differ only by the
association of the Chan updates (<< 1e-9).

The host-side planning functions are pure (numpy / torch-CPU friendly) so that the N > 1 logic is
covered by world_size-2 gloo tests on the CPU (tests/test_dist_cpu.py); kernels run only on CUDA.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

HIST_BITS = 12  # 4096 bins: key ranges balanced to 1/4096 of the key space; the planning kernel walks them in one CTA


# ------------------------------------------------------------------------------------------------
# pure planning helpers (no CUDA)
# ------------------------------------------------------------------------------------------------
def bench_slab_shift(rank: int, extent_x: float) -> float:
    """bench.py's N > 1 layout: rank r holds the slab x in [0.9 r L, 0.9 r L + L): neighbouring chunks overlap
    by 10 %, like adjacent flight strips, so ~10 % of every rank's voxels are shared with a neighbour and the
    all-to-all really carries data."""
    return 0.9 * rank * extent_x


def choose_axis_order(all_bounds) -> tuple:
    """Which coordinate axis leads the packed key (most significant first).  The key space is cut into contiguous key ranges, one
    per rank, so a rank keeps the rows of its own chunk exactly when the chunks are separated along the LEADING axis: chunks that
    are strips along y under an x-major key would send (world - 1) / world of their rows away.  all_bounds: [world, 6] local
    bounding boxes (min x, y, z, max x, y, z).  Measure per axis = sum of the ranks' extents / global extent (1 = the chunks tile the
    axis, world = every chunk spans it), compared in eighths of its range so that chunks which all span everything (shuffled points)
    keep the x, y, z order whatever the last digits of their bounding boxes say; axes in ascending measure, ties in x, y, z order."""
    b = np.asarray(all_bounds, dtype=np.float64).reshape(-1, 6)
    measure = []
    for a in range(3):
        ext = float(b[:, 3 + a].max() - b[:, a].min())
        measure.append(float((b[:, 3 + a] - b[:, a]).sum()) / ext if ext > 0 else float(len(b)))
    world = max(len(b), 1)
    return tuple(sorted(range(3), key=lambda a: (int(round(8.0 * measure[a] / world)), a)))


def plan_splitters(global_hist: np.ndarray, world: int, total_bits: int, hist_bits: int = HIST_BITS) -> List[int]:
    """Key thresholds t_1 <= ... <= t_{world-1}: rank g owns keys in [t_g, t_{g+1}) (t_0 = 0, t_world = inf).
    Chosen on histogram-bin boundaries so that every range holds about the same number of partial rows."""
    hist = np.asarray(global_hist, dtype=np.int64)
    hb = min(hist_bits, max(total_bits, 1))
    shift = max(total_bits - hb, 0)
    cum = np.cumsum(hist)
    total = int(cum[-1]) if len(cum) else 0
    out = []
    for g in range(1, world):
        target = total * g / world
        b = int(np.searchsorted(cum, target, side="left")) + 1  # first bin NOT in range g-1
        b = min(b, len(hist))
        out.append(b << shift)
    for i in range(1, len(out)):
        out[i] = max(out[i], out[i - 1])
    return out


def keys_as_int64(keys: torch.Tensor) -> torch.Tensor:
    """uint32-in-int32 / uint64-in-int64 key storage as non-negative int64 (needs <= 63 key bits)."""
    if keys.dtype == torch.int32:
        return keys.to(torch.int64) & 0xFFFFFFFF
    return keys


def partition_counts(sorted_keys: torch.Tensor, thresholds: Sequence[int]) -> List[int]:
    """Rows of a key-sorted table going to each rank."""
    n = sorted_keys.numel()
    if not thresholds:
        return [n]
    k = keys_as_int64(sorted_keys)
    t = torch.tensor(list(thresholds), dtype=torch.int64, device=k.device)
    cuts = torch.searchsorted(k, t, right=False).cpu().tolist()
    edges = [0] + cuts + [n]
    return [edges[i + 1] - edges[i] for i in range(len(edges) - 1)]


def exchange_rows(columns: Dict[str, torch.Tensor], send_counts: Sequence[int], group=None):
    """all-to-all(v) of row-major columns ([rows] or [rows, k] tensors, rows already grouped by destination).
    Returns (received columns, recv_counts).  One small all-to-all of the counts, then one per column."""
    world = dist.get_world_size(group)
    dev = next(iter(columns.values())).device
    sc = torch.tensor(list(send_counts), dtype=torch.int64, device=dev)
    rc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = rc.cpu().tolist()
    total = int(sum(recv_counts))
    out = {}
    for name, t in columns.items():
        t = t.contiguous()
        r = torch.empty((total,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_to_all_single(r, t, output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
        out[name] = r
    return out, recv_counts


# ------------------------------------------------------------------------------------------------
# CUDA path
# ------------------------------------------------------------------------------------------------
@dataclass
class ShardedVoxelTable:
    """This rank's key range of the global voxel table (rows in key order; ranks in rank order)."""

    grid: object
    nvoxels: int  # rows owned by this rank
    voxel_keys: torch.Tensor
    count: torch.Tensor
    mean3: torch.Tensor  # [3, V]
    m2: torch.Tensor  # [6, V]
    thresholds: List[int]
    partial_rows_sent: int
    partial_rows_received: int
    exchange: str = "nccl"  # "peer": rows written straight into the owner's symmetric receive buffer over NVLink
    local_voxels: int = 0  # rows of this rank's own partial table before the exchange
    # Internal axis order: axis_order[0] is the coordinate axis that leads the packed key (choose_axis_order).  grid, voxel_keys,
    # mean3 and m2 are expressed in THAT order; stats() and voxel_coords() hand everything back in x, y, z order.
    axis_order: tuple = (0, 1, 2)
    fused_stats: object = None  # VoxelStats formed inside the owner merge (the merged moments were then not written: mean3 / m2 None)

    def stats(self, density=True, cog=True, std=True, cov=True, eig=True, eig_n=True, feat=True):
        from . import engine as E

        want = dict(density=density, cog=cog, std=std, cov=cov, eig=eig, eig_n=eig_n, feat=feat)
        if self.fused_stats is not None and all(getattr(self.fused_stats, k) is not None for k, w in want.items() if w):
            return E.VoxelStats(**{k: getattr(self.fused_stats, k) for k, w in want.items() if w})
        if self.mean3 is None:
            raise ValueError("this table was built with its statistics fused into the merge (keep_moments=False): ask voxel_statistics for "
                             "the columns up front, or pass keep_moments=True")

        c = E.VoxelCloud()
        c.grid, c.nvoxels, c.voxel_keys, c.count, c.mean3, c.m2 = self.grid, self.nvoxels, self.voxel_keys, self.count, self.mean3, self.m2
        c.axis_order = tuple(self.axis_order)  # the kernel writes centroid / std / covariance rows in x, y, z order
        return c.stats(density=density, cog=cog, std=std, cov=cov, eig=eig, eig_n=eig_n, feat=feat)

    def voxel_coords(self):
        """(vx, vy, vz) int64 [V] of this rank's rows."""
        from . import engine as E

        c = E.VoxelCloud()
        c.grid, c.nvoxels, c.voxel_keys = self.grid, self.nvoxels, self.voxel_keys
        v = c.voxel_coords()
        out = [None, None, None]
        for i, a in enumerate(self.axis_order):
            out[a] = v[i]
        return out

    def internal_rows(self, t: torch.Tensor) -> torch.Tensor:
        """[3, ...] rows given in x, y, z order -> the table's internal axis order."""
        return t if tuple(self.axis_order) == (0, 1, 2) else t[list(self.axis_order)]


@dataclass
class LocalShard:
    """What a rank keeps of its own chunk so that columns of the merged table can be brought back to its partial rows and
    points (voxel_statistics(..., keep_local=True))."""

    cloud: object  # the rank's own VoxelCloud (own grid) incl. x, y, z and point2voxel
    gkeys: torch.Tensor  # int64 [v]: global keys of the local partial rows (sorted)
    lo: int  # local rows [lo, lo + n_own) are owned by this rank itself
    n_own: int
    send_remote: List[int]  # partial rows sent to every rank (0 for itself)
    recv_remote: List[int]  # partial rows received from every rank (0 for itself)
    recv_keys: torch.Tensor  # int64 [n_in]: keys of the received rows, grouped by source
    group: object = None


def plan_splitters_tensor(global_hist: torch.Tensor, world: int, total_bits: int, hist_bits: int = HIST_BITS) -> torch.Tensor:
    """plan_splitters on a tensor, on whatever device the histogram lives (no host round trip): int64 [world - 1]."""
    hb = min(hist_bits, max(total_bits, 1))
    shift = max(total_bits - hb, 0)
    cum = torch.cumsum(global_hist.to(torch.int64), 0)
    if world <= 1:
        return torch.empty(0, dtype=torch.int64, device=cum.device)
    g = torch.arange(1, world, dtype=torch.float64, device=cum.device)
    target = cum[-1].to(torch.float64) * g / world
    b = torch.searchsorted(cum.to(torch.float64), target, right=False) + 1  # first bin NOT in range g-1
    b = torch.clamp(b, max=cum.numel())
    return torch.cummax(b << shift, 0).values


def destination_counts(sorted_keys_i64: torch.Tensor, thresholds: torch.Tensor) -> torch.Tensor:
    """Rows of a key-sorted table (non-negative int64 keys) going to each rank: int64 [world], same device."""
    n = sorted_keys_i64.numel()
    cuts = torch.searchsorted(sorted_keys_i64, thresholds, right=False)
    edges = torch.cat([cuts.new_zeros(1), cuts, cuts.new_full((1,), n)])
    return edges[1:] - edges[:-1]


class _PeerBuffers:
    """Symmetric (peer-mapped) receive buffers of one process group: two alternating [capacity, 11] fp64 row buffers per rank."""

    def __init__(self, group, device, capacity):
        import torch.distributed._symmetric_memory as symm

        self.capacity = int(capacity)
        self.buf = symm.empty((2, self.capacity, 11), dtype=torch.float64, device=device)
        self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.parity = 0

    def flip(self):
        self.parity ^= 1

    def receive_view(self, rows):
        return self.buf[self.parity, :rows]

    def remote_ptr(self, rank):
        return int(self.handle.buffer_ptrs[rank]) + self.parity * self.capacity * 88

    def barrier(self):
        self.handle.barrier(channel=0)


_PEER = {}
_PEER_BROKEN = False


def _peer_exchange(group, device, matrix_h, world):
    """The peer buffers of this group, grown (collectively: every rank sees the same matrix) when a rank would receive more
    rows than they hold; None when symmetric memory is not available (then NCCL's all-to-all is used)."""
    global _PEER_BROKEN
    if _PEER_BROKEN:
        return None
    need = 0
    for r in range(world):
        need = max(need, int(matrix_h[:, r].sum()) - int(matrix_h[r, r]))
    key = (id(group), device.index)
    cur = _PEER.get(key)
    try:
        if cur is None or cur.capacity < need:
            cap = max(1 << 16, int(need * 1.5))
            cur = _PeerBuffers(group, device, cap)
            _PEER[key] = cur
        cur.flip()
        return cur
    except Exception as exc:  # noqa: BLE001  (no NVLink peer access / symmetric memory not built in)
        import warnings

        warnings.warn(f"peer-memory exchange unavailable ({exc}); using the NCCL all-to-all")
        _PEER_BROKEN = True
        return None


def start_bounds(x, y, z, group=None):
    """Local bbox kernel + the 6-double all-reduce started asynchronously.  Returns (local bounds on the host, finish) where
    finish() waits for the collective and returns the global bounds — the local pipeline does not need them, so it runs
    while the all-reduce is in flight."""
    from . import _lib as L
    from . import engine as E

    n = x.numel()
    out = torch.empty(6, dtype=torch.float64, device=x.device)
    ws_bytes = L.raw("vapc_minmax_workspace_bytes")()
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    L.call("vapc_minmax_f64", E._ptr(x), E._ptr(y), E._ptr(z), n, E._ptr(out), E._ptr(ws), ws_bytes, E._stream())
    red = torch.cat([out[:3], -out[3:]])  # min over (min, -max)
    work = dist.all_reduce(red, op=dist.ReduceOp.MIN, group=group, async_op=True)
    local = out.cpu().numpy()

    def finish():
        work.wait()
        glob = red.cpu().numpy()
        glob[3:] = -glob[3:]
        return glob

    return local, finish


def global_bounds(x, y, z, group=None) -> np.ndarray:
    """Global [min x, min y, min z, max x, max y, max z]."""
    return start_bounds(x, y, z, group)[1]()


def merge_partial_rows(grid, keys: torch.Tensor, count: torch.Tensor, mean3: torch.Tensor, m2: torch.Tensor):
    """Combine partial rows with equal keys (vapc_merge_partials).  keys [R] (any order, runs per source sorted),
    count [R], mean3 [3, R], m2 [6, R].  Returns (voxel_keys, count, mean3, m2) of the merged table."""
    from . import _lib as L
    from . import engine as E

    dev = keys.device
    r = keys.numel()
    ks, order = E.sort_pairs(keys.clone(), grid.total_bits)
    ws_bytes = L.raw("vapc_segment_workspace_bytes")(r)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    nv = torch.zeros(1, dtype=torch.int64, device=dev)
    L.call("vapc_count_voxels", E._ptr(ks), r, grid.key_bytes, E._ptr(nv), E._ptr(ws), ws_bytes, E._stream())
    v = int(nv.item())
    out_keys = torch.empty(v, dtype=keys.dtype, device=dev)
    out_count = torch.empty(v, dtype=torch.int32, device=dev)
    out_mean = torch.empty((3, v), dtype=torch.float64, device=dev)
    out_m2 = torch.empty((6, v), dtype=torch.float64, device=dev)
    if r:
        L.call("vapc_merge_partials", E._ptr(ks), E._ptr(order), r, grid.key_bytes, E._ptr(count.contiguous()), E._ptr(mean3.contiguous()),
               E._ptr(m2.contiguous()), r, v, E._ptr(out_keys), E._ptr(out_count), E._ptr(out_mean), E._ptr(out_m2), E._ptr(ws), ws_bytes,
               E._stream())
    return out_keys, out_count, out_mean, out_m2


def gather_bounds(x, y, z, group=None) -> np.ndarray:
    """[world, 6] local bounding boxes of all ranks (one min/max kernel + one all-gather of 6 doubles per rank)."""
    from . import _lib as L
    from . import engine as E

    world = dist.get_world_size(group)
    out = torch.empty(6, dtype=torch.float64, device=x.device)
    ws_bytes = L.raw("vapc_minmax_workspace_bytes")()
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    L.call("vapc_minmax_f64", E._ptr(x), E._ptr(y), E._ptr(z), x.numel(), E._ptr(out), E._ptr(ws), ws_bytes, E._stream())
    allb = torch.empty(6 * world, dtype=torch.float64, device=x.device)
    dist.all_gather_into_tensor(allb, out, group=group)
    return allb.cpu().numpy().reshape(world, 6)


def voxel_statistics(x, y, z, voxel_size, origin=(0.0, 0.0, 0.0), group=None, with_stats=True, exchange="peer", bounds=None,
                     thresholds=None, keep_local=False, stats_columns=None, axis_order="auto", keep_moments=None, all_bounds=None):
    """The multi-GPU hot path.  x, y, z: this rank's chunk (CUDA fp64).  Returns (ShardedVoxelTable, VoxelStats).

    Every rank voxelises its chunk with its OWN grid (keys relative to its own lowest voxel: fewer key bits and radix passes
    than the global layout, and no dependence on the global bounds, whose all-reduce overlaps the local pipeline).  The
    sorted local table is then re-keyed in the global layout (same order), splitters and destinations are planned on the
    device, only the rows that leave the rank are packed and exchanged (ONE all-to-all), and the owner merges its own rows
    (in place) with the received ones.  Small device -> host copies synchronise the host: the ranks' bounds (not with all_bounds=), local voxel
    count, global bounds, the all-gathered send-count matrix, the merged voxel count.

    keep_moments: keep the merged (mean3, m2) moments in the table (needed to merge it again or to ask for other statistics later).
    Default: only when no statistics are asked for here; otherwise the statistics are formed inside the owner's merge pass and the
    merged moments are never written (72 B per voxel each way saved).
    axis_order: "auto" (choose_axis_order from the ranks' bounding boxes) or a permutation of (0, 1, 2): the coordinate axis that
    leads the packed key.  Results come back in x, y, z order (ShardedVoxelTable.stats / voxel_coords); only the ROW order of the
    sharded table follows the internal order.
    all_bounds (optional): [world, 6] bounding boxes of every rank's chunk as their LAS headers state them (gather_bounds once at load
    time, or the headers themselves): no min/max pass over the points and no collective for the bounds inside the step; a point outside
    its rank's stated box raises VAPC_E_RANGE.
    bounds (optional): global coordinate bounds every rank already agrees on; thresholds (optional): the
    world - 1 key thresholds of an earlier table over the SAME grid — the rows are then sent to the owners of those key
    ranges instead of re-balancing, so that two tables (two epochs) are range-aligned rank by rank (two_epoch_change)."""
    from . import _lib as L
    from . import engine as E

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = x.device
    # the local bounding boxes of all ranks: the global bounds, and which axis leads the key (the one the chunks are separated along)
    local_b = None
    if all_bounds is not None or bounds is None or axis_order == "auto":
        if all_bounds is not None:
            allb = np.asarray(all_bounds, dtype=np.float64).reshape(world, 6)
        else:
            allb = gather_bounds(x, y, z, group)
        local_b = allb[rank]
        if bounds is None:
            bounds = np.concatenate([allb[:, :3].min(0), allb[:, 3:].max(0)])
        if axis_order == "auto":
            axis_order = choose_axis_order(allb)
    axis_order = tuple(int(a) for a in axis_order)
    if sorted(axis_order) != [0, 1, 2]:
        raise ValueError("axis_order must be a permutation of (0, 1, 2)")
    if axis_order != (0, 1, 2):  # from here on everything lives in the internal axis order
        perm6 = list(axis_order) + [3 + a for a in axis_order]
        x, y, z = [(x, y, z)[a] for a in axis_order]
        origin = [origin[a] for a in axis_order]
        bounds = np.asarray(bounds, dtype=np.float64)[perm6]
        local_b = local_b[perm6] if local_b is not None else None
    local = E.VoxelCloud.build(x, y, z, voxel_size, origin, bounds=local_b, want_point2voxel=keep_local, keep_columns=keep_local)
    lg = local.grid
    g = E.grid_from_bounds(voxel_size, origin, bounds)  # the key layout every rank agrees on
    if g.total_bits > 63:
        raise L.VapcError(L.VAPC_E_KEYSPACE, "multi-GPU key ranges need at most 63 key bits")
    v = local.nvoxels
    hb = min(HIST_BITS, max(g.total_bits, 1))
    gkeys = torch.empty(v, dtype=torch.int64, device=dev)
    hist = torch.empty(1 << hb, dtype=torch.int32, device=dev)
    L.call("vapc_global_keys", E._ptr(local.voxel_keys), v, C.byref(lg), C.byref(g), hb, E._ptr(gkeys), E._ptr(hist), E._stream())
    # ONE collective for the planning: with everybody's histogram every rank derives the splitters (global sum) AND the
    # complete send-count matrix (splitters lie on bin boundaries) in one small kernel; one copy tells the host
    if thresholds is None:
        hist_all = torch.empty(world << hb, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(hist_all, hist, group=group)
        plan = torch.empty(world * world + world - 1, dtype=torch.int64, device=dev)
        L.call("vapc_plan_exchange", E._ptr(hist_all), world, hb, g.total_bits, E._ptr(plan), E._stream())
    else:
        # given key ranges: this rank's destination counts from its sorted keys, all-gathered into the same matrix
        th = torch.as_tensor(list(thresholds), dtype=torch.int64, device=dev)
        if th.numel() != world - 1:
            raise ValueError("thresholds must hold world - 1 keys")
        counts = destination_counts(gkeys, th)
        matrix = torch.empty(world * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(matrix, counts, group=group)
        plan = torch.cat([matrix, th])
    host = plan.cpu()
    matrix_h = host[: world * world].view(world, world)
    send_counts = matrix_h[rank].tolist()
    recv_counts = matrix_h[:, rank].tolist()
    # rows [lo, hi) of the local table stay here; the rest leaves, packed row-major, grouped by destination (= key order)
    lo = int(sum(send_counts[:rank]))
    n_own = int(send_counts[rank])
    hi = lo + n_own
    send_remote = list(send_counts)
    recv_remote = list(recv_counts)
    send_remote[rank] = recv_remote[rank] = 0
    n_in = int(sum(recv_remote))
    peer = _peer_exchange(group, dev, matrix_h, world) if exchange == "peer" else None
    if peer is not None:
        # Peer-memory exchange over NVLink: every rank packs the rows that leave it STRAIGHT INTO the owner's receive buffer
        # (symmetric memory; the owner's slot for this source follows from the send-count matrix every rank holds), then one
        # device-side barrier.  No send buffer, no NCCL call; two buffers alternate so that one barrier per step is enough.
        in_rows = peer.receive_view(n_in)
        start = 0
        for dest in range(world):
            cnt = int(send_counts[dest])
            if dest != rank and cnt:
                slot = int(matrix_h[:rank, dest].sum()) - (int(matrix_h[dest, dest]) if dest < rank else 0)  # rows of lower sources, own excluded
                L.call("vapc_pack_partial_rows", E._ptr(gkeys), E._ptr(local.count), E._ptr(local.mean3), E._ptr(local.m2), v, start, start + cnt,
                       C.c_void_p(peer.remote_ptr(dest) + slot * 88), E._stream())
            start += cnt
        peer.barrier()
    else:
        out_rows = torch.empty((v - n_own, 11), dtype=torch.float64, device=dev)
        for begin, end, at in ((0, lo, 0), (hi, v, lo)):
            if end > begin:
                L.call("vapc_pack_partial_rows", E._ptr(gkeys), E._ptr(local.count), E._ptr(local.mean3), E._ptr(local.m2), v, begin, end,
                       C.c_void_p(out_rows.data_ptr() + at * 88), E._stream())
        in_rows = torch.empty((n_in, 11), dtype=torch.float64, device=dev)
        dist.all_to_all_single(in_rows, out_rows, output_split_sizes=recv_remote, input_split_sizes=send_remote, group=group)
    # merge WITHOUT re-sorting the own rows (they are the bulk and already key-sorted): the few received rows are sorted and combined
    # per key on their own (n_in rows), then merged by rank into the own rows in ONE pass (vapc_merge_sorted_tables)
    kdt = torch.int32 if g.key_bytes == 4 else torch.int64
    own_keys = gkeys[lo:hi]
    if n_in:
        rkeys = torch.empty(n_in, dtype=kdt, device=dev)
        L.call("vapc_merge_keys", E._ptr(gkeys), 0, 0, E._ptr(in_rows), n_in, g.key_bytes, E._ptr(rkeys), E._stream())
        ks, order = E.sort_pairs(rkeys, g.total_bits)
        ws_bytes = L.raw("vapc_segment_workspace_bytes")(n_in)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        nv = torch.zeros(1, dtype=torch.int64, device=dev)
        L.call("vapc_count_voxels", E._ptr(ks), n_in, g.key_bytes, E._ptr(nv), E._ptr(ws), ws_bytes, E._stream())
        # the number of distinct received keys (nu) and of those the own rows do not hold (-> vm) reach the host in ONE read: the
        # combined rows are formed at the upper bound n_in (column stride n_in; slots past nu keep the all-ones key, which no own row
        # holds and which the validity mask keeps out of the prefix), then cut to nu
        u_keys = torch.full((n_in,), -1, dtype=kdt, device=dev)
        u_count = torch.empty(n_in, dtype=torch.int32, device=dev)
        u_mean = torch.empty((3, n_in), dtype=torch.float64, device=dev)
        u_m2 = torch.empty((6, n_in), dtype=torch.float64, device=dev)
        L.call("vapc_merge_partial_rows", E._ptr(ks), E._ptr(order), n_in, g.key_bytes, E._ptr(in_rows), None, None, None, 0, 0, 0, n_in,
               E._ptr(u_keys), E._ptr(u_count), E._ptr(u_mean), E._ptr(u_m2), E._ptr(ws), ws_bytes, E._stream())
        u_keys64 = keys_as_int64(u_keys).contiguous()
        valid = torch.arange(n_in, device=dev) < nv
        if n_own:
            at = torch.searchsorted(own_keys, u_keys64)
            held = (own_keys[at.clamp(max=n_own - 1)] == u_keys64) & (at < n_own)
            fresh = valid & ~held
        else:
            fresh = valid
        new_before = torch.zeros(n_in + 1, dtype=torch.int64, device=dev)
        torch.cumsum(fresh.to(torch.int64), 0, out=new_before[1:])
        nu, n_new = torch.stack([nv[0], new_before[-1]]).cpu().tolist()
        vm = n_own + n_new
        if nu < n_in:
            u_keys64, u_count, new_before = u_keys64[:nu], u_count[:nu], new_before[: nu + 1]
            u_mean, u_m2 = u_mean[:, :nu].contiguous(), u_m2[:, :nu].contiguous()
    else:
        nu, u_keys64, u_count, u_mean, u_m2, new_before, vm = 0, None, None, None, None, None, n_own
    if keep_moments is None:
        keep_moments = not with_stats
    out_keys = torch.empty(vm, dtype=kdt, device=dev)
    out_count = torch.empty(vm, dtype=torch.int32, device=dev)
    out_mean = torch.empty((3, vm), dtype=torch.float64, device=dev) if keep_moments else None
    out_m2 = torch.empty((6, vm), dtype=torch.float64, device=dev) if keep_moments else None
    fused = None
    if with_stats:
        names = ("density", "cog", "std", "cov", "eig", "eig_n", "feat")
        cols = names if stats_columns is None else tuple(k for k in names if k in stats_columns)
        shapes = dict(density=(vm,), cog=(3, vm), std=(3, vm), cov=(6, vm), eig=(3, vm), eig_n=(3, vm), feat=(8, vm))
        fused = E.VoxelStats(**{k: torch.empty(shapes[k], dtype=torch.float64, device=dev) for k in cols})
    order_c = (C.c_int32 * 3)(*axis_order)
    if vm:
        stat_ptrs = [E._ptr(getattr(fused, k) if fused is not None else None) for k in ("density", "cog", "std", "cov", "eig", "eig_n", "feat")]
        L.call("vapc_merge_sorted_tables_stats", E._ptr(own_keys.contiguous()), E._ptr(local.count), E._ptr(local.mean3), E._ptr(local.m2), v, lo, n_own,
               E._ptr(u_keys64), E._ptr(u_count), E._ptr(u_mean), E._ptr(u_m2), nu, E._ptr(new_before), g.key_bytes, vm,
               E._ptr(out_keys), E._ptr(out_count), E._ptr(out_mean), E._ptr(out_m2), C.byref(g), order_c, *stat_ptrs, E._stream())
    table = ShardedVoxelTable(grid=g, nvoxels=vm, voxel_keys=out_keys, count=out_count, mean3=out_mean, m2=out_m2,
                              thresholds=host[world * world:].tolist(), partial_rows_sent=int(sum(send_remote)),
                              partial_rows_received=n_in, exchange="peer" if peer is not None else "nccl", axis_order=axis_order)
    table.local_voxels = v
    table.fused_stats = fused
    st = fused
    if keep_local:
        # keep_local: a third value, the LocalShard for rows_from_owner / point_distance / closest_to_cog below
        recv_keys = in_rows[:, 0].contiguous().view(torch.int64).clone() if n_in else torch.empty(0, dtype=torch.int64, device=dev)
        return table, st, LocalShard(cloud=local, gkeys=gkeys, lo=lo, n_own=n_own, send_remote=send_remote, recv_remote=recv_remote,
                                     recv_keys=recv_keys, group=group)
    return table, st


def two_epoch_change(epoch1: Sequence[torch.Tensor], epoch2: Sequence[torch.Tensor], voxel_size, origin=(0.0, 0.0, 0.0), alpha: float = 0.005,
                     group=None, exchange="peer"):
    """Multi-GPU form of the reference's two-epoch change detection (bi_vapc.py:40-47 centroid + covariance per voxel and
    epoch, :78-84 inner join on the voxel triple, :25-36 anti-joins, :86-138 centroid distance + Mahalanobis test).

    epoch1 / epoch2: this rank's (x, y, z) chunk of each epoch.  Both epochs are voxelised over ONE global grid (bounds
    of both) and share ONE set of key ranges (planned on epoch 1), so every voxel of either epoch ends on the same owner:
    the join and the test are local to the owner, no further communication.  Returns a dict for this rank's key range:
    table1 / table2 (ShardedVoxelTable), i1 / i2 (rows of the joined voxels in the two tables, key order), only1 / only2
    (rows occupied in one epoch only: change_type 3 / 4) and the joined rows' distance, d1, d2, p_value, significance,
    change_type."""
    from . import _lib as L
    from . import engine as E

    a1, a2 = gather_bounds(*epoch1, group), gather_bounds(*epoch2, group)
    gb = np.concatenate([np.minimum(a1[:, :3].min(0), a2[:, :3].min(0)), np.maximum(a1[:, 3:].max(0), a2[:, 3:].max(0))])
    order = choose_axis_order(a1)  # both epochs share the key layout AND the key ranges, planned on epoch 1
    t1, _ = voxel_statistics(*epoch1, voxel_size, origin, group=group, with_stats=False, exchange=exchange, bounds=gb, axis_order=order)
    t2, _ = voxel_statistics(*epoch2, voxel_size, origin, group=group, with_stats=False, exchange=exchange, bounds=gb, thresholds=t1.thresholds,
                             axis_order=order)
    dev = t1.voxel_keys.device
    k1, k2 = keys_as_int64(t1.voxel_keys).contiguous(), keys_as_int64(t2.voxel_keys).contiguous()
    m1 = torch.empty(t1.nvoxels, dtype=torch.int32, device=dev)
    m2 = torch.empty(t2.nvoxels, dtype=torch.int32, device=dev)
    L.call("vapc_join_sorted_keys", E._ptr(k1), t1.nvoxels, E._ptr(k2), t2.nvoxels, E._ptr(m1), E._ptr(m2), E._stream())
    m1 = m1.to(torch.int64) & 0xFFFFFFFF
    m2 = m2.to(torch.int64) & 0xFFFFFFFF
    i1 = torch.nonzero(m1 != 0xFFFFFFFF).flatten()
    out = dict(table1=t1, table2=t2, i1=i1, i2=m1[i1], only1=torch.nonzero(m1 == 0xFFFFFFFF).flatten(), only2=torch.nonzero(m2 == 0xFFFFFFFF).flatten())
    only = dict(density=False, cog=True, std=False, cov=True, eig=False, eig_n=False, feat=False)
    s1, s2 = t1.stats(**only), t2.stats(**only)
    out["stats1"], out["stats2"] = s1, s2
    if i1.numel():
        out.update(E.mahalanobis(s1.cog, s1.cov, s2.cog, s2.cov, i1, out["i2"], alpha=alpha))
    return out


# ------------------------------------------------------------------------------------------------
# results of the merged table brought back to the ranks' own rows and points (the reverse exchange)
# ------------------------------------------------------------------------------------------------
def _owner_rows(table: ShardedVoxelTable, keys_i64: torch.Tensor) -> torch.Tensor:
    """Rows of this rank's merged table holding the given (present) global keys."""
    return torch.searchsorted(keys_as_int64(table.voxel_keys).contiguous(), keys_i64.contiguous())


def rows_from_owner(table: ShardedVoxelTable, shard: LocalShard, columns: torch.Tensor) -> torch.Tensor:
    """columns: [k, V_owned] fp64 columns of this rank's merged table (centroids, counts as fp64, ...).  Returns [k, v_local]:
    the same columns for every LOCAL partial row, i.e. the values of the voxel the row was merged into.  Rows a rank owns
    itself are looked up locally; for the rows it sent away the owners answer in ONE all-to-all of k doubles per partial
    row — the reverse of the accumulator exchange."""
    k = columns.shape[0]
    dev = columns.device
    reply = columns[:, _owner_rows(table, shard.recv_keys)].t().contiguous()  # [n_in, k], grouped by source
    v = shard.gkeys.numel()
    back = torch.empty((v - shard.n_own, k), dtype=torch.float64, device=dev)
    dist.all_to_all_single(back, reply, output_split_sizes=shard.send_remote, input_split_sizes=shard.recv_remote, group=shard.group)
    lo, hi = shard.lo, shard.lo + shard.n_own
    own = columns[:, _owner_rows(table, shard.gkeys[lo:hi])]
    return torch.cat([back[:lo].t(), own, back[lo:].t()], dim=1).contiguous()


def broadcast_to_points(table: ShardedVoxelTable, shard: LocalShard, columns: torch.Tensor) -> torch.Tensor:
    """[k, n_local_points]: merged-table columns for every point of this rank's chunk, original point order (the reference's
    per-point broadcast, vapc.py:724-726, 843-850)."""
    per_row = rows_from_owner(table, shard, columns)
    return torch.stack([shard.cloud.broadcast(per_row[j].contiguous()) for j in range(per_row.shape[0])])


def point_distance(table: ShardedVoxelTable, shard: LocalShard, cog: torch.Tensor) -> torch.Tensor:
    """distance of every local point to the GLOBAL centroid of its voxel (vapc.py:776-795)."""
    from . import _lib as L
    from . import engine as E

    c = shard.cloud
    per_row = rows_from_owner(table, shard, table.internal_rows(cog))  # the shard's cloud lives in the internal axis order
    out = torch.empty(c.n, dtype=torch.float64, device=cog.device)
    L.call("vapc_point_distance", E._ptr(c.x), E._ptr(c.y), E._ptr(c.z), c.n, E._ptr(c.point2voxel), E._ptr(per_row), c.nvoxels, E._ptr(out), E._stream())
    return out


def closest_to_cog(table: ShardedVoxelTable, shard: LocalShard, cog: torch.Tensor, point_offset: Optional[int] = None):
    """Closest-to-centroid subsample over all ranks (vapc.py:799-825, 1188-1195).  Every rank proposes, per partial row, its
    nearest point to the voxel's FINAL centroid (vapc_closest_to_cog against the centroids rows_from_owner brought back);
    ONE all-to-all sends the (distance, global point index) candidates of the rows that left to their owners, which keep the
    lexicographic minimum (vapc_reduce_candidates): lowest global index among exactly tied points, as on one GPU.
    point_offset: global index of this rank's first point (default: chunks numbered in rank order).
    Returns (min_distance [V_owned] f64, global point index [V_owned] int64)."""
    from . import _lib as L
    from . import engine as E

    c = shard.cloud
    dev = cog.device
    if point_offset is None:
        sizes = torch.zeros(dist.get_world_size(shard.group), dtype=torch.int64, device=dev)
        sizes[dist.get_rank(shard.group)] = c.n
        dist.all_reduce(sizes, group=shard.group)
        point_offset = int(sizes[: dist.get_rank(shard.group)].sum().item())
    md, am = c.closest_to_cog(cog=rows_from_owner(table, shard, table.internal_rows(cog)))
    gidx = (am.to(torch.int64) & 0xFFFFFFFF) + point_offset
    lo, hi = shard.lo, shard.lo + shard.n_own
    cand = torch.stack([md, gidx.view(torch.float64)], dim=1)  # the index travels as a bit pattern
    out_c = torch.cat([cand[:lo], cand[hi:]]).contiguous()
    in_c = torch.empty((shard.recv_keys.numel(), 2), dtype=torch.float64, device=dev)
    dist.all_to_all_single(in_c, out_c, output_split_sizes=shard.recv_remote, input_split_sizes=shard.send_remote, group=shard.group)
    rows = torch.cat([_owner_rows(table, shard.gkeys[lo:hi]), _owner_rows(table, shard.recv_keys)]).contiguous()
    d = torch.cat([md[lo:hi], in_c[:, 0]]).contiguous()
    g = torch.cat([gidx[lo:hi], in_c[:, 1].contiguous().view(torch.int64)]).contiguous()
    best_d = torch.empty(table.nvoxels, dtype=torch.float64, device=dev)
    best_i = torch.empty(table.nvoxels, dtype=torch.int64, device=dev)
    L.call("vapc_reduce_candidates", E._ptr(rows), E._ptr(d), E._ptr(g), rows.numel(), table.nvoxels, E._ptr(best_d), E._ptr(best_i), E._stream())
    return best_d, best_i


def mode_count(table: ShardedVoxelTable, shard: LocalShard, attr, thresholds: Sequence[float] = (), want_mode=True):
    """mode and mode_count,p of an integer attribute over all ranks (vapc.py:407-433).  Every rank sorts its points by (partial
    row, value) and run-length encodes them (vapc_run_lengths); the (key, value, count) runs of the rows that left the rank go to
    their owners in ONE all-to-all (24 B per run instead of moments), the owner adds the counts of equal (voxel, value) pairs
    (sort + weighted run lengths) and evaluates mode / mode_count on the merged runs (vapc_mode_count_runs).  Counts are
    integers: the result is identical to one GPU.  Returns (mode [V_owned] int64 or None, {p: mode_count [V_owned] int64})."""
    from . import _lib as L
    from . import engine as E

    c = shard.cloud
    dev = c.device
    group = shard.group
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    a = E._dev_f64(attr, dev)
    if a.numel() != c.n:
        raise ValueError("attribute length differs from the chunk")
    amax = a.max().reshape(1).clone() if c.n else torch.zeros(1, dtype=torch.float64, device=dev)
    dist.all_reduce(amax, op=dist.ReduceOp.MAX, group=group)
    attr_bits = max(1, int(max(float(amax.item()), 0.0)).bit_length())  # every rank must pack the values alike
    if attr_bits > 32:
        raise ValueError("attribute values too large for mode / mode_count")
    keys = torch.empty(c.n, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    L.call("vapc_mode_keys", E._ptr(c.point2voxel), E._ptr(a), c.n, attr_bits, 8, E._ptr(keys), E._ptr(status), E._stream())
    ks, _ = E.sort_pairs(keys, max(1, int(c.nvoxels - 1).bit_length()) + attr_bits)
    bad = status.to(torch.int64)
    dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group)
    if int(bad.item()) & 2:
        raise ValueError("Cannot compute 'mode' / 'mode_count' for negative or non-finite attribute values")
    run_keys, run_w = E.run_lengths(ks)
    rows = run_keys >> attr_bits  # partial row of every run; runs are in row order = destination order
    payload = torch.stack([shard.gkeys[rows], run_keys & ((1 << attr_bits) - 1), run_w], dim=1).contiguous()  # [R, 3] int64
    row_edges = [0]
    for d in range(world):
        row_edges.append(row_edges[-1] + (shard.n_own if d == rank else shard.send_remote[d]))
    cuts = torch.searchsorted(rows.contiguous(), torch.tensor(row_edges, dtype=torch.int64, device=dev)).cpu().tolist()
    send = [cuts[d + 1] - cuts[d] for d in range(world)]
    own = payload[cuts[rank]:cuts[rank + 1]]
    send[rank] = 0
    send_t = torch.tensor(send, dtype=torch.int64, device=dev)
    recv_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv_t, send_t, group=group)
    recv = recv_t.cpu().tolist()
    out_p = torch.cat([payload[: cuts[rank]], payload[cuts[rank + 1]:]]).contiguous()
    in_p = torch.empty((int(sum(recv)), 3), dtype=torch.int64, device=dev)
    dist.all_to_all_single(in_p, out_p, output_split_sizes=recv, input_split_sizes=send, group=group)
    runs = torch.cat([own, in_p])
    merged_row = _owner_rows(table, runs[:, 0])
    comp = ((merged_row << attr_bits) | runs[:, 1]).contiguous()
    cs, order = E.sort_pairs(comp, max(1, int(table.nvoxels - 1).bit_length()) + attr_bits)
    weights = runs[:, 2][order.to(torch.int64) & 0xFFFFFFFF].to(torch.int32).contiguous()
    mk, mw = E.run_lengths(cs, weights)
    mode = torch.empty(table.nvoxels, dtype=torch.int64, device=dev) if want_mode else None
    counts = {}
    todo = list(thresholds) or [None]
    for i, p in enumerate(todo):
        mc = torch.empty(table.nvoxels, dtype=torch.int64, device=dev) if p is not None else None
        L.call("vapc_mode_count_runs", E._ptr(mk), E._ptr(mw), mk.numel(), attr_bits, table.nvoxels, float(p if p is not None else 0.0),
               E._ptr(mode if i == 0 else None), E._ptr(mc), E._stream())
        if p is not None:
            counts[p] = mc
    return mode, counts
