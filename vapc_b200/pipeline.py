"""Host-buffer streaming pipeline: clouds arrive as pinned host columns, voxel tables leave as pinned host
columns, and consecutive clouds overlap on the three engines of the GPU —

    copy-in stream   H2D of cloud k+1            (PCIe, host -> device)
    compute stream   keys / sort / moments / statistics of cloud k
    copy-out stream  D2H of the voxel table of cloud k-1 (PCIe is full duplex)

The per-voxel statistics path moves 24 B/point in and ~224 B/voxel out, while the kernels need < 10 ms per
1e8 points, so end to end it is PCIe-bound; overlapping the two copy directions and the compute of neighbouring
clouds is what the hardware allows.  Used by `bench.py` for the `e2e` figure and usable directly:

    pipe = HostStreamPipeline(voxel_size, origin, device)
    for result in pipe.run(iterable_of_(hx, hy, hz)_pinned_tensors):
        result["cog"], result["cov"], ...   # pinned host tensors, valid until the next result is requested

With `las=(scales, offsets)` the clouds are the int32 X, Y, Z of LAS files (`vapc_b200.las.read_las_ints`): 12 B/point over
PCIe instead of 24, the key kernels form X * scale + offset themselves (same doubles, same results).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Iterator, Optional

import torch

from . import engine as E

STAT_NAMES = ("density", "cog", "std", "cov", "eig", "eig_n", "feat")


def full_statistics(dx, dy, dz, voxel_size, origin, las=None, bounds=None, columns=STAT_NAMES) -> Dict[str, torch.Tensor]:
    """The whole single-GPU hot path on device columns -> dict of device tensors (the voxel table).  las = (scales, offsets):
    dx, dy, dz are int32 LAS coordinates.  bounds: [min x, min y, min z, max x, max y, max z] when the caller knows them (a LAS
    header does): the bounding-box pass over the coordinates is skipped.  columns: the statistics to compute and return."""
    if las is not None:
        c = E.VoxelCloud.build(None, None, None, voxel_size, origin, keep_columns=False, las=(dx, dy, dz, las[0], las[1]), bounds=bounds)
    else:
        c = E.VoxelCloud.build(dx, dy, dz, voxel_size, origin, keep_columns=False, bounds=bounds)
    s = c.stats(**{k: True for k in columns})
    out = {"voxel_keys": c.voxel_keys, "count": c.count}
    out.update({k: getattr(s, k) for k in columns})
    out["_grid"] = c.grid
    return out


class HostStreamPipeline:
    def __init__(self, voxel_size, origin, device=None, compute_fn: Optional[Callable] = None, depth: int = 2, las=None, bounds=None,
                 columns=STAT_NAMES):
        """las = (scales, offsets): the clouds are int32 LAS coordinates.  bounds: coordinate bounds valid for every cloud of the
        run (the header of the LAS file the clouds are chunks of).  columns: which statistics come back (every column costs PCIe time:
        the full table is 224 B per voxel)."""
        E._require_cuda()
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.voxel_size, self.origin = voxel_size, list(origin)
        self.compute_fn = compute_fn or (lambda x, y, z: full_statistics(x, y, z, self.voxel_size, self.origin, las=las, bounds=bounds,
                                                                         columns=columns))
        self.depth = depth
        self.s_in = torch.cuda.Stream(self.device)
        self.s_compute = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._dev_in = [None] * depth  # device input column sets
        self._host_out = [dict() for _ in range(depth)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _stage_in(self, slot, cols, after: Optional[torch.cuda.Event]):
        """H2D of one cloud into device slot `slot` on the copy-in stream (after the slot's previous reader is done)."""
        with torch.cuda.stream(self.s_in):
            if after is not None:
                self.s_in.wait_event(after)
            bufs = self._dev_in[slot]
            if bufs is None or bufs[0].numel() != cols[0].numel() or bufs[0].dtype != cols[0].dtype:
                bufs = [torch.empty(c.numel(), dtype=c.dtype, device=self.device) for c in cols]
                self._dev_in[slot] = bufs
            for b, c in zip(bufs, cols):
                b.copy_(c, non_blocking=True)
                self.h2d_bytes += c.numel() * c.element_size()
            ev = torch.cuda.Event()
            ev.record(self.s_in)
        return ev

    def _stage_out(self, slot, table: Dict[str, torch.Tensor], done: torch.cuda.Event):
        """D2H of one voxel table into pinned host slot `slot` on the copy-out stream."""
        host = self._host_out[slot]
        res = {}
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            for k, t in table.items():
                if not isinstance(t, torch.Tensor):
                    res[k] = t
                    continue
                t.record_stream(self.s_out)
                h = host.get(k)
                if h is None or h.shape != t.shape or h.dtype != t.dtype:
                    h = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                    host[k] = h
                h.copy_(t, non_blocking=True)
                self.d2h_bytes += t.numel() * t.element_size()
                res[k] = h
            ev = torch.cuda.Event()
            ev.record(self.s_out)
        return res, ev

    def run(self, clouds: Iterable) -> Iterator[Dict[str, torch.Tensor]]:
        """Yield one host-resident voxel table per input cloud, in order.  A yielded table is valid until the next one
        is requested (its pinned buffers are then reused)."""
        it = iter(clouds)
        pending_out = []  # (result, event)
        compute_done = [None] * self.depth
        nxt = next(it, None)
        k = 0
        h2d_ev = self._stage_in(0, nxt, None) if nxt is not None else None
        while nxt is not None:
            slot = k % self.depth
            cur_ev = h2d_ev
            nxt = next(it, None)
            if nxt is not None:  # prefetch cloud k+1 while cloud k computes
                h2d_ev = self._stage_in((k + 1) % self.depth, nxt, compute_done[(k + 1) % self.depth])
            with torch.cuda.stream(self.s_compute):
                self.s_compute.wait_event(cur_ev)
                table = self.compute_fn(*self._dev_in[slot])
                done = torch.cuda.Event()
                done.record(self.s_compute)
            compute_done[slot] = done
            if len(pending_out) >= self.depth:  # the host slot about to be overwritten must have been delivered
                res, ev = pending_out.pop(0)
                ev.synchronize()
                yield res
            pending_out.append(self._stage_out(slot, table, done))
            k += 1
        for res, ev in pending_out:
            ev.synchronize()
            yield res
