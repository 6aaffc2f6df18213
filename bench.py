#!/usr/bin/env python
"""bench.py — points/s of the voxelise + per-voxel statistics path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--impl reference]

A step is one pass of the hot path over one synthetic cloud: packed voxel keys + records grouped by key prefix -> grouping inside
the buckets (cell table on dense grids, else stable radix passes) -> FP64 (count, mean, M2) moments -> density, centroid, std,
covariance, eigenvalues, 8 eigen-features (+ the point->voxel map).
  N = 1: BASELINE.json configs[1], a 100M-point uniform cloud, voxel 0.1 m, full attribute set.
  N > 1: configs[2], a clustered terrain + trees cloud of 1.25e8 points per GPU sharded in scan-line order, voxel 0.25 m; the same
         cloud shuffled and round 1's slab layout are measured too (extra), and the sharded table is CHECKED after the timed region
         against plain torch on the ranks' own points ("check"; a failed check exits non-zero).

value   whole-job points/s with the inputs already resident in HBM (CUDA events, max over ranks)
e2e     the same metric through the public API with HOST buffers: pinned host -> device copies of the coordinates and the
        device -> pinned-host read of the voxel table inside the timed region; platform_gbs = what plain concurrent copies
        reach on this box, frac_of_platform = the e2e's PCIe traffic against it; e2e.facade = the drop-in `Vapc` class
        (DataFrame in, reduced DataFrame out)
roofline  the kernel with the largest share of the step: algorithmic bytes / CUDA-event time on the launching stream,
        against MEASURED_PEAKS.json; traffic = its DRAM bytes in the committed ncu capture (profiles/r2_ncu_traffic.json)
cpu_baseline  the unmodified reference (oracle/_ref, kind "reference"; oracle/ref_port.py, kind "port", where that copy is absent)
        on a bounded spatial crop, 1 core
--impl reference  times the reference alone on all host cores and prints the same line with "impl": "reference"
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "points/sec voxelise+stats"
UNIT = "points/s"
VOXEL_SIZE = 0.1
VOXEL_SIZE_MULTI = 0.25  # configs[2]
ORIGIN = [0.0, 0.0, 0.0]
EXTENT = (50.0, 50.0, 4.0)  # metres: 500 x 500 x 40 voxels = 1e7 cells -> ~10 points per voxel at 1e8 points
QUANT = 1e-4  # LAS-like coordinate quantisation


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=0, help="points per GPU (default: 1e8 at N = 1 = configs[1], 1.25e8 at N > 1 = configs[2])")
    ap.add_argument("--bounds", default="header", choices=["header", "computed"],
                    help="N = 1: the cloud's bounding box comes with the data, as a LAS header gives it (default), or is computed by one more pass")
    ap.add_argument("--workload", default="configs1", choices=["configs1", "configs2"],
                    help="N = 1 only: configs1 (the bench line) or configs2 = one GPU's chunk of the N > 1 workload, voxelised alone (for profiles)")
    ap.add_argument("--no-extra", action="store_true", help="N > 1: skip the shuffled and slab layouts")
    ap.add_argument("--no-facade", action="store_true", help="skip the drop-in facade e2e (DataFrame in, reduced DataFrame out)")
    ap.add_argument("--no-las-e2e", action="store_true", help="skip the LAS-integer-input variant of the e2e measurement")
    ap.add_argument("--impl", default="vapc_b200", choices=["vapc_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="points in the CPU sample (0 = auto)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic data
# ------------------------------------------------------------------------------------------------
def gen_crop_host(n_sample, seed=2):
    """A spatial crop of the same cloud law at the same density (10 points per voxel at 1e8 points): the box
    [0, lx) x [0, 0.5) x [0, 4) holding n_sample points.  Generated on the host (numpy), used by the CPU legs."""
    import numpy as np

    density = 100_000_000 / (EXTENT[0] * EXTENT[1] * EXTENT[2])  # points per m^3 of config[1]
    lx = n_sample / (density * 0.5 * EXTENT[2])
    rng = np.random.default_rng(seed)
    x = rng.integers(0, max(1, int(round(lx / QUANT))), n_sample) * QUANT
    y = rng.integers(0, int(round(0.5 / QUANT)), n_sample) * QUANT
    z = rng.integers(0, int(round(EXTENT[2] / QUANT)), n_sample) * QUANT
    return x, y, z, f"spatial crop [0,{lx:.2f})x[0,0.5)x[0,4) m of the workload at its density: {n_sample} points"


# ------------------------------------------------------------------------------------------------
# CPU legs
# ------------------------------------------------------------------------------------------------
CPU_KIND_TEXT = {
    "reference": "the UNMODIFIED reference (oracle/_ref = its package directory copied unchanged by oracle/build_ref.py) through its own API: "
                 "Vapc(...).compute_requested_attributes() with point_count, point_density, center_of_gravity, std_of_cog, covariance_matrix, "
                 "eigenvalues, geometric_features; single-threaded pandas/numpy incl. its per-voxel Python eigen step",
    "port": "oracle/ref_port.py (the reference package is not on this box): keeps the reference's single-threaded pandas data flow incl. its "
            "per-voxel Python eigen step",
}


def cpu_reference_once(x, y, z, voxel_size=None):
    """One pass of the reference's CPU path over host columns: (kind, seconds, voxels).  kind "reference" = the unmodified reference
    (oracle/ref_harness.py over oracle/_ref), "port" = oracle/ref_port.py when the reference package is not there."""
    from oracle import ref_harness

    voxel_size = VOXEL_SIZE if voxel_size is None else voxel_size
    if ref_harness.available():
        ref_harness.import_reference()
        t0 = time.perf_counter()
        v = ref_harness.full_attribute_set(x, y, z, voxel_size, ORIGIN)
        dt = time.perf_counter() - t0
        df, kind = v.df, "reference"
    else:
        from oracle import ref_port

        t0 = time.perf_counter()
        df = ref_port.full_attribute_set(x, y, z, voxel_size, ORIGIN, with_eigen=True)
        dt = time.perf_counter() - t0
        kind = "port"
    nvox = int(df[["voxel_x", "voxel_y", "voxel_z"]].drop_duplicates().shape[0])
    return kind, dt, nvox


def gen_crop_host_clustered(n_sample, seed=3):
    """A square crop of the configs[2] cloud law (tools/synth.py: terrain sheet + trees, 1e6 points per 1000 m2) holding n_sample points.
    numpy only: the CPU legs fork worker processes, and torch's thread pools do not survive a fork."""
    import numpy as np

    side = float(np.sqrt(n_sample / 1000.0))
    rng = np.random.default_rng(seed)
    terrain = lambda x, y: 5.0 * np.sin(x / 37.0) * np.cos(y / 53.0) + 0.02 * x  # noqa: E731
    nt = int(0.6 * n_sample)
    tx, ty = rng.random(nt) * side, rng.random(nt) * side
    tz = terrain(tx, ty) + (rng.random(nt) - 0.5) * 0.1
    ntrees = max(1, int(round(0.2 * side * side)))
    cx, cy, ch = rng.random(ntrees) * side, rng.random(ntrees) * side, 5.0 + 20.0 * rng.random(ntrees)
    nb = n_sample - nt
    t = rng.integers(0, ntrees, nb)
    r = rng.standard_normal((nb, 3))
    crown = rng.random(nb) < 0.8
    h = rng.random(nb)
    bx = cx[t] + np.where(crown, r[:, 0] * 1.5, r[:, 0] * 0.1)
    by = cy[t] + np.where(crown, r[:, 1] * 1.5, r[:, 1] * 0.1)
    bz = terrain(cx[t], cy[t]) + np.where(crown, ch[t] * (0.6 + 0.4 * h) + r[:, 2] * 0.8, ch[t] * 0.6 * h)
    x, y, z = (np.round(np.concatenate(c) / QUANT) * QUANT for c in ((tx, bx), (ty, by), (tz, bz)))
    order = np.lexsort((x, np.floor(y / 2.0)))  # scan lines: 2 m strips in y, x ascending inside a strip
    return x[order], y[order], z[order], (f"square crop of {side:.1f} m x {side:.1f} m of the configs[2] cloud law at its density: "
                                          f"{n_sample} points")


def _cpu_worker(job):
    n_sample, seed, clustered = job
    x, y, z, _ = gen_crop_host_clustered(n_sample, seed=seed) if clustered else gen_crop_host(n_sample, seed=seed)
    return cpu_reference_once(x, y, z, VOXEL_SIZE_MULTI if clustered else VOXEL_SIZE)


def run_reference_arm(args):
    """The reference's CPU path (the unmodified reference from oracle/_ref; oracle/ref_port.py where that is absent) on the host cores.  The reference itself is
    single-threaded; its own way to use more cores is to cut the cloud into spatial tiles and process them independently
    (main.py:97-137), so every host core gets its own crop of the workload here and the throughput is the sum."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor

    cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    budget = 150.0 / max(1, args.steps + args.warmup)  # seconds per step
    clustered = args.gpus > 1  # the arm follows the GPU arm's config: configs[1] at N = 1, configs[2] at N > 1
    n_sample = args.cpu_sample or int(min(200_000, max(2_000, budget * (2_000 if clustered else 5_000))))
    _, _, _, what = gen_crop_host_clustered(n_sample) if clustered else gen_crop_host(n_sample)
    what = what.replace(f"{n_sample} points", f"{n_sample} points per core")
    nvox, kind = 0, "port"
    with ProcessPoolExecutor(max_workers=cores) as pool:
        for _ in range(args.warmup):
            list(pool.map(_cpu_worker, [(n_sample, 100 + c, clustered) for c in range(cores)]))
        t0 = time.perf_counter()
        for k in range(args.steps):
            res = list(pool.map(_cpu_worker, [(n_sample, 1000 * (k + 1) + c, clustered) for c in range(cores)]))
            nvox = sum(r[2] for r in res)
            kind = res[0][0]
        total = time.perf_counter() - t0
    value = n_sample * cores * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": (workload_config_multi(args.points or 125_000_000, args.gpus, VOXEL_SIZE_MULTI) if clustered
                   else workload_config(args.points or 100_000_000, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": what + f" ({nvox} voxels per step over {cores} processes, one independent crop each: tiling is the reference's "
                                          "own scaling mechanism, main.py:97-137); " + CPU_KIND_TEXT[kind]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(points_per_gpu, gpus, bounds="header"):
    return {
        "workload": f"configs[1]: {points_per_gpu / 1e6:g}M-point synthetic uniform cloud per GPU, voxel_size={VOXEL_SIZE} m, "
                    "full attribute set (count, density, centroid, std, covariance, eigenvalues, 8 eigen-features) + point->voxel map",
        "points_per_gpu": points_per_gpu, "voxel_size": VOXEL_SIZE, "extent_m": list(EXTENT),
        "bounds": ("the cloud's bounding box is an input (the LAS header states it; the reference's DataHandler reads the same header): no "
                   "min/max pass; extra.bounds_computed is the step with that pass" if bounds == "header" else
                   "computed: one more pass over x, y, z (vapc_minmax_f64)"),
        "l2": "inputs (24 B/point, 2.4 GB at 1e8 points) are larger than the 126 MB L2; no flush needed",
        "parallelism": "one GPU",
    }


def workload_config_multi(points_per_gpu, gpus, vs, bounds="computed"):
    return {
        "workload": f"configs[2]: {points_per_gpu * gpus / 1e9:g}B-point synthetic clustered (terrain sheet + trees) cloud sharded over {gpus} GPUs in "
                    f"scan-line order ({points_per_gpu / 1e6:g}M points per GPU: rank r holds the strip y in [r W, (r+1) W) of a 1000 m wide cloud, W = "
                    f"{125.0 * points_per_gpu / 1.25e8:g} m), voxel_size={vs} m, full attribute set; partial voxel tables exchanged by key range, then merged",
        "points_per_gpu": points_per_gpu, "voxel_size": vs,
        "l2": "inputs (24 B/point, 3 GB per GPU) are larger than the 126 MB L2; no flush needed",
        "parallelism": f"points sharded in contiguous chunks over {gpus} GPUs; every rank voxelises its chunk alone (" +
                       ("the chunks' bounding boxes are an input, as their LAS headers state them: no min/max pass and no collective for the bounds; "
                        "extra.bounds_computed is the step without them" if bounds == "header" else "local bounds by a min/max pass, all-gathered") + "), "
                       "the key space is cut by a global histogram, the partial rows that leave a rank are written into the owner's receive buffer over "
                       "NVLink (peer memory), the owner merges; extra.shuffled = the same cloud dealt at random, extra.slab_configs1 = round 1's layout",
    }


def cpu_baseline_leg(args):
    n_sample = args.cpu_sample or 100_000
    cx, cy, cz, what = gen_crop_host(n_sample)
    kind, dt, cv = cpu_reference_once(cx, cy, cz)
    return {"value": n_sample / dt, "unit": UNIT, "cores": 1, "kind": kind, "seconds": round(dt, 2),
            "sample": what + f" ({cv} voxels); " + CPU_KIND_TEXT[kind] + "; host has %d logical cores" % (os.cpu_count() or 0)}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled in the background from program start; stop(windows) keeps the samples whose arrival time
    falls inside the timed regions (nvidia-smi needs ~0.3 s to start, longer than one timed region)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_first_sample(self, timeout=20.0):
        """nvidia-smi attaches to every GPU of the box while it starts (0.3 s warm, seconds on a fresh box) and CUDA calls of this
        process stall meanwhile (seen: one step of 120 ms among steps of 5 ms): nothing is timed before its first line has arrived."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and self.proc.poll() is None and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if not any(a - 0.03 <= ts <= b + 0.03 for a, b in windows):
                continue
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "windows": "timed region + e2e region"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(name, n, v, key_bytes, passes, low_passes):
    """Algorithmic bytes of one C-ABI call ('bytes per unit'); n points, v voxels, `passes` radix passes for
    all key bits, `low_passes` for the key bits below the bucket digit."""
    k = key_bytes
    return {
        "vapc_minmax_f64": 24 * n,
        "vapc_pack_keys_f64": (24 + k + 32) * n,
        "vapc_pack_records_bucketed": 8 * n + (24 + k + k + 32) * n + 2 * n,
        "vapc_sort_pairs": (k + passes * (2 * k + 8) - 4) * n,
        "vapc_sort_pairs_segmented": (k + low_passes * (2 * k + 8) - 4) * n,
        "vapc_sort_cells_segmented": (2 * k + 8) * n,
        "vapc_count_voxels": k * n,
        "vapc_reduce_moments": (k + 4 + 32) * n + (k + 12 + 72) * v,
        "vapc_point2voxel_dense": (k + 4) * n + (k + 4) * v,
        "vapc_voxel_stats": (k + 4 + 72) * v + (8 + 24 + 24 + 48 + 24 + 24 + 64) * v,
    }.get(name, 0)


def kernel_algorithmic_bytes(kernel, n, v, key_bytes, low_passes):
    """(short name, algorithmic bytes of this kernel per STEP, launches per step that carry them) for a kernel name as recorded by the
    library's launch macro.  The radix pass runs low_passes times over the n points (the first pass takes the payload from the position);
    the multi-GPU step launches it a few more times on the few received rows, whose bytes are left out (their time is not)."""
    k = key_bytes
    table = [
        ("minmax_partial", 24 * n, 1),
        ("tile_digits_kernel", 8 * n + n / 3, 1),                     # x in; the tile's 256 16-bit digit counts (per 1536 points) out
        ("tile_scatter_kernel", (24 + k + k + 32) * n + 2 * n / 3, 1),  # x, y, z + the tile's offsets in; keys (input order), bucketed keys + records out
        ("tile_offsets_kernel", n, 1),                               # tile counts in, per-(tile, digit) first slots out: 6 B per 1.5 points
        ("pack_keys_kernel", (24 + k + 32) * n, 1),
        ("digit_hist_kernel", k * n, 1),
        ("onesweep_kernel", (low_passes * (2 * k + 8) - 4) * n, low_passes),  # keys + payload in and out, per pass
        ("cell_sort_kernel", (2 * k + 8) * n, 1),                     # dense grids: the bucketed keys in twice (count, place), sorted keys and positions out
        ("count_heads_kernel", k * n, 1),
        ("reduce_moments_kernel", (k + 4 + 32) * n + (k + 12 + 72) * v, 1),  # sorted keys + positions + records in, the voxel rows out
        ("p2v_lookup_kernel", (k + 4) * n, 1),
        ("p2v_fill_kernel", (k + 4) * v, 1),
        ("voxel_stats_kernel", (k + 4 + 72) * v + (8 + 24 + 24 + 48 + 24 + 24 + 64) * v, 1),
        # N > 1, the owner's merge fused with the statistics: the own rows in (64-bit global key, count, 9 moments), key + count + the
        # 27 statistic columns out (v = this rank's local voxels; all but ~1 % of them are rows it keeps)
        ("merge_own_rows_kernel", (8 + 4 + 72) * v + (k + 4 + 8 + 24 + 24 + 48 + 24 + 24 + 64) * v, 1),
        ("global_keys_kernel", (k + 8) * v, 1),
    ]
    for key, b, launches in table:
        if key in kernel:
            return key, float(b), launches
    return kernel, 0.0, 1


def time_steps(step, steps, barrier, world, device, dist, profile=False, L=None, settle=3):
    """K calls of step() between CUDA events (max over ranks).  profile: per-kernel / per-call CUDA events inside the same region.
    settle: untimed calls in front, in the SAME loop as the timed ones (the result of the previous call released before the next one
    starts): the caching allocator and the side stream's deferred frees reach the state the timed calls then repeat.  One event per
    step lands in step_ms (the spread of the K steps; the headline is the whole region / K)."""
    import torch

    records = []
    out = None
    for _ in range(settle):
        out = None
        out = step()
    out = None

    def observer(name, phase):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        records.append((name, phase, ev))

    launches0 = L.raw("vapc_launch_count")() if L is not None else 0
    barrier()
    t0 = time.perf_counter()
    if profile:
        L.observer = observer
        L.call("vapc_profile_enable", 1)  # CUDA events around every kernel launch, on the launching stream
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    marks = []
    for _ in range(steps):
        out = None
        out = step()
        if not profile:
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
    t_end.record()
    if profile:
        L.observer = None
        L.call("vapc_profile_enable", 0)
    barrier()
    t1 = time.perf_counter()
    kernel_times = L.profile_collect() if profile else {}
    launches = (L.raw("vapc_launch_count")() - launches0) if L is not None else 0
    ms_total = t_start.elapsed_time(t_end)
    step_ms = [round(a.elapsed_time(b), 4) for a, b in zip([t_start] + marks[:-1], marks)]
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    return {"ms_per_step": ms_total / steps, "out": out, "records": records, "kernel_times": kernel_times, "launches": launches,
            "window": [t0, t1], "step_ms": step_ms}


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel of the committed `ncu --set full` capture
    of this bench (profiles/r2_ncu_traffic.json, written by tools/ncu_summary.py from the .ncu-rep): {kernel: GB}, or {} ."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return d
    except Exception:
        return {}


def platform_copy_bandwidth(device, world, dist, barrier, nbytes=1 << 30, reps=3):
    """What the box gives: every rank copies nbytes host -> device and nbytes device -> host at once (plain pinned cudaMemcpyAsync on
    two streams), all ranks together.  Aggregate GB/s over all ranks (both directions added), wall clock between barriers."""
    import torch

    hin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    din = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dout = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)
    best = 0.0
    for r in range(reps + 1):
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        if r:
            best = max(best, 2 * nbytes * world / dt / 1e9)
    if world > 1:
        t = torch.tensor([best], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        best = float(t.item())
    return best


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    # stdout carries exactly ONE JSON line: anything a library prints there meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist

    from tools import synth
    from vapc_b200 import _lib as L
    from vapc_b200 import engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        from vapc_b200 import dist as D
    multi = world > 1
    chunk_alone = not multi and args.workload == "configs2"
    n = args.points or (125_000_000 if (multi or chunk_alone) else 100_000_000)
    vs = VOXEL_SIZE_MULTI if (multi or chunk_alone) else VOXEL_SIZE
    FULL = dict(density=True, cog=True, std=True, cov=True, eig=True, eig_n=True, feat=True)

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    windows = []
    extra = {}
    check = None
    if chunk_alone:
        # ---- N = 1, one GPU's chunk of configs[2] voxelised alone (what every rank of the N > 1 run does before the exchange) ----
        x, y, z = synth.clustered_chunk(n, 0, 1, device, seed=3, scan_order=True)
        header_bounds = [float(c.min()) for c in (x, y, z)] + [float(c.max()) for c in (x, y, z)]  # what the chunk's LAS header states

        def make_step(bounds):
            def step():
                cloud = E.VoxelCloud.build(x, y, z, vs, ORIGIN, keep_columns=False, bounds=bounds)
                st = cloud.stats(**FULL)
                cloud.point2voxel  # the point->voxel map is formed on a side stream beside the statistics: reading it joins the step's stream
                return cloud, st
            return step

        step_device = make_step(header_bounds if args.bounds == "header" else None)
        workload = dict(workload_config_multi(n, 1, vs, args.bounds), parallelism=f"one GPU: the chunk voxelised alone, no exchange (bounds: {args.bounds})")
    elif not multi:
        # ---- N = 1: configs[1] -----------------------------------------------------------------------------------------
        x, y, z = synth.uniform_cloud(n, seed=2, device=device, extent=EXTENT)
        # the cloud's bounding box as the header of its LAS file states it (the reference's DataHandler reads the same header)
        header_bounds = [0.0, 0.0, 0.0] + [e - QUANT for e in EXTENT]

        def make_step(bounds):
            def step():
                cloud = E.VoxelCloud.build(x, y, z, vs, ORIGIN, keep_columns=False, bounds=bounds)
                st = cloud.stats(**FULL)
                cloud.point2voxel  # the point->voxel map is formed on a side stream beside the statistics: reading it joins the step's stream
                return cloud, st
            return step

        step_device = make_step(header_bounds if args.bounds == "header" else None)
        workload = workload_config(n, world, args.bounds)
    else:
        # ---- N > 1: configs[2], scan-line order ----------------------------------------------------------------------------
        x, y, z = synth.clustered_chunk(n, rank, world, device, seed=3, scan_order=True)

        # every chunk's bounding box as the header of its LAS file states it, exchanged once at load time (outside the step)
        chunk_bounds = D.gather_bounds(x, y, z) if args.bounds == "header" else None

        def step_device():
            return D.voxel_statistics(x, y, z, vs, ORIGIN, all_bounds=chunk_bounds)

        workload = workload_config_multi(n, world, vs, args.bounds)

    sampler.wait_first_sample()
    for _ in range(max(args.warmup, 3)):
        cloud, st = step_device()
    torch.cuda.synchronize()
    nvox = cloud.nvoxels
    exchange_kind = getattr(cloud, "exchange", None)
    cloud_axis_order = tuple(getattr(cloud, "axis_order", (0, 1, 2)))
    key_bytes = cloud.grid.key_bytes
    total_bits = cloud.grid.total_bits
    sorted_by = getattr(cloud, "sorted_by", None)
    passes = (total_bits + 7) // 8
    low_passes = (total_bits - min(8, cloud.grid.bits[0]) + 7) // 8  # inside the key-prefix buckets

    # ---- timed region: K steps between two CUDA events, nothing else on the stream -----------------
    plain = time_steps(step_device, args.steps, barrier, world, device, dist, profile=False, L=L)
    windows.append(plain["window"])
    ms_per_step = plain["ms_per_step"]
    value = n * world / (ms_per_step * 1e-3)
    launches = plain["launches"]
    step_ms = plain["step_ms"]
    plain = None
    # ---- the same K steps once more, instrumented: CUDA events around every kernel launch (on the launching stream) and every C-ABI
    # call, for the roofline and the kernel table.  The ~40 event records per step cost 0.1-0.2 ms of a 5.5 ms step, which is why the
    # headline region above carries none; the point->voxel lookups stay on the step's own stream here so that every kernel is timed alone.
    side_stream = E.P2V_SIDE_STREAM
    E.P2V_SIDE_STREAM = False
    step_device()
    timed = time_steps(step_device, args.steps, barrier, world, device, dist, profile=True, L=L)
    E.P2V_SIDE_STREAM = side_stream
    windows.append(timed["window"])
    ms_instrumented = timed["ms_per_step"]
    cloud, st = timed["out"]
    kernel_times = timed["kernel_times"]
    if multi:
        nv_t = torch.tensor([cloud.nvoxels, cloud.partial_rows_sent], dtype=torch.int64, device=device)
        dist.all_reduce(nv_t)
        nvox_global, rows_exchanged = int(nv_t[0]), int(nv_t[1])

    per_call = {}
    pending = {}
    gaps = {}
    last_after = None
    for name, phase, ev in timed["records"]:
        if phase == "before":
            pending[name] = ev
            if last_after is not None:
                gaps.setdefault(f"{last_after[0]} -> {name}", []).append(last_after[1].elapsed_time(ev))
        else:
            per_call.setdefault(name, []).append(pending.pop(name).elapsed_time(ev))
            last_after = (name, ev)
    gaps = {k: round(sum(v) / len(v), 4) for k, v in gaps.items()}
    timed["records"] = None

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    n_local_vox = getattr(cloud, "local_voxels", nvox)
    breakdown = {}
    for name, ts in per_call.items():
        calls_per_step = len(ts) / args.steps
        avg = sum(ts) / len(ts)
        b = algorithmic_bytes(name, n, n_local_vox, key_bytes, passes, low_passes)
        breakdown[name] = {"ms_per_call": round(avg, 4), "calls_per_step": calls_per_step, "ms_per_step": round(avg * calls_per_step, 4),
                           "algorithmic_GB": round(b / 1e9, 4), "GBps": round(b / 1e9 / (avg * 1e-3), 1) if b else None}
    per_kernel = {}
    for kname, (cnt, ms) in kernel_times.items():
        short, b, big = kernel_algorithmic_bytes(kname, n, n_local_vox, key_bytes, max(1, low_passes))
        e = per_kernel.setdefault(short, {"launches_per_step": 0.0, "ms_per_step": 0.0, "algorithmic_GB_per_step": round(b / 1e9, 4),
                                          "algorithmic_GB_per_launch": round(b / big / 1e9, 4), "launches_with_these_bytes": big})
        e["launches_per_step"] += cnt / args.steps
        e["ms_per_step"] += ms / args.steps
    for e in per_kernel.values():
        # time per launch of the launches that carry the bytes (the multi-GPU step also sorts the few received rows with the same kernel:
        # their microseconds stay in, which can only lower the figure)
        e["ms_per_launch"] = round(e["ms_per_step"] / min(e["launches_per_step"], e["launches_with_these_bytes"]), 4)
        e["ms_per_step"] = round(e["ms_per_step"], 4)
        e["GBps"] = round(e["algorithmic_GB_per_step"] / (e["ms_per_step"] * 1e-3), 1) if e["algorithmic_GB_per_step"] and e["ms_per_step"] else None
    dom = max((k for k in per_kernel if per_kernel[k]["GBps"]), key=lambda k: per_kernel[k]["ms_per_step"])
    achieved = per_kernel[dom]["GBps"]
    traffic_tab = ncu_traffic()
    tkey = "n1" if not (multi or chunk_alone) else "multi"
    traffic = (traffic_tab.get(tkey) or {}).get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "frac_of_nominal_8000": round(achieved / 8000.0, 4),  # also against the nominal 8 TB/s of HBM3e
                "traffic": traffic, "traffic_unit": "GB per launch: ncu dram__bytes_read.sum + dram__bytes_write.sum of this kernel, read from "
                "profiles/r2_ncu_traffic.json (the committed --set full capture of this command" + ("" if not multi else "'s per-GPU workload on one GPU") + ")",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": int(per_kernel[dom]["algorithmic_GB_per_launch"] * 1e9),
                "ms_per_launch": per_kernel[dom]["ms_per_launch"], "launches_per_step": per_kernel[dom]["launches_per_step"],
                "share_of_step": round(per_kernel[dom]["ms_per_step"] / ms_instrumented, 3),
                "measured_in": "the instrumented repeat of the timed region (same K steps, CUDA events around every kernel launch on its stream): "
                               f"{ms_instrumented:.4f} ms per step there against ms_per_step without the events"}
    gpu_ms = sum(v["ms_per_step"] for v in breakdown.values())

    # ---- N = 1 extra: the same step when nobody knows the bounding box (one more pass over x, y, z) ----------------------------
    if not multi and args.bounds == "header":
        step_nb = make_step(None)
        step_nb()
        t2 = time_steps(step_nb, args.steps, barrier, world, device, dist, L=L)
        windows.append(t2["window"])
        extra["bounds_computed"] = {"ms_per_step": t2["ms_per_step"], "value": n / (t2["ms_per_step"] * 1e-3),
                                    "what": "the same step without caller bounds: vapc_minmax_f64 reads x, y, z once more (24 B/point)"}
        t2 = None

    # ---- N = 1 extra: one GPU's chunk of configs[2] (the N > 1 workload) voxelised alone: the N = 1 point of THAT scaling curve -------
    if not multi and not chunk_alone and not args.no_extra:
        keep = (x, y, z)
        del x, y, z
        cx, cy, cz = synth.clustered_chunk(125_000_000, 0, 1, device, seed=3, scan_order=True)

        chunk_hdr = ([float(c.min()) for c in (cx, cy, cz)] + [float(c.max()) for c in (cx, cy, cz)]) if args.bounds == "header" else None

        def step_chunk():
            c = E.VoxelCloud.build(cx, cy, cz, VOXEL_SIZE_MULTI, ORIGIN, keep_columns=False, want_point2voxel=False, bounds=chunk_hdr)
            return c, c.stats(**FULL)
        for _ in range(2):
            step_chunk()
        tc = time_steps(step_chunk, args.steps, barrier, world, device, dist, L=L)
        windows.append(tc["window"])
        extra["configs2_one_gpu"] = {"ms_per_step": tc["ms_per_step"], "value": 125_000_000 / (tc["ms_per_step"] * 1e-3), "points": 125_000_000,
                                     "voxel_size": VOXEL_SIZE_MULTI, "voxels": tc["out"][0].nvoxels, "key_bits": tc["out"][0].grid.total_bits,
                                     "what": "configs[2] at N = 1: the 1.25e8-point clustered chunk one rank of the N > 1 run holds, voxelised alone with "
                                             "the calls every rank issues before the exchange (bounds as in config.bounds, no point->voxel map).  `bench.py --gpus N` "
                                             "(N > 1) reports configs[2]: its value / (N x this value) is that workload's scaling efficiency"}
        tc = None
        del cx, cy, cz
        torch.cuda.empty_cache()
        x, y, z = keep
        del keep

    # ---- N > 1: correctness of the sharded table, then the other layouts ------------------------------------------------------
    if multi:
        from tools.dist_check import verify_sharded_table

        check = verify_sharded_table(cloud, st, x, y, z)
        check["voxels_table"] = nvox_global
        check["ok"] = bool(check["ok"] and check["voxels_torch_unique"] == nvox_global)
        # the per-GPU share of the same workload WITHOUT the other ranks: local pipeline only (N = 1 on this rank's chunk)
        own_bounds = chunk_bounds[rank] if chunk_bounds is not None else None

        def step_local():
            c = E.VoxelCloud.build(x, y, z, vs, ORIGIN, keep_columns=False, want_point2voxel=False, bounds=own_bounds)
            return c, c.stats(**FULL)
        step_local()
        tl = time_steps(step_local, args.steps, barrier, world, device, dist, L=L)
        windows.append(tl["window"])
        extra["one_gpu_same_workload"] = {"ms_per_step": tl["ms_per_step"], "value_per_gpu": n / (tl["ms_per_step"] * 1e-3),
                                          "what": "every rank's own chunk voxelised alone (no exchange, no merge), max over ranks: the N = 1 "
                                                  "figure of THIS workload; value / (N x value_per_gpu) is its scaling efficiency"}
        tl = None
        if chunk_bounds is not None:
            def step_nb():
                return D.voxel_statistics(x, y, z, vs, ORIGIN)
            step_nb()
            t2 = time_steps(step_nb, args.steps, barrier, world, device, dist, L=L)
            windows.append(t2["window"])
            extra["bounds_computed"] = {"ms_per_step": t2["ms_per_step"], "value": n * world / (t2["ms_per_step"] * 1e-3),
                                        "what": "the same step when nobody knows the chunks' bounding boxes: a min/max pass over x, y, z (24 B/point) and an "
                                                "all-gather of the boxes inside the step"}
            t2 = None
        nvlink_bytes = rows_exchanged * 88
        extra["exchange"] = {"partial_rows_sent_all_ranks": rows_exchanged, "bytes_over_nvlink_per_step": nvlink_bytes,
                             "share_of_rows": round(rows_exchanged / max(1, nvox_global), 4)}
        del cloud, st, x, y, z
        torch.cuda.empty_cache()
        if not args.no_extra:
            # shuffled: every rank holds a random sample of the whole cloud -> every rank touches every voxel
            xs, ys, zs = synth.clustered_chunk(n, rank, world, device, seed=3, scan_order=False)

            def step_shuffled():
                return D.voxel_statistics(xs, ys, zs, vs, ORIGIN)
            for _ in range(2):
                step_shuffled()
            tsh = time_steps(step_shuffled, max(3, args.steps // 2), barrier, world, device, dist, profile=True, L=L)
            windows.append(tsh["window"])
            tab = tsh["out"][0]
            sent = torch.tensor([tab.partial_rows_sent, tab.nvoxels], dtype=torch.int64, device=device)
            dist.all_reduce(sent)
            chk = verify_sharded_table(tab, tsh["out"][1], xs, ys, zs)
            ex_ms = sum(ms for k, (c_, ms) in tsh["kernel_times"].items() if "pack_partial_rows" in k) / max(3, args.steps // 2)
            extra["shuffled"] = {"ms_per_step": tsh["ms_per_step"], "value": n * world / (tsh["ms_per_step"] * 1e-3),
                                 "voxels": int(sent[1]), "partial_rows_sent_all_ranks": int(sent[0]),
                                 "bytes_over_nvlink_per_step": int(sent[0]) * 88,
                                 "nvlink_GBps_per_gpu_during_pack": round(int(sent[0]) * 88 / world / 1e9 / (ex_ms * 1e-3), 1) if ex_ms else None,
                                 "nvlink_peak_GBps_per_direction": 900, "check_ok": bool(chk["ok"] and chk["voxels_torch_unique"] == int(sent[1])),
                                 "what": "the same cloud with the points dealt to the ranks at random: every rank holds partial rows of (almost) every "
                                         "voxel, the exchange carries (N-1)/N of all partial rows; limiter = exchange volume + the merge of N partial rows per voxel"}
            tsh = tab = None
            del xs, ys, zs
            torch.cuda.empty_cache()
            # the slab layout of round 1 (configs[1] boxes overlapping by 10 %), for continuity with SCALE_r01
            n_slab = 100_000_000 if n == 125_000_000 else n
            xb, yb, zb = synth.uniform_cloud(n_slab, seed=2 + rank, device=device, extent=EXTENT, x_shift=D.bench_slab_shift(rank, EXTENT[0]))

            def step_slab():
                return D.voxel_statistics(xb, yb, zb, VOXEL_SIZE, ORIGIN)
            for _ in range(2):
                step_slab()
            tsl = time_steps(step_slab, args.steps, barrier, world, device, dist, L=L)
            windows.append(tsl["window"])
            extra["slab_configs1"] = {"ms_per_step": tsl["ms_per_step"], "value": n_slab * world / (tsl["ms_per_step"] * 1e-3), "points_per_gpu": n_slab,
                                      "what": "round 1's N > 1 layout: rank r holds the configs[1] box shifted to x in [0.9 r L, 0.9 r L + L), voxel 0.1 m"}
            tsl = None
            del xb, yb, zb
            torch.cuda.empty_cache()
        x, y, z = synth.clustered_chunk(n, rank, world, device, seed=3, scan_order=True)
    else:
        del cloud, st

    # ---- e2e: host buffers in, host voxel table out -------------------------------------------
    e2e = None
    if not args.no_e2e:
        from vapc_b200.pipeline import HostStreamPipeline, STAT_NAMES

        platform = platform_copy_bandwidth(device, world, dist, barrier)
        # the pipeline's fill (first copy in) and drain (last table out) are inside the timed wall clock: ~2 copies of ~45 ms on top of K steps
        # of ~50 ms; 20 steps keep their share below 10 % at N = 1 (N > 1: 0.7 s per step, the requested K)
        e_steps = max(5, args.steps) if multi else max(20, args.steps)
        if not multi:
            hx, hy, hz = (c.cpu().pin_memory() for c in (x, y, z))
            del x, y, z
            torch.cuda.empty_cache()
            pipe = HostStreamPipeline(vs, ORIGIN, device, bounds=header_bounds if args.bounds == "header" else None)
            cols_in = (hx, hy, hz)
        else:
            # N > 1: the chunk as the int32 X, Y, Z of its LAS file (12 B/point over PCIe); the table comes back without the
            # Eigenvalue_n columns (a left-over of the reference, vapc.py:1132)
            lasq = [torch.round(c / QUANT).to(torch.int32).cpu().pin_memory() for c in (x, y, z)]
            del x, y, z
            torch.cuda.empty_cache()
            e2e_cols = tuple(k for k in STAT_NAMES if k != "eig_n")

            def compute_fn(dx, dy, dz):
                fx, fy, fz = E.las_coordinates(dx, dy, dz, [QUANT] * 3, [0.0] * 3, device)
                c, s_ = D.voxel_statistics(fx, fy, fz, vs, ORIGIN, stats_columns=e2e_cols, all_bounds=chunk_bounds)
                out = {"voxel_keys": c.voxel_keys, "count": c.count}
                out.update({k: getattr(s_, k) for k in e2e_cols})
                return out

            pipe = HostStreamPipeline(vs, ORIGIN, device, compute_fn=compute_fn)
            cols_in = tuple(lasq)
        for _res in pipe.run([cols_in] * 2):  # warm-up: allocates device / pinned buffers
            pass
        torch.cuda.synchronize()
        pipe.h2d_bytes = pipe.d2h_bytes = 0
        barrier()
        t0 = time.perf_counter()
        nres = 0
        for _res in pipe.run([cols_in] * e_steps):
            nres += 1
        torch.cuda.synchronize()
        barrier()
        windows.append([t0, time.perf_counter()])
        dt = (windows[-1][1] - t0) / e_steps
        assert nres == e_steps
        h2d, d2h = pipe.h2d_bytes // e_steps, pipe.d2h_bytes // e_steps
        if multi:
            t = torch.tensor([dt, float(h2d), float(d2h)], dtype=torch.float64, device=device)
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dt, h2d, d2h = float(tmax[0].item()), int(t[1].item()), int(t[2].item())
        e2e = {"value": n * world / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3,
               "steps": e_steps, "platform_gbs": round(platform, 1), "pcie_gbs": round((h2d + d2h) / dt / 1e9, 1),
               "frac_of_platform": round((h2d + d2h) / dt / 1e9 / platform, 3) if platform else None,
               "platform_what": "aggregate GB/s of all ranks each copying 1 GiB host -> device and 1 GiB device -> host at once (plain pinned "
                                "cudaMemcpyAsync, two streams per GPU, wall clock between barriers): what this box's PCIe / host memory gives",
               "api": ("vapc_b200.pipeline.HostStreamPipeline over VoxelCloud.build + .stats: every step copies its fp64 X, Y, Z from pinned host "
                       "memory and its complete voxel table (224 B/voxel) back to pinned host memory" if not multi else
                       "vapc_b200.pipeline.HostStreamPipeline over dist.voxel_statistics: every step copies the rank's chunk as int32 LAS X, Y, Z "
                       "(12 B/point) from pinned host memory and its key range of the voxel table (count, density, centroid, std, covariance, "
                       "eigenvalues, 8 features: 200 B/voxel) back to pinned host memory") +
                      "; H2D of step k+1, compute of step k and D2H of step k-1 overlap on three streams; wall clock over all steps incl. "
                      "pipeline fill and drain, max over ranks"}

        # the same cloud as the int32 coordinates of a LAS file (scale 1e-4, offset 0): 12 B/point over PCIe; reported next to e2e
        if not multi and not args.no_las_e2e:
            lasq = [torch.round(c / QUANT).to(torch.int32) for c in (hx, hy, hz)]
            assert all(bool((q.to(torch.float64) * QUANT + 0.0 == c).all()) for q, c in zip(lasq, (hx, hy, hz))), "not the same cloud"
            lasq = [q.pin_memory() for q in lasq]
            del hx, hy, hz
            lpipe = HostStreamPipeline(vs, ORIGIN, device, las=([QUANT] * 3, [0.0] * 3), bounds=header_bounds if args.bounds == "header" else None)
            for _res in lpipe.run([tuple(lasq)] * 2):
                pass
            torch.cuda.synchronize()
            lpipe.h2d_bytes = lpipe.d2h_bytes = 0
            t0 = time.perf_counter()
            for _res in lpipe.run([tuple(lasq)] * e_steps):
                pass
            torch.cuda.synchronize()
            windows.append([t0, time.perf_counter()])
            ldt = (windows[-1][1] - t0) / e_steps
            lh, ld = lpipe.h2d_bytes // e_steps, lpipe.d2h_bytes // e_steps
            e2e["las_int32"] = {"value": n / ldt, "unit": UNIT, "ms_per_step": ldt * 1e3, "h2d_bytes_per_step": lh, "d2h_bytes_per_step": ld,
                                "frac_of_platform": round((lh + ld) / ldt / 1e9 / platform, 3) if platform else None,
                                "what": "the same cloud entering as the int32 X, Y, Z of a LAS file (scale 1e-4, offset 0; identical doubles, "
                                        "identical voxel table): vapc_pack_records_bucketed_i32 forms X * scale + offset in the key kernels"}
            del lasq
        # the drop-in facade: DataFrame in -> reduced DataFrame out
        if not multi and not args.no_facade:
            try:
                from tools.facade_bench import facade_e2e

                e2e["facade"] = facade_e2e(sizes=(10_000_000, 100_000_000) if n >= 100_000_000 else (n,))
            except Exception as exc:  # noqa: BLE001
                e2e["facade"] = {"error": repr(exc)}

    clocks = sampler.stop(windows)

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        cpu_baseline = cpu_baseline_leg(args)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(workload, voxels_per_gpu=nvox, key_bits=total_bits, radix_passes=passes,
                                                 **({"grouping": sorted_by} if sorted_by else {}),
                                                 **({"exchange": exchange_kind, "voxels_all_gpus": nvox_global,
                                                     "key_axis_order": list(cloud_axis_order)} if multi else {})),
            **({"n1_same_workload": {"value_per_gpu": extra["one_gpu_same_workload"]["value_per_gpu"],
                                     "efficiency_inputs": "this line's value / (n_gpus x value_per_gpu): every rank's chunk voxelised alone in the same "
                                                          "run; the N = 1 bench line measures configs[1] (a different cloud) and carries the same figure as "
                                                          "extra.configs2_one_gpu"}} if multi else {}),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "check": check, "extra": extra, "kernels": per_kernel, "abi_calls": breakdown,
            "step_ms": step_ms, "ms_per_step_instrumented": ms_instrumented, "gpu_ms_in_abi_calls_per_step": round(gpu_ms, 3), "gaps_between_calls_ms": gaps,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if multi:
        ok = check is None or check["ok"]
        dist.destroy_process_group()
        if not ok:
            raise SystemExit("bench.py: the sharded voxel table failed its check: " + json.dumps(check))


if __name__ == "__main__":
    main()
