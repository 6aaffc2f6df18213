#!/usr/bin/env python
"""bench.py — points/s of the voxelise + per-voxel statistics path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--impl reference]

A step is one pass of the hot path over one synthetic cloud: bounding box -> packed voxel keys ->
stable radix sort -> run-length segmentation -> FP64 (count, mean, M2) moments -> density, centroid,
std, covariance, eigenvalues, 8 eigen-features (+ the point->voxel map).  At N = 1 the workload is
BASELINE.json configs[1]: a 100M-point uniform cloud, voxel 0.1 m, full attribute set.

value   whole-job points/s with the inputs already resident in HBM (CUDA events, max over ranks)
e2e     the same metric through the public API with HOST buffers: pinned host -> device copies of X, Y, Z
        and the device -> pinned-host read of the complete voxel table inside the timed region
roofline  the dominant C-ABI call: algorithmic bytes / CUDA-event time on the launching stream, against
        MEASURED_PEAKS.json
cpu_baseline  the CPU port of the reference's pandas path (oracle/ref_port.py, "port": the reference is pure
        Python and /root/reference does not exist on the GPU box) on a bounded spatial crop, 1 core
--impl reference  times that CPU port alone and prints the same line with "impl": "reference"
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
ORIGIN = [0.0, 0.0, 0.0]
EXTENT = (50.0, 50.0, 4.0)  # metres: 500 x 500 x 40 voxels = 1e7 cells -> ~10 points per voxel at 1e8 points
QUANT = 1e-4  # LAS-like coordinate quantisation
# DRAM traffic per launch of the big kernels at the default workload, from the committed `ncu --set full` capture
# (profiles/README.md says which file); reported next to the algorithmic bytes of the dominant kernel
NCU_TRAFFIC_GB = {  # profiles/r1_f_ncu_full.csv (1e8 points, 9 999 543 voxels, 24-bit keys)
    "bucket_scatter_kernel": 6.42, "reduce_moments_kernel": 5.27, "onesweep_kernel": 1.40, "voxel_stats_kernel": 2.90,
    "keys_hist_kernel": 2.75, "p2v_lookup_kernel": 0.76, "minmax_partial": 2.40, "digit_hist_kernel": 0.40, "count_heads_kernel": 0.40,
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=100_000_000, help="points per GPU")
    ap.add_argument("--no-las-e2e", action="store_true", help="skip the LAS-integer-input variant of the e2e measurement")
    ap.add_argument("--impl", default="vapc_b200", choices=["vapc_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="points in the CPU sample (0 = auto)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic data
# ------------------------------------------------------------------------------------------------
def gen_cloud_device(n, seed, device, x_shift=0.0):
    """n quantised uniform points in EXTENT (x shifted by x_shift metres), fp64 device columns."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cols = []
    for k, ext in enumerate(EXTENT):
        q = torch.randint(0, int(round(ext / QUANT)), (n,), generator=g, device=device, dtype=torch.int64)
        c = q.to(torch.float64) * QUANT
        if k == 0 and x_shift:
            c += x_shift
        cols.append(c)
        del q
    return cols


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
def cpu_port_once(x, y, z):
    from oracle import ref_port

    t0 = time.perf_counter()
    df = ref_port.full_attribute_set(x, y, z, VOXEL_SIZE, ORIGIN, with_eigen=True)
    dt = time.perf_counter() - t0
    nvox = int(df[["voxel_x", "voxel_y", "voxel_z"]].drop_duplicates().shape[0])
    return dt, nvox


def _cpu_worker(job):
    n_sample, seed = job
    x, y, z, _ = gen_crop_host(n_sample, seed=seed)
    dt, nvox = cpu_port_once(x, y, z)
    return dt, nvox


def run_reference_arm(args):
    """The reference's CPU path (oracle/ref_port.py keeps its pandas data flow) on the host cores.  The reference itself is
    single-threaded; its own way to use more cores is to cut the cloud into spatial tiles and process them independently
    (main.py:97-137), so every host core gets its own crop of the workload here and the throughput is the sum."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor

    cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    budget = 150.0 / max(1, args.steps + args.warmup)  # seconds per step
    n_sample = args.cpu_sample or int(min(200_000, max(2_000, budget * 5_000)))
    _, _, _, what = gen_crop_host(n_sample)
    what = what.replace(f"{n_sample} points", f"{n_sample} points per core")
    nvox = 0
    with ProcessPoolExecutor(max_workers=cores) as pool:
        for _ in range(args.warmup):
            list(pool.map(_cpu_worker, [(n_sample, 100 + c) for c in range(cores)]))
        t0 = time.perf_counter()
        for k in range(args.steps):
            res = list(pool.map(_cpu_worker, [(n_sample, 1000 * (k + 1) + c) for c in range(cores)]))
            nvox = sum(r[1] for r in res)
        total = time.perf_counter() - t0
    value = n_sample * cores * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.points, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": what + f" ({nvox} voxels per step over {cores} processes, one independent crop each: tiling is the reference's "
                                          "own scaling mechanism); the reference is single-threaded pandas/numpy; per-voxel Python eigen step included"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(points_per_gpu, gpus):
    return {
        "workload": f"configs[1]: {points_per_gpu / 1e6:g}M-point synthetic uniform cloud per GPU, voxel_size={VOXEL_SIZE} m, "
                    "full attribute set (count, density, centroid, std, covariance, eigenvalues, 8 eigen-features) + point->voxel map",
        "points_per_gpu": points_per_gpu, "voxel_size": VOXEL_SIZE, "extent_m": list(EXTENT),
        "l2": "inputs (24 B/point, 2.4 GB at 1e8 points) are larger than the 126 MB L2; no flush needed",
        "parallelism": f"points sharded in contiguous chunks over {gpus} GPU(s)" + ("" if gpus == 1 else
                       "; rank r holds the x-slab [0.9 r L, 0.9 r L + L) (10 % overlap with its neighbour); partial voxel tables "
                       "exchanged by key range with one NCCL all-to-all, then merged"),
    }


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
        "vapc_pack_records_bucketed": (24 + k) * n + (k + 24 + k + 32) * n,
        "vapc_sort_pairs": (k + passes * (2 * k + 8) - 4) * n,
        "vapc_sort_pairs_segmented": (k + low_passes * (2 * k + 8) - 4) * n,
        "vapc_count_voxels": k * n,
        "vapc_reduce_moments": (k + 4 + 32 + 4) * n + (k + 12 + 72) * v,
        "vapc_point2voxel_dense": (k + 4) * n + (k + 4) * v,
        "vapc_voxel_stats": (k + 4 + 72) * v + (8 + 24 + 24 + 48 + 24 + 24 + 64) * v,
    }.get(name, 0)


def kernel_algorithmic_bytes(kernel, n, v, key_bytes, launches_per_step):
    """Algorithmic bytes of ONE launch of a kernel (name as recorded by the library's launch macro)."""
    k = key_bytes
    table = [
        ("minmax_partial", 24 * n),
        ("keys_hist_kernel", (24 + k) * n),                      # x, y, z in; keys in input order out
        ("bucket_scatter_kernel", (k + 24 + k + 32) * n),        # keys + x, y, z in; bucketed keys + records out
        ("pack_keys_kernel", (24 + k + 32) * n),
        ("digit_hist_kernel", k * n),
        # one radix pass: keys + payload in and out; the first pass of a sort takes the payload from the position
        ("onesweep_kernel", (2 * k + 8) * n - (4 * n) / max(1.0, launches_per_step)),
        ("count_heads_kernel", k * n),
        ("reduce_moments_kernel", (k + 4 + 32 + 4) * n + (k + 12 + 72) * v),
        ("p2v_lookup_kernel", (k + 4) * n),
        ("p2v_fill_kernel", (k + 4) * v),
        ("voxel_stats_kernel", (k + 4 + 72) * v + (8 + 24 + 24 + 48 + 24 + 24 + 64) * v),
    ]
    for key, b in table:
        if key in kernel:
            return key, float(b)
    return kernel, 0.0


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    # stdout carries exactly ONE JSON line: anything a library prints there meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist

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
    n = args.points

    if world > 1:
        from vapc_b200 import dist as D

        x, y, z = gen_cloud_device(n, seed=2 + rank, device=device, x_shift=D.bench_slab_shift(rank, EXTENT[0]))
    else:
        x, y, z = gen_cloud_device(n, seed=2, device=device)

    def step_device():
        if world > 1:
            return D.voxel_statistics(x, y, z, VOXEL_SIZE, ORIGIN)
        cloud = E.VoxelCloud.build(x, y, z, VOXEL_SIZE, ORIGIN, keep_columns=False)
        st = cloud.stats(density=True, cog=True, std=True, cov=True, eig=True, eig_n=True, feat=True)
        return cloud, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        cloud, st = step_device()
    torch.cuda.synchronize()
    nvox = cloud.nvoxels
    exchange_kind = getattr(cloud, "exchange", None)
    key_bytes = cloud.grid.key_bytes
    total_bits = cloud.grid.total_bits
    passes = (total_bits + 7) // 8
    low_passes = (total_bits - min(8, cloud.grid.bits[0]) + 7) // 8  # inside the key-prefix buckets

    # ---- timed region: K steps, CUDA events, per-call events for the roofline -----------------
    records = []

    def observer(name, phase):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        records.append((name, phase, ev))

    launches0 = L.raw("vapc_launch_count")()
    barrier()
    windows = [[time.perf_counter(), None]]
    L.observer = observer
    L.call("vapc_profile_enable", 1)  # CUDA events around every kernel launch, on the launching stream
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        cloud, st = step_device()
    t_end.record()
    L.observer = None
    L.call("vapc_profile_enable", 0)
    barrier()
    kernel_times = L.profile_collect()
    launches = L.raw("vapc_launch_count")() - launches0
    windows[0][1] = time.perf_counter()
    ms_total = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = n * world / (ms_per_step * 1e-3)

    per_call = {}
    pending = {}
    gaps = {}
    last_after = None
    for name, phase, ev in records:
        if phase == "before":
            pending[name] = ev
            if last_after is not None:
                gaps.setdefault(f"{last_after[0]} -> {name}", []).append(last_after[1].elapsed_time(ev))
        else:
            per_call.setdefault(name, []).append(pending.pop(name).elapsed_time(ev))
            last_after = (name, ev)
    gaps = {k: round(sum(v) / len(v), 4) for k, v in gaps.items()}
    del records

    # ---- roofline of the dominant call --------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    breakdown = {}
    for name, ts in per_call.items():
        calls_per_step = len(ts) / args.steps
        avg = sum(ts) / len(ts)
        b = algorithmic_bytes(name, n, nvox, key_bytes, passes, low_passes)
        breakdown[name] = {"ms_per_call": round(avg, 4), "calls_per_step": calls_per_step, "ms_per_step": round(avg * calls_per_step, 4),
                           "algorithmic_GB": round(b / 1e9, 4), "GBps": round(b / 1e9 / (avg * 1e-3), 1) if b else None}
    # per kernel (CUDA events around every launch inside the timed region): the dominant one gives the roofline
    per_kernel = {}
    for kname, (cnt, ms) in kernel_times.items():
        short, b = kernel_algorithmic_bytes(kname, n, nvox, key_bytes, cnt / args.steps)
        e = per_kernel.setdefault(short, {"launches_per_step": 0.0, "ms_per_step": 0.0, "algorithmic_GB_per_launch": round(b / 1e9, 4)})
        e["launches_per_step"] += cnt / args.steps
        e["ms_per_step"] += ms / args.steps
    for e in per_kernel.values():
        e["ms_per_launch"] = round(e["ms_per_step"] / e["launches_per_step"], 4)
        e["ms_per_step"] = round(e["ms_per_step"], 4)
        e["GBps"] = round(e["algorithmic_GB_per_launch"] / (e["ms_per_launch"] * 1e-3), 1) if e["algorithmic_GB_per_launch"] and e["ms_per_launch"] else None
    dom = max((k for k in per_kernel if per_kernel[k]["GBps"]), key=lambda k: per_kernel[k]["ms_per_step"])
    achieved = per_kernel[dom]["GBps"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": NCU_TRAFFIC_GB.get(dom) if (n == 100_000_000 and world == 1) else None, "traffic_unit": "GB per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": int(per_kernel[dom]["algorithmic_GB_per_launch"] * 1e9),
                "ms_per_launch": per_kernel[dom]["ms_per_launch"], "launches_per_step": per_kernel[dom]["launches_per_step"],
                "share_of_step": round(per_kernel[dom]["ms_per_step"] / ms_per_step, 3)}
    gpu_ms = sum(v["ms_per_step"] for v in breakdown.values())

    # ---- e2e: host buffers in, host voxel table out -------------------------------------------
    e2e = None
    if not args.no_e2e:
        hx, hy, hz = (c.cpu().pin_memory() for c in (x, y, z))
        del x, y, z, cloud, st
        torch.cuda.empty_cache()
        from vapc_b200.pipeline import HostStreamPipeline, STAT_NAMES

        def compute_fn(dx, dy, dz):
            if world > 1:
                c, s_ = D.voxel_statistics(dx, dy, dz, VOXEL_SIZE, ORIGIN)
            else:
                c = E.VoxelCloud.build(dx, dy, dz, VOXEL_SIZE, ORIGIN, keep_columns=False)
                s_ = c.stats(density=True, cog=True, std=True, cov=True, eig=True, eig_n=True, feat=True)
            out = {"voxel_keys": c.voxel_keys, "count": c.count}
            out.update({k: getattr(s_, k) for k in STAT_NAMES})
            return out

        pipe = HostStreamPipeline(VOXEL_SIZE, ORIGIN, device, compute_fn=compute_fn)
        for _res in pipe.run([(hx, hy, hz)] * 2):  # warm-up: allocates device / pinned buffers
            pass
        torch.cuda.synchronize()
        e_steps = max(5, args.steps)
        pipe.h2d_bytes = pipe.d2h_bytes = 0
        barrier()
        t0 = time.perf_counter()
        nres = 0
        for _res in pipe.run([(hx, hy, hz)] * e_steps):
            nres += 1
        torch.cuda.synchronize()
        barrier()
        windows.append([t0, time.perf_counter()])
        dt = (windows[-1][1] - t0) / e_steps
        assert nres == e_steps
        h2d, d2h = pipe.h2d_bytes // e_steps, pipe.d2h_bytes // e_steps
        if world > 1:
            t = torch.tensor([dt, float(h2d), float(d2h)], dtype=torch.float64, device=device)
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dt, h2d, d2h = float(tmax[0].item()), int(t[1].item()), int(t[2].item())
        e2e = {"value": n * world / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3,
               "steps": e_steps,
               "api": "vapc_b200.pipeline.HostStreamPipeline over VoxelCloud.build + .stats (N = 1) / dist.voxel_statistics (N > 1): every step "
                      "copies its X, Y, Z from pinned host memory and its complete voxel table back to pinned host memory; H2D of step k+1, "
                      "compute of step k and D2H of step k-1 overlap on three streams; wall clock over all steps incl. pipeline fill and "
                      "drain, max over ranks"}

        # the same cloud as the int32 coordinates of a LAS file (scale 1e-4, offset 0): 12 B/point over PCIe; reported next to e2e
        if world == 1 and not args.no_las_e2e:
            lasq = [torch.round(c / QUANT).to(torch.int32) for c in (hx, hy, hz)]
            assert all(bool((q.to(torch.float64) * QUANT + 0.0 == c).all()) for q, c in zip(lasq, (hx, hy, hz))), "not the same cloud"
            lasq = [q.pin_memory() for q in lasq]
            del hx, hy, hz
            lpipe = HostStreamPipeline(VOXEL_SIZE, ORIGIN, device, las=([QUANT] * 3, [0.0] * 3))
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
            e2e["las_int32"] = {"value": n / ldt, "unit": UNIT, "ms_per_step": ldt * 1e3, "h2d_bytes_per_step": lpipe.h2d_bytes // e_steps,
                                "d2h_bytes_per_step": lpipe.d2h_bytes // e_steps,
                                "what": "the same cloud entering as the int32 X, Y, Z of a LAS file (scale 1e-4, offset 0; identical doubles, "
                                        "identical voxel table): vapc_pack_records_bucketed_i32 forms X * scale + offset in the key kernels"}

    clocks = sampler.stop(windows)

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        n_sample = args.cpu_sample or 100_000
        cx, cy, cz, what = gen_crop_host(n_sample)
        dt, cv = cpu_port_once(cx, cy, cz)
        cpu_baseline = {"value": n_sample / dt, "unit": UNIT, "cores": 1, "kind": "port", "seconds": round(dt, 2),
                        "sample": what + f" ({cv} voxels); oracle/ref_port.py keeps the reference's single-threaded pandas data flow "
                                         "incl. its per-voxel Python eigen step; host has %d logical cores" % (os.cpu_count() or 0)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(workload_config(n, world), voxels_per_gpu=nvox, key_bits=total_bits, radix_passes=passes,
                                                 **({"exchange": exchange_kind} if world > 1 else {})),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "kernels": per_kernel, "abi_calls": breakdown, "gpu_ms_in_abi_calls_per_step": round(gpu_ms, 3),
            "gaps_between_calls_ms": gaps,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
