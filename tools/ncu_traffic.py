"""DRAM traffic per launch of every kernel in an `ncu --set full` capture -> profiles/r2_ncu_traffic.json (read by bench.py for
`roofline.traffic`).  Runs in the build container on the .ncu-rep brought back from the GPU box.

    python tools/ncu_traffic.py <key> <rep.ncu-rep> [<key> <rep> ...]      key: "n1" (the bench line) or "multi" (one GPU's share of configs[2])
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("VAPC_NCU_TRAFFIC_OUT") or os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")


def traffic(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kn, rd, wr, tm = (hdr.index(c) for c in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
    scale = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
    acc = {}
    for r in rows[2:]:
        name = r[kn].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        name = re.sub(r"^\(?void\s+", "", name)          # "void name<...>(args)" / "(void name<...>)(args)"
        name = re.split(r"[<(]", name.strip())[0].strip()
        gb = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
        e = acc.setdefault(name, [0, 0.0, 0.0])
        e[0] += 1
        e[1] += gb
        e[2] += float(r[tm])
    return {k: round(v[1] / v[0], 4) for k, v in acc.items()}, {k: [v[0], round(v[2] / v[0], 4), units[tm]] for k, v in acc.items()}


def traffic_from_summary(path):
    """The same from a tools/ncu_summary.py --csv table (when only the summary travelled back from the GPU box)."""
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    col = {h.split(" [")[0]: (i, h.split("[")[1].rstrip("]") if "[" in h else "") for i, h in enumerate(hdr)}
    scale = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
    acc = {}
    for r in rows[1:]:
        name = r[0].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        name = re.sub(r"^\(?void\s+", "", name)
        name = re.split(r"[<(]", name.strip())[0].strip()
        gb = sum(float(r[col[c][0]]) * scale[col[c][1]] for c in ("dram_rd", "dram_wr"))
        e = acc.setdefault(name, [0, 0.0, 0.0])
        e[0] += 1
        e[1] += gb
        e[2] += float(r[col["time"][0]])
    return {k: round(v[1] / v[0], 4) for k, v in acc.items()}, {k: [v[0], round(v[2] / v[0], 4), col["time"][1]] for k, v in acc.items()}


def main():
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    args = sys.argv[1:]
    for key, rep in zip(args[0::2], args[1::2]):
        gb, times = traffic_from_summary(rep) if rep.endswith(".csv") else traffic(rep)
        data[key] = gb
        data.setdefault("_source", {})[key] = {"rep": os.path.basename(rep), "launches_and_time_under_ncu": times,
                                               "what": "GB per launch: dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches of the kernel"}
    json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps({k: v for k, v in data.items() if not k.startswith("_")}, indent=1))


if __name__ == "__main__":
    main()
