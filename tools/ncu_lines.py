"""Per-source-line stall samples of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py rep.ncu-rep <kernel regex> [launch index] [top]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{rx}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # split into launches at "Function Name"-preceded "File Path" groups; a new launch restarts with a "Kernel Name"-less layout,
    # so detect launches by the first 'File Path' after a non-File block
    launches, cur, cur_file, hdr = [], None, None, None
    prev_fn = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Function Name":
            if cur is None or (prev_fn is not None and r[1] != prev_fn and False):
                pass
            prev_fn = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            if cur is None:
                cur = {}
                launches.append(cur)
            continue
        if r[0] == "Kernel Name":
            cur = None
            continue
        if hdr and r[0].isdigit():
            key = (cur_file.split("/")[-1], int(r[0]), r[1].strip())
            try:
                samples = int(r[4] or 0)
            except ValueError:
                samples = 0
            d = cur.setdefault(key, [0, {}])
            d[0] += samples
            for name, v in zip(hdr[5:], r[5:]):
                if name.startswith("stall_") and "Not Issued" not in name and v not in ("", "0", "-"):
                    try:
                        d[1][name] = d[1].get(name, 0) + int(v)
                    except ValueError:
                        pass
    if not launches:
        print("no data")
        return
    L = launches[min(which, len(launches) - 1)]
    total = sum(v[0] for v in L.values()) or 1
    print(f"{len(launches)} launch block(s); showing #{which}; total samples {total}")
    for (f, ln, src), (s, st) in sorted(L.items(), key=lambda kv: -kv[1][0])[:top]:
        tops = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{100 * s / total:5.1f}%  {f}:{ln:<4d} {src[:90]:90s} {tops}")


if __name__ == "__main__":
    main()
