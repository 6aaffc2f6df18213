"""Opcode histogram of the SASS in libvapc_b200.so (cuobjdump -sass; runs in the build container, no GPU needed).

    python tools/sass_summary.py [vapc_b200/libvapc_b200.so] > profiles/r2_sass_summary.txt

Per kernel: instruction count and the memory / synchronisation opcodes that matter for an HBM-bound integer / fp64 code:
wide global loads and stores (LDG / STG .128 / .256: the 256-bit path of sm_100), shared-memory atomics, cluster barriers and
distributed-shared-memory accesses, bulk (TMA) copies, fp64 arithmetic.  No tensor-core opcode (UTC*MMA, HMMA) is expected:
nothing on this path is a contraction.
"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "vapc_b200/libvapc_b200.so"
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kernels = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"^void ", "", name)
        cur = kernels.setdefault(name.split("(")[0], collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1

GROUPS = [
    ("LDG.256", lambda op: op.startswith("LDG") and ".256" in op), ("LDG.128", lambda op: op.startswith("LDG") and ".128" in op),
    ("LDG other", lambda op: op.startswith("LDG") and ".128" not in op and ".256" not in op),
    ("STG.256", lambda op: op.startswith("STG") and ".256" in op), ("STG.128", lambda op: op.startswith("STG") and ".128" in op),
    ("STG other", lambda op: op.startswith("STG") and ".128" not in op and ".256" not in op),
    ("LDS", lambda op: op.startswith("LDS")), ("STS", lambda op: op.startswith("STS")), ("ATOMS", lambda op: op.startswith("ATOMS")),
    ("ATOMG/RED", lambda op: op.startswith("ATOMG") or op.startswith("RED")),
    ("DSMEM (LD/ST .CLUSTER? via generic)", lambda op: "MAPA" in op or op.startswith("LDC.CLUSTER")),
    ("cluster barrier", lambda op: op.startswith("UCGABAR") or "CGABAR" in op), ("BAR", lambda op: op.startswith("BAR")),
    ("bulk copy (TMA)", lambda op: op.startswith("UBLKCP") or op.startswith("UTMA")),
    ("tensor core", lambda op: "MMA" in op), ("FP64 (DADD/DMUL/DFMA/DSETP/MUFU.RCP64H...)", lambda op: op.startswith("D") and op[:4] in ("DADD", "DMUL", "DFMA", "DSET", "DMNM")),
    ("SHFL/VOTE/MATCH/REDUX", lambda op: op.split(".")[0] in ("SHFL", "VOTE", "MATCH", "REDUX")),
]
print(f"# SASS of {LIB}: architectures {arch}, {len(kernels)} functions, {sum(sum(c.values()) for c in kernels.values())} instructions")
total = collections.Counter()
for c in kernels.values():
    total.update(c)
print("\n## whole library: opcode groups")
for label, pred in GROUPS:
    n = sum(v for op, v in total.items() if pred(op))
    print(f"{label:48s} {n}")
print("\n## whole library: the 40 most frequent opcodes")
for op, v in total.most_common(40):
    print(f"{op:32s} {v}")
print("\n## per kernel (functions of the benchmarked path and the cell sort)")
want = ("tile_digits_kernel", "tile_scatter_kernel", "digit_hist_kernel", "onesweep_kernel", "count_heads_kernel", "reduce_moments_kernel",
        "p2v_lookup_kernel", "p2v_fill_kernel", "voxel_stats_kernel", "minmax_partial", "cell_sort_kernel", "pack_partial_rows_kernel", "merge_runs_kernel")
for name, c in kernels.items():
    if not any(w in name for w in want):
        continue
    parts = [f"{label.split(' ')[0]}={sum(v for op, v in c.items() if pred(op))}" for label, pred in GROUPS if sum(v for op, v in c.items() if pred(op))]
    print(f"{name[:110]:110s} {sum(c.values()):6d} instr  " + " ".join(parts))
