"""Top stalled SASS instructions of a captured kernel: python profiles/ncu_hot.py <rep.ncu-rep> [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
data = []
for r in rows[hi + 1:]:          # first kernel of the report only (a second capture repeats the header)
    if r and r[0] in ("Address", "Kernel Name"):
        break
    if len(r) == len(h):
        data.append(r)
ci = {k: i for i, k in enumerate(h)}
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
ex = sum(int(r[ci["Instructions Executed"]] or 0) for r in data)
keys = ["stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_not_selected", "stall_selected", "stall_math", "stall_branch_resolving", "stall_dispatch", "stall_no_inst", "stall_mio"]
print(f"# total samples {tot}, executed warp instructions {ex}, {len(data)} SASS instructions")
print("# shares: " + ", ".join(f"{k[6:]} {100*sum(int(r[ci[k]] or 0) for r in data)/tot:.1f}%" for k in keys))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ci["# Samples"]] or 0))
for i in order[:n]:
    r = data[i]
    s = int(r[ci["# Samples"]] or 0)
    print(f"{100*s/tot:6.2f}%  idx={i:5d} exec={int(r[ci['Instructions Executed']] or 0):9d} long_sb={r[ci['stall_long_sb']]:>6} barrier={r[ci['stall_barrier']]:>5} wait={r[ci['stall_wait']]:>5} short={r[ci['stall_short_sb']]:>5}  {r[ci['Source']][:90]}")
if len(sys.argv) > 3:   # dump all rows: idx, samples, executed, source
    with open(sys.argv[3], "w") as f:
        for i, r in enumerate(data):
            f.write(f"{i:5d} {int(r[ci['# Samples']] or 0):6d} {int(r[ci['Instructions Executed']] or 0):9d} {r[ci['stall_long_sb']]:>6} {r[ci['Source']]}\n")
