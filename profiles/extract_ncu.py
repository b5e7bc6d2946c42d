#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the small text summary committed under profiles/.

    python profiles/extract_ncu.py gpurun_out/prof_k1_v3.ncu-rep profiles/k1_r1_v3_ncu.txt [profiles/k1_traffic.json]
"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum",
]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * mult.get(unit, 1)


def main(rep, out_txt, out_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {rep}", f"# kernels: {[r[ci['Kernel Name']][:80] for r in data]}"]
    for w in WANT:
        if w in ci:
            lines.append(f"{w:72s} {units[ci[w]]:16s} " + "  ".join(r[ci[w]] for r in data))
    open(out_txt, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if out_json:
        tr = [to_bytes(r[ci["dram__bytes_read.sum"]], units[ci["dram__bytes_read.sum"]]) +
              to_bytes(r[ci["dram__bytes_write.sum"]], units[ci["dram__bytes_write.sum"]]) for r in data]
        json.dump({"source": rep, "kernel": data[0][ci["Kernel Name"]][:80], "dram_bytes_per_launch": sum(tr) / len(tr),
                   "launches": len(tr)}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
