#!/usr/bin/env python
"""Condense an .ncu-rep (read here, no GPU needed) into a small per-launch CSV for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_kernels.csv [100_q1_32 "spmm_brb_kernel<4, 1"]
"""
import csv
import io
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_insts"),
]


def traffic(rows, idx, units, key, pattern, source):
    """profiles/ncu_traffic.json[key] = mean DRAM read + write bytes per launch of the kernels matching `pattern` (bench.py's
    roofline.traffic reads it), with the commit the capture was taken at"""
    import json
    import os
    import re

    def to_bytes(v, u):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return float(v.replace(",", "")) * scale

    tot, cnt = 0.0, 0
    for r in rows[2:]:
        if re.search(pattern, r[idx["Kernel Name"]]):
            tot += to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                   to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            cnt += 1
    if cnt == 0:
        raise SystemExit("no launch matches " + pattern)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "profiles", "ncu_traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    commit = subprocess.run(["git", "-C", root, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    d[key] = {"dram_bytes_per_launch": tot / cnt, "launches": cnt, "kernel": pattern, "source": source, "commit": commit}
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, d[key])


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + ["%s[%s]" % (short, units[idx[name]]) for name, short in KEEP if name in idx])
        for r in rows[2:]:
            w.writerow([r[idx["Kernel Name"]]] + [r[idx[name]] for name, _ in KEEP if name in idx])
    print("wrote", out)
    if len(sys.argv) >= 5:  # ... <traffic key, e.g. 100_q1_32> <kernel regex>
        traffic(rows, idx, units, sys.argv[3], sys.argv[4], out)


if __name__ == "__main__":
    main()
