#!/usr/bin/env python
"""Reads `ncu --set full --csv --page raw` exports of the step kernel and writes profiles/kernel_traffic.json, the file bench.py takes
`roofline.traffic` and `launches_per_step` from (no literals in bench.py).

    python scripts/ncu_traffic.py <task>/<control>/<envs>=<raw.csv>:<step_kernel launches per step>:<description> ...

Per capture: dram__bytes_read.sum + dram__bytes_write.sum of the (first) step_kernel row, per launch."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def read_raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            d[h] = (v, u)
        out.append(d)
    return out


def main():
    dst = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    table = json.load(open(dst)) if os.path.exists(dst) else {}
    for arg in sys.argv[1:]:
        key, rest = arg.split("=", 1)
        path, launches, desc = rest.split(":", 2)
        rows = [r for r in read_raw(path) if "step_kernel" in r["Kernel Name"][0]]
        if not rows:
            raise SystemExit(f"no step_kernel row in {path}")
        r = rows[0]
        rd = float(r["dram__bytes_read.sum"][0]) * UNIT[r["dram__bytes_read.sum"][1]]
        wr = float(r["dram__bytes_write.sum"][0]) * UNIT[r["dram__bytes_write.sum"][1]]
        t = float(r["gpu__time_duration.sum"][0]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[r["gpu__time_duration.sum"][1]]
        def pct(name):
            return float(r[name][0]) if name in r else None
        table[key] = {"fma_pipe_pct": pct("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), "issue_active_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      "threads_per_instruction": pct("smsp__thread_inst_executed_per_inst_executed.ratio"), "warps_active_pct": pct("sm__warps_active.avg.pct_of_peak_sustained_active"),
                      "dram_bytes_per_launch": rd + wr, "step_kernel_launches_per_step": int(launches), "launches_per_step": desc, "kernel_us_under_ncu": t,
                      "grid": r["launch__grid_size"][0], "kernel": r["Kernel Name"][0], "source": os.path.relpath(os.path.abspath(path), ROOT)}
    json.dump(table, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(table, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
