#!/usr/bin/env python
"""Reads the `ncu --set full --page raw --csv` export of scripts/her_prof.py (scripts/prof_r2.sh) and writes profiles/her_kernel_traffic.json:
per HER kernel launch, in the order her_prof.py issues them, the DRAM bytes ncu counted and the kernel time under ncu.  bench.py attaches
them to its `her_compute_reward` entries (`dram_bytes_ncu`, the fraction of the measured HBM peak in DRAM-traffic terms)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_her_kernels_ncu_full_raw.csv")
rows = list(csv.reader(open(src))); hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def val(r, name, table):
    return float(r[col[name]]) * table[units[col[name]]]
# her_prof.py: 2 x reward 32M x 3-D, 2 x reward 1M x 6-D, 2 x relabel random, 2 x relabel episode-local unsorted, 2 x relabel episode-local sorted
names = ["reach_32M_rows_3d", "stack_1M_rows_6d", "her_relabel_random_indices", "her_relabel_episode_local_goals_unsorted_batch", "her_relabel_episode_local_goals_index_sorted_batch"]
out = {}
for k, nm in enumerate(names):
    r = rows[2 + 2 * k + 1]         # the second (warm) launch of each pair
    out[nm] = {"kernel": r[col["Kernel Name"]], "kernel_us_under_ncu": val(r, "gpu__time_duration.sum", TIME),
               "dram_bytes_read": val(r, "dram__bytes_read.sum", UNIT), "dram_bytes_write": val(r, "dram__bytes_write.sum", UNIT),
               "lts_sector_hit_rate_pct": float(r[col["lts__t_sector_hit_rate.pct"]]), "grid": r[col["launch__grid_size"]],
               "warps_active_pct": float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]]), "source": os.path.relpath(os.path.abspath(src), ROOT)}
    out[nm]["dram_gbs_under_ncu"] = (out[nm]["dram_bytes_read"] + out[nm]["dram_bytes_write"]) / out[nm]["kernel_us_under_ncu"] / 1e3
json.dump(out, open(os.path.join(ROOT, "profiles", "her_kernel_traffic.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
