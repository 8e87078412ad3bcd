#!/usr/bin/env python
"""
profiles/r02_traffic.json from an `ncu --set full` capture of ONE launch of the rows kernel (tools/r2_call_ncu.sh):
    ncu -i gpurun_out/r02_rows_full74.ncu-rep --page raw --csv > raw.csv ; python tools/make_traffic_json.py raw.csv
The file records the sha256 of the kernel source it was captured from; bench.py reports `roofline.traffic` only while
that still matches the source in the tree.
"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "miosqp_b200/csrc/bqp_rows.cu"
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def num(key, scale={"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "%": 1.0, "": 1.0}):
    v, u = d[key]
    return float(v.replace(",", "")) * scale.get(u, 1.0)


peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
rd, wr, ms = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), num("gpu__time_duration.sum")
out = {
    "kernel": "admm_rows_kernel (8 nodes per tile)", "kernel_source": SRC,
    "kernel_source_sha256": hashlib.sha256(open(os.path.join(ROOT, SRC), "rb").read()).hexdigest(),
    "grid": int(num("launch__grid_size")), "block": int(num("launch__block_size")), "cluster": int(num("launch__cluster_size")),
    "registers_per_thread_at_launch": int(num("launch__registers_per_thread")),
    "scope": "ONE full-occupancy launch (74 tiles x 8 leaves on cluster pairs, one round of 100 ADMM iterations incl. 4 termination "
             "checks, prologue A pass and objective P pass) of the 32 launches of a bench step",
    "dram_bytes_read": rd, "dram_bytes_write": wr, "gpu_time_ms_under_ncu": ms,
    "dram_gbs": (rd + wr) / ms / 1e6,
    "dram_frac_of_measured_peak": (rd + wr) / ms / 1e6 / peaks["hbm_gbs"] if peaks else None,
    "dmma_pipe_active_pct": num("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "lts_hit_rate_pct": num("lts__t_sector_hit_rate.pct"),
    "shared_wavefronts": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "source": "ncu --set full --clock-control none --import-source on -k regex:admm_rows -s 52 -c 1 python bench.py --steps 1 --warmup 1 "
              "--no-cpu-baseline --mode frontier (tools/r2_call_ncu.sh; summary in profiles/r02_summary.md)",
}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
