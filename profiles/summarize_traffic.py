# -*- coding: utf-8 -*-
"""Per-step DRAM bytes / instructions / launches of a whole strategy step from
gpurun_out/<tag>_traffic_all.csv(.gz) (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum,smsp__inst_executed.sum` over every launch of the small bench command, see
tools/gpu_round2_profile.sh lists) -> profiles/<tag>_traffic.json, which bench.py reads for
roofline.traffic (DRAM bytes per launch of the dominant kernel family).

usage: python profiles/summarize_traffic.py r02 [steps_in_capture=3] [baselines=16]"""
import collections
import csv
import gzip
import json
import os
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
STEPS = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
BASELINES = int(sys.argv[3]) if len(sys.argv) > 3 else 16
DOMINANT = "k_box5b"          # second filtered axis of the 2-D masked filter

path = "gpurun_out/%s_traffic_all.csv" % TAG
op = gzip.open if not os.path.exists(path) else open
if not os.path.exists(path):
    path += ".gz"
lines = [l for l in op(path, "rt") if not l.startswith("==")]
per = collections.defaultdict(lambda: collections.defaultdict(float))   # launch id -> metric -> value
name = {}
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0,
          "inst": 1.0}.get(u, 1.0)
    per[row["ID"]][row["Metric Name"]] = v
    name[row["ID"]] = row["Kernel Name"].split("(")[0]
by = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i, m in per.items():
    b = by[name[i]]
    b[0] += 1
    b[1] += m["gpu__time_duration.sum"]
    b[2] += m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
    b[3] += m["smsp__inst_executed.sum"]
nvis = BASELINES * 4 * 512 * 4096
tot_b = sum(b[2] for b in by.values()) / STEPS
tot_i = sum(b[3] for b in by.values()) / STEPS
tot_l = sum(b[0] for b in by.values()) / STEPS
tot_ms = sum(b[1] for b in by.values()) / STEPS
dom = [m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"] for i, m in per.items()
       if DOMINANT in name[i] and m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"] > 1e8]
out = {
    "command": "python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-light --parity-planes 0 "
               "--baselines %d (every launch under ncu --metrics; %g strategy steps in the capture)" % (BASELINES, STEPS),
    "visibilities_per_step": nvis, "launches_per_step_in_capture": tot_l,
    "launches_note": "every launch of the process divided by the steps: includes torch's input-generation kernels; the library counts its own launches in bench.py's gpu_launches (672 per 8-pass step)", "dram_bytes_per_step": tot_b,
    "dram_bytes_per_visibility": tot_b / nvis, "algorithmic_bytes_per_visibility": 10,
    "warp_instructions_per_step": tot_i, "thread_instructions_per_visibility": tot_i * 32 / nvis,
    "gpu_time_ms_per_step_under_ncu": tot_ms,
    "by_kernel": {k: {"launches_per_step": b[0] / STEPS, "ms_per_step": round(b[1] / STEPS, 3),
                      "dram_gb_per_step": round(b[2] / STEPS / 1e9, 3), "warp_instr_M_per_step": round(b[3] / STEPS / 1e6, 3)}
                  for k, b in sorted(by.items(), key=lambda kv: -kv[1][1])},
    "note": "ncu serialises and cold-starts every launch, so times are larger than in a free run",
    "kernel": "box_filter (%s, second filtered axis)" % DOMINANT,
    "dram_bytes_per_launch": sum(dom) / max(len(dom), 1), "launches_sampled": len(dom), "nvis_per_launch": nvis,
    "source": "ncu --metrics over every launch of a %d-baseline step: profiles/%s_traffic.json" % (BASELINES, TAG),
}
json.dump(out, open("profiles/%s_traffic.json" % TAG, "w"), indent=1)
print("launches/step %.0f, DRAM %.1f GB/step (%.0f B/vis), %.0f thread-instr/vis, %s %.2f GB/launch (%d)" % (
    tot_l, tot_b / 1e9, tot_b / nvis, tot_i * 32 / nvis, DOMINANT, out["dram_bytes_per_launch"] / 1e9, len(dom)))
