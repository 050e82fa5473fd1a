# -*- coding: utf-8 -*-
"""Turns the ncu outputs brought back in gpurun_out/ into the small text/CSV
summaries committed under profiles/ (run in the build container, no GPU)."""
import collections
import csv
import gzip
import json
import shutil
import subprocess
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
SRC = "gpurun_out"

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches():
    with open("%s/%s_launches.csv" % (SRC, TAG)) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3}.get(row["Metric Unit"], 1.0)
        k = row["Kernel Name"].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open("profiles/%s_launch_summary.csv" % TAG, "w") as f:
        f.write("kernel,launches,total_ms,share_pct,avg_ms\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s",%d,%.3f,%.2f,%.4f\n' % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
    with open("%s/%s_launches.csv" % (SRC, TAG), "rb") as fi, gzip.open("profiles/%s_launches.csv.gz" % TAG, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    return tot


def raw(rep, out):
    txt = subprocess.run(["ncu", "-i", "%s/%s.ncu-rep" % (SRC, rep), "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h = rows[0]
    idx = [h.index(k) for k in KEEP if k in h]
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow([h[i] for i in idx])
        w.writerow([rows[1][i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    return rows


def stalls(rep, out):
    txt = subprocess.run(["ncu", "-i", "%s/%s.ncu-rep" % (SRC, rep), "--page", "source", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    seen = set()
    with open(out, "w") as f:
        for b in blocks:
            if b["name"] in seen or not b["rows"]:
                continue
            seen.add(b["name"])
            h, data = b["rows"][0], b["rows"][1:]
            isamp, isrc, iex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
            st = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
            tot = sum(int(r[isamp] or 0) for r in data) or 1
            agg = {s: sum(int(r[h.index(s)] or 0) for r in data) for s in st}
            mix = collections.Counter()
            for r in data:
                op = [o for o in r[isrc].split() if not o.startswith("@")]
                if op:
                    mix[op[0].split(".")[0]] += int(r[iex] or 0)
            te = sum(mix.values()) or 1
            f.write("## %s\n" % b["name"])
            f.write("stall reasons (%% of samples): %s\n" % ", ".join(
                "%s %.1f" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
            f.write("executed instruction mix (%%): %s\n" % ", ".join(
                "%s %.1f" % (k, 100.0 * v / te) for k, v in mix.most_common(14)))
            f.write("hottest instructions:\n")
            for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:8]:
                f.write("  %5.1f%%  %s\n" % (100.0 * int(r[isamp]) / tot, " ".join(r[isrc].split())[:90]))
            f.write("\n")


if __name__ == "__main__":
    import os
    tot = launches()
    if not os.path.exists("%s/%s_box_filter.ncu-rep" % (SRC, TAG)):
        # launch list only: keep the committed --set full summaries as they are
        print("total ms in launch list:", tot)
        sys.exit(0)
    rows = raw(TAG + "_box_filter", "profiles/%s_box_filter_metrics.csv" % TAG)
    stalls(TAG + "_box_filter", "profiles/%s_box_filter_stalls.txt" % TAG)
    if os.path.exists("%s/%s_other.ncu-rep" % (SRC, TAG)):
        raw(TAG + "_other", "profiles/%s_other_kernels_metrics.csv" % TAG)
        stalls(TAG + "_other", "profiles/%s_other_kernels_stalls.txt" % TAG)
    # (roofline.traffic of bench.py comes from profiles/summarize_traffic.py: every launch of a step)
    print("total ms in launch list:", tot)
