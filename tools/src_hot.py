#!/usr/bin/env python
"""Hot spots of one kernel launch from `ncu --page source --csv --print-source sass` output
(plain or .gz): stall samples and executed-instruction counts per SASS instruction.

usage: src_hot.py capture_src0.csv[.gz] [top N] [--loop]   (--loop prints the hottest loop body in order)"""
import collections
import csv
import gzip
import re
import sys


def load(path):
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(op(path, "rt")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    data, seen = [], set()
    for r in rows[hi + 1:]:
        if len(r) <= iex or r[ia] in seen:
            continue
        try:
            data.append((r[ia], r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0)))
            seen.add(r[ia])
        except ValueError:
            pass
    return rows[0][1] if len(rows[0]) > 1 else "?", data


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 25
    name, data = load(path)
    tot_s = sum(d[2] for d in data)
    tot_i = sum(d[3] for d in data)
    print("##", name, "| samples", tot_s, "| warp instructions", tot_i)
    mix, smp = collections.Counter(), collections.Counter()
    for _, src, s, n in data:
        op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0].split(".")[0]
        mix[op] += n
        smp[op] += s
    print("mix (% executed / % samples): " + ", ".join(
        "%s %.1f/%.1f" % (k, 100.0 * v / max(tot_i, 1), 100.0 * smp[k] / max(tot_s, 1)) for k, v in mix.most_common(16)))
    print("hottest instructions (% of samples, executed):")
    for a, src, s, n in sorted(data, key=lambda d: -d[2])[:top]:
        print("  %5.1f%% %10d  %s" % (100.0 * s / max(tot_s, 1), n, src))
    if "--loop" in sys.argv:
        c = collections.Counter(d[3] for d in data if d[3] > 0)
        # the execution count that carries the most instructions x executions
        best = max(c, key=lambda k: k * c[k])
        print("loop body (executed %d times):" % best)
        for a, src, s, n in data:
            if n >= 0.9 * best:
                print("  %5d %9d  %s" % (s, n, src))


if __name__ == "__main__":
    main()
