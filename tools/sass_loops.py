#!/usr/bin/env python
"""List the loops (backward branches) of a kernel in a cuobjdump -sass listing with
their instruction mix.  usage: sass_loops.py listing.sass kernel-substring [min_len]"""
import re, sys, collections
lst, pat = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 20
cur, ins = None, []
funcs = {}
for line in open(lst):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name: continue
    print("==", name, len(ins), "instructions")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
        if not m: continue
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr and i - addr[tgt] >= minlen:
            body = ins[addr[tgt]:i + 1]
            mix = collections.Counter()
            for _, x in body:
                x = re.sub(r"^@!?U?P\d+\s+", "", x)
                op = x.split()[0].split(".")[0]
                mix[op] += 1
            print("  loop %#x..%#x: %d instr  " % (tgt, a, len(body)) + " ".join("%s=%d" % kv for kv in mix.most_common(24)))
