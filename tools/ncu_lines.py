#!/usr/bin/env python
"""Attribute an `ncu --page source --csv` export (SASS level) to source lines / named line ranges.
usage: tools/ncu_lines.py <sass.csv> <nvdisasm -g -c output> [ranges: name:lo-hi ...]
The i-th SASS instruction of the kernel in the ncu export is matched with the i-th instruction of the disassembly."""
import csv, re, sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[h]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = [(r[1], int(r[ci] or 0), int(r[cs] or 0)) for r in rows[h + 1:] if len(r) > cs]
# nvdisasm -gi prints the inline chain of an instruction as consecutive "//## File" lines, innermost first; an
# instruction is attributed to the innermost frame that lies in the kernel's own source file (MAIN_FILE env, default:
# the file most instructions come from is found in a second pass)
import os
lines, chain, cur, fresh = [], [], None, True
MAIN = os.environ.get("MAIN_FILE")
for ln in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        if chain:
            pick = next((c for c in chain if MAIN and c[0] == MAIN), chain[-1])
            cur = pick
        lines.append(cur)
        fresh = True
assert len(lines) >= len(ins), (len(lines), len(ins))
by = defaultdict(lambda: [0, 0])
for (src, n, s), loc in zip(ins, lines):
    by[loc][0] += n
    by[loc][1] += s
tot_i = sum(v[0] for v in by.values()); tot_s = sum(v[1] for v in by.values())
ranges = []
for a in sys.argv[3:]:
    nm, r = a.split(":"); lo, hi = r.split("-"); ranges.append((nm, int(lo), int(hi)))
if ranges:
    main = max(set(f for f, _ in by if f), key=lambda f: sum(v[0] for k, v in by.items() if k[0] == f))
    acc = defaultdict(lambda: [0, 0])
    for (f, l), v in by.items():
        nm = f if f != main else next((n for n, lo, hi in ranges if lo <= l <= hi), "other")
        acc[nm][0] += v[0]; acc[nm][1] += v[1]
    for nm, v in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print(f"{nm:28s} instr {v[0]:12d} {100*v[0]/tot_i:5.1f}%   samples {v[1]:8d} {100*v[1]/max(tot_s,1):5.1f}%")
else:
    for (f, l), v in sorted(by.items(), key=lambda kv: -kv[1][0])[:60]:
        print(f"{f}:{l:5d} instr {v[0]:12d} {100*v[0]/tot_i:5.1f}%   samples {v[1]:8d} {100*v[1]/max(tot_s,1):5.1f}%")
print("total instr", tot_i, "samples", tot_s)
