#!/usr/bin/env python
"""Join an ncu SASS source page (csv) with nvdisasm -g line info -> per-source-line instruction counts.
usage: sass_lines.py <src.csv from `ncu --page source --csv`> <nvdisasm -g output> <kernel substring> [top]"""
import csv, re, sys, collections
srccsv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]
iI, iT, iS = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
inst = [(r[1].strip(), int(r[iI]), int(r[iT]), int(r[iS])) for r in rows[2:] if len(r) > iS and r[iI].isdigit()]
# nvdisasm: collect (line) per instruction of the function
lines, cur, infn = [], None, False
for l in open(sass, errors='ignore'):
    if l.startswith('.text.'):
        infn = kname in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
        lines.append(cur)
n = min(len(lines), len(inst))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for k in range(n):
    a = agg[lines[k]]
    a[0] += inst[k][1]; a[1] += inst[k][2]; a[2] += inst[k][3]
    tot += inst[k][1]
print("instructions matched: %d sass / %d profiled; total warp-inst %d" % (len(lines), len(inst), tot))
srcs = {}
for key, (wi, ti, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, ln = key if key else ("?", 0)
    try:
        if f not in srcs:
            srcs[f] = open('/root/repo/non_matching_test_suite_b200/csrc/' + f).read().split('\n')
        text = srcs[f][ln - 1].strip()[:100]
    except Exception:
        text = ""
    print("%6.2f%%  lanes %5.1f  samples %6d  %s:%d  %s" % (100.0 * wi / tot, ti / max(wi, 1), sm, f, ln, text))
