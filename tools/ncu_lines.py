#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples of one kernel from an .ncu-rep captured with --import-source on.
usage: tools/ncu_lines.py <rep> <kernel-regex> [top_n]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-count", "1",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; agg = {}; cur = None; fname = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != "":    # a source line row (aggregated over its SASS)
        key = (fname, int(r[0]))
        def num(x):
            try: return int(x)
            except ValueError: return 0
        ie = num(r[hdr.index("Instructions Executed")])
        smp = num(r[hdr.index("# Samples")])
        a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += ie; a[1] += smp
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print("total warp instructions", tot_i, "samples", tot_s)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-18s %5d  inst %9d (%4.1f%%)  samples %6d (%4.1f%%)  %s" % (key[0], key[1], a[0], 100.0 * a[0] / max(tot_i, 1), a[1], 100.0 * a[1] / max(tot_s, 1), a[2]))
