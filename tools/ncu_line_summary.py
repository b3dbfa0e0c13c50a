#!/usr/bin/env python3
"""Per-CUDA-source-line totals from `ncu --page source --print-source cuda,sass --csv`:
warp instructions and stall samples by (file, line), per kernel.
usage: ncu_line_summary.py X.ncu-rep KERNEL_REGEX [top]"""
import csv, io, re, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fname = kern = None; hdr = None
agg = collections.defaultdict(lambda: [0.0, 0.0, ""])   # (kernel, file, line) -> [instr, samples, text]
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": kern = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try: ins = float(r[hdr.index("Instructions Executed")]); smp = float(r[hdr.index("# Samples")])
        except ValueError: continue
        a = agg[(kern, fname, int(r[0]))]; a[0] += ins; a[1] += smp; a[2] = r[1].strip()[:90]
for k in sorted({k[0] for k in agg}):
    if not re.search(pat, k): continue
    rows = [(f, l, v) for (kk, f, l), v in agg.items() if kk == k]
    ti = sum(v[0] for _, _, v in rows) or 1; ts = sum(v[1] for _, _, v in rows) or 1
    print("=" * 110); print(re.sub(r"asc::|\(int\)", "", k)[:110]); print("total %.1fM warp instr, %d samples" % (ti / 1e6, ts))
    for f, l, v in sorted(rows, key=lambda x: -x[2][1])[:top]:
        print("%-18s %4d  %7.2fMi %5.1f%% | %6d smp %5.1f%% | %s" % (f, l, v[0] / 1e6, 100 * v[0] / ti, v[1], 100 * v[1] / ts, v[2]))
