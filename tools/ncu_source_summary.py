#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv` per kernel: instruction mix by opcode,
stall samples by opcode, shared-memory wavefront excess.  usage: ncu_source_summary.py X.ncu-rep [top]"""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
kern = None; hdr = None; per = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        kern = re.sub(r"asc::|\(int\)", "", r[1])[:110]; per[kern] = []; hdr = None; continue
    if r[0] == "Address": hdr = r; continue
    if hdr and kern: per[kern].append(dict(zip(hdr, r)))
def num(x):
    try: return float(x)
    except: return 0.0
for k, lines in per.items():
    tot_i = sum(num(l["Instructions Executed"]) for l in lines)
    tot_s = sum(num(l["# Samples"]) for l in lines)
    print("=" * 100); print(k); print("warp instructions %.3fM, stall samples %d, SASS lines %d" % (tot_i / 1e6, tot_s, len(lines)))
    byop_i = collections.Counter(); byop_s = collections.Counter()
    sh_w = sh_ideal = 0
    for l in lines:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", l["Source"]); op = m.group(2) if m else "?"
        byop_i[op] += num(l["Instructions Executed"]); byop_s[op] += num(l["# Samples"])
        sh_w += num(l.get("L1 Wavefronts Shared", 0)); sh_ideal += num(l.get("L1 Wavefronts Shared Ideal", 0))
    print("shared wavefronts %.2fM (ideal %.2fM)" % (sh_w / 1e6, sh_ideal / 1e6))
    print("%-12s %8s %6s | %8s %6s" % ("opcode", "instrM", "%", "samples", "%"))
    for op, v in byop_i.most_common(top):
        print("%-12s %8.3f %6.1f | %8d %6.1f" % (op, v / 1e6, 100 * v / tot_i, byop_s[op], 100 * byop_s[op] / max(1, tot_s)))
    # stall reason columns
    reasons = [h for h in hdr if h.startswith("stall_")] if hdr else []
    if reasons:
        agg = {h: sum(num(l.get(h, 0)) for l in lines) for h in reasons}
        s = sum(agg.values()) or 1
        print("stall reasons:", ", ".join("%s %.1f%%" % (h[6:], 100 * v / s) for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    # phase view: SASS order split at barriers (kernels are fully unrolled pass by pass)
    seg = []; cur = {"i": 0.0, "s": 0.0, "n": 0, "ldg": 0.0, "fp": 0.0}
    for l in lines:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", l["Source"]); op = m.group(2) if m else "?"
        cur["i"] += num(l["Instructions Executed"]); cur["s"] += num(l["# Samples"]); cur["n"] += 1
        if op in ("LDG", "STG", "LDGSTS"): cur["ldg"] += num(l["Instructions Executed"])
        if op in ("FADD2", "FMUL2", "FFMA2", "FADD", "FMUL", "FFMA"): cur["fp"] += num(l["Instructions Executed"])
        if op == "BAR":
            seg.append(cur); cur = {"i": 0.0, "s": 0.0, "n": 0, "ldg": 0.0, "fp": 0.0}
    seg.append(cur)
    print("phases (between barriers): " + " | ".join("%d: %.1f%%t %.1fMi fp%.1fM g%.2fM" % (j, 100 * g["s"] / max(1, tot_s), g["i"] / 1e6, g["fp"] / 1e6, g["ldg"] / 1e6) for j, g in enumerate(seg)))
