#!/usr/bin/env python3
"""Builds the committed profile summary of one GPU visit from gpurun_out/:
   profiles/<tag>_ncu.md   -- launch list (gpu__time_duration), per-kernel ncu metrics, instruction mix / phases
   profiles/traffic.json   -- ncu DRAM bytes per pair for each kernel class (bench.py reads it for roofline.traffic)
usage: make_profile_summary.py TAG PAIRS_PER_LAUNCH_IN_FULL_CAPTURE [PAIRS_IN_LAUNCH_LIST] [ROUND]   (e.g. v6 64 128 r02)"""
import collections, csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, ppl = sys.argv[1], int(sys.argv[2])
ppl_list = int(sys.argv[3]) if len(sys.argv) > 3 else 128
rnd = sys.argv[4] if len(sys.argv) > 4 else "r02"
G = os.path.join(ROOT, "gpurun_out")
out = ["# Round %d, %s kernels: ncu evidence" % (int(rnd[1:]), tag), ""]

def cls(name):
    for k, c in (("ColFwd", "col_fwd"), ("RowFused", "row_fused"), ("ColInv", "col_inv_argmax"), ("pearson", "pearson"),
                 ("synth", "synth")):
        if k in name: return c
    return name[:30]

# ---- launch list
lp = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(lp):
    rows = [r for r in csv.reader(l for l in open(lp) if l.startswith('"'))]
    hdr = rows[0]; ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = (cls(r[ki]), r[gi], r[bi]); agg.setdefault(k, []).append(float(r[vi].replace(",", "")))
    ours = ("col_fwd", "row_fused", "col_inv_argmax", "pearson")
    tot = sum(sum(v) for k, v in agg.items() if k[0] in ours)
    out += ["## Launch list", "", ("`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --pairs %d --steps 1 --warmup 1` "
             "(cold-cache, serialised: compare SHARES).") % ppl_list, "", "| kernel | grid | block | launches | avg us | share of path |", "|---|---|---|---|---|---|"]
    for (k, g, b), v in agg.items():
        out.append("| %s | %s | %s | %d | %.1f | %s |" % (k, g, b, len(v), sum(v) / len(v) / 1e3,
                                                       "%.3f" % (sum(v) / tot) if k in ours else "- (setup)"))
    out.append("")

# ---- full capture
rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
if os.path.exists(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt))); hdr, units, data = rows[0], rows[1], rows[2:]
    def col(name): return hdr.index(name) if name in hdr else None
    metrics = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
               ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
               ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fp32 pipe %"),
               ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
               ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
               ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/shared data pipe %"),
               ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
               ("smsp__inst_executed.sum", "warp instructions"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
               ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "global-load wavefronts"),
               ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum", "global-store wavefronts"),
               ("launch__registers_per_thread", "registers"), ("launch__occupancy_limit_shared_mem", "CTA/SM (smem)"),
               ("launch__occupancy_limit_registers", "CTA/SM (regs)")]
    names = [cls(d[col("Kernel Name")]) for d in data]
    out += ["## Full capture (`ncu --set full --clock-control none`, %d pairs per launch)" % ppl, "",
            "| metric | " + " | ".join(names) + " |", "|---|" + "---|" * len(names)]
    traffic = {}
    for m, label in metrics:
        i = col(m)
        if i is None: continue
        out.append("| %s (%s) | " % (label, units[i]) + " | ".join(d[i][:12] for d in data) + " |")
    ir, iw = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    def tobytes(v, u):
        v = float(v.replace(",", "")); return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    out += ["", "DRAM traffic per pair (read + write) / algorithmic bytes (U = L*4 = 5.76 MB):", ""]
    algU = {"col_fwd": 7, "row_fused": 6, "col_inv_argmax": 2, "pearson": 2}
    for n, d in zip(names, data):
        b = (tobytes(d[ir], units[ir]) + tobytes(d[iw], units[iw])) / ppl
        traffic[n] = {"bytes_per_pair": b, "source": "profiles/%s_%s_ncu.md" % (rnd, tag)}
        out.append("- %s: %.2f MB / pair measured, %.2f MB algorithmic (%d U)" % (n, b / 1e6, algU.get(n, 0) * 5.76, algU.get(n, 0)))
    tot = sum(v["bytes_per_pair"] for v in traffic.values())
    out += ["- whole path: %.2f MB / pair measured; 97.92 MB moved by design (17 U); 120.96 MB in the prescribed accounting (21 U)" % (tot / 1e6), ""]
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    src = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_summary.py"), rep, "12"], capture_output=True, text=True).stdout
    out += ["## Instruction mix, stall reasons and phase shares (source page)", "", "```", src.rstrip(), "```", ""]
open(os.path.join(ROOT, "profiles", "%s_%s_ncu.md" % (rnd, tag)), "w").write("\n".join(out))
print("\n".join(out[:60]))
