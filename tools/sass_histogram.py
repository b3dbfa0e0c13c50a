#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libaudiosync_cuda.so (cuobjdump -sass; no GPU needed).
Writes profiles/<round>_sass_opcodes.md: the evidence for TMA staging (UTMALDG / UBLKCP), packed
fp32 arithmetic (FADD2 / FMUL2 / FFMA2), programmatic dependent launch (ACQBULK / PREEXIT ...) and
the absence of tensor-core / library code.   usage: sass_histogram.py [round tag, default r02]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "old-audiosync_b200", "libaudiosync_cuda.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern = None; per = collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); per[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m and kern:
        op = m.group(2); mod = m.group(3) or ""
        key = op
        if op in ("UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "UTMASTG", "ACQBULK", "REDUX"):
            key = op + ".".join(mod.split(".")[:3])
        per[kern][key] += 1
demangle = subprocess.run(["c++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
def short(name):
    name = re.sub(r"asc::|\(.*$|void ", "", name)
    return name[:120]
watch = ["UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "FADD2", "FMUL2", "FFMA2", "FADD", "FMUL", "FFMA", "DFMA", "DADD", "DMUL",
         "LDS", "STS", "LDG", "STG", "BAR", "HMMA", "UTCHMMA", "IMMA", "ATOMG", "REDUX", "SHFL", "ACQBULK"]
out = ["# Round %d: SASS opcode histogram of libaudiosync_cuda.so" % int(rnd[1:]), "",
       "`cuobjdump -sass old-audiosync_b200/libaudiosync_cuda.so` (sm_100a is the only image), static instruction counts per kernel.",
       "TMA staging shows as `UTMALDG.3D` (tensor-map boxes of the column tiles) and `UBLKCP.S.G` / `UBLKCP.G.S` (bulk copies of",
       "rows, global->shared / shared->global) with `SYNCS` (mbarrier) beside them; the packed fp32 forms are `FADD2 / FMUL2 / FFMA2`;",
       "the fp64 validation kernels are `DFMA / DADD / DMUL`; there is no `HMMA / UTCHMMA / IMMA` (no tensor cores: nothing here is a GEMM)",
       "and no library code (the .so links only libcudart and libpthread, see tests/test_boundary.py).", "",
       "| kernel | total | " + " | ".join(watch) + " |", "|---|---|" + "---|" * len(watch)]
tot = collections.Counter()
for (k, c), dn in zip(per.items(), demangle):
    agg = collections.Counter()
    for op, n in c.items():
        base = op.split(".")[0]
        agg[base] += n
        tot[op] += n
    out.append("| `%s` | %d | " % (short(dn), sum(c.values())) + " | ".join(str(agg.get(w, 0)) for w in watch) + " |")
out += ["", "## Whole library, TMA / async-copy / barrier opcodes with their modifiers", ""]
for op, n in sorted(tot.items()):
    if op.split(".")[0] in ("UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "UTMASTG", "ACQBULK", "REDUX"):
        out.append("- `%s`: %d" % (op, n))
out += ["", "Packed fp32 in the whole library: FADD2 %d, FMUL2 %d, FFMA2 %d; scalar FADD %d, FMUL %d, FFMA %d; fp64 DFMA %d, DADD %d, DMUL %d." % tuple(
    sum(v for k, v in tot.items() if k.split(".")[0] == o) for o in ("FADD2", "FMUL2", "FFMA2", "FADD", "FMUL", "FFMA", "DFMA", "DADD", "DMUL"))]
open(os.path.join(ROOT, "profiles", "%s_sass_opcodes.md" % rnd), "w").write("\n".join(out) + "\n")
print("\n".join(out[-14:]))
