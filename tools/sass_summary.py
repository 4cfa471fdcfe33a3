"""Mnemonic counts per kernel + three inner-loop excerpts from the SASS of libkbp.so -> profiles/r02_sass_evidence.txt.
usage: python tools/sass_summary.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "kagomeperiodicbp_b200", "libkbp.so")], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
out = ["SASS evidence for libkbp.so (cuobjdump -sass, sm_100a build of this commit; regenerate: tools/sass_summary.py)",
       "mnemonic counts per kernel: DMMA = FP64 tensor core (mma.sync.m16n8k8.f64; tcgen05 has no FP64 kind), LDGSTS = cp.async global->shared,",
       "STAS / SYNCS / UCGABAR = distributed shared memory stores + mbarrier + cluster barrier, MUFU64 = FP64 rsqrt / reciprocal seeds, SHFL = warp shuffles,",
       "UBLKCP = cp.async.bulk (TMA unit, 1-D; the opt-in BULK variants of the small-tile GEMM)", ""]
rows = []
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", f)
    c = collections.Counter(i.split(".")[0] for i in ins)
    c2 = collections.Counter(ins)
    dem = re.sub(r"\(.*", "", subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip())
    rows.append((dem, len(ins), c["DMMA"], c["DFMA"] + c["DMUL"] + c["DADD"], c["LDGSTS"], c["LDS"], c["STS"], c["SHFL"], c["BAR"], c["STAS"], c["SYNCS"],
                 c["UCGABAR"], c2.get("MUFU.RSQ64H", 0) + c2.get("MUFU.RCP64H", 0), c["UBLKCP"]))
out.append(f"{'kernel':78s} {'instr':>6s} {'DMMA':>5s} {'DFP':>5s} {'LDGSTS':>6s} {'LDS':>5s} {'STS':>5s} {'SHFL':>5s} {'BAR':>4s} {'STAS':>4s} {'SYNCS':>5s} {'UCGABAR':>7s} {'MUFU64':>6s} {'UBLKCP':>6s}")
for r in sorted(rows, key=lambda r: -r[1]):
    out.append(f"{r[0][:78]:78s} {r[1]:6d} {r[2]:5d} {r[3]:5d} {r[4]:6d} {r[5]:5d} {r[6]:5d} {r[7]:5d} {r[8]:4d} {r[9]:4d} {r[10]:5d} {r[11]:7d} {r[12]:6d} {r[13]:6d}")


def excerpt(pattern, anchor, before, after, title):
    for f in funcs:
        name = f.split("\n", 1)[0]
        if pattern not in name:
            continue
        lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/\s", l) and not re.match(r"\s*/\* 0x", l)]
        idx = [i for i, l in enumerate(lines) if anchor in l]
        if not idx:
            continue
        i = idx[len(idx) // 2]
        out.append("")
        out.append(f"--- {title} ({name.strip()[:70]})")
        for l in lines[max(0, i - before):i + after]:
            out.append("   " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", l).strip())
        return


excerpt("zgemm_dmma_kernelILi16ELi16ELi3ELi2ELb1ELb0E", "DMMA", 14, 14, "ZGEMM 16x16 tile inner step: LDS.128 fragment loads feeding DMMA, LDGSTS prefetch")
excerpt("chol_reg_kernelILi2E", "MUFU.RCP64H", 10, 24, "register Cholesky, publishing a pivot row: SHFL pivot broadcast, RCP64H + Newton, STS.128 of the row and the multipliers, BAR")
excerpt("svd_cluster_kernelILi2ELi8E", "STAS", 10, 10, "cluster Jacobi: partial Gram sums pushed to the peer CTAs (STAS = st.async to distributed shared memory, completing the receiver's mbarrier)")
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_evidence.txt")
open(path, "w").write("\n".join(out) + "\n")
print(f"wrote {path}: {len(rows)} kernels")
