"""SASS evidence for profiles/: the bulk-copy (TMA) staging prologue and the packed FP32x2 static-object loop of the
thread-per-candidate sweep, cut out of the built library with cuobjdump + nvdisasm (no GPU needed).
usage: python tools/sass_excerpt.py <tag>   ->  profiles/<tag>_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
lib = os.path.join(ROOT, "humap_local_planner_b200", "lib", "libhmp_planner.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
sass = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, "hmp_kernels.sm_100a.cubin")], capture_output=True, text=True).stdout.splitlines()
kern, name = [], None
for ln in sass:
    if ln.startswith(".text."):
        name = ln
    elif name and "sweep_tpc_kernelILi2ELb1E" in name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln) or (name and "sweep_tpc_kernelILi2ELb1E" in name and ln.startswith(".L_")):
        kern.append(ln)
ops = collections.Counter()
for ln in kern:
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m:
        ops[m.group(1)] += 1
out = [f"SASS excerpt of sweep_tpc_kernel<2> (sm_100a) from {os.path.relpath(lib, ROOT)} [{tag}]",
       f"instructions: {sum(ops.values())}; packed FP32x2: FFMA2 {ops['FFMA2']} FMUL2 {ops['FMUL2']} FADD2 {ops['FADD2']}; "
       f"scalar FP32: FFMA {ops['FFMA']} FMUL {ops['FMUL']} FADD {ops['FADD']}; MUFU {ops['MUFU']}; bulk copies UBLKCP {ops['UBLKCP']}; "
       f"mbarrier SYNCS {ops['SYNCS']}; tensor-core / TMEM opcodes (UTC*, HMMA, UTMALDG): "
       f"{sum(v for k, v in ops.items() if k.startswith('UTC') or k in ('HMMA', 'UTMALDG'))}", ""]
# (a) staging prologue: the three cp.async.bulk copies onto one mbarrier
idx = [i for i, ln in enumerate(kern) if "UBLKCP" in ln]
if idx:
    out.append("---- staging prologue: parameters | scene blob | costmap window by cp.async.bulk (UBLKCP) onto one mbarrier ----")
    out += kern[max(0, idx[0] - 12): idx[-1] + 8]
    out.append("")
# (b) the packed static-object loop: the densest run of FFMA2 / FMUL2 / FADD2 that ends with a backward branch
dense = [i for i, ln in enumerate(kern) if re.search(r"\b(FFMA2|FMUL2|FADD2)\b", ln)]
if dense:
    # longest cluster of packed instructions with gaps < 12
    clusters, cur = [], [dense[0]]
    for i in dense[1:]:
        if i - cur[-1] < 12:
            cur.append(i)
        else:
            clusters.append(cur)
            cur = [i]
    clusters.append(cur)
    best = max(clusters, key=len)
    lo, hi = best[0], best[-1]
    while hi < len(kern) - 1 and not re.search(r"\bBRA\b", kern[hi]):
        hi += 1
    body = kern[max(0, lo - 6): hi + 1]
    n_packed = sum(1 for ln in body if re.search(r"\b(FFMA2|FMUL2|FADD2)\b", ln))
    n_mufu = sum(1 for ln in body if "MUFU" in ln)
    out.append(f"---- packed static-object loop (two objects per iteration): {len(body)} lines, {n_packed} packed FP32x2 instructions, {n_mufu} MUFU ----")
    out += body
open(os.path.join(ROOT, "profiles", f"{tag}_sass_excerpt.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:3]))
print("lines", len(out))
