"""Joins an `ncu --page source --csv` (per-SASS-instruction) dump with nvdisasm line info so that executed
instructions and stall samples can be read per CUDA source line.
usage: ncu_by_line.py <src.csv> <cubin> <kernel mangled substring> [top N] [first line] [source file of the kernel body]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

src_csv, cubin, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
KERNEL_FIRST_LINE = int(sys.argv[5]) if len(sys.argv) > 5 else 0   # attribute to the outermost frame at/after this line
KERNEL_FILE = sys.argv[6] if len(sys.argv) > 6 else "hmp_kernels.cu"   # file that holds the kernel body
addr2line = {}
frames = []
last = None
in_k = False
for ln in dis:
    if ln.startswith(".text."):
        in_k = kname in ln
    if not in_k:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        frames.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        pick = None
        for f, l in frames:            # innermost first; keep the outermost frame that lies in the kernel body
            if f.endswith(KERNEL_FILE) and l >= KERNEL_FIRST_LINE:
                pick = l
        if pick is None and frames:
            pick = frames[-1][1]
        if pick is None and not frames:
            pick = last          # nvdisasm prints line info only when it changes
        if pick is not None:
            addr2line[int(m.group(1), 16)] = pick
            last = pick
        frames = []
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
inst = defaultdict(float)
samp = defaultdict(float)
stall = defaultdict(lambda: defaultdict(float))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = None
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
    if base is None:
        base = a
    line = addr2line.get(a - base, -1)
    f = lambda x: float(x) if x not in ("", "-") else 0.0
    inst[line] += f(r[ci["Instructions Executed"]])
    samp[line] += f(r[ci["# Samples"]])
    for s in stall_cols:
        stall[line][s] += f(r[ci[s]])
ti, ts = sum(inst.values()), sum(samp.values())
srclines = open("humap_local_planner_b200/csrc/" + KERNEL_FILE).read().splitlines()
print(f"total warp instructions {ti:.3e}, samples {ts:.0f}, mapped lines {len(inst)}")
top = sorted(inst, key=lambda l: -samp[l])[:topn]
for l in sorted(top):
    st = sorted(stall[l].items(), key=lambda kv: -kv[1])[:2]
    sts = " ".join(f"{k.replace('stall_', '')}:{100 * v / max(1, samp[l]):.0f}%" for k, v in st)
    text = srclines[l - 1].strip()[:80] if 0 < l <= len(srclines) else "?"
    print(f"{l:5d} inst {100 * inst[l] / ti:5.1f}% samp {100 * samp[l] / ts:5.1f}% [{sts:32s}] {text}")
