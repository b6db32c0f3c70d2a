"""Executed FP work of one kernel from an `ncu --page source --csv` dump (per-SASS-instruction counts): thread-level executed
instructions summed per opcode class. Packed FP32x2 opcodes (FFMA2 / FMUL2 / FADD2) perform two operations per thread
instruction and are counted twice. Together with issue-active and the pipe utilisations of the raw page this is what
bench.py's roofline.executed_fp32_frac / issue_active come from.
usage: ncu_opcounts.py <src.csv> <raw.csv> <out.json> [capture description]"""
import csv
import json
import re
import sys
from collections import defaultdict

src_csv, raw_csv, out_json = sys.argv[1:4]
what = sys.argv[4] if len(sys.argv) > 4 else ""
rows = list(csv.reader(open(src_csv)))
hdr = rows[1] if "Address" in rows[1] else rows[0]
start = rows.index(hdr) + 1
ci = {h: i for i, h in enumerate(hdr)}
col_src = next(h for h in hdr if h in ("Source", "SASS", "Instruction"))
col_thr = next((h for h in hdr if h.startswith("Predicated-On Thread Instructions Executed")), None) or \
    next(h for h in hdr if h.startswith("Thread Instructions Executed"))
col_warp = next(h for h in hdr if h == "Instructions Executed")


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


thr = defaultdict(float)
warp = defaultdict(float)
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", r[ci[col_src]])
    if not m:
        continue
    op = m.group(1)
    base = op.split(".")[0]
    key = "MUFU" if base == "MUFU" else base
    thr[key] += num(r[ci[col_thr]])
    warp[key] += num(r[ci[col_warp]])
g = lambda k: thr.get(k, 0.0)
out = {
    "what": what,
    "ffma": g("FFMA") + 2.0 * g("FFMA2"), "fmul": g("FMUL") + 2.0 * g("FMUL2"), "fadd": g("FADD") + 2.0 * g("FADD2"),
    "mufu": g("MUFU"), "fmnmx": g("FMNMX") + g("FMNMX3"), "fsetp_fsel": g("FSETP") + g("FSEL"),
    "dfma": g("DFMA"), "dmul": g("DMUL"), "dadd": g("DADD"),
    "conversions": sum(v for k, v in thr.items() if k in ("F2F", "F2I", "I2F", "FRND", "F2FP", "I2FP")),
    "thread_inst_total": sum(thr.values()), "warp_inst_total": sum(warp.values()),
    "thread_inst_by_opcode": {k: v for k, v in sorted(thr.items(), key=lambda kv: -kv[1])[:40]},
}
raw = list(csv.reader(open(raw_csv)))
d = dict(zip(raw[0], raw[2]))
for name, key in (("issue_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  ("fma_pipe", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                  ("alu_pipe", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                  ("xu_pipe", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                  ("fp64_pipe", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                  ("warps_active", "sm__warps_active.avg.pct_of_peak_sustained_active")):
    if key in d:
        out[name] = num(d[key]) / 100.0
if "gpu__time_duration.sum" in d:
    out["capture_duration"] = d["gpu__time_duration.sum"] + " " + dict(zip(raw[0], raw[1])).get("gpu__time_duration.sum", "")
out["kernel"] = d.get("Kernel Name", "?")
json.dump(out, open(out_json, "w"), indent=1)
print(json.dumps({k: out[k] for k in out if k != "thread_inst_by_opcode"}))
