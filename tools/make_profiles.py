"""Turns the raw outputs of one measurement run under gpurun_out/ into the tracked summaries under profiles/.

Expected inputs (written on the GPU box by the commands in DESIGN.md section 6 / the round's run script):
  gpurun_out/<tag>_full.ncu-rep          one `ncu --set full --import-source on -k regex:sweep_tpc -c 1` capture of bench.py
  gpurun_out/<tag>_launches.csv          `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py
  gpurun_out/bench_<tag>[_cfg1|_batched|_reference].json   bench.py lines
usage: python tools/make_profiles.py <tag>      (e.g. r01j; needs ncu + cuobjdump + nvdisasm, no GPU)"""
import csv
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = os.path.join(ROOT, "profiles")
src = os.path.join(ROOT, "gpurun_out")
tmp = tempfile.mkdtemp()
rep = os.path.join(src, f"{tag}_full.ncu-rep")
raw, page = os.path.join(tmp, "raw.csv"), os.path.join(tmp, "src.csv")
for kind, dst in (("raw", raw), ("source", page)):
    with open(dst, "w") as f:
        subprocess.run(["ncu", "-i", rep, "--page", kind, "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=True)
with open(os.path.join(out, f"{tag}_ncu_summary.txt"), "w") as f:
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), raw], stdout=f, check=True)
lib = os.path.join(ROOT, "humap_local_planner_b200", "lib", "libhmp_planner.so")
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = os.path.join(tmp, "hmp_kernels.sm_100a.cubin")
with open(os.path.join(out, f"{tag}_ncu_by_line.txt"), "w") as f:
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), page, cubin, "sweep_tpc_kernelILi2ELb1E", "60", "0", "hmp_sweep_tpc.inl"],
                   stdout=f, stderr=subprocess.STDOUT, cwd=ROOT)
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_opcounts.py"), page, raw, os.path.join(out, f"{tag}_ncu_counts.json"),
                f"ncu --set full --clock-control none, bench.py --steps 1 --warmup 3 ({tag}), sweep_tpc_kernel"], check=True, stdout=subprocess.DEVNULL)
rows = list(csv.reader(open(raw)))
d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tr = {"kernel": d["Kernel Name"], "capture": f"ncu --set full --clock-control none, bench.py --steps 1 --warmup 3 ({tag})"}
for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    tr[k.replace("dram__bytes_", "dram_bytes_").replace(".sum", "")] = int(float(d[k].replace(",", "")) * scale[u[k]])
json.dump(tr, open(os.path.join(out, f"{tag}_traffic.json"), "w"))
for suffix, name in (("", "bench"), ("_cfg1", "bench_cfg1"), ("_batched", "bench_batched_cfg3"), ("_reference", "bench_reference")):
    p = os.path.join(src, f"bench_{tag}{suffix}.json")
    if os.path.exists(p):
        line = open(p).read().strip().splitlines()[-1]
        json.loads(line)
        open(os.path.join(out, f"{tag}_{name}.json"), "w").write(line + "\n")
shutil.copy(os.path.join(src, f"{tag}_launches.csv"), os.path.join(out, f"{tag}_launches.csv"))
print("profiles written for", tag)
