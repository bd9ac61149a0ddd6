#!/usr/bin/env python
"""Turn the raw files of one `tools/gpu_profiles.sh TAG` visit (gpurun_out/TAG_*) into the tracked evidence under profiles/:
launch list + per-kernel totals, per-launch DRAM traffic of the step kernel for both layouts (profiles/traffic_per_launch.json,
read by bench.py for roofline.traffic), the ncu summary / per-line / per-region tables of the full capture, the bench lines.
    python tools/collect_profiles.py TAG OUTPREFIX        e.g.  r01g r01_final2"""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, out = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def rows(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(lines))


# launch list
shutil.copy(os.path.join(G, f"{tag}_launches.csv"), os.path.join(P, f"{out}_launches.csv"))
agg = OrderedDict()
for r in rows(os.path.join(G, f"{tag}_launches.csv")):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    v = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else (v * 1e3 if r["Metric Unit"] in ("ms", "msecond") else v)
    name = r["Kernel Name"].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
with open(os.path.join(P, f"{out}_launches_by_kernel.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none of: python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 8\n")
    for k, (n, t) in agg.items():
        f.write(f"{n:5d} launches {t:10.1f} us total {t / n:9.1f} us/launch  {k}\n")

# traffic
tr = {}
for L in ("mv", "tiled"):
    p = os.path.join(G, f"{tag}_traffic_{L}.csv")
    shutil.copy(p, os.path.join(P, f"{out}_traffic_{L}.csv"))
    rd = [float(r["Metric Value"].replace(",", "")) for r in rows(p) if r["Metric Name"] == "dram__bytes_read.sum"]
    wr = [float(r["Metric Value"].replace(",", "")) for r in rows(p) if r["Metric Name"] == "dram__bytes_write.sum"]
    unit = [r["Metric Unit"] for r in rows(p) if r["Metric Name"] == "dram__bytes_read.sum"][0]
    mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    tr[f"{L}_read_bytes_per_launch"] = mul * sum(rd) / len(rd)
    tr[f"{L}_write_bytes_per_launch"] = mul * sum(wr) / len(wr)
    tr[f"{L}_bytes_per_launch"] = tr[f"{L}_read_bytes_per_launch"] + tr[f"{L}_write_bytes_per_launch"]
tr["how"] = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ipp_step_async -s 8 -c 4, bench.py C3 workload "
             f"at the full batch (65536 envs), mean of 4 launches; tools/gpu_profiles.sh; raw: profiles/{out}_traffic_*.csv")
json.dump(tr, open(os.path.join(P, "traffic_per_launch.json"), "w"), indent=1)

# full capture
rep = os.path.join(G, f"{tag}_async_tiled.ncu-rep")
kern = "ipp_step_async_kernelILb1ELb0ELb0ELb1E"
for tool, suffix, args in (("ncu_summary.py", "ncu_summary", []), ("ncu_lines.py", "hotlines", [kern, "90"]), ("ncu_groups.py", "groups", ["65536", kern])):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), rep] + args, capture_output=True, text=True).stdout
    open(os.path.join(P, f"{out}_async_tiled_{suffix}.txt"), "w").write(txt)
for src, dst in ((f"{tag}_bench.json", f"{out}_bench.json"), (f"{tag}_bench_ref.json", f"{out}_bench_reference_arm.json"), (f"{tag}_dram_probe2.txt", f"{out}_dram_probe2.txt")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
print(json.dumps(tr, indent=1))
print(open(os.path.join(P, f"{out}_launches_by_kernel.txt")).read())
