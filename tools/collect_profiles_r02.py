#!/usr/bin/env python
"""Turn the raw files of one `tools/gpu_r2_final.sh TAG` visit (gpurun_out/TAG_*) into the tracked evidence under profiles/r02_*.
    python tools/collect_profiles_r02.py TAG"""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()


def rows(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(lines))


def val(r):
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    return v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


# launch list of the bench command
shutil.copy(os.path.join(G, f"{tag}_launches.csv"), os.path.join(P, f"{out}_launches.csv"))
agg = OrderedDict()
for r in rows(os.path.join(G, f"{tag}_launches.csv")):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    a = agg.setdefault(r["Kernel Name"].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += val(r)
with open(os.path.join(P, f"{out}_launches_by_kernel.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none of: python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 4 --mcts-sims 8\n")
    f.write("# (cold-cache, serialised launches: compare shares, not absolutes; the timed step legs launch ipp_step_bulk_kernel only)\n")
    for k, (n, t) in agg.items():
        f.write(f"{n:5d} launches {t:10.1f} us total {t / n:9.1f} us/launch  {k}\n")

# DRAM traffic + instructions per launch of the step kernel (first 3 launches: trace mode; all six: both reward modes)
p = os.path.join(G, f"{tag}_traffic_super.csv")
shutil.copy(p, os.path.join(P, f"{out}_traffic_super.csv"))
m = defaultdict(list)
for r in rows(p):
    m[r["Metric Name"]].append(val(r))
rd, wr = sum(m["dram__bytes_read.sum"]) / len(m["dram__bytes_read.sum"]), sum(m["dram__bytes_write.sum"]) / len(m["dram__bytes_write.sum"])
tj_path = os.path.join(P, "traffic_per_launch.json")
tj = json.load(open(tj_path)) if os.path.exists(tj_path) else {}
tj.setdefault("kernels", {})["ipp_step_bulk_kernel:super"] = {
    "read_bytes_per_launch": rd, "write_bytes_per_launch": wr, "bytes_per_launch": rd + wr,
    "warp_instructions_per_launch": sum(m["smsp__inst_executed.sum"]) / len(m["smsp__inst_executed.sum"]),
    "us_per_launch_under_ncu": sum(m["gpu__time_duration.sum"]) / len(m["gpu__time_duration.sum"]), "commit": commit,
    "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none "
           f"-k regex:ipp_step_bulk -s 4 -c 6, bench.py C3 workload at the full batch (65536 envs), mean of 6 launches; raw: profiles/{out}_traffic_super.csv"}
if "tiled_bytes_per_launch" in tj:  # round-1 numbers of the cp.async kernel, kept for the A/B
    tj["kernels"].setdefault("ipp_step_async_kernel:tiled", {"bytes_per_launch": tj["tiled_bytes_per_launch"], "commit": "bb84100 (round 1)", "how": tj.get("how", "")})
    tj["kernels"].setdefault("ipp_step_async_kernel:mv", {"bytes_per_launch": tj["mv_bytes_per_launch"], "commit": "bb84100 (round 1)", "how": tj.get("how", "")})
json.dump(tj, open(tj_path, "w"), indent=1)

# the covariance-only persistent kernel on the split layout (committing)
pp = os.path.join(G, f"{tag}_traffic_predict_split.csv")
if os.path.exists(pp):
    shutil.copy(pp, os.path.join(P, f"{out}_traffic_predict_split.csv"))
    m2 = defaultdict(list)
    for r in rows(pp):
        m2[r["Metric Name"]].append(val(r))
    mean = lambda xs: sum(xs) / len(xs)
    tj["kernels"]["ipp_step_bulk_kernel<MODE_PREDICT>:split"] = {
        "read_bytes_per_launch": mean(m2["dram__bytes_read.sum"]), "write_bytes_per_launch": mean(m2["dram__bytes_write.sum"]),
        "bytes_per_launch": mean(m2["dram__bytes_read.sum"]) + mean(m2["dram__bytes_write.sum"]),
        "warp_instructions_per_launch": mean(m2["smsp__inst_executed.sum"]), "us_per_launch_under_ncu": mean(m2["gpu__time_duration.sum"]),
        "commit": commit, "how": "ncu (same metrics) -k regex:ipp_step_bulk -s 114 -c 6 of tools/predict_probe.py split: committing whole-batch "
                                 f"prediction steps, 65536 envs, mean of 6 launches; raw: profiles/{out}_traffic_predict_split.csv"}
    json.dump(tj, open(tj_path, "w"), indent=1)
rep2 = os.path.join(G, f"{tag}_predict_split.ncu-rep")
if os.path.exists(rep2):
    for tool, suffix, args in (("ncu_summary.py", "ncu_summary", []), ("ncu_lines.py", "hotlines", ["ipp_step_bulk_kernelILi1ELb0ELb0ELb0ELb1E", "60"])):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), rep2] + args, capture_output=True, text=True).stdout
        open(os.path.join(P, f"{out}_predict_split_{suffix}.txt"), "w").write(txt)
for kname, mangled in (("select", "mcts_select_kernelILi4E"), ("expand", "mcts_expand_kernel")):
    rep3 = os.path.join(G, f"{tag}_mcts_{kname}.ncu-rep")
    if os.path.exists(rep3):
        for tool, suffix, args in (("ncu_summary.py", "ncu_summary", []), ("ncu_lines.py", "hotlines", [mangled, "50"])):
            txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), rep3] + args, capture_output=True, text=True).stdout
            open(os.path.join(P, f"{out}_mcts_{kname}_{suffix}.txt"), "w").write(txt)

# full capture of the step kernel
rep = os.path.join(G, f"{tag}_bulk_step.ncu-rep")
kern = "ipp_step_bulk_kernelILi0ELb0ELb0ELb0ELb0E"
for tool, suffix, args in (("ncu_summary.py", "ncu_summary", []), ("ncu_lines.py", "hotlines", [kern, "90"]), ("ncu_groups_bulk.py", "groups", [kern, "65536"])):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), rep] + args, capture_output=True, text=True).stdout
    open(os.path.join(P, f"{out}_bulk_step_{suffix}.txt"), "w").write(txt)

# tree-search kernels
mm = defaultdict(lambda: defaultdict(list))
for r in rows(os.path.join(G, f"{tag}_mcts_launches.csv")):
    mm[r["Kernel Name"].split("(")[0]][r["Metric Name"]].append(val(r))
with open(os.path.join(P, f"{out}_mcts_launches.txt"), "w") as f:
    f.write("# per launch, 16 384 growing trees (tools/mcts_probe.py split 100 peaked: the 100 simulations of the first search; the rollout runs inside the select launch); ncu --clock-control none, mean (max) over the launches\n")
    for k, d in mm.items():
        f.write(f"{k}\n")
        for name, xs in d.items():
            f.write(f"    {name:60s} n={len(xs):3d} mean {sum(xs) / len(xs):14.1f} max {max(xs):14.1f}\n")

for src, dst in ((f"{tag}_bench.json", f"{out}_bench.json"), (f"{tag}_bench_k20.json", f"{out}_bench_steps20_warmup5.json"), (f"{tag}_bench_ref.json", f"{out}_bench_reference_arm.json"),
                 (f"{tag}_sanitizer_memcheck.log", f"{out}_sanitizer_memcheck.log"), (f"{tag}_sanitizer_racecheck.log", f"{out}_sanitizer_racecheck.log"),
                 (f"{tag}_sanitizer_synccheck.log", f"{out}_sanitizer_synccheck.log"), (f"{tag}_sanitizer_memcheck.out", f"{out}_sanitizer_workload.out"),
                 (f"{tag}_pytest.log", f"{out}_pytest_gpu.log"), ("c9_predict_probe.log", f"{out}_predict_probe.txt"), ("c6_probe.log", f"{out}_e2e_probe.txt"),
                 ("n2_bench.json", f"{out}_bench_2gpus.json"), ("n8_bench.json", f"{out}_bench_8gpus.json"), ("n8_ref.json", f"{out}_bench_reference_arm_8gpus.json"),
                 (f"{tag}_predict_probe_split.txt", f"{out}_predict_probe.txt"), (f"{tag}_predict_probe_super.txt", f"{out}_predict_probe_super.txt"),
                 (f"{tag}_e2e_ab.txt", f"{out}_e2e_ab.txt"), (f"{tag}_mcts_probe.txt", f"{out}_mcts_probe.txt"),
                 (f"{tag}_configs_throughput.jsonl", f"{out}_configs_throughput.jsonl")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
open(os.path.join(P, f"{out}_sass_opcodes.txt"), "w").write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_opcodes.py")], capture_output=True, text=True).stdout)
print(json.dumps(tj["kernels"]["ipp_step_bulk_kernel:super"], indent=1))
print(open(os.path.join(P, f"{out}_launches_by_kernel.txt")).read())
print(open(os.path.join(P, f"{out}_bulk_step_groups.txt")).read())
