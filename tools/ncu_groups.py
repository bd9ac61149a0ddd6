#!/usr/bin/env python
"""Group the per-line executed-instruction counts of an ncu report by code region of the async step kernel.
    python tools/ncu_groups.py gpurun_out/prof.ncu-rep [n_envs=16384] [kernel-substring]"""
import os, re, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
n_env = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
kern = sys.argv[3] if len(sys.argv) > 3 else "ipp_step_async_kernel"
txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, kern, "5000"], capture_output=True, text=True).stdout
src = open(os.path.join(ROOT, "ipp_rl_b200/csrc/step_async.cuh")).read().splitlines()
qm = open(os.path.join(ROOT, "ipp_rl_b200/csrc/quad_math.cuh")).read().splitlines()
def find(lines, pat):
    for i, l in enumerate(lines):
        if pat in l:
            return i + 1
    return None
marks = sorted(m for m in [(find(src, 'auto fill = [&]'), 'fill'), (find(src, '// ---- prologue: one ticket chunk'), 'prologue'),
    (find(src, 'while (true) {'), 'loop head/ticket/wait'), (find(src, '// (B) fuse the env'), 'env setup'),
    (find(src, 'for (int q = lane; q < nq; q += 32)'), 'quad: index+LDS'), (find(src, '// ---- measurement ---'), 'quad: noise+z'),
    (find(src, 'if (EXTRAS && p.z_out != nullptr)'), 'quad: z_out'), (find(src, '// ---- fusion + reward'), 'quad: mask+kalman call+stores'),
    (find(src, '// per-env information gain'), 'reduce/reward'), (find(src, '// (C) refill slot s'), 'refill/ticket rotate')] if m[0])
qmarks = sorted(m for m in [(find(qm, 'void philox4x32_10('), 'philox'), (find(qm, 'float u01('), 'box_muller'), (find(qm, 'void draw_normals('), 'draw_normals glue'),
    (find(qm, 'int fdiv('), 'fdiv'), (find(qm, 'struct Geom'), 'decode/geom'), (find(qm, 'float fast_sqrt('), 'cost'), (find(qm, 'constexpr int TAPS_FAST'), 'tap build'),
    (find(qm, 'struct TapView'), 'downsample'), (find(qm, 'float bernoulli_entropy('), 'bern'), (find(qm, 'struct FuseCtx'), 'kalman_quad')] if m[0])
def grp(marks, l, default):
    g = default
    for ln, name in marks:
        if l >= ln:
            g = name
    return g
g = defaultdict(lambda: [0, 0, 0.0])
tot = 0
for ln in txt.splitlines():
    m = re.match(r"\s*(\d+)\s+([\d.]+)%\s+samp\s+([\d.]+)%\s+sass\s+(\d+)\s+\('([^']+)', (\d+)\)", ln)
    if not m:
        continue
    d, s, n, f, l = int(m.group(1)), float(m.group(3)), int(m.group(4)), m.group(5), int(m.group(6))
    k = grp(marks, l, 'kernel head') if f == 'step_async.cuh' else (grp(qmarks, l, 'qm head') if f == 'quad_math.cuh' else f)
    g[k][0] += d; g[k][1] += n; g[k][2] += s; tot += d
print(txt.splitlines()[0] if txt else "")
for k, v in sorted(g.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:34s} dyn {v[0]/n_env:8.1f}/env {100*v[0]/max(tot,1):5.1f}%  static {v[1]:5d}  samples {v[2]:5.1f}%")
print('total warp-instructions per env-step:', round(tot / n_env, 1))
