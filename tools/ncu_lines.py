#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep 'ipp_step_kernelILi1ELi0E' [top_n]

ncu's CSV source page carries SASS only; the line mapping comes from ``nvdisasm --print-line-info`` of
the cubin inside ipp_rl_b200/csrc/libipp_b200.so (built with -lineinfo).  Instructions are matched by
order within the kernel (same cubin => same sequence).
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ipp_rl_b200", "csrc", "libipp_b200.so")


def sass_lines(mangled_substr):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=d, check=True, capture_output=True)
        txt = ""
        for cubin in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):  # one cubin per translation unit
            txt += subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=d, check=True, capture_output=True, text=True).stdout
    out, cur, active = [], None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            active = mangled_substr in m.group(1)
            continue
        if not active:
            continue
        if ln.strip().startswith(".section"):
            active = False
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    csv_txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(csv_txt.splitlines()))
    # split per kernel
    blocks, cur = [], None
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    demangled_key = re.sub(r"ILi(\d)ELi(\d)E.*", r"<(int)\1, (int)\2>", kern).replace("ipp_step_kernel", "ipp_step_kernel")
    blk = None
    for b in blocks:
        if demangled_key in b["name"] or kern in b["name"]:
            blk = b
            break
    if blk is None:
        blk = blocks[0]
    hdr = blk["rows"][0]
    ie, ns, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    insts = [(r[src].strip(), int(r[ie]), int(r[ns])) for r in blk["rows"][1:] if len(r) > ie and r[ie].isdigit()]
    lines = sass_lines(kern)
    print(f"kernel: {blk['name']}  ncu insts: {len(insts)}  nvdisasm insts: {len(lines)}")
    n = min(len(insts), len(lines))
    agg = {}
    tot_i = tot_s = 0
    for k in range(n):
        key = lines[k][0]
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += insts[k][1]
        a[1] += insts[k][2]
        a[2] += 1
        tot_i += insts[k][1]
        tot_s += insts[k][2]
    srcs = {}
    print(f"total warp-instructions {tot_i}, samples {tot_s}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if key:
            f = os.path.join(ROOT, "ipp_rl_b200", "csrc", key[0])
            if os.path.exists(f):
                if f not in srcs:
                    srcs[f] = open(f).read().splitlines()
                text = srcs[f][key[1] - 1].strip()[:90]
        print(f"{a[0]:>11} {100 * a[0] / tot_i:5.1f}%  samp {100 * a[1] / max(tot_s, 1):5.1f}%  sass {a[2]:>4}  {key}: {text}")


if __name__ == "__main__":
    main()
