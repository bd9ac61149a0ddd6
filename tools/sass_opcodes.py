#!/usr/bin/env python
"""SASS opcode histogram of the step kernels in ipp_rl_b200/csrc/libipp_b200.so (evidence that the staging is Blackwell-native:
UBLKCP = cp.async.bulk, SYNCS.* = mbarrier; LDGSTS = cp.async of the round-1 kernel).   python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ipp_rl_b200", "csrc", "libipp_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, per = None, collections.OrderedDict()
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        per[cur][m.group(1)] += 1


def show(name_sub, title):
    for fn, c in per.items():
        if name_sub in fn:
            tot = sum(c.values())
            print(f"== {title}: {fn}  ({tot} SASS instructions)")
            base = collections.Counter()
            for op, n in c.items():
                base[op.split(".")[0]] += n
            for op, n in base.most_common(24):
                print(f"   {n:5d} {100.0 * n / tot:5.1f}%  {op}")
            special = {op: n for op, n in c.items() if op.startswith(("UBLKCP", "SYNCS", "LDGSTS", "UTMA", "REDUX", "ELECT", "R2UR"))}
            print("   staging / sync opcodes:", dict(sorted(special.items())))
            return


show("ipp_step_bulk_kernelILi0ELb0ELb0ELb0", "bulk-copy step kernel, full step, trace-reduction reward (bench headline)")
show("ipp_step_bulk_kernelILi1ELb0ELb0ELb0", "bulk-copy step kernel, predict-only")
show("ipp_step_async_kernelILb0ELb0ELb0ELb1", "round-1 cp.async step kernel, tiled layout")
show("mcts_expand_kernel", "MCTS expand + backup")
show("mcts_select_kernel", "MCTS select")
tot = collections.Counter()
for c in per.values():
    for op, n in c.items():
        if op.startswith(("UBLKCP", "SYNCS", "LDGSTS", "UTMA", "REDUX")):
            tot[op] += n
print("== whole library, staging / sync / reduction opcodes:", dict(sorted(tot.items())))
