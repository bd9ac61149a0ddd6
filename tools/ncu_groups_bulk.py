#!/usr/bin/env python
"""Executed warp-instructions of the bulk-copy step kernel (csrc/step_bulk.cuh) grouped by code region, per env-step.
    python tools/ncu_groups_bulk.py gpurun_out/prof.ncu-rep 'ipp_step_bulk_kernelILi0ELb0ELb0ELb0' [n_envs=65536]"""
import sys, csv, subprocess, re, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as nl
rep, kern = sys.argv[1], sys.argv[2]
csv_txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csv_txt.splitlines()))
blocks, cur = [], None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None: cur["rows"].append(r)
blk = blocks[0]
hdr = blk["rows"][0]
ie, src = hdr.index("Instructions Executed"), hdr.index("Source")
insts = [(r[src].strip(), int(r[ie])) for r in blk["rows"][1:] if len(r) > ie and r[ie].isdigit()]
lines = nl.sass_lines(kern)
B = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
srcl=open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'ipp_rl_b200', 'csrc', 'step_bulk.cuh')).read().split('\n')
def find(s, last=False): return [i+1 for i,l in enumerate(srcl) if s in l][-1 if last else 0]
def find_all(s): return [i+1 for i,l in enumerate(srcl) if s in l]
# the quad loop exists twice (SPLIT first, super-tiles second): this tool analyses the super-tile instantiation, i.e. the LAST copy
L_loop=find('for (int q = lane; q < nq; q += 32, ++it)', True); L_end=find('// per-env information gain'); L_col=find('const uint32_t magic_x = (uint32_t)pc.x;'); L_B=find('// (B) fuse the oldest staged env')
L_tf0=find('auto try_fill'); L_tf1=find('// The first pass of the loop'); L_meas=find('// ---- measurement ---', True); L_fus=find('// ---- fusion + reward', True)
L_lds=set(find_all('asm volatile("ld.shared.')); L_gts=set(range(find('struct GtSuperShared'), find('struct GtSuperShared') + 8))
L_fillasm=set(find_all('mbarrier.') + find_all('cp.async.bulk.shared::cluster'))
L_plan0=find('__device__ __forceinline__ void bulk_plan_env'); L_plan1=find('// MODE: MODE_KALMAN (full step)')
L_C=find('// (C) release the footprint'); L_D=find('// (D) plan the chunk')
qml=open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'ipp_rl_b200', 'csrc', 'quad_math.cuh')).read().split('\n')
def qfind(t): return [i+1 for i,l in enumerate(qml) if t in l][0]
Q_rng0=qfind('counter-based RNG'); Q_rng1=qfind('floor(n / d) for 0 <= n'); Q_geo0=qfind('per-job geometry'); Q_tap0=qfind('One axis of cv2 INTER_AREA')
Q_ds0=qfind('D[pr, pc] of the down-sampled measurement'); Q_kal0=qfind('Per-quad fusion.')
def group(key):
    if key is None: return 'none'
    f,l = key
    if f=='step_bulk.cuh':
        if l in L_lds: return 'loop: lds asm'
        if l in L_gts: return 'clipped: GtSuperShared::at'
        if l in L_fillasm or L_tf0<=l<L_tf1: return 'fill'
        if L_plan0<=l<L_plan1: return 'plan_env'
        if L_loop<=l<L_meas: return 'loop: index + belief'
        if L_meas<=l<L_fus: return 'loop: measurement'
        if L_fus<=l<L_end: return 'loop: mask + stores'
        if L_col<=l<L_loop: return 'per-env: column setup (+ dead SPLIT loop lines)'
        if L_B<=l<L_col: return 'per-env: plan unpack, wait, weights'
        if L_end<=l<L_C: return 'per-env: reduce + reward'
        if L_C<=l<L_D: return 'per-env: release'
        if l>=L_D: return 'per-env: ticket/plan chunk'
        return 'kernel prologue/other %d'%0
    if f=='quad_math.cuh':
        if Q_rng0<=l<Q_rng1: return 'philox + box-muller + draw'
        if l>=Q_kal0-3: return 'kalman_quad'
        if Q_rng1<=l<Q_tap0: return 'plan_env'
        if Q_tap0<=l<Q_ds0: return 'clipped: tap table build'
        if Q_ds0<=l<Q_kal0-3: return 'clipped: downsample table paths'
        return 'quad_math other'
    return f
agg={}
for k in range(min(len(insts),len(lines))):
    g=group(lines[k][0]); agg[g]=agg.get(g,0)+insts[k][1]
tot=sum(agg.values())
for g,c in sorted(agg.items(), key=lambda kv:-kv[1]): print(f"{c/B:8.1f} /env {100*c/tot:5.1f}%  {g}")
print('total warp-instructions per env-step: %.1f' % (tot / B))
