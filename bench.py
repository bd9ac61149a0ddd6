#!/usr/bin/env python
"""bench.py — env steps/s of the batched IPP environment engine on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[2], "C3"): per GPU 65 536 env instances, 200x200 grid at 1 m/cell,
FoV 60x60 deg, 3-altitude action set {8, 14, 20} m (footprints 9x9 / 17x17 / 23x23 cells, resolution
factor 1 / 2 / 2), uniform-random valid action per env per step, Gaussian sensor noise from the
device Philox stream, per-cell Kalman fusion, entropy-reduction reward (trace-reduction timed too).
A "step" = one pass of the fused step kernel over the whole env batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1: launched under torchrun, one rank per GPU, env-batch sharded (weak scaling: 65 536 envs per
GPU), no data-path collective in `value`; `e2e` adds the per-step NCCL all-gather of the rewards that
a trainer's experience buffer needs.

`e2e.value` is the synchronous call (host ids in, rewards back on the host before the next step is
submitted); `e2e.pipelined_value` the same per-step transfers over two slots (ipp_step_submit /
ipp_step_wait at N = 1; exchange + D2H on a side stream at N > 1).  `mcts_rollouts` is the secondary
leg at BASELINE.json configs[3]'s per-GPU size (16 384 trees).

`--impl reference` times the CPU port of the reference's algorithm (oracle/; the reference itself
is pure Python and cannot travel to the GPU box) on all host cores, on a bounded sample.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(x_dim=200, y_dim=200, resolution=1.0, min_altitude=8.0, max_altitude=20.0, altitude_spacing=6.0,
                angle_x=60.0, angle_y=60.0, coeff_a=0.05, coeff_b=0.2, max_v=2.0, max_a=2.0)
PRIOR_MEAN, PRIOR_VAR = 0.5, 1.82


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def footprint_cells(ids, X, Y, radii):
    """Covered cells per job for action ids (planning/common/actions.py id layout), clipped."""
    N = X * Y
    lvl = ids // N
    i = ids - lvl * N
    col, row = i // X, i % X
    r = np.asarray(radii)[lvl]
    nx = np.minimum(col + r, X - 1) - np.maximum(col - r, 0) + 1
    ny = np.minimum(row + r, Y - 1) - np.maximum(row - r, 0) + 1
    return nx.astype(np.int64) * ny


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference algorithm on all host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import cpu_baseline

    res = cpu_baseline.run(WORKLOAD, steps=args.steps, warmup=args.warmup, envs=args.cpu_envs, reward_mode=1)
    line = {
        "metric": "env_steps_per_sec", "value": res["steps_per_sec"], "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": "C3: batch 65536 envs/GPU, 200x200 grid, 3-altitude action set {8,14,20} m, Kalman fusion, "
                               "entropy-reduction reward (Gaussian-entropy extension; trace-reduction timed alongside)",
                   "batch_per_step_sample": res["envs"], "grid": [200, 200], "noise": "Philox4x32-10 (same stream definition)",
                   "note": "CPU port of the reference algorithm (oracle/ipp_oracle.c, fp64, OpenMP over envs); each step = one pass over a "
                           "bounded env sample of the same workload"},
        "cpu_baseline": {"value": res["steps_per_sec"], "unit": "env-steps/s", "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"]},
        "e2e": {"value": res["steps_per_sec"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch

    from ipp_rl_b200 import BatchedEngine, EngineConfig, _capi as capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    B, K, W = args.batch, args.steps, args.warmup
    layout = capi.LAYOUT_NAMES[args.layout]
    stream = torch.cuda.Stream(device=local_rank)
    cfg = EngineConfig(batch=B, layout=layout, device=local_rank, seed=20260925, env_id_offset=rank * B, stream=stream.cuda_stream,
                       **WORKLOAD)
    eng = BatchedEngine(cfg)
    eng.reset(PRIOR_MEAN, PRIOR_VAR)
    eng.synth_ground_truth(seed=1000)
    info = eng.info
    radii = [info.radius_x[k] for k in range(info.num_altitude_levels)]
    X, Y = cfg.x_dim, cfg.y_dim

    # synthetic action streams: uniform-random valid (cell, altitude) per env per step
    rng = np.random.RandomState(777 + rank)
    total = W + K
    POOL = min(total, 64)  # distinct action sets, cycled (each env still sees a different action every step)
    ids_host = rng.randint(0, eng.num_actions, size=(POOL, B)).astype(np.int32)
    ids_e2e = rng.randint(0, eng.num_actions, size=(POOL, B)).astype(np.int32)
    cells_pool = footprint_cells(ids_host.astype(np.int64), X, Y, radii).sum(axis=1)  # (POOL,)
    cells_timed_total = float(sum(cells_pool[t % POOL] for t in range(W, W + K)))
    alg_bytes_per_launch = (20.0 * cells_timed_total + 16.0 * B * K) / K

    reward_modes = {"entropy": capi.REWARD_GAUSS_ENTROPY, "trace": capi.REWARD_TRACE}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        ids_dev = torch.from_numpy(ids_host).cuda(non_blocking=False)
        reward_dev = torch.empty(B, dtype=torch.float32, device="cuda")
        gathered = torch.empty(world * B, dtype=torch.float32, device="cuda") if dist is not None else None
        torch.cuda.synchronize()

        def device_loop(mode, lo, hi):
            for t in range(lo, hi):
                eng.step_device(action_ids_ptr=ids_dev[t % POOL].data_ptr(), reward_ptr=reward_dev.data_ptr(), reward_mode=mode)

        results = {}
        clocks = None
        for name, mode in reward_modes.items():
            eng.reset(PRIOR_MEAN, PRIOR_VAR)
            device_loop(mode, 0, W)
            barrier()
            sampler = ClockSampler(local_rank) if (name == "entropy" and rank == 0) else None
            if sampler:
                sampler.start()
            launches0 = eng.launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            device_loop(mode, W, W + K)
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
            if sampler:
                clocks = sampler.stop()
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            results[name] = dict(ms_total=float(t.item()), launches=eng.launches - launches0)
            eng.sync()

        # ---- e2e: host buffers through the public step call, H2D + kernel + (gather) + D2H per step
        ids_pinned = torch.from_numpy(ids_e2e).pin_memory()
        out_pinned = torch.empty(world * B if dist is not None else B, dtype=torch.float32).pin_memory()
        ids_np, out_np = ids_pinned.numpy(), out_pinned.numpy()
        id_rows = [ids_np[k] for k in range(POOL)]  # stable array objects (the engine caches their ctypes pointers)
        eng.reset(PRIOR_MEAN, PRIOR_VAR)
        if args.zero_copy is not None:
            eng.set_zero_copy(rewards="r" in args.zero_copy, ids="i" in args.zero_copy)
        zc0 = eng.zero_copy_steps

        def e2e_step(t):
            if dist is None:
                eng.step(id_rows[t % POOL], reward_mode=capi.REWARD_GAUSS_ENTROPY, out=out_np)
            else:
                staged = ids_pinned[t % POOL].cuda(non_blocking=True)
                eng.step_device(action_ids_ptr=staged.data_ptr(), reward_ptr=reward_dev.data_ptr(), reward_mode=capi.REWARD_GAUSS_ENTROPY)
                dist.all_gather_into_tensor(gathered, reward_dev)
                out_pinned.copy_(gathered, non_blocking=True)
                stream.synchronize()

        KE = min(K, args.e2e_steps)  # the e2e leg is host-latency bound; keep the default run short
        for t in range(W):
            e2e_step(t)
        barrier()
        launches_e2e0 = eng.launches
        t0 = time.perf_counter()
        for t in range(W, W + KE):
            e2e_step(t)
        barrier()
        e2e_s = time.perf_counter() - t0
        tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
        launches_e2e = eng.launches - launches_e2e0
        zero_copy_steps = eng.zero_copy_steps - zc0
        assert np.isfinite(out_np).all()

        # ---- e2e, pipelined (N = 1): the same per-step transfers through ipp_step_submit / ipp_step_wait, two slots — the host
        #      uploads step t+1 while step t computes and waits for step t-1 before reusing its buffers (reported beside e2e.value)
        pipe_s = None
        if dist is None:
            outs2 = [torch.empty(B, dtype=torch.float32).pin_memory().numpy() for _ in range(2)]

            def pipe_loop(lo, hi):
                for t in range(lo, hi):
                    sl = t & 1
                    eng.step_wait(sl)
                    eng.step_submit(sl, id_rows[t % POOL], outs2[sl], reward_mode=capi.REWARD_GAUSS_ENTROPY)
                eng.step_wait(0)
                eng.step_wait(1)

            pipe_loop(0, W)
            barrier()
            t0 = time.perf_counter()
            pipe_loop(W, W + KE)
            barrier()
            pipe_s = time.perf_counter() - t0
            assert np.isfinite(outs2[0]).all() and np.isfinite(outs2[1]).all()
        else:
            # N > 1: same idea with the exchange in the pipeline — the rewards all-gather (NCCL) and the D2H of the gathered vector
            # of step t run on a side stream under the kernel of step t+1; the host waits for step t-1 before reusing its buffers
            comm = torch.cuda.Stream()
            ids_dev2 = [torch.empty(B, dtype=torch.int32, device="cuda") for _ in range(2)]
            rew_dev2 = [torch.empty(B, dtype=torch.float32, device="cuda") for _ in range(2)]
            gath2 = [torch.empty(world * B, dtype=torch.float32, device="cuda") for _ in range(2)]
            outs2 = [torch.empty(world * B, dtype=torch.float32).pin_memory() for _ in range(2)]
            ev_k = [torch.cuda.Event() for _ in range(2)]
            ev_d = [torch.cuda.Event() for _ in range(2)]
            used = [False, False]

            def pipe_loop(lo, hi):
                for t in range(lo, hi):
                    sl = t & 1
                    if used[sl]:
                        ev_d[sl].synchronize()
                    ids_dev2[sl].copy_(ids_pinned[t % POOL], non_blocking=True)
                    eng.step_device(action_ids_ptr=ids_dev2[sl].data_ptr(), reward_ptr=rew_dev2[sl].data_ptr(), reward_mode=capi.REWARD_GAUSS_ENTROPY)
                    ev_k[sl].record(stream)
                    with torch.cuda.stream(comm):
                        comm.wait_event(ev_k[sl])
                        dist.all_gather_into_tensor(gath2[sl], rew_dev2[sl])
                        outs2[sl].copy_(gath2[sl], non_blocking=True)
                        ev_d[sl].record(comm)
                    used[sl] = True
                for sl in (0, 1):
                    if used[sl]:
                        ev_d[sl].synchronize()

            pipe_loop(0, W)
            barrier()
            t0 = time.perf_counter()
            pipe_loop(W, W + KE)
            barrier()
            pipe_s = time.perf_counter() - t0
            tp = torch.tensor([pipe_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
            pipe_s = float(tp.item())
            assert np.isfinite(outs2[0].numpy()).all() and np.isfinite(outs2[1].numpy()).all()

    # ---- secondary leg (BASELINE.json configs[3], "C4"): mcts_zero rollouts on the same beliefs.  Lock-step search over
    #      `--mcts-trees` envs, `--mcts-sims` simulations, episode_horizon 5, max_valid_action_distance 11.5 m, uniform
    #      priors / zero values (no network: the policy/value net is outside this library).  Reported beside the headline.
    mcts_res = None
    if args.mcts_trees > 0:
        from ipp_rl_b200.planning.mcts_zero import BatchedMCTS

        Tm, Sm = min(args.mcts_trees, B), args.mcts_sims
        hyper = dict(puct_init=15.0, puct_base=10000, num_mcts_simulations=Sm, gamma=1.0, dirichlet_alpha=0.3, dirichlet_eps=0.25,
                     forced_playout_factor=2.0, max_valid_action_distance=11.5)
        meta = dict(episode_horizon=5, scenario_info=None)
        budgets = np.full(Tm, 150.0, np.float32)
        with torch.cuda.stream(stream):
            with BatchedMCTS(eng, hyper, meta, n_trees=Tm) as mcts:
                # pass 1 (untimed, host-synchronous): count the prediction steps of the search (it is deterministic)
                mcts.begin(budgets)
                edges = 0
                cells_by_level = np.array([(2 * r + 1) ** 2 for r in radii], np.float64)
                cells = 0.0
                for i in range(Sm):
                    leaf = mcts.simulate(lambda lf: (None, None))
                    edges += int(leaf.path_len.sum())
                path_ids = None
                # pass 2 (timed): device-only loop, CUDA events on the engine stream
                mcts.begin(budgets)
                l0 = mcts.launches
                barrier()
                m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                m0.record(stream)
                for i in range(Sm):
                    mcts.simulate(None)
                m1.record(stream)
                barrier()
                ms_m = m0.elapsed_time(m1)
                tm = torch.tensor([ms_m], dtype=torch.float64, device="cuda")
                if dist is not None:
                    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                ms_m = float(tm.item())
                st = mcts.root_stats()
                assert np.all(st["Ns"] == Sm - 1)
                mcts_res = {"trees_per_gpu": Tm, "simulations": Sm, "episode_horizon": 5, "window_slots": mcts.window_slots,
                            "tree_simulations_per_sec": world * Tm * Sm / (ms_m * 1e-3),
                            "prediction_steps_per_sec": world * edges / (ms_m * 1e-3), "prediction_steps": edges,
                            "ms_per_lockstep_simulation": ms_m / Sm, "gpu_launches": int(mcts.launches - l0),
                            "tree_bytes_per_gpu": int(mcts.info.device_bytes),
                            "evaluator": "uniform priors, zero values (network outside this library)"}

    # ---- CPU baseline (rank 0, N == 1 only): oracle port on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import cpu_baseline

            r = cpu_baseline.run(WORKLOAD, steps=3, warmup=1, envs=args.cpu_envs, reward_mode=1, min_seconds=1.0)
            cpu = {"value": r["steps_per_sec"], "unit": "env-steps/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        except Exception as exc:  # report, never hide
            cpu = {"value": None, "unit": "env-steps/s", "cores": 0, "kind": "port", "sample": f"failed: {exc!r}"}

    if rank == 0:
        peak, peak_src = load_peaks()
        ms = results["entropy"]["ms_total"]
        per_launch_ms = ms / K
        achieved = alg_bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
        ms_tr = results["trace"]["ms_total"]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_per_launch.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(f"{args.layout}_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "env_steps_per_sec", "value": world * B * K / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": per_launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3: batch 65536 envs/GPU, 200x200 grid, 3-altitude action set {8,14,20} m, Kalman fusion, "
                                   "entropy-reduction reward (Gaussian-entropy extension; trace-reduction timed alongside)",
                       "batch_per_gpu": B, "grid": [Y, X], "layout": args.layout, "noise": "device Philox4x32-10",
                       "cache": "working set 31.5 GB/GPU >> 126 MB L2 (inputs larger than L2, no flush needed)",
                       "parallelism": f"env-batch sharded x{world}"},
            "value_trace_reduction": world * B * K / (ms_tr * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                         "mean_cells_per_env_step": cells_timed_total / (B * K), "kernel": "ipp_step_async_kernel" if eng.step_path == "async" else "ipp_step_kernel<MV, KALMAN>", "step_path": eng.step_path},
            "e2e": {"value": world * B * KE / e2e_s, "steps": KE, "unit": "env-steps/s", "h2d_bytes_per_step": 4 * B * world,
                    "d2h_bytes_per_step": 4 * B * world * (world if dist is not None else 1),
                    "pipelined_value": (world * B * KE / pipe_s) if pipe_s else None,
                    "pipelined_path": ("ipp_step_submit / ipp_step_wait, 2 slots: same per-step H2D ids + rewards to pinned host memory, upload of "
                                       "step t+1 under the kernel of step t") if dist is None else
                                      ("2 slots: H2D ids -> ipp_step_device; NCCL all_gather(rewards) + D2H of step t on a side stream under the "
                                       "kernel of step t+1"),
                    "path": ("BatchedEngine.step (ipp_step: pinned host ids -> H2D -> fused kernel -> "
                             + ("rewards written by the kernel into the caller's pinned buffer [zero-copy D2H, 4 B/env over PCIe])"
                                if zero_copy_steps > 0 else "D2H rewards)")) if dist is None else
                            "pinned host ids -> H2D -> ipp_step_device -> NCCL all_gather(rewards) -> D2H"},
            "gpu_launches": int(results["entropy"]["launches"]),
            "gpu_launches_e2e": int(launches_e2e),
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if mcts_res is not None:
            line["mcts_rollouts"] = mcts_res
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--layout", default=os.environ.get("IPP_LAYOUT", "tiled"), choices=["planes", "mv", "tiled", "super"])
    ap.add_argument("--cpu-envs", type=int, default=4096, help="env sample of the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=500, help="steps of the host-buffer (e2e) leg (<= --steps)")
    ap.add_argument("--zero-copy", default=None, choices=["", "r", "i", "ri"],
                    help="e2e leg: host buffers the kernel accesses in place (r = rewards, i = action ids); default = the engine's (r)")
    ap.add_argument("--mcts-trees", type=int, default=16384, help="trees of the secondary mcts_zero rollout leg = BASELINE.json C4 per-GPU share (0 = skip)")
    ap.add_argument("--mcts-sims", type=int, default=32)
    ap.add_argument("--max-altitude", type=float, default=None, help="experiments only: override the top altitude of the action set")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.max_altitude is not None:
        WORKLOAD["max_altitude"] = args.max_altitude
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
