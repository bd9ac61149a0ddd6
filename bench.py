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
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu"

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
        sm, load, mx, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            try:
                if len(r) > 7 and float(r[7]) >= 10.0:
                    load.append(sm[-1])
            except ValueError:
                pass
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        use = load if load else sm  # median over the samples that saw the GPU busy (the legs alternate with host-side set-up)
        return {"sm_mhz": float(np.median(use)) if use else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_under_load": len(load), "reasons": sorted(reasons)}


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

    res = cpu_baseline.run(WORKLOAD, steps=args.steps, warmup=args.warmup, envs=args.cpu_envs, reward_mode=0)
    line = {
        "metric": "env_steps_per_sec", "value": res["steps_per_sec"], "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": "C3: batch 65536 envs/GPU, 200x200 grid, 3-altitude action set {8,14,20} m, Kalman fusion, information-gain reward: "
                               "trace (variance) reduction = the reference's reward, parity-pinned [headline]",
                   "reward_mode": "trace_reduction",
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
class SharedRewards:
    """One pinned host segment shared by all ranks of the node (POSIX shared memory + cudaHostRegister): every rank's step
    kernel writes the rewards of its env slice straight into it (4 B per env over its own PCIe link) and bumps its step
    counter; the learner (rank 0) sees the whole batch's rewards without any collective on the step path.  Two slots, so a
    rank may run one step ahead of the learner."""

    SLOTS = 2

    def __init__(self, rank, world, per_rank, tag):
        from multiprocessing import shared_memory

        import torch

        self.rank, self.world, self.n = rank, world, per_rank
        self.bytes = self.SLOTS * world * per_rank * 4 + 4096
        name = f"ipp_b200_rewards_{tag}"
        if rank == 0:
            try:
                shared_memory.SharedMemory(name=name).unlink()
            except Exception:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=self.bytes)
            self.shm.buf[: self.bytes] = bytes(self.bytes)
        import torch.distributed as dist

        dist.barrier()
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name)
            try:  # attached, not owned: keep Python's resource tracker from unlinking (and warning about) the creator's segment
                from multiprocessing import resource_tracker

                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        arr = np.ndarray((self.bytes,), dtype=np.uint8, buffer=self.shm.buf)
        self.rewards = arr[: self.SLOTS * world * per_rank * 4].view(np.float32).reshape(self.SLOTS, world, per_rank)
        self.flags = arr[self.SLOTS * world * per_rank * 4 :].view(np.int64)[: 64]  # flags[r] = steps rank r has completed, flags[32] = consumed by the learner
        rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, self.bytes, 1 | 2)  # portable | mapped
        self.registered = (getattr(rc, "value", rc) == 0) or ("success" in str(rc).lower())
        self._ptr = arr.ctypes.data
        dist.barrier()

    def my_slice(self, step):
        return self.rewards[step % self.SLOTS, self.rank]

    def publish(self, step):
        self.flags[self.rank] = step + 1

    def _spin(self, cond, what):
        t0 = time.perf_counter()
        n = 0
        while not cond():
            n += 1
            if (n & 0xFFFF) == 0 and time.perf_counter() - t0 > 120.0:  # a rank died: fail instead of hanging the node
                raise RuntimeError(f"SharedRewards: timed out waiting for {what} (flags {self.flags[: self.world].tolist()}, consumed {int(self.flags[32])})")

    def learner_collect(self, step):
        """rank 0: wait until every rank has published `step`, return the (world * per_rank,) view."""
        self._spin(lambda: int(self.flags[: self.world].min()) >= step + 1, f"step {step} of every rank")
        out = self.rewards[step % self.SLOTS].reshape(-1)
        self.flags[32] = step + 1
        return out

    def wait_for_slot(self, step):
        """any rank: the slot of `step` was last used by step - SLOTS; the learner must have consumed that one."""
        self._spin(lambda: int(self.flags[32]) >= step + 1 - self.SLOTS, f"the learner to consume step {step - self.SLOTS}")

    def close(self):
        import torch

        try:
            torch.cuda.cudart().cudaHostUnregister(self._ptr)
        except Exception:
            pass
        self.rewards = self.flags = None
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:
            pass


def run_ours(args, rank, world, local_rank):
    import torch

    from ipp_rl_b200 import BatchedEngine, EngineConfig, _capi as capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    nvtx = torch.cuda.nvtx
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
    alg_bytes_per_launch = (20.0 * cells_timed_total + 16.0 * B * K) / K          # full step: 20 B per covered cell + 16 B per env
    alg_bytes_predict = (8.0 * cells_timed_total + 16.0 * B * K) / K              # covariance-only step: read + write the variance

    # headline = the reference-pinned reward (trace reduction, planning/common/rewards.py:15-31); the Gaussian-entropy
    # extension (BASELINE.json's "entropy-reduction" wording, parity unpinned) is timed and reported beside it
    reward_modes = {"trace_reduction": capi.REWARD_TRACE, "gauss_entropy": capi.REWARD_GAUSS_ENTROPY}
    HEAD = "trace_reduction"

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()

    with torch.cuda.stream(stream):
        ids_dev = torch.from_numpy(ids_host).cuda(non_blocking=False)
        reward_dev = torch.empty(B, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()

        def timed_device_loop(fn, tag, eng=eng):
            """W warm-up launches, then K launches between two CUDA events on the engine's stream; max over ranks."""
            eng.reset(PRIOR_MEAN, PRIOR_VAR)
            for t in range(W):
                fn(t)
            barrier()
            launches0 = eng.launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nvtx.range_push(tag)
            e0.record(stream)
            for t in range(W, W + K):
                fn(t)
            e1.record(stream)
            barrier()
            nvtx.range_pop()
            ms = max_over_ranks(e0.elapsed_time(e1))
            eng.sync()
            return dict(ms_total=ms, launches=eng.launches - launches0)

        results = {}
        for name, mode in reward_modes.items():
            results[name] = timed_device_loop(
                lambda t, mode=mode: eng.step_device(action_ids_ptr=ids_dev[t % POOL].data_ptr(), reward_ptr=reward_dev.data_ptr(), reward_mode=mode),
                f"step[{name}]")
        # predict-only legs: the covariance-only step the planners' rollouts run (simulate_prediction_step,
        # planning/common/optimization.py:14-30), whole batch — the persistent kernel's MODE_PREDICT.  Committing, and the
        # evaluate-only form (IPP_FLAG_NO_COMMIT: rewards of candidate actions, nothing written — what greedy_search and the
        # tree search ask for).  Timed on the step engine's layout here and on the search engine's (below).
        def predict_legs(e, suffix):
            pl0 = e.path_launches("async")
            results["predict" + suffix] = timed_device_loop(
                lambda t: e.predict_device(B, action_ids_ptr=ids_dev[t % POOL].data_ptr(), reward_ptr=reward_dev.data_ptr(), commit=True,
                                           reward_mode=capi.REWARD_TRACE),
                "predict" + suffix, eng=e)
            persistent = e.path_launches("async") - pl0 >= K
            results["predict_eval" + suffix] = timed_device_loop(
                lambda t: e.predict_device(B, action_ids_ptr=ids_dev[t % POOL].data_ptr(), reward_ptr=reward_dev.data_ptr(), commit=False,
                                           reward_mode=capi.REWARD_TRACE),
                "predict[no commit]" + suffix, eng=e)
            return persistent

        predict_persistent = predict_legs(eng, "@step_layout")

        # the clock sampler polls nvidia-smi, which takes driver locks for ~1 ms at a time: stop it before the latency-bound legs
        clocks = sampler.stop() if sampler else None

        # ---- e2e: host buffers through the public step call, H2D + kernel + rewards back on the host, every step -----------
        ids_pinned = torch.from_numpy(ids_e2e).pin_memory()
        ids_np = ids_pinned.numpy()
        id_rows = [ids_np[k] for k in range(POOL)]  # stable array objects (the engine caches their ctypes pointers)
        shared = None
        if dist is None:
            out_np = torch.empty(B, dtype=torch.float32).pin_memory().numpy()
        else:
            shared = SharedRewards(rank, world, B, tag=os.environ.get("MASTER_PORT", "0"))
        eng.reset(PRIOR_MEAN, PRIOR_VAR)
        if args.zero_copy is not None:
            eng.set_zero_copy(rewards="r" in args.zero_copy, ids="i" in args.zero_copy, ids_fetch="f" in args.zero_copy)
        zc0, if0 = eng.zero_copy_steps, eng.ids_fetch_steps
        e2e_mode = reward_modes[HEAD]
        checksum = [0.0]

        def e2e_step(t, base):
            if shared is None:
                eng.step(id_rows[t % POOL], reward_mode=e2e_mode, out=out_np)
            else:  # every rank: ids up, kernel, rewards straight into the node's shared pinned segment; the learner reads them all
                s = base + t
                shared.wait_for_slot(s)
                eng.step(id_rows[t % POOL], reward_mode=e2e_mode, out=shared.my_slice(s))
                shared.publish(s)
                if rank == 0:
                    checksum[0] = float(shared.learner_collect(s)[:: 4096].sum())

        # The e2e legs are host-latency bound (~25 us of copy / launch / wake-up around a ~120 us kernel) and any other process that
        # polls the driver (nvidia-smi, ~1 ms per poll) shows up in them: time a few hundred steps whatever K is, so that one such
        # stall is not a third of the measurement.
        KE = args.e2e_steps
        for t in range(W):
            e2e_step(t, 0)
        barrier()
        launches_e2e0 = eng.launches
        nvtx.range_push("e2e[sync]")
        t0 = time.perf_counter()
        for t in range(W, W + KE):
            e2e_step(t, 0)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        nvtx.range_pop()
        launches_e2e = eng.launches - launches_e2e0
        zero_copy_steps = eng.zero_copy_steps - zc0
        ids_fetch_steps = eng.ids_fetch_steps - if0
        if shared is None:
            assert np.isfinite(out_np).all()
        else:
            assert np.isfinite(shared.rewards).all()

        # ---- e2e, pipelined: the same per-step transfers through ipp_step_submit / ipp_step_wait, two slots — the host uploads
        #      step t+1 while step t computes and waits for step t-1 before reusing its buffers (reported beside e2e.value)
        if shared is None:
            outs2 = [torch.empty(B, dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
            slot_out = lambda s: outs2[s & 1]  # noqa: E731
        else:
            slot_out = lambda s: shared.my_slice(s)  # noqa: E731
        base2 = W + KE  # continues the shared segment's step numbering

        def pipe_loop(lo, hi):
            for t in range(lo, hi):
                sl = t & 1
                eng.step_wait(sl)
                if shared is not None:
                    s = base2 + t
                    if t - 2 >= lo:  # step t-2 (same slot) is complete: publish it, the learner collects
                        shared.publish(s - 2)
                        if rank == 0:
                            checksum[0] = float(shared.learner_collect(s - 2)[:: 4096].sum())
                    shared.wait_for_slot(s)
                eng.step_submit(sl, id_rows[t % POOL], slot_out(base2 + t), reward_mode=e2e_mode)
            eng.step_wait(0)
            eng.step_wait(1)
            if shared is not None:
                for s in range(max(lo, hi - 2), hi):
                    shared.publish(base2 + s)
                    if rank == 0:
                        checksum[0] = float(shared.learner_collect(base2 + s)[:: 4096].sum())

        if shared is not None:  # restart the segment's counters at a common point
            barrier()
            shared.flags[rank] = base2
            if rank == 0:
                shared.flags[32] = base2
            barrier()
        pipe_loop(0, W)
        base2 += W if shared is not None else 0
        barrier()
        nvtx.range_push("e2e[pipelined]")
        t0 = time.perf_counter()
        pipe_loop(0, KE) if shared is not None else pipe_loop(W, W + KE)
        barrier()
        pipe_s = max_over_ranks(time.perf_counter() - t0)
        nvtx.range_pop()
        if shared is not None:
            shared.close()

        # ---- search engine: the covariance-only paths (whole-batch prediction steps, tree search) only read / write the
        #      variance, so they run on IPP_LAYOUT_SPLIT (variance tiles in their own array: 4 + 4 B per cell move, not the
        #      ~27 B of the interleaved super-tiles).  Same workload, its own 31.5 GB of maps; a few executed steps first so
        #      that the beliefs are not uniform.
        seng = eng
        if args.search_layout != args.layout:
            scfg = EngineConfig(batch=B, layout=capi.LAYOUT_NAMES[args.search_layout], device=local_rank, seed=20260925, env_id_offset=rank * B,
                                stream=stream.cuda_stream, **WORKLOAD)
            seng = BatchedEngine(scfg)
            seng.reset(PRIOR_MEAN, PRIOR_VAR)
            seng.synth_ground_truth(seed=1000)
            predict_persistent_search = predict_legs(seng, "")
        else:
            results["predict"], results["predict_eval"] = results["predict@step_layout"], results["predict_eval@step_layout"]
            predict_persistent_search = predict_persistent

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
        # The search runs on the search engine (first Tm envs), after a few executed steps so that the beliefs are not uniform.
        meng = seng
        meng.reset(PRIOR_MEAN, PRIOR_VAR)
        with torch.cuda.stream(stream):
            for t in range(4):
                meng.step_device(action_ids_ptr=ids_dev[t % POOL].data_ptr(), reward_ptr=reward_dev.data_ptr(), reward_mode=capi.REWARD_TRACE)
            meng.sync()

        def search_leg(mcts, pri, val, tag):
            """Two passes of one search: host-synchronous to count its path edges / expansions (it is deterministic), then timed."""
            step = (lambda leaf: mcts.simulate_device(priors_window_ptr=pri.data_ptr(), values_ptr=val.data_ptr(), want_leaf=leaf)) if pri is not None \
                else (lambda leaf: mcts.simulate(lambda lf: (None, None)) if leaf else mcts.simulate(None))
            mcts.begin(budgets)
            edges, expansions = 0, 0
            for i in range(Sm):
                leaf = step(True)
                edges += int(leaf.path_len.sum())
                expansions += int(leaf.needs_eval.sum())
            mcts.begin(budgets)
            l0 = mcts.launches
            barrier()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nvtx.range_push(tag)
            m0.record(stream)
            for i in range(Sm):
                step(False)
            m1.record(stream)
            barrier()
            nvtx.range_pop()
            ms_m = max_over_ranks(m0.elapsed_time(m1))
            launches = int(mcts.launches - l0)
            st = mcts.root_stats()
            assert np.all(st["Ns"] == Sm - 1)
            # algorithmic bytes of the memoised search: per COMPUTED prediction step (= new edge) its footprint's variance read and
            # its overlay written (4 + 4 B per cell; mean footprint of the 3-altitude set, unclipped), per expansion one row of
            # evaluator priors read and one row of masked priors written, 32 B per path edge walked.  `prediction_steps` counts
            # what the reference computes for the same search: one simulate_prediction_step per level of every simulation
            # (mcts.py:239-246).
            mean_cells = float(np.mean([float((2 * r + 1) ** 2) for r in radii]))
            computed = int(mcts.info.edges)
            alg = 8.0 * mean_cells * computed + (8.0 if pri is not None else 4.0) * mcts.window_slots * expansions + 32.0 * edges
            return {"tree_simulations_per_sec": world * Tm * Sm / (ms_m * 1e-3), "ms_per_lockstep_simulation": ms_m / Sm,
                    "prediction_steps_per_sec": world * edges / (ms_m * 1e-3), "prediction_steps": edges, "prediction_steps_computed": computed,
                    "expansions": expansions, "edges_per_tree": computed / Tm, "gpu_launches": launches,
                    "algorithmic_bytes": alg, "achieved_gbs": alg / (ms_m * 1e-3) / 1e9}

        with torch.cuda.stream(stream):
            with BatchedMCTS(meng, hyper, meta, n_trees=Tm) as mcts:
                # (1) growing trees: synthetic network outputs resident in device memory — peaked priors (soft-max of random logits
                #     per tree) and small random values: about one new edge, one rollout and one expansion per simulation, as under
                #     a trained policy / value network;  (2) no evaluator: uniform priors, zero values — the search then exploits ONE
                #     path per tree (6 edges after 100 simulations) and measures the descent / backup alone (round 1's configuration)
                g = torch.Generator(device="cuda").manual_seed(5 + rank)
                pri = torch.softmax(4.0 * torch.randn(Tm, mcts.window_slots, device="cuda", generator=g), dim=1).contiguous()
                val = (0.05 * torch.rand(Tm, device="cuda", generator=g)).contiguous()
                torch.cuda.synchronize()
                mcts_res = {"trees_per_gpu": Tm, "simulations": Sm, "episode_horizon": 5, "window_slots": mcts.window_slots}
                mcts_res.update(search_leg(mcts, pri, val, "mcts[growing]"))
                mcts_res.update({
                    "evaluator": "synthetic network outputs in device memory: soft-max(4 * N(0,1)) priors over the window, values U(0, 0.05) "
                                 "(the policy / value network is outside this library)",
                    "no_evaluator": dict(search_leg(mcts, None, None, "mcts[uniform]"),
                                         evaluator="none: uniform priors, zero values (the search exploits one path per tree)"),
                    "tree_bytes_per_gpu": int(mcts.info.device_bytes), "layout": args.search_layout,
                    "rollouts": "memoised: an edge caches its reward, a node the variances its prediction step left behind; a simulation "
                                "computes at most ONE prediction step (the reference replays one per level: prediction_steps counts those)"})
                del pri, val
    if seng is not eng:
        seng.close()

    # ---- CPU baseline (rank 0, N == 1 only): oracle port on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import cpu_baseline

            r = cpu_baseline.run(WORKLOAD, steps=3, warmup=1, envs=args.cpu_envs, reward_mode=0, min_seconds=1.0)
            cpu = {"value": r["steps_per_sec"], "unit": "env-steps/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            tiers = os.path.join(ROOT, "profiles", "r02_reference_cpu_tiers.json")
            if os.path.exists(tiers):  # the reference's OWN code (pure Python, cannot travel): timed in the build container
                cpu["reference_tiers_build_container"] = json.load(open(tiers))
        except Exception as exc:  # report, never hide
            cpu = {"value": None, "unit": "env-steps/s", "cores": 0, "kind": "port", "sample": f"failed: {exc!r}"}

    if rank == 0:
        peak, peak_src = load_peaks()

        def mode_block(name, alg):
            ms = results[name]["ms_total"]
            ach = alg / (ms / K * 1e-3) / 1e9
            return {"value": world * B * K / (ms * 1e-3), "ms_per_launch": ms / K, "achieved": ach, "frac": ach / peak,
                    "algorithmic_bytes_per_launch": alg}

        modes = {n: mode_block(n, alg_bytes_per_launch) for n in reward_modes}
        head = modes[HEAD]
        kernel = {"super": "ipp_step_bulk_kernel", "split": "ipp_step_bulk_kernel", "tiled": "ipp_step_async_kernel", "mv": "ipp_step_async_kernel"}.get(args.layout, "ipp_step_kernel") \
            if eng.step_path == "async" else "ipp_step_kernel"
        traffic, traffic_note = None, "no ncu capture committed for this kernel"
        tp = os.path.join(ROOT, "profiles", "traffic_per_launch.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                ent = tj.get("kernels", {}).get(kernel + ":" + args.layout)
                if ent:
                    traffic = ent["bytes_per_launch"]
                    traffic_note = f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, captured at commit {ent.get('commit', '?')} ({ent.get('how', '')})"
            except Exception:
                pass
        def predict_block(suffix, persistent, layout, note):
            pb = mode_block("predict" + suffix, alg_bytes_predict)
            pb.update({"unit": "predict-steps/s", "kernel": "ipp_step_bulk_kernel<MODE_PREDICT>" if persistent else "ipp_step_kernel<MODE_PREDICT>",
                       "layout": layout, "bytes_per_cell": 8, "note": note})
            pe = mode_block("predict_eval" + suffix, (4.0 * cells_timed_total + 16.0 * B * K) / K)
            pb["evaluate_only"] = {"value": pe["value"], "ms_per_launch": pe["ms_per_launch"], "achieved": pe["achieved"], "frac": pe["frac"],
                                   "bytes_per_cell": 4, "note": "IPP_FLAG_NO_COMMIT: the variance is read, nothing is written"}
            return pb

        note_inter = ("covariance-only step on the interleaved layout: the staged run and the written sectors carry the mean (and, in the "
                      "super-tile layout, the ground truth) too, so ~27 B per cell move where 8 B are algorithmic (DESIGN.md 3.1)")
        note_split = ("covariance-only step on IPP_LAYOUT_SPLIT: the variance tiles have their own array, the kernel stages and writes "
                      "them alone (DESIGN.md 3.1)")
        pred = predict_block("", predict_persistent_search, args.search_layout, note_split if args.search_layout == "split" else note_inter)
        if args.search_layout != args.layout:
            pred["on_step_layout"] = predict_block("@step_layout", predict_persistent, args.layout, note_inter)
        line = {
            "metric": "env_steps_per_sec", "value": head["value"], "unit": "env-steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": head["ms_per_launch"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3: batch 65536 envs/GPU, 200x200 grid, 3-altitude action set {8,14,20} m, Kalman fusion, information-gain reward: "
                                   "trace (variance) reduction = the reference's reward, parity-pinned [headline]; Gaussian-entropy reduction = "
                                   "extension, parity unpinned [roofline.modes / e2e in the same mode as the headline]",
                       "batch_per_gpu": B, "grid": [Y, X], "layout": args.layout, "noise": "device Philox4x32-10",
                       "reward_mode": HEAD,
                       "cache": "working set 31.5 GB/GPU >> 126 MB L2 (inputs larger than L2, no flush needed)",
                       "parallelism": f"env-batch sharded x{world}"},
            "roofline": {"bound": "hbm", "achieved": head["achieved"], "peak": peak, "unit": "GB/s", "frac": head["frac"],
                         "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                         "mean_cells_per_env_step": cells_timed_total / (B * K), "kernel": kernel, "step_path": eng.step_path,
                         "modes": modes, "predict": pred},
            "e2e": {"value": world * B * KE / e2e_s, "steps": KE, "unit": "env-steps/s", "reward_mode": HEAD,
                    "us_per_step": 1e6 * e2e_s / KE, "h2d_bytes_per_step": 4 * B * world, "d2h_bytes_per_step": 4 * B * world,
                    "pipelined_value": world * B * KE / pipe_s,
                    "pipelined_path": "ipp_step_submit / ipp_step_wait, 2 slots: same per-step H2D ids + rewards to pinned host memory, upload of "
                                      "step t+1 under the kernel of step t",
                    "path": ("BatchedEngine.step (ipp_step: pinned host ids -> "
                             + ("fetched by the fused kernel itself, 512-byte slices over PCIe under its first footprints -> "
                                if ids_fetch_steps > 0 else "H2D -> fused kernel -> ")
                             + ("rewards written by the kernel into the caller's pinned buffer [zero-copy D2H, 4 B/env over PCIe])"
                                if zero_copy_steps > 0 else "D2H rewards)"))
                            + ("" if dist is None else "; every rank writes its slice of ONE pinned segment shared by the node's ranks "
                               "(POSIX shm + cudaHostRegister), the learner rank reads the whole batch's rewards there: no collective on the step path")},
            "gpu_launches": int(results[HEAD]["launches"]),
            "gpu_launches_e2e": int(launches_e2e),
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if mcts_res is not None:
            mcts_res["frac_of_hbm_peak"] = mcts_res["achieved_gbs"] / peak / world
            line["roofline"]["mcts"] = mcts_res
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
    ap.add_argument("--layout", default=os.environ.get("IPP_LAYOUT", "super"), choices=["planes", "mv", "tiled", "super", "split"])
    ap.add_argument("--cpu-envs", type=int, default=4096, help="env sample of the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=300, help="steps of the host-buffer (e2e) legs (independent of --steps)")
    ap.add_argument("--zero-copy", default=None, choices=["", "r", "i", "ri", "f", "rf"],
                    help="e2e leg: host buffers the kernel accesses in place (r = rewards, i = action ids read in place, f = action ids fetched "
                         "by the persistent kernel); default = the engine's (rf)")
    ap.add_argument("--mcts-trees", type=int, default=16384, help="trees of the secondary mcts_zero rollout leg = BASELINE.json C4 per-GPU share (0 = skip)")
    ap.add_argument("--mcts-sims", type=int, default=100, help="simulations per decision (BASELINE.json C4: 100)")
    ap.add_argument("--search-layout", "--mcts-layout", dest="search_layout", default="split", choices=["planes", "mv", "tiled", "super", "split"],
                    help="belief layout of the search engine (whole-batch prediction steps + tree search legs)")
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
