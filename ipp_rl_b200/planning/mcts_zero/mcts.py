"""Batched MCTS-zero rollout loop — host side of ``include/ipp_mcts.h`` (``csrc/mcts.cu``).

Reference: ``planning/mcts_zero/mcts.py`` — ``MCTS.get_policy`` (:83-143), ``simulate`` (:166-265),
``compute_uct`` (:280-296), ``get_next_actions_mask`` (:148-158), ``normalize_q_values`` (:267-278),
``add_exploration_noise`` (:160-164) — one tree per worker process, one dense covariance copy per tree
level.  Here: one tree per env of a :class:`~ipp_rl_b200.engine.BatchedEngine`, all trees advanced in
lock-step on the GPU; the prediction steps of a tree path run as one warp-level path rollout that never
writes the belief (``ipp_rollout_device``).

Same hyper-parameter / meta-data keys as the reference (``config/example.yaml:54-62``).  The
policy/value network stays outside (stock PyTorch in the reference): ``evaluator(leaf)`` is called once
per simulation with a :class:`LeafBatch` and returns ``(priors, values)``; ``None`` = uniform priors,
zero values (pure reward-driven search).

Differences from the reference (include/ipp_mcts.h): nodes are keyed by their path, candidate actions
live in a window of ``levels x D x D`` slots around a node, arg-max ties go to the lowest action id, a
fresh tree per ``get_policy``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Dict, Optional, Tuple

import numpy as np

from ... import _capi as capi


@dataclass
class LeafBatch:
    """Leaves of one lock-step simulation (one row per tree)."""

    kind: np.ndarray    # capi.MCTS_LEAF_TERMINAL / MCTS_LEAF_EVAL
    node: np.ndarray    # node index inside the tree (-1: no node)
    col: np.ndarray     # centre cell of the leaf's action window
    row: np.ndarray
    level: np.ndarray   # altitude level of the leaf (-1: the root, an arbitrary pose)
    depth: np.ndarray
    budget: np.ndarray  # remaining budget at the leaf (float32)
    path_len: np.ndarray
    mcts: "BatchedMCTS"

    @property
    def needs_eval(self) -> np.ndarray:
        return self.kind == capi.MCTS_LEAF_EVAL

    def window_action_ids(self) -> np.ndarray:
        """(n_trees, W) action id of every slot of each leaf's window (-1 outside the grid)."""
        return self.mcts.window_action_ids(self.col, self.row)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class BatchedMCTS:
    def __init__(self, engine, hyper_params: Dict, meta_data: Dict, n_trees: Optional[int] = None, first_env: int = 0,
                 reward_mode: int = capi.REWARD_TRACE):
        self.engine = engine
        self.hyper_params = dict(hyper_params)
        self.meta_data = dict(meta_data)
        self._lib = capi.load_library()
        self.n_trees = engine.batch - first_env if n_trees is None else int(n_trees)
        self.first_env = int(first_env)
        try:
            self.num_simulations = int(hyper_params["num_mcts_simulations"])
            self.puct_init = float(hyper_params["puct_init"])
            self.puct_base = float(hyper_params["puct_base"])
            self.gamma = float(hyper_params["gamma"])
            self.forced_playout_factor = float(hyper_params["forced_playout_factor"])
            self.max_valid_action_distance = float(hyper_params["max_valid_action_distance"])
            self.episode_horizon = int(meta_data["episode_horizon"])
        except KeyError as exc:  # same behaviour as the reference's dict accesses
            raise ValueError(f"Cannot find {exc} specification in the MCTS hyper-parameters / meta data!") from exc
        self.dirichlet_alpha = float(hyper_params.get("dirichlet_alpha", 0.0))
        self.dirichlet_eps = float(hyper_params.get("dirichlet_eps", 0.0))
        self.adaptive = meta_data.get("scenario_info") is not None
        cfg = capi.ipp_mcts_config()
        cfg.struct_bytes = C.sizeof(capi.ipp_mcts_config)
        cfg.n_trees, cfg.first_env = self.n_trees, self.first_env
        cfg.num_simulations, cfg.episode_horizon = self.num_simulations, self.episode_horizon
        cfg.step_flags = (int(reward_mode) & 3) | (capi.FLAG_ADAPTIVE if self.adaptive else 0)
        cfg.puct_init, cfg.puct_base, cfg.gamma = self.puct_init, self.puct_base, self.gamma
        cfg.forced_playout_factor = self.forced_playout_factor
        cfg.max_valid_action_distance = self.max_valid_action_distance
        cfg.dirichlet_eps = self.dirichlet_eps
        self._h = C.c_void_p()
        rc = self._lib.ipp_mcts_create(engine._h, C.byref(cfg), C.byref(self._h))
        if rc != capi.IPP_OK:
            msg = self._lib.ipp_mcts_last_error(None)
            self._h = C.c_void_p()
            raise capi.IppError(rc, msg.decode() if msg else "ipp_mcts_create failed")
        info = self.info
        self.window_slots, self.window_dim, self.window_radius = info.window_slots, info.window_dim, info.window_radius
        self.levels, self.max_path = info.levels, info.max_path
        self.num_actions = engine.num_actions
        self._leaf = np.zeros((self.n_trees, capi.MCTS_LEAF_WORDS), np.int32)

    # -- plumbing -------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ipp_mcts_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int) -> None:
        if rc != capi.IPP_OK:
            msg = self._lib.ipp_mcts_last_error(self._h)
            raise capi.IppError(rc, msg.decode() if msg else "unknown error")

    @property
    def info(self) -> capi.ipp_mcts_info:
        out = capi.ipp_mcts_info()
        self._ck(self._lib.ipp_mcts_get_info(self._h, C.byref(out)))
        return out

    @property
    def launches(self) -> int:
        return int(self.info.launches)

    def device_ptr(self, which: int) -> int:
        p = self._lib.ipp_mcts_device_ptr(self._h, which)
        return int(p) if p else 0

    # -- geometry ---------------------------------------------------------------------------------
    def window_action_ids(self, col, row) -> np.ndarray:
        """Action ids (planning/common/actions.py:73-91: level*N + x_dim*col + row) of the window slots around the
        given centre cells; -1 outside the grid.  Slot = (level*D + dcol + r)*D + drow + r."""
        D, r, L = self.window_dim, self.window_radius, self.levels
        X, Y = self.engine.x_dim, self.engine.y_dim
        col = np.asarray(col)[:, None, None, None]
        row = np.asarray(row)[:, None, None, None]
        lv = np.arange(L)[None, :, None, None]
        c = col + np.arange(-r, r + 1)[None, None, :, None]
        w = row + np.arange(-r, r + 1)[None, None, None, :]
        ids = lv * (X * Y) + X * c + w
        ok = (c >= 0) & (c < X) & (w >= 0) & (w < Y)
        return np.where(ok, ids, -1).reshape(-1, L * D * D).astype(np.int32)

    # -- one search ------------------------------------------------------------------------------
    def begin(self, budgets, previous_actions=None) -> None:
        """Start a fresh tree per env from its current belief: ``previous_actions`` (n_trees, 3) poses (None = the
        engine's stored previous actions), ``budgets`` (n_trees,) remaining budgets."""
        b = np.ascontiguousarray(np.broadcast_to(np.asarray(budgets, np.float32), (self.n_trees,)))
        pp = None
        if previous_actions is not None:
            pp = np.ascontiguousarray(np.broadcast_to(np.asarray(previous_actions, np.float64), (self.n_trees, 3)))
        self._ck(self._lib.ipp_mcts_begin(self._h, _ptr(pp), _ptr(b)))

    def sample_root_noise(self, rng: np.random.Generator) -> Optional[np.ndarray]:
        """Dirichlet(alpha) over ALL actions (mcts.py:160-164), returned for the root window's slots only: the window's
        gamma variates divided by (their sum + one Gamma(alpha * #other actions) variate)."""
        if not (self.dirichlet_eps > 0 and self.dirichlet_alpha > 0):
            return None
        W = self.window_slots
        g = rng.gamma(self.dirichlet_alpha, 1.0, size=(self.n_trees, W))
        ids = self._root_ids()
        g = np.where(ids >= 0, g, 0.0)
        rest = self.num_actions - (ids >= 0).sum(axis=1)
        g_rest = np.where(rest > 0, rng.gamma(self.dirichlet_alpha * np.maximum(rest, 1), 1.0), 0.0)
        return (g / (g.sum(axis=1) + g_rest)[:, None]).astype(np.float32)

    def _root_ids(self) -> np.ndarray:
        ids = np.empty((self.n_trees, self.window_slots), np.int32)
        self._ck(self._lib.ipp_mcts_root_stats(self._h, None, None, None, _ptr(ids), None))
        return ids

    def simulate(self, evaluator: Optional[Callable] = None, root_noise: Optional[np.ndarray] = None) -> Optional[LeafBatch]:
        """One lock-step simulation of every tree (mcts.py:166-265)."""
        if evaluator is None:
            self._ck(self._lib.ipp_mcts_simulate_begin(self._h, None))
            rn = None if root_noise is None else np.ascontiguousarray(root_noise, np.float32)
            self._ck(self._lib.ipp_mcts_simulate_end(self._h, None, None, None, _ptr(rn), 0))
            return None
        self._ck(self._lib.ipp_mcts_simulate_begin(self._h, _ptr(self._leaf)))
        lf = self._leaf
        leaf = LeafBatch(kind=lf[:, 0].copy(), node=lf[:, 1].copy(), col=lf[:, 2].copy(), row=lf[:, 3].copy(), level=lf[:, 4].copy(),
                         depth=lf[:, 5].copy(), budget=lf[:, 6].copy().view(np.float32), path_len=lf[:, 7].copy(), mcts=self)
        priors, values = evaluator(leaf)
        pw = pd = None
        if priors is not None:
            priors = np.ascontiguousarray(priors, np.float32)
            if priors.shape == (self.n_trees, self.window_slots):
                pw = priors
            elif priors.shape == (self.n_trees, self.num_actions):
                pd = priors
            else:
                raise ValueError(f"priors must be (n_trees, {self.window_slots}) window slots or (n_trees, {self.num_actions}) dense")
        v = None if values is None else np.ascontiguousarray(np.broadcast_to(np.asarray(values, np.float32), (self.n_trees,)))
        rn = None if root_noise is None else np.ascontiguousarray(root_noise, np.float32)
        self._ck(self._lib.ipp_mcts_simulate_end(self._h, _ptr(pw), _ptr(pd), _ptr(v), _ptr(rn), 0))
        return leaf

    def _leaf_batch(self) -> "LeafBatch":
        lf = self._leaf
        return LeafBatch(kind=lf[:, 0].copy(), node=lf[:, 1].copy(), col=lf[:, 2].copy(), row=lf[:, 3].copy(), level=lf[:, 4].copy(),
                         depth=lf[:, 5].copy(), budget=lf[:, 6].copy().view(np.float32), path_len=lf[:, 7].copy(), mcts=self)

    def simulate_device(self, priors_window_ptr: int = 0, priors_dense_ptr: int = 0, values_ptr: int = 0, root_noise_ptr: int = 0,
                        want_leaf: bool = False) -> Optional["LeafBatch"]:
        """One lock-step simulation with the evaluator's outputs already in device memory (raw pointers, 0 = absent: uniform
        priors / zero values): float32 (n_trees, window_slots) or (n_trees, num_actions) priors, (n_trees,) values.  Nothing
        crosses the host (unless ``want_leaf`` asks for the leaf records): the form a GPU-resident policy / value network uses."""
        self._ck(self._lib.ipp_mcts_simulate_begin(self._h, _ptr(self._leaf) if want_leaf else None))
        leaf = self._leaf_batch() if want_leaf else None
        vp = lambda p: C.c_void_p(p) if p else None  # noqa: E731
        self._ck(self._lib.ipp_mcts_simulate_end(self._h, vp(priors_window_ptr), vp(priors_dense_ptr), vp(values_ptr), vp(root_noise_ptr), 1))
        return leaf

    def paths(self):
        """(actions, rewards) of the simulation in flight — callable from an evaluator: action ids root -> leaf (-1 padded)
        and the rewards of their prediction steps, both (n_trees, max_path)."""
        P = self.max_path
        a, r = np.empty((self.n_trees, P), np.int32), np.empty((self.n_trees, P), np.float32)
        self._ck(self._lib.ipp_mcts_get_paths(self._h, _ptr(a), _ptr(r)))
        return a, r

    def root_stats(self) -> Dict[str, np.ndarray]:
        T, W = self.n_trees, self.window_slots
        ps, qsa = np.empty((T, W), np.float32), np.empty((T, W), np.float32)
        nsa, ids, ns = np.empty((T, W), np.int32), np.empty((T, W), np.int32), np.empty(T, np.int32)
        self._ck(self._lib.ipp_mcts_root_stats(self._h, _ptr(ps), _ptr(qsa), _ptr(nsa), _ptr(ids), _ptr(ns)))
        return dict(Ps=ps, Qsa=qsa, Nsa=nsa, action_ids=ids, Ns=ns)

    # -- reference-shaped helpers (vectorised over trees) -------------------------------------------
    @staticmethod
    def normalize_q_values(values: np.ndarray) -> np.ndarray:
        """mcts.py:267-278 per row; rows are window slices of the dense action vector, whose other entries are 0."""
        v = np.asarray(values, np.float64)
        lo = np.minimum(v.min(axis=-1, keepdims=True), 0.0)
        hi = np.maximum(v.max(axis=-1, keepdims=True), 0.0)
        span = hi - lo
        return np.where(span > 0, (v - lo) / np.where(span > 0, span, 1.0), v)

    def compute_uct(self, stats: Dict[str, np.ndarray], force_playouts: bool = False, visits: Optional[np.ndarray] = None) -> np.ndarray:
        """mcts.py:280-296 on root statistics."""
        ps = stats["Ps"].astype(np.float64)
        valid = ps >= 0
        p = np.where(valid, ps, 0.0)
        ns = stats["Ns"].astype(np.float64)[:, None]
        nsa = (stats["Nsa"] if visits is None else visits).astype(np.float64)
        prior = self.puct_init + np.log((ns + self.puct_base + 1) / self.puct_base)
        with np.errstate(divide="ignore", invalid="ignore"):
            uct = self.normalize_q_values(stats["Qsa"]) + prior * p * (np.sqrt(ns + 1) / (1 + nsa))
        if force_playouts:
            nf = np.ceil(np.sqrt(self.forced_playout_factor * p * ns))
            nf[stats["Nsa"] == 0] = 0
            uct[stats["Nsa"] < nf] = np.inf
        uct[~valid] = -np.inf
        return uct

    def get_next_actions_mask(self, stats: Dict[str, np.ndarray]) -> np.ndarray:
        """mcts.py:148-158 for the root (evaluated on the device at expansion): valid window slots."""
        return stats["Ps"] >= 0

    def get_policy(self, budgets, previous_actions=None, evaluator: Optional[Callable] = None, temperature: float = 1.0,
                   deploy_time: bool = False, rng: Optional[np.random.Generator] = None,
                   root_noise: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """``num_mcts_simulations`` lock-step simulations from every env's current belief, then the reference's policy
        extraction (mcts.py:94-143).  Returns ``(policy, action_ids, visits)``, each (n_trees, W) over the root window's
        slots; rows of trees without any valid visited action are all zero (reference: ``None``)."""
        self.begin(budgets, previous_actions)
        # the reference adds exploration noise to the root's first expansion whenever dirichlet_eps > 0, at deploy time
        # too (mcts.py:225-226); here the caller opts in by passing a generator
        noise = root_noise if root_noise is not None else (self.sample_root_noise(rng) if rng is not None else None)
        for i in range(self.num_simulations):
            self.simulate(evaluator, root_noise=noise if i == 0 else None)
        stats = self.root_stats()
        visits = stats["Nsa"].astype(np.float64)
        if not deploy_time:
            visits = self._prune_forced_playouts(stats, visits)
        total = visits.sum(axis=1, keepdims=True)
        if temperature == 0:
            best = np.argmax(visits, axis=1)
            policy = np.zeros_like(visits)
            policy[np.arange(self.n_trees), best] = 1.0
        else:
            vt = visits ** (1.0 / temperature)
            s = vt.sum(axis=1, keepdims=True)
            policy = vt / np.where(s > 0, s, 1.0)
        policy[total[:, 0] == 0] = 0.0
        return policy, stats["action_ids"], visits

    def _prune_forced_playouts(self, stats: Dict[str, np.ndarray], visits: np.ndarray) -> np.ndarray:
        """Policy-target pruning of forced playouts (mcts.py:99-128)."""
        T = self.n_trees
        ps = np.where(stats["Ps"] >= 0, stats["Ps"], 0.0).astype(np.float64)
        ns = stats["Ns"].astype(np.float64)[:, None]
        best = np.argmax(visits, axis=1)
        nf = np.ceil(np.sqrt(self.forced_playout_factor * ps * ns))
        nf[stats["Nsa"] == 0] = 0
        max_puct = self.compute_uct(stats, force_playouts=False)[np.arange(T), best][:, None]
        qn = self.normalize_q_values(stats["Qsa"])
        prior_c = self.puct_init + np.log((ns + self.puct_base + 1) / self.puct_base)
        not_best = np.ones_like(visits, bool)
        not_best[np.arange(T), best] = False
        done = np.zeros_like(visits, bool)
        for it in range(int(nf.max()) if nf.size else 0):
            active = (nf > it) & not_best & ~done
            if not active.any():
                break
            visits = np.where(active, visits - 1, visits)
            with np.errstate(divide="ignore", invalid="ignore"):
                pruned = qn + prior_c * ps * (np.sqrt(ns + 1) / (1 + visits))
            hit = active & (pruned >= max_puct)
            visits = np.where(hit, visits + 1, visits)
            done |= hit
        visits[visits == 1] = 0
        return visits
