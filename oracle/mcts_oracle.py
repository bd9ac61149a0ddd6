"""CPU oracle of the MCTS-zero rollout loop (TEST INFRASTRUCTURE — NOT PRODUCT CODE).

A float64 NumPy restatement of ``planning/mcts_zero/mcts.py`` (reference @ 25dfb33): ``simulate``
(:166-265), ``compute_uct`` (:280-296), ``normalize_q_values`` (:267-278),
``get_next_actions_mask`` (:148-158), ``add_exploration_noise`` (:160-164) and ``get_policy``
(:83-143), dense over the whole action set like the reference, driving the diagonal prediction step of
``oracle/ipp_oracle.py`` (``simulate_prediction_step``, planning/common/optimization.py:14-30).

Deliberate differences from the reference (the same as the product's, include/ipp_mcts.h):

* the policy/value network + feature planes + request/reply queues (:187-218) are replaced by an
  ``evaluator(node_info) -> (policy over all actions, value)`` callable;
* a node is keyed by its PATH from the root (tuple of action ids), not by ``hash(str(state))``
  (:20-21) — the reference's key collides for arrays > 1000 elements (NumPy summarises the repr) and
  merges transpositions regardless of the previous action;
* ties in ``arg max`` go to the lowest action id (reference: ``np.random.choice``).

Pinning status: PINNED for the parts that do not depend on those differences —
``tests/golden/make_golden.py`` runs the REAL reference ``MCTS`` (stub evaluator answering the queue
synchronously, ``np.random.choice`` -> first candidate) on a grid small enough that the reference's
state keys do not collide and no transposition occurs, and stores root visit counts / Q values / priors /
policy; ``tests/test_oracle_golden.py`` re-checks this module against them.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import numpy as np

from oracle import ipp_oracle as orc


def normalize_q_values(values: np.ndarray) -> np.ndarray:
    """mcts.py:267-278."""
    if np.all(values == 0):
        return values
    lo, hi = np.min(values), np.max(values)
    if lo == hi:
        return values / hi
    return (values - lo) / (hi - lo)


class OracleMCTS:
    """One tree; ``evaluator(info) -> (policy[num_actions], value)`` with ``info`` = dict(path, depth, budget,
    previous_action, var)."""

    def __init__(self, cfg: orc.OracleConfig, hyper_params: Dict, episode_horizon: int, evaluator: Optional[Callable] = None,
                 mean: Optional[np.ndarray] = None, adaptive: bool = False, reward_mode: int = orc.REWARD_TRACE):
        self.cfg = cfg
        self.hp = hyper_params
        self.H = episode_horizon
        self.evaluator = evaluator
        self.mean = mean
        self.adaptive = adaptive
        self.reward_mode = reward_mode
        self.actions_np = orc.enumerate_actions(cfg)
        self.num_actions = self.actions_np.shape[0]
        self.Qsa, self.Nsa, self.Ns, self.Vs, self.Ps = {}, {}, {}, {}, {}
        self.inference_counter = 0

    # mcts.py:148-158.  Every call site passes (position, budget) only, so ``uav_specificaion`` is None and the
    # mask compares the Euclidean DISTANCE with the budget even when costs are flight times (reference quirk,
    # reproduced: the budget itself is still decremented by action_costs, mcts.py:247).
    def get_next_actions_mask(self, position, budget) -> np.ndarray:
        distances = np.linalg.norm(self.actions_np - np.asarray(position, float), ord=2, axis=1)
        return (distances > 0) & (distances <= budget) & (distances < self.hp["max_valid_action_distance"])

    # mcts.py:280-296
    def compute_uct(self, key, force_playouts: bool = False) -> np.ndarray:
        qn = normalize_q_values(self.Qsa[key])
        prior = self.hp["puct_init"] + np.log((self.Ns[key] + self.hp["puct_base"] + 1) / self.hp["puct_base"])
        prior = prior * self.Ps[key] * (np.sqrt(self.Ns[key] + 1) / (1 + self.Nsa[key]))
        uct = qn + prior
        if force_playouts:
            nf = np.ceil(np.sqrt(self.hp["forced_playout_factor"] * self.Ps[key] * self.Ns[key]))
            nf[self.Nsa[key] == 0] = 0
            uct[self.Nsa[key] < nf] = np.inf
        uct[~self.Vs[key]] = -np.inf
        return uct

    # mcts.py:166-265
    def simulate(self, key: Tuple[int, ...], var: np.ndarray, depth: int, budget: float, previous_action, num_sim: int,
                 root_noise: Optional[np.ndarray] = None) -> float:
        if depth > self.H or budget <= 0:
            return 0.0
        if key not in self.Nsa:
            self.Nsa[key] = np.zeros(self.num_actions)
            self.Qsa[key] = np.zeros(self.num_actions)
        if key not in self.Ps:  # leaf
            msk = self.get_next_actions_mask(previous_action, budget)
            if msk.sum() == 0:
                return 0.0
            info = dict(path=key, depth=depth, budget=budget, previous_action=np.asarray(previous_action, float), var=var)
            if self.evaluator is None:
                policy, value = np.ones(self.num_actions), 0.0
            else:
                policy, value = self.evaluator(info)
            ps = np.asarray(policy, float) * msk
            self.inference_counter += 1
            if depth == 0 and num_sim == 0 and root_noise is not None:  # add_exploration_noise, :160-164
                eps = self.hp["dirichlet_eps"]
                ps = (1 - eps) * ps + eps * root_noise
                ps = ps / np.sum(ps)
            if np.sum(ps) > 0:
                ps = ps / np.sum(ps)
            else:
                ps = ps + msk
                ps = ps / np.sum(ps)
            self.Ps[key] = ps
            self.Vs[key] = msk
            self.Ns[key] = 0
            return float(value)
        uct = self.compute_uct(key, force_playouts=(depth == 0))
        a = int(np.flatnonzero(uct == np.max(uct))[0])  # lowest id among ties
        action = self.actions_np[a].copy()
        reward, var_next = orc.simulate_prediction_step(self.cfg, var, previous_action, action, mean=self.mean, adaptive=self.adaptive,
                                                        reward_mode=self.reward_mode)
        budget_next = budget - orc.action_costs(action, previous_action, self.cfg.uav)
        value = reward + self.hp["gamma"] * self.simulate(key + (a,), var_next, depth + 1, budget_next, action, num_sim)
        if self.Nsa[key][a] > 0:
            self.Qsa[key][a] = (self.Nsa[key][a] * self.Qsa[key][a] + value) / (self.Nsa[key][a] + 1)
            self.Nsa[key][a] += 1
        else:
            self.Qsa[key][a] = value
            self.Nsa[key][a] = 1
        self.Ns[key] += 1
        return value

    def search(self, var: np.ndarray, previous_action, budget: float, num_simulations: int, root_noise: Optional[np.ndarray] = None):
        for i in range(num_simulations):
            self.simulate((), var, 0, budget, previous_action, i, root_noise=root_noise)
        return self.Nsa.get((), np.zeros(self.num_actions)).copy()

    # mcts.py:94-143 (after the simulations)
    def policy_from_root(self, temperature: float = 1.0, deploy_time: bool = False):
        key = ()
        visits = self.Nsa[key].copy()
        if not deploy_time:
            best = int(np.flatnonzero(visits == np.max(visits))[0])
            nf = np.ceil(np.sqrt(self.hp["forced_playout_factor"] * self.Ps[key] * self.Ns[key]))
            nf[self.Nsa[key] == 0] = 0
            max_puct = self.compute_uct(key, force_playouts=False)[best]
            for a in range(len(nf)):
                if a == best or nf[a] <= 0:
                    continue
                for _ in range(int(nf[a])):
                    visits[a] -= 1
                    qn = normalize_q_values(self.Qsa[key])[a]
                    prior = self.hp["puct_init"] + np.log((self.Ns[key] + self.hp["puct_base"] + 1) / self.hp["puct_base"])
                    with np.errstate(divide="ignore", invalid="ignore"):
                        prior = prior * self.Ps[key][a] * (np.sqrt(self.Ns[key] + 1) / (1 + visits[a]))
                    if qn + prior >= max_puct:
                        visits[a] += 1
                        break
            visits[visits == 1] = 0
        if np.sum(visits) == 0:
            return None, visits
        if temperature == 0:
            best = int(np.flatnonzero(visits == np.max(visits))[0])
            policy = np.zeros(len(visits))
            policy[best] = 1
            return policy, visits
        vt = visits ** (1.0 / temperature)
        return vt / np.sum(vt), visits
