"""MCTSZeroMission — the deploy-time planner of the reference (planning/mcts_zero/mcts_zero_mission.py:93-680) as a host
loop over the batched engine.

What is kept: the constructor signature, ``get_meta_data`` (:201-214), ``get_next_actions_mask`` (:457-467), ``replan``
(:469-523: root-parallel search — ``num_workers`` independent trees from the same belief, visit-count policies SUMMED over
the workers, arg-max action) and the mission loop of ``execute`` (:596-650).  What replaces the reference's machinery:

* the ``num_workers`` forked MCTS processes become ``num_workers`` trees of ONE ``BatchedMCTS`` advanced in lock-step on the
  GPU (each tree its own Dirichlet root noise; seeds ``42 * planning_step`` as in ``run_deploy_time_mcts_worker`` :36-55);
* with ``torch.distributed`` initialised the per-rank policy sums are all-reduced (``distributed.all_reduce_policy``), i.e.
  ``world_size x num_workers`` root-parallel trees;
* the policy/value network is outside this library: pass ``evaluator(leaf) -> (priors, values)`` (e.g. a torch module
  wrapped around ``BatchedEngine.observe``); ``None`` searches with uniform priors and zero values.  Self-play training
  (``learn``, arenas, inference server processes) is trainer orchestration and not provided.
"""
import logging
import time
from typing import Callable, Dict, Optional

import numpy as np

from ... import _capi as capi
from ...constants import MissionType
from ...engine import BatchedEngine, EngineConfig
from ..common.actions import action_costs, action_dict_to_np_array, compute_flight_time, compute_flight_times, enumerate_actions
from ..missions import Mission
from .mcts import BatchedMCTS

logger = logging.getLogger(__name__)


class MCTSZeroMission(Mission):
    def __init__(self, mapping, uav_specifications: Dict, hyper_params: Dict, dist_to_boundaries: float = 10, min_altitude: float = 5,
                 max_altitude: float = 30, episode_horizon: int = 10, altitude_spacing: float = 5, budget: float = 100,
                 model_deployment_filename: str = "best.pth.tar", train_examples_iter: int = 0, restart_training: bool = False,
                 adaptive: bool = False, value_threshold: float = 0.5, interval_factor: float = 2, telegram_notifications: bool = False,
                 config_name: str = "standard", use_effective_mission_time: bool = False, evaluator: Optional[Callable] = None):
        super().__init__(mapping, uav_specifications, dist_to_boundaries, min_altitude, max_altitude, budget, adaptive, value_threshold,
                         interval_factor, config_name, use_effective_mission_time)
        self.episode_horizon = episode_horizon
        self.altitude_spacing = altitude_spacing
        self.hyper_params = hyper_params
        self.initial_budget = budget
        self.model_deployment_filename = model_deployment_filename
        self.train_examples_iter = train_examples_iter
        self.restart_training = restart_training
        self.telegram_notifications = telegram_notifications
        self.evaluator = evaluator
        self.actions = enumerate_actions(self.mapping.grid_map, self.min_altitude, self.max_altitude, self.altitude_spacing)
        self.actions_np = action_dict_to_np_array(self.actions)
        self.mission_name = "Ours"
        self.mission_type = MissionType.MCTS_ZERO
        self.meta_data = self.get_meta_data()
        self._planner: Optional[BatchedEngine] = None

    def get_meta_data(self) -> Dict:
        return {
            "budget": self.budget,
            "initial_budget": self.initial_budget,
            "episode_horizon": self.episode_horizon,
            "max_episode_steps": self.hyper_params["max_episode_steps"],
            "min_altitude": self.min_altitude,
            "max_altitude": self.max_altitude,
            "altitude_spacing": self.altitude_spacing,
            "cov_matrix_shape": self.mapping.grid_map.cov_matrix.shape,
            "num_grid_cells": self.mapping.grid_map.num_grid_cells,
            "uav_specifications": self.uav_specifications,
            "scenario_info": self.get_adaptive_info(),
        }

    def create_waypoints(self) -> np.array:
        raise NotImplementedError("MCTS zero planning mission does not implement 'create_waypoints' function!")

    def learn(self):
        raise NotImplementedError("self-play training is trainer orchestration outside this library (DESIGN.md, out of scope); "
                                  "use BatchedMCTS + ExperienceRing from a training script")

    def get_next_actions_mask(self, position: np.array, budget: float) -> np.array:
        """reference :457-467"""
        distances = np.linalg.norm(self.actions_np - np.asarray(position, float), ord=2, axis=1)
        flight_times = compute_flight_times(self.actions_np, position, self.uav_specifications)
        return (flight_times > 0) & (flight_times <= budget) & (distances < self.hyper_params["max_valid_action_distance"])

    # -- planner engine: num_workers copies of the current belief, one tree each ---------------------------------------
    def _planner_engine(self) -> BatchedEngine:
        if self._planner is None:
            workers = max(1, int(self.hyper_params.get("num_workers", 1)))
            cfg = EngineConfig.from_params(self.mapping.grid_map.params, batch=workers, layout=capi.LAYOUT_MV,
                                           min_altitude=float(self.min_altitude), max_altitude=float(self.max_altitude),
                                           altitude_spacing=float(self.altitude_spacing), value_threshold=float(self.value_threshold),
                                           interval_factor=float(self.interval_factor),
                                           max_v=float(self.uav_specifications["max_v"]), max_a=float(self.uav_specifications["max_a"]))
            self._planner = BatchedEngine(cfg)
            self._planner.reset(0.5, 1.0)
        return self._planner

    def close(self) -> None:
        if self._planner is not None:
            self._planner.close()
            self._planner = None

    def replan(self, budget: float, previous_action: np.array, planning_step: int) -> Optional[np.ndarray]:
        """reference :469-523.  Returns the next waypoint, or None when no action is affordable."""
        eng = self._planner_engine()
        g = self.mapping.grid_map
        W = eng.batch
        eng.set_state(np.broadcast_to(np.asarray(g.mean, np.float32), (W, g.y_dim, g.x_dim)),
                      np.broadcast_to(np.asarray(g.var, np.float32), (W, g.y_dim, g.x_dim)))
        meta = dict(self.meta_data, scenario_info=self.get_adaptive_info())
        with BatchedMCTS(eng, self.hyper_params, meta) as mcts:
            rng = np.random.default_rng(42 * planning_step)  # run_deploy_time_mcts_worker seeds 42 * planning_step + worker_id
            policy, ids, _ = mcts.get_policy(np.full(W, budget, np.float32), np.asarray(previous_action, np.float64), evaluator=self.evaluator,
                                             temperature=1, deploy_time=True, rng=rng)
        policy_total = np.zeros(len(self.actions_np))
        ok = ids >= 0
        np.add.at(policy_total, ids[ok], policy[ok])
        from ... import distributed

        policy_total = distributed.all_reduce_policy(policy_total)
        if policy_total.sum() <= 0:
            return None
        policy_total /= np.sum(policy_total)
        return self.actions_np[int(np.argmax(policy_total)), :]

    def execute(self):
        """The deploy-time loop (reference :596-650): replan, fly, measure, fuse, evaluate."""
        remaining_budget = self.budget
        previous_action = self.init_action
        self.eval(run_time=0, flight_time=0)
        try:
            while remaining_budget >= self.mapping.grid_map.resolution:
                logger.info(f"\nREMAINING BUDGET: {remaining_budget}")
                start_time = time.time()
                action = self.replan(remaining_budget, previous_action, planning_step=len(self.waypoints))
                finish_time = time.time()
                if action is None:
                    break
                simulated_raw_measurement = self.mapping.sensor.take_measurement(action)
                self.mapping.update_grid_map(action, simulated_raw_measurement)
                self.waypoints = np.vstack((self.waypoints, action))
                flight_time = compute_flight_time(action, previous_action, self.uav_specifications)
                run_time = finish_time - start_time
                remaining_budget -= action_costs(action, previous_action, self.uav_specifications)
                if self.use_effective_mission_time:
                    remaining_budget -= run_time
                previous_action = action
                self.eval(run_time=run_time, flight_time=flight_time)
        finally:
            self.close()
