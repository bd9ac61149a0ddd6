"""MissionFactory (reference planning/mission_factories.py:19-130): string dispatch on ``mission.type`` with the same
required-key validation (``logger.error`` + bare ``ValueError``).  Missions on the hot path of this engine are provided —
``greedy`` and the deploy-time ``mcts_zero`` planner; the baselines and the classical MCTS / CMA-ES planners of the reference
are control flow over the same Mapping surface (out of scope) and raise NotImplementedError by name."""
import logging
from typing import Dict

from ..constants import MISSION_TYPES, UAV_PARAMS, MissionParams, MissionType
from .missions import Mission

logger = logging.getLogger(__name__)

_PARAM_NAMES = {
    MissionType.LAWNMOWER: MissionParams.LAWNMOWER,
    MissionType.CONICAL_SPIRAL: MissionParams.CONICAL_SPIRAL,
    MissionType.RANDOM_CONTINUOUS: MissionParams.RANDOM_CONTINUOUS,
    MissionType.RANDOM_DISCRETE: MissionParams.RANDOM_DISCRETE,
    MissionType.GREEDY: MissionParams.GREEDY,
    MissionType.MCTS: MissionParams.MCTS,
    MissionType.IPP_MASHA: MissionParams.IPP_MASHA,
    MissionType.MCTS_ZERO: MissionParams.MCTS_ZERO,
}


class MissionFactory:
    def __init__(self, params: Dict, mapping, use_effective_mission_time: bool):
        self.params = params
        self.mapping = mapping
        self.use_effective_mission_time = use_effective_mission_time
        self.mission_params = self.get_mission_params()

    def get_mission_params(self) -> Dict:
        if self.mission_type not in MISSION_TYPES:
            logger.error(f"'{self.mission_type}' not in list of known missions: {MISSION_TYPES}")
            raise ValueError
        param_names = MissionParams.STATIC_MISSION + _PARAM_NAMES[self.mission_type]
        params = dict()
        for param in param_names:
            if isinstance(param, dict):
                for key in param.keys():
                    params[key] = {}
                    if key not in self.params["mission"].keys():
                        logger.error(f"Cannot find '{key}' section for mission '{self.mission_type}' in config file!")
                        raise ValueError
                    for sub_param in param[key]:
                        if sub_param not in self.params["mission"][key].keys():
                            logger.error(f"Cannot find '{sub_param}' in '{key}' for mission '{self.mission_type}' in config file!")
                            raise ValueError
                        params[key][sub_param] = self.params["mission"][key][sub_param]
            else:
                if param not in self.params["mission"].keys():
                    logger.error(f"Cannot find '{param}' parameter for mission '{self.mission_type}' in config file!")
                    raise ValueError
                params[param] = self.params["mission"][param]
        params["mapping"] = self.mapping
        params["uav_specifications"] = self.get_uav_params()
        params["use_effective_mission_time"] = self.use_effective_mission_time
        return params

    def get_uav_params(self) -> Dict:
        params = dict()
        for param in UAV_PARAMS:
            if param not in self.uav_specifications.keys():
                logger.error(f"Cannot find '{param}' parameter for uav specification in config file!")
                raise ValueError
            params[param] = self.uav_specifications[param]
        return params

    @property
    def mission_type(self) -> str:
        if "mission" not in self.params.keys():
            logger.error("Cannot find mission specification in config file!")
            raise ValueError
        if "type" not in self.params["mission"].keys():
            logger.error("Cannot find mission type specification in config file!")
            raise ValueError
        return self.params["mission"]["type"]

    @property
    def uav_specifications(self) -> Dict:
        if "experiment" not in self.params.keys():
            logger.error("Cannot find experiment specification in config file!")
            raise ValueError
        if "uav" not in self.params["experiment"].keys():
            logger.error("Cannot find uav specification in config file!")
            raise ValueError
        return self.params["experiment"]["uav"]

    def create_mission(self) -> Mission:
        if self.mission_type not in MISSION_TYPES:
            logger.error(f"'{self.mission_type}' not in list of known mission types: {MISSION_TYPES}")
            raise ValueError
        if self.mission_type == MissionType.GREEDY:
            from .greedy_mission import GreedyMission

            return GreedyMission(**self.mission_params)
        if self.mission_type == MissionType.MCTS_ZERO:
            from .mcts_zero.mcts_zero_mission import MCTSZeroMission

            return MCTSZeroMission(**self.mission_params)
        raise NotImplementedError(
            f"mission type '{self.mission_type}' is planner control flow outside the engine's hot path (SURVEY.md section 2); "
            "run the reference's mission class over this package's Mapping / Sensor objects instead")
