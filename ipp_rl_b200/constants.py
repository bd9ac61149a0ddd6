"""String registries of the hot-path factories (same keys and values as the reference's
constants.py:56-100, so one YAML drives both)."""


class SensorType:
    RGB_CAMERA = "rgb_camera"


class SensorModelType:
    ALTITUDE_DEPENDENT = "altitude_dependent"


class SensorSimulationType:
    GAUSSIAN_RANDOM_FIELD = "gaussian_random_field"
    HOTSPOT_RANDOM_FIELD = "hotspot_random_field"
    SPLIT_RANDOM_FIELD = "split_random_field"
    TEMPERATURE_DATA_FIELD = "temperature_data_field"


SENSOR_TYPES = [SensorType.RGB_CAMERA]
SENSOR_MODELS = [SensorModelType.ALTITUDE_DEPENDENT]
SENSOR_SIMULATIONS = [
    SensorSimulationType.GAUSSIAN_RANDOM_FIELD,
    SensorSimulationType.HOTSPOT_RANDOM_FIELD,
    SensorSimulationType.SPLIT_RANDOM_FIELD,
    SensorSimulationType.TEMPERATURE_DATA_FIELD,
]

# required config keys per registry entry (reference constants.py SensorParams / SensorModelParams /
# SensorSimulationParams)
REQUIRED_KEYS = {
    ("sensor", SensorType.RGB_CAMERA): ["field_of_view", "encoding"],
    ("model", SensorModelType.ALTITUDE_DEPENDENT): ["coeff_a", "coeff_b"],
    ("simulation", SensorSimulationType.GAUSSIAN_RANDOM_FIELD): ["cluster_radius"],
    ("simulation", SensorSimulationType.HOTSPOT_RANDOM_FIELD): ["cluster_radius"],
    ("simulation", SensorSimulationType.SPLIT_RANDOM_FIELD): ["cluster_radius"],
    ("simulation", SensorSimulationType.TEMPERATURE_DATA_FIELD): ["filename"],
}


# ---- missions (reference constants.py:92-236) -------------------------------------------------------------
class MissionType:
    CONICAL_SPIRAL = "conical_spiral"
    LAWNMOWER = "lawnmower"
    RANDOM_CONTINUOUS = "random_continuous"
    RANDOM_DISCRETE = "random_discrete"
    GREEDY = "greedy"
    MCTS = "mcts"
    IPP_MASHA = "ipp_masha"
    MCTS_ZERO = "mcts_zero"


class MissionParams:
    STATIC_MISSION = ["dist_to_boundaries", "min_altitude", "max_altitude", "budget", "adaptive", "value_threshold", "interval_factor",
                      "config_name"]
    CONICAL_SPIRAL = ["num_waypoints", "slope_factor"]
    LAWNMOWER = ["step_size", "altitude_spacing"]
    RANDOM_CONTINUOUS = []
    RANDOM_DISCRETE = ["altitude_spacing"]
    GREEDY = ["num_waypoints", "altitude_spacing"]
    MCTS = ["altitude_spacing", "num_simulations", "gamma", "c", "episode_horizon", "k", "alpha", "epsilon_expand", "epsilon_rollout",
            "max_greedy_radius", "use_gcb_rollout"]
    IPP_MASHA = ["episode_horizon", "altitude_spacing", "cmaes_max_iter", "cmaes_sigma0", "cmaes_population_size"]
    MCTS_ZERO = [
        "altitude_spacing", "episode_horizon", "model_deployment_filename", "train_examples_iter", "restart_training",
        "telegram_notifications",
        {"hyper_params": [
            "gamma", "puct_init", "puct_init_decay", "puct_init_min", "puct_base", "forced_playout_factor", "num_mcts_simulations",
            "max_valid_action_distance", "max_episode_steps", "temperature_threshold", "num_self_play_iterations", "num_episodes",
            "start_train_examples_history", "train_examples_history_step", "max_train_examples_history", "num_arena_games",
            "network_update_threshold", "learning_rate", "max_learning_rate", "weight_decay", "num_epochs", "batch_size",
            "input_channels", "use_fov_input", "use_action_costs_input", "num_channels", "num_encoder_res_blocks",
            "num_policy_head_conv_bn_blocks", "num_value_head_conv_bn_blocks", "shared_network", "dropout", "max_grad_norm",
            "lr_step_size", "lr_decay", "policy_loss_coeff", "value_loss_coeff", "reward_loss_coeff", "reconstruction_loss_coeff",
            "entropy_regularization_coeff", "dirichlet_alpha", "dirichlet_alpha_decay", "dirichlet_alpha_min", "dirichlet_eps",
            "continuous_network_update", "reset_mcts_each_step", "momentum", "temperature_scale", "shuffle_train_env_intervals",
            "shuffle_budget", "shuffle_prior_cov", "num_workers", "max_inference_batch_size", "max_waiting_time", "non_blocking_read",
            "use_autoencoder", "use_reward_target", "replay_alpha", "replay_beta0", "use_per", "mask_policy_head", "use_silu",
            "use_separable_conv_layers", "num_augmented_samples", "input_history_length", "log_network_parameters",
            "use_global_context_mixing", "num_global_pooling_channels",
        ]},
    ]


BASELINE_MISSION_TYPES = ["conical_spiral", "lawnmower", "random_continuous", "random_discrete"]
MISSION_TYPES = BASELINE_MISSION_TYPES + ["greedy", "mcts", "ipp_masha", "mcts_zero"]
UAV_PARAMS = ["max_v", "max_a", "sampling_time"]
