"""Map-quality metrics (reference planning/evaluation_metrics.py:4-58) for diagonal covariances.

These host functions keep the reference signatures for single maps; ``BatchedEngine.eval`` computes
the same quantities for every env of a batch on the device (eval_kernel in csrc/ipp_engine.cu)."""
import numpy as np

from ..mapping.grid_maps import covariance_diagonal


def root_mean_squared_error(ground_truth_map, estimated_map, adaptive_msk=None) -> float:
    err = np.square(np.asarray(ground_truth_map) - np.asarray(estimated_map))
    if adaptive_msk is not None:
        err = err.flatten(order="C")[adaptive_msk]
    return np.sqrt(np.mean(err))


def map_uncertainty(estimated_map_covariance_matrix, adaptive_msk=None) -> float:
    d = covariance_diagonal(estimated_map_covariance_matrix)
    return np.sum(d if adaptive_msk is None else d[adaptive_msk])


def map_uncertainty_difference(estimated_map_covariance_matrix, adaptive_msk) -> float:
    d = covariance_diagonal(estimated_map_covariance_matrix)
    inside, outside = np.mean(d[adaptive_msk]), np.mean(d[~adaptive_msk])
    return (outside - inside) / outside


def _value_weights(ground_truth_map, estimated_map):
    w = (ground_truth_map - np.min(estimated_map)) / (np.max(ground_truth_map) - np.min(ground_truth_map))
    return w / np.sum(w)


def weighted_root_mean_squared_error(ground_truth_map, estimated_map) -> float:
    return np.sqrt(np.mean(_value_weights(ground_truth_map, estimated_map) * np.square(ground_truth_map - estimated_map)))


def _log_loss(ground_truth_map, estimated_map, cov):
    p = covariance_diagonal(cov, int(np.size(estimated_map))).reshape(np.shape(estimated_map))
    return 0.5 * np.log(2 * np.pi * p) + np.square(ground_truth_map - estimated_map) / 2 * p  # "* p" as in the reference (:44)


def mean_log_loss(ground_truth_map, estimated_map, estimated_map_covariance_matrix) -> float:
    return np.mean(_log_loss(ground_truth_map, estimated_map, estimated_map_covariance_matrix))


def weighted_mean_log_loss(ground_truth_map, estimated_map, estimated_map_covariance_matrix) -> float:
    return np.mean(_value_weights(ground_truth_map, estimated_map) * _log_loss(ground_truth_map, estimated_map, estimated_map_covariance_matrix))
