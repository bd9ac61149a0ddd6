"""CPU oracle for the IPP per-step hot path (TEST INFRASTRUCTURE — NOT PRODUCT CODE).

This module is a float64 NumPy restatement of the algorithm the reference
(dmar-bonn/ipp-rl, /root/reference @ 25dfb33) runs on its per-step hot path,
restricted to a *diagonal* covariance ("re-diagonalised Kalman step", DESIGN.md
section 1).  Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product path (``ipp_rl_b200``)
never does: it fails loudly when the CUDA extension is missing.

Pinning status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference in the build container, runs it, and checks this restatement against it
(worst |err| ~1e-15 variance/reward, ~5e-8 measurement — cv2's float32 weights);
the resulting vectors are committed under ``tests/golden/*.npz`` and re-checked
by ``tests/test_oracle_golden.py`` on every run.  The log-odds / Shannon-entropy
and Gaussian-entropy modes have NO reference counterpart ("parity unpinned",
they are this repo's own extension definitions; see the functions' docstrings).

Conventions (reference): pose ``[x, y, h]``; x <-> column, y <-> row; arrays are
row-major ``[row, col]``; flat cell index ``x_dim*row + col``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# configuration mirror (the keys of config/example.yaml that the hot path consumes)
# --------------------------------------------------------------------------------------


@dataclass
class OracleConfig:
    x_dim: int = 10
    y_dim: int = 10
    resolution: float = 4.0
    angle_x: float = 60.0
    angle_y: float = 60.0
    coeff_a: float = 0.05
    coeff_b: float = 0.2
    min_altitude: float = 8.0
    max_altitude: float = 14.0
    altitude_spacing: float = 6.0
    max_v: Optional[float] = 2.0  # None -> Euclidean distance cost (actions.py:8-12)
    max_a: Optional[float] = 2.0
    value_threshold: float = 0.4
    interval_factor: float = 0.0
    rf_altitude: float = 10.0  # cameras.py:125 hard-codes 10.0

    @classmethod
    def from_params(cls, params: Dict) -> "OracleConfig":
        env = params["environment"]
        sen = params["sensor"]
        exp = params.get("experiment", {})
        con = exp.get("constraints", {})
        sce = exp.get("scenario", {})
        uav = exp.get("uav", None)
        return cls(
            x_dim=int(env["x_dim"]),
            y_dim=int(env["y_dim"]),
            resolution=float(env["resolution"]),
            angle_x=float(sen["field_of_view"]["angle_x"]),
            angle_y=float(sen["field_of_view"]["angle_y"]),
            coeff_a=float(sen["model"]["coeff_a"]),
            coeff_b=float(sen["model"]["coeff_b"]),
            min_altitude=float(con.get("min_altitude", 8.0)),
            max_altitude=float(con.get("max_altitude", 14.0)),
            altitude_spacing=float(con.get("altitude_spacing", 6.0)),
            max_v=None if uav is None else float(uav["max_v"]),
            max_a=None if uav is None else float(uav["max_a"]),
            value_threshold=float(sce.get("value_threshold", 0.4)),
            interval_factor=float(sce.get("interval_factor", 0.0)),
        )

    @property
    def uav(self) -> Optional[Dict]:
        if self.max_v is None:
            return None
        return {"max_v": self.max_v, "max_a": self.max_a}


# --------------------------------------------------------------------------------------
# a2/a3/a4  camera footprint                                     sensors/cameras.py
# --------------------------------------------------------------------------------------


def field_of_view_range(cfg: OracleConfig, height: float) -> Tuple[float, float]:
    """Ground footprint [m] from altitude.  Follows sensors/cameras.py:34-47
    (``2 * height * tan(0.5 * radians(angle))``, same operation order)."""
    x_m = 2 * height * np.tan(0.5 * np.radians(cfg.angle_x))
    y_m = 2 * height * np.tan(0.5 * np.radians(cfg.angle_y))
    return float(x_m), float(y_m)


def project_field_of_view(cfg: OracleConfig, position: Sequence[float]) -> Tuple[int, int, int, int]:
    """(xl, xr, yu, yd) inclusive cell rectangle.  Follows sensors/cameras.py:49-75:
    floor(range/res), floor(pos/res), radius=floor(0.5*range_cells), clip to the grid."""
    x_m, y_m = field_of_view_range(cfg, float(position[2]))
    wx = math.floor(x_m / cfg.resolution)
    wy = math.floor(y_m / cfg.resolution)
    cx = math.floor(float(position[0]) / cfg.resolution)
    cy = math.floor(float(position[1]) / cfg.resolution)
    rx = math.floor(0.5 * wx)
    ry = math.floor(0.5 * wy)
    xl = min(max(cx - rx, 0), cfg.x_dim - 1)
    xr = min(max(cx + rx, 0), cfg.x_dim - 1)
    yu = min(max(cy - ry, 0), cfg.y_dim - 1)
    yd = min(max(cy + ry, 0), cfg.y_dim - 1)
    return int(xl), int(xr), int(yu), int(yd)


def resolution_factor(cfg: OracleConfig, position: Sequence[float]) -> int:
    """sensors/cameras.py:122-125: ``2 if altitude > 10.0 else 1``."""
    return 2 if float(position[2]) > cfg.rf_altitude else 1


# --------------------------------------------------------------------------------------
# a5/a6/a7  altitude sensor model                      sensors/models/sensor_models.py
# --------------------------------------------------------------------------------------


def noise_variance(cfg: OracleConfig, position: Sequence[float]) -> float:
    """sensor_models.py:27-30: ``a * (1 - exp(-b * h))``."""
    return float(cfg.coeff_a * (1 - np.exp(-cfg.coeff_b * float(position[2]))))


def measurement_variance(cfg: OracleConfig, position: Sequence[float], rf: int) -> float:
    """Diagonal entry of R.  sensor_models.py:32-36: ``rf**3 * sigma2(h)``."""
    return float(rf ** 3 * noise_variance(cfg, position))


def num_measurements(fov: Tuple[int, int, int, int], rf: int) -> int:
    """mapping/mappings.py:126."""
    xl, xr, yu, yd = fov
    return int(np.ceil((xr - xl + 1) / rf) * np.ceil((yd - yu + 1) / rf))


def measurement_blocks(fov: Tuple[int, int, int, int], rf: int):
    """Row structure of the measurement matrix H as a list of
    ``(row_slice, col_slice, weight)`` in measurement order i.

    Follows sensor_models.py:54-81: ``nbx = floor((xr-xl)/rf)+1``; measurement i ->
    ``(by, bx) = (i // nbx, i % nbx)``; covered cells rows ``yu+[by*rf, min(by*rf+rf, ny))``,
    cols ``xl+[bx*rf, min(bx*rf+rf, nx))``; weight ``1/rf**2``, or ``1/rf`` when the block
    holds fewer than ``rf**2`` cells (:76-79)."""
    xl, xr, yu, yd = fov
    nx, ny = xr - xl + 1, yd - yu + 1
    nbx = math.floor((xr - xl) / rf) + 1
    m = num_measurements(fov, rf)
    out = []
    for i in range(m):
        by = i // nbx
        bx = i - nbx * by
        x_end = min(bx * rf + rf, nx)
        y_end = min(by * rf + rf, ny)
        x_start = min(bx * rf, x_end)
        y_start = min(by * rf, y_end)
        count = (x_end - x_start) * (y_end - y_start)
        w = 1.0 / rf ** 2 if count >= rf ** 2 else 1.0 / rf
        out.append((slice(yu + y_start, yu + y_end), slice(xl + x_start, xl + x_end), w))
    return out


def measurement_model_matrix(cfg: OracleConfig, fov: Tuple[int, int, int, int], rf: int) -> np.ndarray:
    """Dense H (m, N) — only used by tests to compare with the reference's matrix
    (sensor_models.py:38-85)."""
    H = np.zeros((num_measurements(fov, rf), cfg.x_dim * cfg.y_dim))
    for i, (rs, cs, w) in enumerate(measurement_blocks(fov, rf)):
        rows, cols = np.meshgrid(np.arange(rs.start, rs.stop), np.arange(cs.start, cs.stop), indexing="ij")
        H[i, (cfg.x_dim * rows + cols).ravel()] = w
    return H


# --------------------------------------------------------------------------------------
# a8-a11  simulated measurement                simulations/{simulations,sensor_manipulations}.py
# --------------------------------------------------------------------------------------


def inter_area_weights(n_in: int, n_out: int) -> np.ndarray:
    """(n_out, n_in) weights of OpenCV's INTER_AREA decimation along one axis.

    The algorithm is the third-party dependency opencv-python (requirements.txt pins
    4.5.2.54; 4.13 here) — ``cv::computeResizeAreaTab`` / ``resizeAreaFast_``: output
    sample d integrates input over ``[d*s, (d+1)*s)``, ``s = n_in/n_out``; partial overlaps
    below 1e-3 are dropped; weights are float32 divided by the cell width
    ``min(s, n_in - d*s)``.  Anchored on the reference's call site
    simulations/sensor_manipulations.py:20-22 and pinned against cv2 itself in
    tests/test_oracle_golden.py."""
    assert n_out >= 1 and n_in >= n_out, "INTER_AREA decimation needs scale >= 1"
    scale = n_in / n_out
    W = np.zeros((n_out, n_in))
    if n_in % n_out == 0:  # integer scale: cv2's fast path is the plain block mean
        k = n_in // n_out
        for d in range(n_out):
            W[d, d * k : (d + 1) * k] = 1.0 / k
        return W
    for d in range(n_out):
        f1 = d * scale
        f2 = f1 + scale
        cell = min(scale, n_in - f1)
        s1 = math.ceil(f1)
        s2 = min(math.floor(f2), n_in - 1)
        s1 = min(s1, s2)
        if s1 - f1 > 1e-3:
            W[d, s1 - 1] = np.float32((s1 - f1) / cell)
        for s in range(s1, s2):
            W[d, s] = np.float32(1.0 / cell)
        if f2 - s2 > 1e-3:
            W[d, s2] = np.float32(min(min(f2 - s2, 1.0), cell) / cell)
    return W


def downsample_measurement(submap: np.ndarray, rf: int, dsize_quirk: bool = True) -> np.ndarray:
    """simulations/sensor_manipulations.py:7-26.  rf == 1: identity.  rf > 1:
    ``cv2.resize(G, dsize=(ceil(ny/rf), ceil(nx/rf)), INTER_AREA)``; cv2's dsize is
    (width, height), so the output has ``ceil(nx/rf)`` ROWS and ``ceil(ny/rf)`` COLUMNS
    (SURVEY Appendix C #2).  ``dsize_quirk=False`` gives the un-swapped shape."""
    if rf <= 1:
        return submap
    ny, nx = submap.shape
    out_rows = math.ceil(nx / rf) if dsize_quirk else math.ceil(ny / rf)
    out_cols = math.ceil(ny / rf) if dsize_quirk else math.ceil(nx / rf)
    if out_rows > ny or out_cols > nx:
        raise NotImplementedError(
            "INTER_AREA with an up-sampling axis (cv2 switches to its bilinear branch); "
            "cannot occur for square FoV on grids no smaller than the footprint radius"
        )
    Wr = inter_area_weights(ny, out_rows)
    Wc = inter_area_weights(nx, out_cols)
    return Wr @ submap @ Wc.T


def take_measurement(
    cfg: OracleConfig, gt: np.ndarray, position: Sequence[float], eps: np.ndarray, dsize_quirk: bool = True
) -> np.ndarray:
    """simulations/simulations.py:26-34 + sensor_manipulations.py:44-57.
    ``z = clip(D + N(0, scale=sigma2(h)))`` — the noise *variance* is used as the std
    (SURVEY Appendix C #1).  ``eps`` holds the standard normals (C order, same shape as D or
    flat with >= D.size entries): ``np.random.normal(0, s, shape) == s * standard_normal(shape)``
    bitwise under the same seed."""
    xl, xr, yu, yd = project_field_of_view(cfg, position)
    sub = gt[yu : yd + 1, xl : xr + 1]
    D = downsample_measurement(sub, resolution_factor(cfg, position), dsize_quirk)
    e = np.asarray(eps, dtype=np.float64).ravel()[: D.size].reshape(D.shape)
    return np.clip(D + noise_variance(cfg, position) * e, 0.0, 1.0)


def measurement_shape(cfg: OracleConfig, position: Sequence[float], dsize_quirk: bool = True) -> Tuple[int, int]:
    xl, xr, yu, yd = project_field_of_view(cfg, position)
    rf = resolution_factor(cfg, position)
    nx, ny = xr - xl + 1, yd - yu + 1
    if rf == 1:
        return ny, nx
    a, b = math.ceil(nx / rf), math.ceil(ny / rf)
    return (a, b) if dsize_quirk else (b, a)


# --------------------------------------------------------------------------------------
# a12/a13  belief update restricted to a diagonal covariance        mapping/mappings.py
# --------------------------------------------------------------------------------------


def kalman_update_diag(
    cfg: OracleConfig,
    mean: np.ndarray,
    var: np.ndarray,
    position: Sequence[float],
    z: Optional[np.ndarray] = None,
) -> Tuple[Optional[np.ndarray], np.ndarray]:
    """Diagonal of mapping/mappings.py:114-197 (``update_grid_map`` ->
    ``kalman_filter_update``) when the prior covariance is ``diag(var)``.

    With diagonal P, ``S = H P H^T + R`` is diagonal because measurement blocks are disjoint,
    so ``P' = P - P H^T S^-1 H P`` has, for cell j of block i (weight w_i):
    ``v'_j = v_j - (w_i v_j)^2 / S_i``, ``S_i = w_i^2 sum_k v_k + R`` and
    ``mu'_j = mu_j + (w_i v_j / S_i) (z_i - w_i sum_k mu_k)`` (:188-197).  Off-diagonals that
    the dense update creates inside rf=2 blocks are dropped (re-diagonalisation).
    ``z=None`` is the reference's ``cov_only=True`` branch (:199)."""
    fov = project_field_of_view(cfg, position)
    rf = resolution_factor(cfg, position)
    R = measurement_variance(cfg, position, rf)
    var_n = np.array(var, dtype=np.float64, copy=True)
    mean_n = None if z is None else np.array(mean, dtype=np.float64, copy=True)
    zf = None if z is None else np.asarray(z, dtype=np.float64).flatten(order="C")
    for i, (rs, cs, w) in enumerate(measurement_blocks(fov, rf)):
        v = np.asarray(var, dtype=np.float64)[rs, cs]
        S = w * w * v.sum() + R
        var_n[rs, cs] = v - (w * v) ** 2 / S
        if zf is not None:
            mu = np.asarray(mean, dtype=np.float64)[rs, cs]
            mean_n[rs, cs] = mu + (w * v / S) * (zf[i] - w * mu.sum())
    return mean_n, var_n


# --------------------------------------------------------------------------------------
# a14-a18  reward, cost, action table                   planning/common/{rewards,actions}.py
# --------------------------------------------------------------------------------------


def compute_adaptive_msk(mean: np.ndarray, var: np.ndarray, value_threshold: float, interval_factor: float):
    """planning/common/rewards.py:8-12 (variance, not std: Appendix C #5); returned in grid shape."""
    return np.asarray(mean) + interval_factor * np.asarray(var) >= value_threshold


def compute_distance(action, previous_action) -> float:
    """planning/common/actions.py:15-16."""
    return float(np.linalg.norm(np.asarray(action, float) - np.asarray(previous_action, float), ord=2))


def compute_flight_time(action, previous_action, uav: Dict) -> float:
    """planning/common/actions.py:32-41 (trapezoidal velocity profile)."""
    dist_total = compute_distance(action, previous_action)
    dist_acc = min(dist_total * 0.5, np.square(uav["max_v"]) / (2 * uav["max_a"]))
    dist_const = dist_total - 2 * dist_acc
    time_acc = np.sqrt(2 * dist_acc / uav["max_a"])
    time_const = dist_const / uav["max_v"]
    return float(time_const + 2 * time_acc)


def action_costs(action, previous_action, uav: Optional[Dict]) -> float:
    """planning/common/actions.py:8-12."""
    if uav is None:
        return compute_distance(action, previous_action)
    return compute_flight_time(action, previous_action, uav)


def compute_reward(var: np.ndarray, var_next: np.ndarray, cost: float, mask: Optional[np.ndarray] = None) -> float:
    """planning/common/rewards.py:15-31: ``(sum diag P - sum diag P') / (cost + 1)`` over
    the adaptive mask when given."""
    v0 = np.asarray(var, dtype=np.float64)
    v1 = np.asarray(var_next, dtype=np.float64)
    if mask is not None:
        v0, v1 = v0[mask], v1[mask]
    return float((np.sum(v0) - np.sum(v1)) / (cost + 1))


def altitude_levels(cfg: OracleConfig) -> np.ndarray:
    """planning/common/actions.py:74."""
    n = int((cfg.max_altitude - cfg.min_altitude) / cfg.altitude_spacing) + 1
    return np.linspace(cfg.min_altitude, cfg.max_altitude, n)


def enumerate_actions(cfg: OracleConfig) -> np.ndarray:
    """(A, 3) action table, ``A = levels * N``.  planning/common/actions.py:73-100.

    The reference builds positions row-major over (row, col), then files them under
    ``flatten_grid_index(pos_idx2d=[col, row]) = x_dim*col + row`` — a column-major id
    (``acts[1] == [2, 6, 8]`` on example.yaml).  Hence for id
    ``k = h*N + i``: ``col = i // x_dim``, ``row = i % x_dim`` (square grids; for non-square
    grids the reference's ids collide — this oracle reproduces the formula as written)."""
    N = cfg.x_dim * cfg.y_dim
    lv = altitude_levels(cfg)
    acts = np.zeros((len(lv) * N, 3))
    res = cfg.resolution
    for h, alt in enumerate(lv):
        for row in range(cfg.y_dim):
            for col in range(cfg.x_dim):
                i = cfg.x_dim * col + row
                acts[h * N + i] = (res * col + 0.5 * res, res * row + 0.5 * res, alt)
    return acts


# --------------------------------------------------------------------------------------
# a17  rollout step and the executed step
# --------------------------------------------------------------------------------------

REWARD_TRACE = 0  # reference (rewards.py:15-31)
REWARD_GAUSS_ENTROPY = 1  # extension, parity unpinned
REWARD_BERNOULLI_ENTROPY = 2  # extension, parity unpinned (log-odds belief)


def gaussian_entropy_reduction(var, var_next, mask=None) -> float:
    """EXTENSION (no reference counterpart — parity unpinned).  Differential entropy of a
    diagonal Gaussian is ``0.5 * sum ln(2 pi e v)``; the reduction is ``0.5 * sum ln(v / v')``."""
    v0 = np.asarray(var, dtype=np.float64)
    v1 = np.asarray(var_next, dtype=np.float64)
    t = 0.5 * np.log(v0 / v1)
    if mask is not None:
        t = t[mask]
    return float(np.sum(t))


def simulate_prediction_step(
    cfg: OracleConfig,
    var: np.ndarray,
    previous_action,
    action,
    mean: Optional[np.ndarray] = None,
    adaptive: bool = False,
    reward_mode: int = REWARD_TRACE,
):
    """planning/common/optimization.py:14-30 on a diagonal state: adaptive mask from the
    pre-update state (:21-26), cov-only predict (:28), reward (:29).  Returns
    ``(reward, var_next)``."""
    mask = None
    if adaptive:
        mask = compute_adaptive_msk(mean, var, cfg.value_threshold, cfg.interval_factor)
    _, var_next = kalman_update_diag(cfg, mean, var, action, None)
    cost = action_costs(action, previous_action, cfg.uav)
    if reward_mode == REWARD_GAUSS_ENTROPY:
        return gaussian_entropy_reduction(var, var_next, mask) / (cost + 1), var_next
    return compute_reward(var, var_next, cost, mask), var_next


def full_step(
    cfg: OracleConfig,
    gt: np.ndarray,
    mean: np.ndarray,
    var: np.ndarray,
    previous_action,
    action,
    eps: np.ndarray,
    adaptive: bool = False,
    reward_mode: int = REWARD_TRACE,
    dsize_quirk: bool = True,
):
    """One executed step as every planner runs it (e.g. planning/greedy_mission.py:99-103):
    ``z = sensor.take_measurement(a)``; ``mapping.update_grid_map(a, z)``; plus the
    information-gain reward of that same action (rewards.py:15-31) on the pre/post state.
    Returns ``(reward, mean', var', z)``."""
    z = take_measurement(cfg, gt, action, eps, dsize_quirk)
    mask = None
    if adaptive:
        mask = compute_adaptive_msk(mean, var, cfg.value_threshold, cfg.interval_factor)
    mean_n, var_n = kalman_update_diag(cfg, mean, var, action, z)
    cost = action_costs(action, previous_action, cfg.uav)
    if reward_mode == REWARD_GAUSS_ENTROPY:
        r = gaussian_entropy_reduction(var, var_n, mask) / (cost + 1)
    else:
        r = compute_reward(var, var_n, cost, mask)
    return r, mean_n, var_n, z


# --------------------------------------------------------------------------------------
# EXTENSION: log-odds occupancy fusion + Shannon-entropy reward (parity unpinned)
# --------------------------------------------------------------------------------------


def logodds_step(
    cfg: OracleConfig,
    gt: np.ndarray,
    logodds: np.ndarray,
    previous_action,
    action,
    eps: np.ndarray,
    dsize_quirk: bool = True,
    clamp: float = 30.0,
):
    """EXTENSION — the reference has no occupancy / log-odds / Shannon-entropy code
    (SURVEY 0.3), so this is this repo's own definition; parity unpinned.

    Belief: per-cell log-odds ``l`` of "cell value is 1" for a field in [0, 1].  The
    measurement is the reference's own ``take_measurement`` (noisy, block-averaged GT).  With a
    Gaussian likelihood of std ``s`` around the two hypotheses 0 and 1, the likelihood ratio of a
    block reading z is ``exp((2 z - 1) / (2 s^2))``; each covered cell receives
    ``l += (2 z_i - 1) / (2 R)`` with ``R = rf^3 sigma2(h)`` the reference's measurement variance
    (sensor_models.py:32-36), clamped to ``[-clamp, clamp]``.
    Reward: Shannon entropy reduction ``sum H(sigmoid(l)) - H(sigmoid(l'))`` over the footprint,
    divided by ``cost + 1`` like rewards.py:31.  Returns ``(reward, logodds', z)``."""
    z = take_measurement(cfg, gt, action, eps, dsize_quirk)
    fov = project_field_of_view(cfg, action)
    rf = resolution_factor(cfg, action)
    R = measurement_variance(cfg, action, rf)
    zf = z.flatten(order="C")
    l0 = np.asarray(logodds, dtype=np.float64)
    l1 = l0.copy()
    for i, (rs, cs, _w) in enumerate(measurement_blocks(fov, rf)):
        l1[rs, cs] = np.clip(l0[rs, cs] + (2.0 * zf[i] - 1.0) / (2.0 * R), -clamp, clamp)
    dH = float(np.sum(bernoulli_entropy(l0) - bernoulli_entropy(l1)))
    cost = action_costs(action, previous_action, cfg.uav)
    return dH / (cost + 1), l1, z


def bernoulli_entropy(logodds: np.ndarray) -> np.ndarray:
    """Shannon entropy [nats] of Bernoulli(sigmoid(l)), computed stably:
    ``H = softplus(|l|) - |l| * sigmoid(|l|)`` with ``softplus(a) = log1p(exp(-a)) + a``."""
    a = np.abs(np.asarray(logodds, dtype=np.float64))
    e = np.exp(-a)
    return np.log1p(e) + a * e / (1.0 + e)


# --------------------------------------------------------------------------------------
# a19  evaluation metrics                                   planning/evaluation_metrics.py
# --------------------------------------------------------------------------------------

METRIC_NAMES = ("rmse", "wrmse", "mll", "wmll", "uncertainty", "uncertainty_difference", "rmse_masked", "uncertainty_masked")


def evaluation_metrics(gt: np.ndarray, mean: np.ndarray, var: np.ndarray, mask: Optional[np.ndarray] = None) -> np.ndarray:
    """All reductions of planning/evaluation_metrics.py:4-58 as called by
    planning/missions.py:176-203, on a diagonal covariance; order = METRIC_NAMES.
    Quirks kept as written (Appendix C #6): MLL multiplies the squared error by P_ii (:44,:57);
    W* weights use ``min(estimated_map)`` (:34,:53) and may go negative -> NaN in WRMSE."""
    gt = np.asarray(gt, np.float64)
    mean = np.asarray(mean, np.float64)
    var = np.asarray(var, np.float64)
    out = np.full(len(METRIC_NAMES), np.nan)
    sq = np.square(gt - mean)
    out[0] = np.sqrt(np.mean(sq))
    rng = np.max(gt) - np.min(gt)
    with np.errstate(all="ignore"):
        w = (gt - np.min(mean)) / rng
        w = w / np.sum(w)
        out[1] = np.sqrt(np.mean(w * sq))
        ll = 0.5 * np.log(2 * np.pi * var) + sq / 2 * var
        out[2] = np.mean(ll)
        out[3] = np.mean(w * ll)
        out[4] = np.sum(var)
        if mask is not None:
            m = np.asarray(mask, bool)
            vi, vu = var[m], var[~m]
            out[5] = (np.mean(vu) - np.mean(vi)) / np.mean(vu) if vu.size and vi.size else np.nan
            out[6] = np.sqrt(np.mean(sq[m])) if m.any() else np.nan
            out[7] = np.sum(var[m])
    return out


# --------------------------------------------------------------------------------------
# device RNG definition (throughput mode): Philox4x32-10 + Box-Muller
# --------------------------------------------------------------------------------------

_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox4x32-10 (Salmon et al., SC'11 — the Random123 definition; not part of the
    reference, whose RNG is NumPy's global MT19937 at sensor_manipulations.py:57).
    ``counter`` (..., 4) uint32, ``key`` (..., 2) uint32 -> (..., 4) uint32."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    k = np.array(np.broadcast_to(key, c.shape[:-1] + (2,)), dtype=np.uint32, copy=True)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c[..., 0].astype(np.uint64) * _PHILOX_M0
            p1 = c[..., 2].astype(np.uint64) * _PHILOX_M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            n0 = hi1 ^ c[..., 1] ^ k[..., 0]
            n2 = hi0 ^ c[..., 3] ^ k[..., 1]
            c = np.stack([n0, lo1, n2, lo0], axis=-1)
            k = np.stack([k[..., 0] + _PHILOX_W0, k[..., 1] + _PHILOX_W1], axis=-1)
    return c


def device_normals(seed: int, env: int, step: int, n_groups: int) -> np.ndarray:
    """Standard normals of the engine's throughput-mode RNG: (n_groups, 4) float64.

    Group g (one 2x2 cell quad at rf=1, one measurement block at rf=2 — see
    ``device_noise_field``) draws ``philox4x32_10(counter=(g, env, step, 0), key=seed)``;
    uniforms ``u = (x + 0.5) * 2^-32`` in (0, 1); Box-Muller pairs:
    ``n0, n1 = r(u0) * (cos, sin)(pi (2 u1 - 1))``, ``n2, n3 = r(u2) * (cos, sin)(pi (2 u3 - 1))``,
    ``r(u) = sqrt(-2 ln u)`` (angle in [-pi, pi): the range where the GPU's SFU sin/cos are most
    accurate)."""
    ctr = np.zeros((n_groups, 4), dtype=np.uint32)
    ctr[:, 0] = np.arange(n_groups, dtype=np.uint32)
    ctr[:, 1] = np.uint32(env)
    ctr[:, 2] = np.uint32(step & 0xFFFFFFFF)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    x = philox4x32_10(ctr, key).astype(np.float64)
    u = (x + 0.5) * (2.0 ** -32)
    r0 = np.sqrt(-2.0 * np.log(u[:, 0]))
    r1 = np.sqrt(-2.0 * np.log(u[:, 2]))
    a0 = np.pi * (2.0 * u[:, 1] - 1.0)
    a1 = np.pi * (2.0 * u[:, 3] - 1.0)
    return np.stack([r0 * np.cos(a0), r0 * np.sin(a0), r1 * np.cos(a1), r1 * np.sin(a1)], axis=-1)


def device_noise_field(cfg: OracleConfig, position, seed: int, env: int, step: int) -> np.ndarray:
    """The eps array (measurement shape, C order) the engine's Philox mode applies for this
    (env, step).  The footprint is tiled by 2x2 cell quads anchored at (yu, xl), quad index
    ``g = qy * ceil(nx/2) + qx``.  rf=1: cell (2qy+dy, 2qx+dx) uses normal ``[g, 2*dy+dx]``.
    rf=2: measurement i (flat index into z) uses normal ``[(i & 31) + 32 * (i >> 7), (i >> 5) & 3]`` — four
    consecutive 32-blocks of measurements share one Philox call per lane position (quad_math.cuh draw_normals)."""
    xl, xr, yu, yd = project_field_of_view(cfg, position)
    nx, ny = xr - xl + 1, yd - yu + 1
    rf = resolution_factor(cfg, position)
    nqx, nqy = (nx + 1) // 2, (ny + 1) // 2
    nrm = device_normals(seed, env, step, nqx * nqy)
    if rf == 1:
        eps = np.zeros((ny, nx))
        for r in range(ny):
            for c in range(nx):
                eps[r, c] = nrm[(r // 2) * nqx + (c // 2), 2 * (r % 2) + (c % 2)]
        return eps
    shape = measurement_shape(cfg, position)
    m = shape[0] * shape[1]
    i = np.arange(m)
    groups, comp = (i & 31) + 32 * (i >> 7), (i >> 5) & 3
    nrm = device_normals(seed, env, step, int(groups.max()) + 1)
    return nrm[groups, comp].reshape(shape)


# --------------------------------------------------------------------------------------
# batched driver used by tests / CPU baseline
# --------------------------------------------------------------------------------------


@dataclass
class BatchState:
    gt: np.ndarray  # (B, Y, X)
    mean: np.ndarray
    var: np.ndarray
    prev: np.ndarray  # (B, 3) previous action
    step: int = 0


def batched_full_step(
    cfg: OracleConfig,
    st: BatchState,
    actions: np.ndarray,
    eps: Optional[np.ndarray] = None,
    seed: Optional[int] = None,
    env_offset: int = 0,
    adaptive: bool = False,
    reward_mode: int = REWARD_TRACE,
) -> np.ndarray:
    """Apply ``full_step`` to every env in place; returns rewards (B,).  ``eps`` (B, m_max)
    host normals (parity mode) or ``seed`` for the Philox definition above."""
    B = st.gt.shape[0]
    rewards = np.zeros(B)
    for b in range(B):
        if eps is not None:
            e = eps[b]
        else:
            e = device_noise_field(cfg, actions[b], seed, env_offset + b, st.step)
        r, m, v, _ = full_step(cfg, st.gt[b], st.mean[b], st.var[b], st.prev[b], actions[b], e, adaptive, reward_mode)
        rewards[b] = r
        st.mean[b], st.var[b] = m, v
    st.prev = np.array(actions, dtype=np.float64, copy=True)
    st.step += 1
    return rewards
