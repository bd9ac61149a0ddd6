/*
 * ipp_b200.h — C ABI of the B200-native batched IPP environment engine.
 *
 * Drop-in boundary for the per-step hot path of dmar-bonn/ipp-rl.  The reference has no FFI of
 * its own (it is pure Python); each entry point below names the reference interface it replaces
 * (path:line under the reference tree) and is what a ctypes binding on the reference side binds
 * (see INTEGRATION.md).  Plain pointers and sizes only — no torch / numpy types.
 *
 * One engine <-> one CUDA device <-> one stream.  Not thread-safe, not fork-safe.  An engine owns
 * `batch` independent environment instances ("envs"), each with three row-major (y_dim, x_dim)
 * fp32 maps resident in HBM: ground truth, belief mean, belief variance (the diagonal of the
 * reference's covariance matrix, mapping/grid_maps.py:10-11).
 *
 * Conventions (reference): pose = [x, y, h] metres; x <-> column, y <-> row; flat cell index
 * x_dim*row + col (sensors/models/sensor_models.py:85).
 *
 * All functions return IPP_OK (0) or a negative error code; ipp_last_error() gives the message.
 * "host" entry points take host pointers, copy in/out and return after the work has completed;
 * "_device" entry points take device pointers, enqueue on the engine stream and return at once.
 */
#ifndef IPP_B200_H
#define IPP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPP_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------------------------ */
#define IPP_OK 0
#define IPP_ERR_INVALID (-1)     /* bad argument / configuration (reference: logger.error + ValueError) */
#define IPP_ERR_CUDA (-2)        /* CUDA runtime failure (sticky errors surface at ipp_sync) */
#define IPP_ERR_NOMEM (-3)       /* device or pinned-host allocation failed */
#define IPP_ERR_UNSUPPORTED (-4) /* e.g. INTER_AREA footprint with an up-sampling axis */

/* ---- step flags -------------------------------------------------------------------------- */
#define IPP_REWARD_TRACE 0u          /* planning/common/rewards.py:15-31  (reference, pinned) */
#define IPP_REWARD_GAUSS_ENTROPY 1u  /* 0.5*sum ln(v/v') / (cost+1)       (extension, unpinned) */
#define IPP_REWARD_MASK 3u
#define IPP_FLAG_ADAPTIVE 4u         /* adaptive mask, planning/common/rewards.py:8-12 */
#define IPP_FLAG_NO_DSIZE_QUIRK 8u   /* un-swap cv2 dsize (simulations/sensor_manipulations.py:20-22) */
#define IPP_FLAG_LOGODDS 16u         /* log-odds occupancy fusion + Shannon entropy (extension) */
#define IPP_FLAG_NO_COMMIT 32u       /* predict: compute rewards only, leave the variance untouched */
#define IPP_FLAG_KEEP_PREV 64u       /* do not advance the stored previous action */

/* ---- memory layouts of the belief in HBM --------------------------------------------------- */
#define IPP_LAYOUT_PLANES 0 /* mean[B][Y][X], var[B][Y][X], gt[B][Y][X] */
#define IPP_LAYOUT_MV 1     /* {mean,var}[B][Y][X] interleaved float2, gt[B][Y][X] */
#define IPP_LAYOUT_TILED 2  /* every 128-byte line of HBM holds a 2-D tile: {mean,var} float2 in 4x4-cell tiles, gt float
                               in 8(x) x 4(y)-cell tiles; tiles row-major over the map (padded to whole tiles), cells
                               row-major inside a tile.  L2 fetches whole 128 B lines from DRAM, so a footprint pulls in
                               ~1.4x its own bytes instead of ~2.1x with row-major maps.  Internal to the engine: the
                               state / ground-truth entry points still exchange dense [n][y_dim][x_dim] arrays. */
#define IPP_LAYOUT_SUPER 3  /* 192-byte super-tiles: the {mean,var} float2 of a 4x4-cell tile (128 B) followed by the tile's
                               16 ground-truth floats (64 B); super-tiles row-major over the map.  The tiles a footprint
                               touches in one tile row are ONE contiguous run of ~0.6-1.3 KB holding everything the fused
                               step reads: the persistent step kernel stages it with one cp.async.bulk (TMA) per tile row
                               and DRAM serves ~1 KB bursts instead of 128-byte lines.  Internal like IPP_LAYOUT_TILED. */
#define IPP_LAYOUT_SPLIT 4  /* the super-tile cut in two arrays: var[B][tiles][16] (64 B per 4x4-cell tile) and
                               {mean x 16 | gt x 16}[B][tiles] (128 B per tile, line aligned).  The fused step stages two
                               runs per tile row (the same bytes as IPP_LAYOUT_SUPER); the covariance-only paths
                               (simulate_prediction_step: ipp_predict, the tree search's rollouts) stage and write the
                               variance run alone: 4 + 4 bytes per cell instead of dragging mean and ground truth along.
                               Internal like IPP_LAYOUT_TILED. */

/* ---- cost model (planning/common/actions.py:8-41) ---------------------------------------- */
#define IPP_COST_DISTANCE 0    /* uav_specifications is None -> Euclidean distance */
#define IPP_COST_FLIGHT_TIME 1 /* trapezoidal velocity profile */

#define IPP_MAX_ALTITUDE_LEVELS 32
#define IPP_NUM_METRICS 8

typedef struct ipp_engine ipp_engine;

/* Configuration = the keys of the reference's YAML that the hot path consumes
 * (config/example.yaml; SURVEY.md section 5 "Config / flags"). */
typedef struct ipp_config {
    uint32_t struct_bytes; /* = sizeof(ipp_config): ABI guard */
    uint32_t abi_version;  /* = IPP_ABI_VERSION */
    int32_t device;        /* CUDA device ordinal */
    int32_t batch;         /* number of envs on this device */
    int32_t x_dim;         /* environment.x_dim  [cells]  (mapping/grid_maps.py:13-24) */
    int32_t y_dim;         /* environment.y_dim  [cells]  (mapping/grid_maps.py:26-37) */
    int32_t layout;        /* IPP_LAYOUT_* */
    int32_t cost_mode;     /* IPP_COST_* */
    double resolution;     /* environment.resolution [m/cell] (mapping/grid_maps.py:39-50) */
    double angle_x_deg;    /* sensor.field_of_view.angle_x (sensors/cameras.py:25-27) */
    double angle_y_deg;    /* sensor.field_of_view.angle_y (sensors/cameras.py:29-31) */
    double tan_half_x;     /* tan(0.5*radians(angle_x)); 0 -> computed here with libm.  The Python
                              host passes NumPy's value so floor() never flips (cameras.py:44-45) */
    double tan_half_y;
    double coeff_a;        /* sensor.model.coeff_a (sensors/models/sensor_models.py:27-30) */
    double coeff_b;        /* sensor.model.coeff_b */
    double rf_altitude;    /* resolution factor 2 above this altitude; reference hard-codes 10.0
                              (sensors/cameras.py:122-125) */
    double min_altitude;   /* experiment.constraints.* -> action table (planning/common/actions.py:73-91) */
    double max_altitude;
    double altitude_spacing;
    double max_v;          /* experiment.uav.max_v  (planning/common/actions.py:32-41) */
    double max_a;          /* experiment.uav.max_a */
    double value_threshold; /* experiment.scenario.value_threshold (planning/common/rewards.py:8-12) */
    double interval_factor; /* experiment.scenario.interval_factor */
    uint64_t seed;         /* counter-based device RNG seed (throughput mode) */
    int64_t env_id_offset; /* global id of local env 0 (env-batch sharding: the RNG stream of an env
                              does not depend on how the batch is split across GPUs) */
    void *stream;          /* optional caller-owned cudaStream_t; NULL -> the engine creates one */
} ipp_config;

typedef struct ipp_info {
    int32_t batch, x_dim, y_dim, layout;
    int32_t num_altitude_levels; /* int((max-min)/spacing)+1 */
    int32_t num_actions;         /* levels * x_dim * y_dim */
    int32_t max_measurements;    /* upper bound of measurements per step (noise row length) */
    int32_t sm_count;
    uint64_t launches;           /* kernels launched by this engine so far */
    uint64_t steps;              /* ipp_step calls so far (RNG counter) */
    uint64_t device_bytes;       /* HBM held by the engine */
    double altitude[IPP_MAX_ALTITUDE_LEVELS];
    int32_t radius_x[IPP_MAX_ALTITUDE_LEVELS]; /* footprint half-width in cells per level */
    int32_t radius_y[IPP_MAX_ALTITUDE_LEVELS];
} ipp_info;

/* ---- life cycle ---------------------------------------------------------------------------- */

/* Replaces GridMap(params) + SensorModelFactory/SensorFactory/SimulationFactory/Mapping
 * construction (experiments/experiments.py:154-168) for `batch` envs.  Allocates HBM; maps are
 * uninitialised until ipp_reset + ipp_set_ground_truth / ipp_synth_ground_truth. */
int ipp_create(const ipp_config *cfg, ipp_engine **out);
void ipp_destroy(ipp_engine *e);
const char *ipp_last_error(const ipp_engine *e); /* e == NULL: message of the last failed ipp_create */
int ipp_get_info(const ipp_engine *e, ipp_info *out);
int ipp_sync(ipp_engine *e);

/* Replaces Mapping.init_priors (mapping/mappings.py:217-261) in its diagonal restriction:
 * mean <- prior_mean (reference: 0.5), var <- prior_var (diag of the GP prior = signal_variance),
 * or per-env prior_var_per_env[batch] (host, shuffle_prior_cov) when not NULL.  Also sets every
 * env's previous action to init_pose[3] (host; NULL -> [2, 2, 14], planning/missions.py:69) and
 * zeroes the step counter. */
int ipp_reset(ipp_engine *e, float prior_mean, float prior_var, const float *prior_var_per_env,
              const double *init_pose);

/* Replaces Simulation.ground_truth_map assignment (simulations/__init__.py:15-16,
 * simulations/simulations.py:41): upload n_env maps [n_env][y_dim][x_dim] fp32 for envs
 * [first_env, first_env+n_env).  src_is_device != 0 -> gt is a device pointer. */
int ipp_set_ground_truth(ipp_engine *e, const float *gt, int32_t first_env, int32_t n_env,
                         int32_t src_is_device);

/* Piecewise-constant ground truths generated ON THE DEVICE for envs [first_env, first_env + n_env): replaces
 * HotspotRandomField / SplitRandomField.create_ground_truth_map (simulations/simulations.py:57-92, :102-125).  Every env
 * takes the reference's sequence of decisions (values, centres / cut, rejection of overlapping hot spots) with uniforms
 * from Philox4x32-10 keyed by (seed; global env id, draw index) instead of NumPy's global MT19937: statistical parity,
 * independent of how the batch is sharded. */
#define IPP_FIELD_HOTSPOT 1
#define IPP_FIELD_SPLIT 2
int ipp_generate_field(ipp_engine *e, int32_t kind, int32_t cluster_radius, uint64_t seed, int32_t first_env, int32_t n_env);

/* ipp_reset with the per-env prior of Mapping.init_priors(shuffle_prior_cov=True) (mapping/mappings.py:219-240), drawn on
 * the device (Philox keyed by (seed; global env id)): gp_mode != 0: variance = U(0.8, 1.2) * signal_variance per env (the
 * diagonal of the shuffled Matern prior; p0 = signal_variance); gp_mode == 0: prior_cov_mean ~ U(0.1, p0 = prior_cov_mean),
 * prior_cov_std = prior_cov_mean, per-cell variance = the diagonal of A A^T / ||A||_F in its large-N normal limit
 * (level sqrt(2) mu, relative spread sqrt(6 / N) / 2 per cell).  Statistical parity. */
int ipp_reset_shuffled(ipp_engine *e, float prior_mean, int32_t gp_mode, float p0, uint64_t seed, const double *init_pose);

/* Synthetic ground truth generated on the device for benchmarking (smooth random harmonic field in
 * [0,1] per env, a stand-in for simulations/ground_truths.py:14-33; not a parity item). */
int ipp_synth_ground_truth(ipp_engine *e, uint64_t seed);

/* Gaussian-random-field ground truth generated ON THE DEVICE for envs [first_env, first_env + n_env) — replaces
 * GaussianRandomField.create_ground_truth_map (simulations/simulations.py:43-48) over
 * gaussian_random_field (simulations/ground_truths.py:14-33): white noise -> fft2 -> * sqrt(k^-cluster_radius)
 * (0 at k = 0) -> ifft2 -> real part -> min-max normalisation to [0, 1].  white_noise: host array
 * [n_env][y_dim][x_dim] of standard normals (the reference's np.random.normal draw; parity mode), or NULL ->
 * counter-based Philox4x32-10 normals keyed by (seed, global env id, cell), independent of how the batch is sharded.
 * The FFTs are cuFFT (library; loaded with dlopen at first use -> IPP_ERR_UNSUPPORTED if absent). */
int ipp_generate_ground_truth(ipp_engine *e, double cluster_radius, uint64_t seed, const float *white_noise,
                              int32_t first_env, int32_t n_env);

/* Read / write belief state as dense [n_env][y_dim][x_dim] fp32 arrays (either may be NULL).
 * Replaces reads/writes of grid_map.mean / np.diag(grid_map.cov_matrix)
 * (mapping/grid_maps.py:10-11). */
int ipp_get_state(ipp_engine *e, float *mean, float *var, int32_t first_env, int32_t n_env,
                  int32_t dst_is_device);
int ipp_set_state(ipp_engine *e, const float *mean, const float *var, int32_t first_env,
                  int32_t n_env, int32_t src_is_device);
int ipp_get_ground_truth(ipp_engine *e, float *gt, int32_t first_env, int32_t n_env,
                         int32_t dst_is_device);
int ipp_set_prev_pose(ipp_engine *e, const double *poses /* [batch][3] host */);
int ipp_get_prev_pose(ipp_engine *e, double *poses /* [batch][3] host */);

/* ---- the hot path ---------------------------------------------------------------------------- */

/* One executed step for every env, ONE fused kernel:
 *   Camera.project_field_of_view            sensors/cameras.py:49-75
 *   ScalarFieldSimulation.take_measurement  simulations/simulations.py:26-34
 *     (+ downsample_measurement / add_model_dependent_gaussian_noise,
 *        simulations/sensor_manipulations.py:7-26,44-57)
 *   Mapping.update_grid_map                 mapping/mappings.py:114-153 (diagonal restriction)
 *   compute_adaptive_msk / compute_reward   planning/common/rewards.py:8-31
 *   action_costs                            planning/common/actions.py:8-41
 * Exactly one of action_ids[batch] (ids of planning/common/actions.py:73-91:
 * id = level*N + x_dim*col + row) and poses[batch][3] (fp64 metres) must be non-NULL.
 * noise: standard normals, row b holds the measurement noise of env b in C order of the
 * measurement array (row stride noise_stride floats); NULL -> counter-based Philox4x32-10 on the
 * device.  reward[batch] receives the information-gain reward of the step.  flags: IPP_REWARD_*,
 * IPP_FLAG_*.  measurements (optional, may be NULL): z of every env, row stride noise_stride. */
int ipp_step(ipp_engine *e, const int32_t *action_ids, const double *poses, const float *noise,
             int32_t noise_stride, float *reward, float *measurements, uint32_t flags);
int ipp_step_device(ipp_engine *e, const int32_t *action_ids, const double *poses,
                    const float *noise, int32_t noise_stride, float *reward, float *measurements,
                    uint32_t flags);

/* Pipelined host steps (action ids only): ipp_step_submit queues, for one of IPP_STEP_SLOTS slots, the upload of the ids
 * (on a copy stream, under whatever kernel is running), the fused step kernel and the delivery of the rewards (written by
 * the kernel straight into `reward` when that is pinned + mapped, else a device->host copy), and returns at once;
 * ipp_step_wait blocks until the step submitted to `slot` is complete (status errors surface here).  Steps execute in
 * submission order.  The caller keeps `action_ids` / `reward` untouched between submit and wait.  With two slots the host
 * prepares and uploads step t+1 while step t computes — what an actor loop that does not need reward t to choose
 * action t+1 (random / open-loop / one-step-stale policies, data collection) gets over the synchronous ipp_step. */
#define IPP_STEP_SLOTS 2
int ipp_step_submit(ipp_engine *e, int32_t slot, const int32_t *action_ids, float *reward, uint32_t flags);
int ipp_step_wait(ipp_engine *e, int32_t slot);

/* The two halves of ipp_step as separate calls (host pointers), for callers that keep the
 * reference's two-call protocol (planning/greedy_mission.py:101-102):
 *   ipp_measure = Sensor.take_measurement(position)            sensors/cameras.py:108-116
 *   ipp_update  = Mapping.update_grid_map(position, data)      mapping/mappings.py:114-153
 * measurements: [batch][stride] fp32, row b = z of env b flattened in C order. */
int ipp_measure(ipp_engine *e, const int32_t *action_ids, const double *poses, const float *noise,
                int32_t stride, float *measurements, uint32_t flags);
int ipp_update(ipp_engine *e, const int32_t *action_ids, const double *poses,
               const float *measurements, int32_t stride, float *reward, uint32_t flags);

/* Rollout / prediction step — replaces simulate_prediction_step
 * (planning/common/optimization.py:14-30): covariance-only update, no ground truth, no noise.
 * n_jobs independent jobs; job j acts on env env_index[j] (NULL -> j, requires n_jobs == batch)
 * with action action_ids[j] / poses[j] and previous action prev_poses[j] (NULL -> the env's
 * stored previous action).  With IPP_FLAG_NO_COMMIT the variance is left untouched (the
 * reference's predict_only=True contract: greedy_search evaluates every candidate from the same
 * state, optimization.py:82-98); otherwise the env's variance advances (rollout descent) and jobs
 * must reference distinct envs. */
int ipp_predict(ipp_engine *e, int32_t n_jobs, const int32_t *env_index, const int32_t *action_ids,
                const double *poses, const double *prev_poses, float *reward, uint32_t flags);
int ipp_predict_device(ipp_engine *e, int32_t n_jobs, const int32_t *env_index,
                       const int32_t *action_ids, const double *poses, const double *prev_poses,
                       float *reward, uint32_t flags);

/* Path rollout — `horizon` (<= 8) chained prediction steps per job, starting from the env's CURRENT
 * belief, without writing any state: what MCTS.simulate does along one tree path, one
 * simulate_prediction_step per level (planning/mcts_zero/mcts.py:239-246 ->
 * planning/common/optimization.py:14-30), where the reference copies the full covariance per level.
 * path_action_ids[n_jobs][horizon]: action ids, a negative id ends the path;
 * prev_poses[n_jobs][3]: pose before the first step (NULL -> the env's stored previous action);
 * rewards[n_jobs][horizon]: per-step reward (0 past the end of a path).  Later steps of a path see the
 * variance the earlier ones produced (kept in shared memory).  flags: IPP_REWARD_*, IPP_FLAG_ADAPTIVE. */
int ipp_rollout(ipp_engine *e, int32_t n_jobs, int32_t horizon, const int32_t *env_index,
                const int32_t *path_action_ids, const double *prev_poses, float *rewards,
                uint32_t flags);
int ipp_rollout_device(ipp_engine *e, int32_t n_jobs, int32_t horizon, const int32_t *env_index,
                       const int32_t *path_action_ids, const double *prev_poses, float *rewards,
                       uint32_t flags);

/* Evaluation metrics per env — replaces Mission.eval (planning/missions.py:176-203) over
 * planning/evaluation_metrics.py:4-58.  metrics[batch][IPP_NUM_METRICS] =
 * {rmse, wrmse, mll, wmll, trace, uncertainty_difference, rmse_masked, trace_masked}; the mask is
 * gt >= value_threshold as in missions.py:179. */
int ipp_eval(ipp_engine *e, float *metrics);
int ipp_eval_device(ipp_engine *e, float *metrics);

/* Observation (network-input) planes of the current belief for envs [first_env, first_env + n_env), NCHW fp32
 * out[n_env][C][y_dim][x_dim] — the per-cell restriction of generate_input_feature_planes
 * (planning/common/features.py:83-151) for one history entry: {variance / max variance (adaptive-masked with
 * IPP_FLAG_ADAPTIVE, :94-99), x / (x_dim*res), y / (x_dim*res) [sic :51], (h - min_alt)/(max_alt - min_alt),
 * budget_ratio} and, with IPP_OBS_COSTS, the min-max normalised action-cost plane (:61-71); C = 5 or 6.
 * poses[n_env][3] (host; NULL -> the envs' stored previous actions), budget_ratio[n_env] (host; NULL -> 1). */
#define IPP_OBS_COSTS 256u
int ipp_observe(ipp_engine *e, int32_t first_env, int32_t n_env, const double *poses, const float *budget_ratio,
                uint32_t flags, float *out, int32_t out_is_device);

/* ---- interop ----------------------------------------------------------------------------------- */
#define IPP_PTR_MEAN 0   /* PLANES: float[B][Y][X];  MV: float2[B][Y][X] base (mean at .x);  TILED / SUPER: float2 tiles;
                            SPLIT: {mean x 16 | gt x 16} tiles */
#define IPP_PTR_VAR 1    /* PLANES: float[B][Y][X];  MV / TILED / SUPER: same base + 1 float (stride 2);  SPLIT: float[16] tiles */
#define IPP_PTR_GT 2
#define IPP_PTR_REWARD 3 /* engine-owned float[batch] staging of the last host-API step */
#define IPP_PTR_STREAM 4 /* the cudaStream_t the engine launches on */
void *ipp_device_ptr(ipp_engine *e, int32_t which);

/* Runtime options.
 * IPP_OPT_STEP_PATH selects the kernel ipp_step uses for Kalman steps on action ids:
 *   IPP_PATH_ASYNC (default) persistent kernel, footprints staged with cp.async, double buffered
 *                            per warp (MV and TILED layouts; falls back to LSU otherwise)
 *   IPP_PATH_LSU             general warp-per-env gather kernel (every mode / input form)
 * ipp_get_option(IPP_OPT_STEP_PATH) returns the path in effect; IPP_OPT_LAUNCHES_* (read only) count
 * the step launches per path.  The environment variable IPP_STEP_PATH=lsu|async sets the default. */
#define IPP_PATH_LSU 0
#define IPP_PATH_ASYNC 1
#define IPP_OPT_STEP_PATH 1
#define IPP_OPT_LAUNCHES_LSU 2
#define IPP_OPT_LAUNCHES_ASYNC 3
/* IPP_OPT_ZERO_COPY: bit mask of the ipp_step host buffers that the kernel accesses IN PLACE when the caller
 * passes page-locked, mapped host memory (ipp_host_alloc, cudaHostAlloc/cudaHostRegister, torch pin_memory):
 * rewards are then written by the fused kernel straight into the caller's buffer (4 B per env over PCIe, inside
 * the kernel) and action ids are read from it, instead of separate stream copies around the launch.  Pageable
 * buffers always take the copy path.  IPP_ZERO_COPY_IDS_FETCH: the persistent step kernel pulls the ids itself, in
 * 512-byte slices over PCIe into device memory while its first footprints are already being planned — no separate
 * H2D copy for the kernel to wait for (16 us at 65 536 envs); needs a 16-byte aligned buffer, other kernels ignore it.
 * Default IPP_ZERO_COPY_REWARDS | IPP_ZERO_COPY_IDS_FETCH; env IPP_ZERO_COPY = any of "r", "i", "f" (or "").
 * IPP_OPT_ZERO_COPY_STEPS (read only) counts the steps whose rewards went out that way, IPP_OPT_IDS_FETCH_STEPS those
 * whose ids the kernel fetched. */
#define IPP_ZERO_COPY_REWARDS 1
#define IPP_ZERO_COPY_IDS 2
#define IPP_ZERO_COPY_IDS_FETCH 4
#define IPP_OPT_IDS_FETCH_STEPS 6
#define IPP_OPT_ZERO_COPY 4
#define IPP_OPT_ZERO_COPY_STEPS 5
int ipp_set_option(ipp_engine *e, int32_t option, int64_t value);
int64_t ipp_get_option(const ipp_engine *e, int32_t option);

/* Static Kalman update on a DIAGONAL covariance for a measurement model given as disjoint blocks (CSR rows: the cells of
 * measurement i are cols[row_ptr[i] .. row_ptr[i+1]), all with weight[i]; noise_var[i] = R_ii).  Replaces the reference's
 * static Mapping.kalman_filter_update(P, H, R, grid_mean, observation, cov_only) (mapping/mappings.py:155-215) for the H / R
 * that AltitudeSensorModel builds (sensors/models/sensor_models.py:32-81).  Stateless (no engine handle), host buffers, fp64
 * on the device.  var[n_cells] is updated in place; mean[n_cells] too when mean and obs are given (cov_only: pass NULL). */
int ipp_kalman_blocks(int32_t device, int32_t n_cells, int32_t n_meas, const int32_t *row_ptr, const int32_t *cols, const double *weight,
                      const double *noise_var, const double *obs, double *var, double *mean);

/* Pinned host memory for the host entry points (cudaHostAlloc). */
int ipp_host_alloc(void **ptr, size_t bytes);
int ipp_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* IPP_B200_H */
