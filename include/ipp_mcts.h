/*
 * ipp_mcts.h — C ABI of the batched MCTS-zero rollout loop on top of the batched IPP engine
 * (ipp_b200.h).  One tree per env, all trees advanced in lock-step, tree statistics and the
 * prediction-step rollouts resident on the GPU.
 *
 * Replaces, for `n_trees` envs at once, the reference's per-process search
 *   MCTS.get_policy / simulate / compute_uct / get_next_actions_mask / normalize_q_values
 *   (planning/mcts_zero/mcts.py:83-296)
 * whose inner step is simulate_prediction_step (planning/common/optimization.py:14-30; here:
 * ipp_rollout_device, a whole tree path per warp, no state written).  The policy/value network is
 * NOT part of this library (stock PyTorch in the reference, planning/mcts_zero/networks/): leaf
 * evaluation is a call-out between ipp_mcts_simulate_begin and ipp_mcts_simulate_end.
 *
 * Differences from the reference, on purpose:
 *   - a node is identified by its PATH from the root, not by hash(str(covariance))
 *     (mcts.py:20-21: NumPy summarises arrays > 1000 elements, so unrelated states collide);
 *   - candidate actions of a node are the lattice cells within max_valid_action_distance of it
 *     (the reference masks all other actions anyway, mcts.py:148-158), stored as a fixed window of
 *     levels x D x D slots, D = 2*(floor(max_valid_action_distance / resolution) + 1) + 1;
 *     slot order = action-id order restricted to the window: ((level * D) + dcol) * D + drow;
 *   - ties in arg max are broken towards the lowest action id (reference: np.random.choice);
 *   - a fresh tree per ipp_mcts_begin (reference: hyper_params["reset_mcts_each_step"] = true).
 */
#ifndef IPP_MCTS_H
#define IPP_MCTS_H

#include "ipp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define IPP_MCTS_MAX_PATH 8 /* edges per simulation = episode_horizon + 1 <= 8 */

/* leaf kinds written by ipp_mcts_simulate_begin */
#define IPP_MCTS_LEAF_TERMINAL 0 /* depth > episode_horizon, or budget <= 0: value 0 (mcts.py:175-176) */
#define IPP_MCTS_LEAF_EVAL 1     /* node needs (policy, value) from the evaluator (mcts.py:185-237) */

typedef struct ipp_mcts ipp_mcts;

typedef struct ipp_mcts_config {
    uint32_t struct_bytes;            /* = sizeof(ipp_mcts_config) */
    int32_t n_trees;                  /* trees <-> envs [first_env, first_env + n_trees) of the engine */
    int32_t first_env;
    int32_t num_simulations;          /* hyper_params["num_mcts_simulations"]; node capacity = this + 1 */
    int32_t episode_horizon;          /* meta_data["episode_horizon"] (mcts.py:45), 0 .. IPP_MCTS_MAX_PATH-1 */
    uint32_t step_flags;              /* IPP_REWARD_* | IPP_FLAG_ADAPTIVE used by the rollouts */
    double puct_init;                 /* hyper_params["puct_init"]  (mcts.py:282) */
    double puct_base;                 /* hyper_params["puct_base"] */
    double gamma;                     /* hyper_params["gamma"] (mcts.py:248) */
    double forced_playout_factor;     /* hyper_params["forced_playout_factor"] (mcts.py:286-291) */
    double max_valid_action_distance; /* hyper_params["max_valid_action_distance"] (mcts.py:148-158) */
    double dirichlet_eps;             /* hyper_params["dirichlet_eps"]; noise itself is supplied by the caller */
} ipp_mcts_config;

typedef struct ipp_mcts_info {
    int32_t n_trees, max_nodes, levels, window_dim /* D */, window_radius, window_slots /* W = levels*D*D */;
    int32_t max_path;       /* episode_horizon + 1 */
    int32_t simulations;    /* simulations run since ipp_mcts_begin */
    uint64_t device_bytes;  /* HBM held by the trees */
    uint64_t launches;      /* kernels launched by this object (rollouts included) */
    uint64_t edges;         /* edges in the trees' pools = prediction steps actually computed since ipp_mcts_begin (the rollouts are
                               memoised: one per new edge; the reference replays one per level of every simulation).  Filling it
                               reads n_trees counters back: ipp_mcts_get_info synchronises the engine's stream once begun */
} ipp_mcts_info;

/* leaf record per tree, int32[8]: {kind, node, centre col, centre row, level (-1: root off-lattice), depth,
 * float bits of the remaining budget, number of edges on the path} */
#define IPP_MCTS_LEAF_WORDS 8

int ipp_mcts_create(ipp_engine *env, const ipp_mcts_config *cfg, ipp_mcts **out);
void ipp_mcts_destroy(ipp_mcts *m);
const char *ipp_mcts_last_error(const ipp_mcts *m);
int ipp_mcts_get_info(const ipp_mcts *m, ipp_mcts_info *out);

/* Start a search from every env's current belief.  root_poses[n_trees][3] (host, fp64; NULL -> the
 * engine's stored previous actions) = previous_action, budgets[n_trees] (host) = remaining budget
 * (MCTS.get_policy arguments, mcts.py:83-92).  Root depth is 0 as at every reference call site. */
int ipp_mcts_begin(ipp_mcts *m, const double *root_poses, const float *budgets);

/* First half of one lock-step simulation (mcts.py:166-265): PUCT descent with forced playouts at
 * the root, creation of the child node of a new edge, and the path's prediction-step rewards — memoised:
 * an edge keeps the reward of its step, a node the variances its step left on its footprint, so a simulation
 * computes the path's NEW step only (the reference replays one simulate_prediction_step per level, mcts.py:239-246;
 * same values).  The belief of the searched envs must not change between ipp_mcts_begin and the last simulation.
 * leaf_info (host, may be NULL): int32[n_trees][IPP_MCTS_LEAF_WORDS]. */
int ipp_mcts_simulate_begin(ipp_mcts *m, int32_t *leaf_info);

/* Second half: expand the leaves with the evaluator's output and back the values up.
 * Exactly one of priors_window[n_trees][W] (slot order of the LEAF node's window) and
 * priors_dense[n_trees][num_actions] may be non-NULL; both NULL -> uniform priors.  values[n_trees]
 * (NULL -> 0).  root_noise[n_trees][W] (NULL -> none): Dirichlet sample restricted to the root window
 * (entries of a draw over ALL actions), applied to the root's first expansion with weight
 * dirichlet_eps (mcts.py:160-164,225-226).  *_is_device != 0 -> device pointers. */
int ipp_mcts_simulate_end(ipp_mcts *m, const float *priors_window, const float *priors_dense, const float *values,
                          const float *root_noise, int32_t inputs_are_device);

/* Root statistics (host arrays [n_trees][W], any may be NULL): masked normalised priors Ps (-1 for
 * invalid slots, mcts.py:220-234), Qsa, Nsa, and the action id of every slot (-1 outside the grid);
 * ns[n_trees] = Ns of the root. */
int ipp_mcts_root_stats(ipp_mcts *m, float *ps, float *qsa, int32_t *nsa, int32_t *action_ids, int32_t *ns);

/* The paths of the simulation in flight (between simulate_begin and simulate_end; host arrays, either may be NULL):
 * actions[n_trees][max_path] action ids root -> leaf (-1 padded), rewards[n_trees][max_path] of their prediction steps
 * (entries past a path's length are unspecified). */
int ipp_mcts_get_paths(ipp_mcts *m, int32_t *actions, float *rewards);

/* Device views for evaluators that stay on the GPU (valid until ipp_mcts_destroy). */
#define IPP_MCTS_PTR_LEAF_INFO 0    /* int32[n_trees][IPP_MCTS_LEAF_WORDS] */
#define IPP_MCTS_PTR_PATH_ACTIONS 1 /* int32[n_trees][max_path], -1 padded: root -> leaf action ids */
#define IPP_MCTS_PTR_PATH_REWARDS 2 /* float[n_trees][max_path] */
void *ipp_mcts_device_ptr(ipp_mcts *m, int32_t which);

#ifdef __cplusplus
}
#endif
#endif /* IPP_MCTS_H */
