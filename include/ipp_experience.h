/*
 * ipp_experience.h — C ABI of the device-resident experience store of the batched IPP engine
 * ("experience format + inference batching").
 *
 * Replaces, for whole batches of envs at once, the reference's self-play data path
 *   EpisodeGenerator.execute value targets          planning/mcts_zero/episode_generators.py:158-164
 *   scale_value_target                               planning/common/rewards.py:34-35
 *   save_sample_to_disk (one bz2 pickle per sample)  planning/mcts_zero/episode_generators.py:186-192
 *   ReplayBuffer / ExperienceReplayBuffer /
 *   PrioritizedExperienceReplayBuffer                planning/mcts_zero/replay_buffers.py:15-141
 *   augment_random_crop (ReplicationPad2d(4) + RandomCrop)          replay_buffers.py:58-77
 * by a ring of samples in HBM: a sample = {observation planes float[C][Y][X] (ipp_observe), policy float[P],
 * valid-action mask uint8[P], value target, reward, priority}.  P is the caller's policy width (all actions, or
 * the search window of ipp_mcts.h).  Sampling draws indices on the device from the priorities
 * (np.random.choice semantics: inverse CDF, searchsorted side='right'), gathering copies the sampled rows —
 * optionally shifted with edge replication, the reference's augmentation — into one contiguous training batch.
 * The policy/value network and its optimiser are NOT part of this library (stock PyTorch in the reference).
 *
 * Every array argument is a HOST pointer unless `is_device` is non-zero.  Not thread-safe; one ring <-> one device.
 */
#ifndef IPP_EXPERIENCE_H
#define IPP_EXPERIENCE_H

#include "ipp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ipp_ring ipp_ring;

typedef struct ipp_ring_config {
    uint32_t struct_bytes; /* = sizeof(ipp_ring_config) */
    int32_t device;
    int64_t capacity;      /* samples held; pushes beyond it overwrite the oldest (window_size of replay_buffers.py:33-45) */
    int32_t channels, y_dim, x_dim; /* observation planes per sample */
    int32_t policy_slots;  /* P */
    void *stream;          /* cudaStream_t to launch on (NULL -> own stream) */
} ipp_ring_config;

typedef struct ipp_ring_info {
    int64_t capacity, size, head; /* head = slot the next push writes */
    uint64_t pushed;              /* samples pushed since creation */
    uint64_t device_bytes, launches;
} ipp_ring_info;

int ipp_ring_create(const ipp_ring_config *cfg, ipp_ring **out);
void ipp_ring_destroy(ipp_ring *r);
const char *ipp_ring_last_error(const ipp_ring *r);
int ipp_ring_get_info(const ipp_ring *r, ipp_ring_info *out);

/* Value targets of finished episodes (episode_generators.py:158-164), rewards[n_episodes][max_steps] row-major,
 * lengths[n_episodes] <= max_steps:
 *   values[e][i] = sqrt(1 + sum_{j=i}^{min(i+horizon, len)-1} gamma^j * rewards[e][j]) - 1
 * (the exponent is the ABSOLUTE step j as in the reference, :164; scaling = scale_value_target), 0 past the end;
 * totals[e] (may be NULL) = sum_j gamma^j rewards[e][j] (:158, the value execute() returns).  Stream: r's. */
int ipp_ring_value_targets(ipp_ring *r, const float *rewards, const int32_t *lengths, int32_t n_episodes, int32_t max_steps, double gamma,
                           int32_t horizon, float *values, float *totals, int32_t is_device);

/* Append n samples (any of policy / valid_mask may be NULL -> zeros / ones).  New samples get priority
 * `priority` (<= 0 -> the current maximum priority, 1 for an empty ring). */
int ipp_ring_push(ipp_ring *r, int32_t n, const float *obs, const float *policy, const uint8_t *valid_mask, const float *values,
                  const float *rewards, float priority, int32_t is_device);

/* priorities <- 1 / size for every held sample (PrioritizedExperienceReplayBuffer.__init__, replay_buffers.py:115). */
int ipp_ring_reset_priorities(ipp_ring *r);

/* Draw n sample indices.  alpha < 0: uniform over the held samples (ExperienceReplayBuffer.sample, :90); else
 * probabilities = priorities^alpha / sum (:121-122), indices = searchsorted(cumsum(probabilities), u, 'right')
 * (np.random.choice), weights[k] = (probabilities[idx] * size)^(-beta) / max_k(...) (:131-132; uniform: 1).
 * uniforms[n] in [0,1) fp64 (host, or device with is_device): the caller's np.random.random_sample stream for
 * parity; NULL -> Philox4x32-10 keyed by (seed, draw counter).  indices / weights (may be NULL) are also kept on
 * the device for ipp_ring_gather(indices = NULL). */
int ipp_ring_sample(ipp_ring *r, int32_t n, double alpha, double beta, const double *uniforms, uint64_t seed, int64_t *indices,
                    float *weights, int32_t is_device);

/* Gather samples into a contiguous batch: indices[n] (NULL -> the last ipp_ring_sample draw).  shifts[n][2]
 * (int8 {dy, dx}, each in [-pad, pad]; NULL -> none): out[c][y][x] = obs[c][clamp(y + dy)][clamp(x + dx)], i.e.
 * ReplicationPad2d(pad) followed by a crop at offset (pad + dy, pad + dx) (replay_buffers.py:71-73).  Any output
 * may be NULL. */
int ipp_ring_gather(ipp_ring *r, int32_t n, const int64_t *indices, const int8_t *shifts, float *obs, float *policy, uint8_t *valid_mask,
                    float *values, float *rewards, int32_t is_device);

/* priorities[indices[k]] = priorities[k] (PrioritizedExperienceReplayBuffer.update, :140-141). */
int ipp_ring_update_priorities(ipp_ring *r, int32_t n, const int64_t *indices, const float *priorities, int32_t is_device);

/* priorities of slots [0, size) (host), for inspection / checkpointing. */
int ipp_ring_get_priorities(ipp_ring *r, float *priorities);

/* Device views (valid until ipp_ring_destroy). */
#define IPP_RING_PTR_OBS 0        /* float[capacity][C][Y][X] */
#define IPP_RING_PTR_POLICY 1     /* float[capacity][P] */
#define IPP_RING_PTR_MASK 2       /* uint8[capacity][P] */
#define IPP_RING_PTR_VALUE 3      /* float[capacity] */
#define IPP_RING_PTR_REWARD 4     /* float[capacity] */
#define IPP_RING_PTR_PRIORITY 5   /* float[capacity] */
#define IPP_RING_PTR_LAST_INDICES 6 /* int64[last n] */
#define IPP_RING_PTR_LAST_WEIGHTS 7 /* float[last n] */
#define IPP_RING_PTR_STREAM 8
void *ipp_ring_device_ptr(ipp_ring *r, int32_t which);

#ifdef __cplusplus
}
#endif
#endif /* IPP_EXPERIENCE_H */
