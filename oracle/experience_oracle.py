"""CPU restatement of the reference's experience path (TEST INFRASTRUCTURE — only tests/, smoke() and the CPU-baseline
leg of bench.py may import this; the product never does).

Pinned against the real reference by tests/golden/make_golden_experience.py -> golden_experience.npz:
  value_targets            planning/mcts_zero/episode_generators.py:158-164 (+ planning/common/rewards.py:34-35)
  prioritized_sample       planning/mcts_zero/replay_buffers.py:120-134 (np.random.choice = inverse CDF, side='right')
  uniform_sample           the ring's own definition, floor(u * N)  (reference: np.random.choice(N) = randint; PARITY
                           UNPINNED for the uniform index stream — a uniform draw has no arithmetic to reproduce)
  shift_with_replication   replay_buffers.py:69-74 (nn.ReplicationPad2d(4) + torchvision RandomCrop at a given offset)
"""
import numpy as np


def scale_value_target(value):
    """planning/common/rewards.py:34-35."""
    return np.sqrt(value + 1) - 1


def value_targets(rewards, gamma, horizon):
    """One episode.  episode_generators.py:158-164: total = sum_j gamma^j r_j; value_i = scale(sum_{j=i}^{min(i+H,L)-1}
    gamma^j r_j) — the exponent is the absolute step j, as written in the reference."""
    rewards = [float(r) for r in rewards]
    total = sum([gamma ** j * rewards[j] for j in range(len(rewards))])
    values = []
    for i in range(len(rewards)):
        hi = min(i + horizon, len(rewards))
        values.append(scale_value_target(sum([gamma ** j * rewards[j] for j in range(i, hi)])))
    return np.array(values, np.float64), total


def prioritized_sample(priorities, alpha, beta, uniforms):
    """replay_buffers.py:120-134 with np.random.choice(N, size, p) spelled out (numpy/random/mtrand.pyx `choice`:
    cdf = p.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(uniform_samples, side='right'))."""
    pr = np.asarray(priorities, np.float64)
    prob = pr ** alpha
    prob /= prob.sum()
    cdf = prob.cumsum()
    cdf /= cdf[-1]
    idx = cdf.searchsorted(np.asarray(uniforms, np.float64), side="right")
    w = (prob[idx] * len(pr)) ** (-beta)
    return idx.astype(np.int64), np.array(w / w.max(), dtype=np.float32)


def uniform_sample(size, uniforms):
    return np.minimum((np.asarray(uniforms, np.float64) * size).astype(np.int64), size - 1)


def shift_with_replication(states, dy, dx, pad=4):
    """ReplicationPad2d(pad) followed by a (Y, X) crop whose top-left corner is (pad + dy, pad + dx)."""
    s = np.asarray(states)
    Y, X = s.shape[-2:]
    padded = np.pad(s, [(0, 0)] * (s.ndim - 2) + [(pad, pad), (pad, pad)], mode="edge")
    return padded[..., pad + dy:pad + dy + Y, pad + dx:pad + dx + X]
