"""GPU tests of the device-resident experience store (include/ipp_experience.h) against the golden
vectors produced by the real reference's replay-buffer / value-target code and against the NumPy oracle.

Bars: sampled indices, gathered rows and shifted (augmented) planes bit-exact; value targets and importance weights are
fp64 results rounded to float32 — 1e-6 relative."""
import numpy as np
import pytest

from oracle import experience_oracle as xo
from tests._util import golden

pytestmark = pytest.mark.gpu


class _ScriptedRng:
    """np.random stand-in that replays stored draws (the reference consumed exactly these under its seed)."""

    def __init__(self, uniforms=(), ints=()):
        self.uniforms, self.ints = list(uniforms), list(ints)

    def random_sample(self, n):
        u = self.uniforms.pop(0)
        assert len(u) == n
        return u

    def randint(self, lo, hi, size):
        v = np.asarray(self.ints.pop(0))
        assert v.shape == tuple(size) and v.min() >= lo and v.max() < hi
        return v


def _ring_from_golden(g, capacity=None):
    from ipp_rl_b200.planning.mcts_zero.replay_buffers import ExperienceRing

    st = g["per_states"]
    ring = ExperienceRing(capacity or st.shape[0], st.shape[1:], g["per_policies"].shape[1])
    ring.push(st, g["per_values"], g["per_rewards"], g["per_policies"], g["per_masks"])
    return ring


def test_value_targets_match_reference():
    from ipp_rl_b200.planning.mcts_zero.replay_buffers import ExperienceRing

    g = golden("golden_experience.npz")
    with ExperienceRing(4, (1, 4, 4), 2) as ring:
        for k in g["vt_cases"]:
            gamma, H = g[f"vt{k}_params"]
            rw = g[f"vt{k}_rewards"].astype(np.float32)
            v, tot = ring.value_targets(rw, gamma=float(gamma), horizon=int(H))
            ref_v, ref_t = xo.value_targets(rw, float(gamma), int(H))  # same float32-rounded rewards
            assert np.allclose(v[0], ref_v, rtol=1e-6, atol=1e-7)
            assert np.allclose(v[0], g[f"vt{k}_values"], rtol=1e-5, atol=1e-6)  # reference ran on fp64 rewards
            assert abs(tot[0] - ref_t) <= 1e-6 * max(1.0, abs(ref_t))
        # ragged batch: many episodes in one launch, zeros past each episode's end
        rng = np.random.RandomState(3)
        n, T = 300, 24
        rw = rng.uniform(0, 2, (n, T)).astype(np.float32)
        ln = rng.randint(0, T + 1, n).astype(np.int32)
        v, tot = ring.value_targets(rw, ln, gamma=0.97, horizon=5)
        for e in range(n):
            ref_v, ref_t = xo.value_targets(rw[e, :ln[e]], 0.97, 5)
            assert np.allclose(v[e, :ln[e]], ref_v, rtol=1e-6, atol=1e-7)
            assert np.all(v[e, ln[e]:] == 0)
            assert abs(tot[e] - ref_t) <= 1e-6 * max(1.0, abs(ref_t))


def test_prioritized_replay_reproduces_the_reference_buffer():
    from ipp_rl_b200.planning.mcts_zero.replay_buffers import PrioritizedExperienceReplayBuffer

    g = golden("golden_experience.npz")
    rounds = int(g["per_rounds"])
    with _ring_from_golden(g) as ring:
        rng = _ScriptedRng(uniforms=[g[f"per{t}_uniforms"] for t in range(rounds)])
        buf = PrioritizedExperienceReplayBuffer(ring, batch_size=32, alpha=float(g["per_alpha"]), beta0=0.5, num_epochs=3, rng=rng)
        assert len(buf) == 257 and buf.sample_size == 32
        for t in range(rounds):
            assert np.allclose(buf.priorities, g[f"per{t}_priorities"], rtol=1e-6)
            assert abs(float(buf.beta) - float(g[f"per{t}_beta"])) < 1e-12
            states, policies, values, rewards, msk, idx, w = buf.sample()
            assert np.array_equal(idx, g[f"per{t}_indices"])
            assert np.allclose(w, g[f"per{t}_weights"], rtol=2e-6)
            assert np.array_equal(states, g[f"per{t}_states"])
            assert np.array_equal(values, g[f"per{t}_values"])
            assert np.array_equal(policies, g["per_policies"][idx]) and np.array_equal(msk, g["per_masks"][idx])
            assert np.array_equal(rewards, g["per_rewards"][idx])
            buf.update(idx, g[f"per{t}_new_priorities"])
            buf.step()
        assert np.allclose(buf.priorities, g["per_final_priorities"], rtol=1e-6)
        assert abs(float(buf.beta) - float(g["per_final_beta"])) < 1e-12


def test_shift_augmentation_matches_reference_random_crop():
    from ipp_rl_b200.planning.mcts_zero.replay_buffers import ExperienceReplayBuffer

    g = golden("golden_experience.npz")
    sel = g["aug_sel"]
    with _ring_from_golden(g) as ring:
        # offsets the reference drew, as the randint stream of the buffer (values in [0, 2 * pad])
        rng = _ScriptedRng(ints=[g["aug_offsets"] + 4])
        buf = ExperienceReplayBuffer(ring, batch_size=16, num_augmented_samples=3, rng=rng)
        assert buf.sample_size == 4
        states, policies, msk, values, rewards = buf._gather_augmented(sel.astype(np.int64))
        assert np.array_equal(states, g["aug_states"])
        assert np.array_equal(values, g["aug_values"]) and np.array_equal(policies, g["aug_policies"])
        # the public call: uniform draw (floor(u N)), originals first then the shifted blocks
        u = np.array([0.0, 0.5, 0.99999, 0.25])
        buf.rng = _ScriptedRng(uniforms=[u], ints=[g["aug_offsets"] + 4])
        out = buf.sample()
        idx = xo.uniform_sample(len(ring), u)
        assert np.array_equal(out[5], idx) and out[0].shape[0] == 16 and np.all(out[6] == 1)
        base = g["per_states"][idx]
        expect = np.vstack([base] + [xo.shift_with_replication(base, int(dy), int(dx)) for dy, dx in g["aug_offsets"]])
        assert np.array_equal(out[0], expect)


@pytest.mark.parametrize("shape", [(6, 200, 200), (5, 33, 21)])
def test_ring_wraparound_gather_and_device_path(shape):
    """Ring semantics at network-input size: wrap-around pushes overwrite the oldest rows, gathers with per-sample
    shifts equal NumPy's edge-padded crops, the device-pointer path fills torch tensors, priorities of new rows."""
    import torch

    from ipp_rl_b200 import _capi as capi
    from ipp_rl_b200.planning.mcts_zero.replay_buffers import ExperienceRing

    C, Y, X = shape
    cap, P = 96, 50
    rng = np.random.RandomState(1)
    mirror = np.zeros((cap,) + shape, np.float32)
    mval = np.zeros(cap, np.float32)
    mrew = np.zeros(cap, np.float32)
    with ExperienceRing(cap, shape, P) as ring:
        head = 0
        for n in (40, 40, 40, 7):  # 127 rows through a 96-slot ring
            obs = rng.uniform(-1, 1, (n,) + shape).astype(np.float32)
            val = rng.uniform(0, 1, n).astype(np.float32)
            ring.push(obs, val, val * 2)
            for k in range(n):
                mirror[(head + k) % cap] = obs[k]
                mval[(head + k) % cap] = val[k]
                mrew[(head + k) % cap] = 2 * val[k]
            head = (head + n) % cap
        assert len(ring) == cap and ring.head == head
        assert np.all(ring.priorities() == 1.0)  # first push: 1, later pushes: the running maximum
        ring.update_priorities([5, 9], [3.5, 0.25])
        ring.push(mirror[:2], mval[:2], mval[:2])  # new rows take the maximum priority
        mirror[head:head + 2] = mirror[:2].copy()
        mval[head:head + 2] = mval[:2]
        mrew[head:head + 2] = mval[:2]
        pr = ring.priorities()
        assert pr[5] == 3.5 and pr[9] == 0.25 and np.all(pr[head:head + 2] == 3.5)

        n = 48
        idx = rng.randint(0, cap, n).astype(np.int64)
        sh = rng.randint(-4, 5, (n, 2))
        sh[:6] = [[0, 0], [4, 4], [-4, -4], [0, 3], [-2, 0], [4, -4]]
        obs, pol, msk, val, rew = ring.gather(idx, shifts=sh)
        for k in range(n):
            assert np.array_equal(obs[k], xo.shift_with_replication(mirror[idx[k]], int(sh[k, 0]), int(sh[k, 1]))), k
        assert np.array_equal(val, mval[idx]) and np.array_equal(rew, mrew[idx])
        assert np.all(msk == 1) and np.all(pol == 0)  # defaults of push(policies=None, valid_actions_msk=None)

        # device path: draw on the device (Philox), gather straight into torch tensors on the ring's stream
        ring.sample_indices(n, alpha=0.6, beta=0.4, seed=123)
        t_obs = torch.empty((n,) + shape, dtype=torch.float32, device="cuda")
        t_val = torch.empty(n, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        ring.gather_device(n, obs_ptr=t_obs.data_ptr(), values_ptr=t_val.data_ptr())
        stream = torch.cuda.ExternalStream(ring.device_ptr(capi.RING_PTR_STREAM))
        stream.synchronize()
        obs2, _, _, val2, _ = ring.gather(n=n, with_policy=False)  # indices == None -> the same last draw
        assert np.array_equal(t_obs.cpu().numpy(), obs2) and np.array_equal(t_val.cpu().numpy(), val2)


def test_device_draws_follow_the_priorities():
    """Philox-driven prioritised draws: empirical frequencies match p^alpha / sum within 5 sigma; same seed + counter
    gives the same draw only for the same call index (the draw counter advances)."""
    from ipp_rl_b200.planning.mcts_zero.replay_buffers import ExperienceRing

    N, n, alpha = 64, 4096, 0.75
    rng = np.random.RandomState(0)
    pri = rng.uniform(0.05, 2.0, N).astype(np.float32)
    with ExperienceRing(N, (1, 4, 4), 3) as ring:
        ring.push(np.zeros((N, 1, 4, 4), np.float32), np.zeros(N), np.zeros(N))
        ring.update_priorities(np.arange(N), pri)
        counts = np.zeros(N)
        draws = []
        for _ in range(8):
            idx, w = ring.sample_indices(n, alpha=alpha, beta=0.5, seed=42)
            draws.append(idx)
            counts += np.bincount(idx, minlength=N)
            prob = pri.astype(np.float64) ** alpha
            prob /= prob.sum()
            wref = (prob[idx] * N) ** -0.5
            assert np.allclose(w, wref / wref.max(), rtol=2e-6)
        assert not np.array_equal(draws[0], draws[1])
        total = 8 * n
        sigma = np.sqrt(total * prob * (1 - prob))
        assert np.all(np.abs(counts - total * prob) <= 5 * sigma + 1)
        iu, _ = ring.sample_indices(n, alpha=-1.0, seed=7)
        assert iu.min() >= 0 and iu.max() < N and len(np.unique(iu)) == N


def test_ring_error_paths():
    from ipp_rl_b200 import IppError
    from ipp_rl_b200.planning.mcts_zero.replay_buffers import ExperienceRing

    with pytest.raises(IppError):
        ExperienceRing(0, (1, 4, 4), 2)
    with ExperienceRing(8, (1, 4, 4), 2) as ring:
        with pytest.raises(IppError):
            ring.sample_indices(4)  # empty
        ring.push(np.zeros((3, 1, 4, 4), np.float32), np.zeros(3), np.zeros(3))
        with pytest.raises(IppError):
            ring.gather(np.array([0, 3]))  # row 3 not held yet
        with pytest.raises(ValueError):
            ring.push(np.zeros((1, 2, 4, 4), np.float32), 0.0, 0.0)
        with pytest.raises(IppError):
            ring.push(np.zeros((9, 1, 4, 4), np.float32), np.zeros(9), np.zeros(9))  # more than the capacity at once
