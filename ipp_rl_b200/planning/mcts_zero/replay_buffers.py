"""Device-resident experience store — host side of ``include/ipp_experience.h`` (``csrc/experience.cu``).

Reference: ``planning/mcts_zero/replay_buffers.py`` (``ReplayBuffer`` :15-80, ``ExperienceReplayBuffer`` :83-101,
``PrioritizedExperienceReplayBuffer`` :104-141) reads one bz2 pickle per sample from disk
(``EpisodeGenerator.save_sample_to_disk``, ``episode_generators.py:186-192``) and the value targets are computed by
a Python loop per episode (``episode_generators.py:158-164``).  Here the samples of whole env batches are rows of a
ring in HBM (:class:`ExperienceRing`); sampling, importance weights, the random-shift augmentation and the batch
gather are CUDA kernels.  The two buffer classes keep the reference's names, constructor hyper-parameters and
``sample() / step() / update() / __len__`` contract, with the ring in place of the file list:

    states, policies, values, rewards, valid_actions_msk, sample_indices, weights = buffer.sample()

No CPU fallback: everything below goes through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from ... import _capi as capi


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ExperienceRing:
    """Ring of ``capacity`` samples {obs (C, Y, X) f32, policy (P,) f32, valid mask (P,) u8, value, reward, priority}."""

    def __init__(self, capacity: int, obs_shape: Tuple[int, int, int], policy_slots: int, device: int = 0, stream: Optional[int] = None):
        self._lib = capi.load_library()
        self._h = C.c_void_p()
        cfg = capi.ipp_ring_config()
        cfg.struct_bytes = C.sizeof(capi.ipp_ring_config)
        cfg.device = device
        cfg.capacity = capacity
        cfg.channels, cfg.y_dim, cfg.x_dim = (int(v) for v in obs_shape)
        cfg.policy_slots = int(policy_slots)
        cfg.stream = C.c_void_p(stream) if stream else None
        rc = self._lib.ipp_ring_create(C.byref(cfg), C.byref(self._h))
        if rc != capi.IPP_OK:
            msg = self._lib.ipp_ring_last_error(None)
            self._h = C.c_void_p()
            raise capi.IppError(rc, msg.decode() if msg else "ipp_ring_create failed")
        self.capacity, self.obs_shape, self.policy_slots = int(capacity), tuple(int(v) for v in obs_shape), int(policy_slots)

    # -- plumbing -----------------------------------------------------------------------------------
    def _ck(self, rc: int) -> None:
        if rc != capi.IPP_OK:
            msg = self._lib.ipp_ring_last_error(self._h)
            raise capi.IppError(rc, msg.decode() if msg else "")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ipp_ring_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _info(self) -> capi.ipp_ring_info:
        i = capi.ipp_ring_info()
        self._ck(self._lib.ipp_ring_get_info(self._h, C.byref(i)))
        return i

    def __len__(self) -> int:
        return int(self._info().size)

    @property
    def head(self) -> int:
        return int(self._info().head)

    @property
    def launches(self) -> int:
        return int(self._info().launches)

    @property
    def device_bytes(self) -> int:
        return int(self._info().device_bytes)

    def device_ptr(self, which: int) -> int:
        p = self._lib.ipp_ring_device_ptr(self._h, which)
        return int(p) if p else 0

    # -- value targets (episode_generators.py:158-164) ------------------------------------------------
    def value_targets(self, rewards, lengths=None, gamma: float = 1.0, horizon: int = 1) -> Tuple[np.ndarray, np.ndarray]:
        """``rewards`` (n_episodes, max_steps) -> (scaled n-step value targets (n_episodes, max_steps), total episode
        values (n_episodes,)); steps past ``lengths`` give 0."""
        rw = np.ascontiguousarray(rewards, dtype=np.float32)
        if rw.ndim == 1:
            rw = rw[None]
        n, T = rw.shape
        ln = np.full(n, T, np.int32) if lengths is None else np.ascontiguousarray(lengths, dtype=np.int32).reshape(n)
        values = np.empty((n, T), np.float32)
        totals = np.empty(n, np.float32)
        self._ck(self._lib.ipp_ring_value_targets(self._h, _ptr(rw), _ptr(ln), n, T, float(gamma), int(horizon), _ptr(values), _ptr(totals), 0))
        return values, totals

    # -- storage ------------------------------------------------------------------------------------
    def push(self, obs, values, rewards, policies=None, valid_actions_msk=None, priority: float = 0.0) -> None:
        obs = np.ascontiguousarray(obs, dtype=np.float32)
        if obs.ndim == 3:
            obs = obs[None]
        n = obs.shape[0]
        if obs.shape[1:] != self.obs_shape:
            raise ValueError(f"obs must be (n, {self.obs_shape}), got {obs.shape}")
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(values, np.float32), (n,)))
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(rewards, np.float32), (n,)))
        pol = None if policies is None else np.ascontiguousarray(policies, dtype=np.float32).reshape(n, self.policy_slots)
        msk = None if valid_actions_msk is None else np.ascontiguousarray(valid_actions_msk, dtype=np.uint8).reshape(n, self.policy_slots)
        self._ck(self._lib.ipp_ring_push(self._h, n, _ptr(obs), _ptr(pol), _ptr(msk), _ptr(v), _ptr(r), float(priority), 0))

    def push_device(self, n: int, obs_ptr: int, values_ptr: int, rewards_ptr: int, policies_ptr: int = 0, mask_ptr: int = 0,
                    priority: float = 0.0) -> None:
        """Append rows that already live in HBM (e.g. ``ipp_observe_device`` output) — device-to-device stream copies."""
        self._ck(self._lib.ipp_ring_push(self._h, n, obs_ptr or None, policies_ptr or None, mask_ptr or None, values_ptr or None,
                                         rewards_ptr or None, float(priority), 1))

    def reset_priorities(self) -> None:
        self._ck(self._lib.ipp_ring_reset_priorities(self._h))

    def priorities(self) -> np.ndarray:
        p = np.empty(len(self), np.float32)
        self._ck(self._lib.ipp_ring_get_priorities(self._h, _ptr(p)))
        return p

    def update_priorities(self, indices, priorities) -> None:
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        pr = np.ascontiguousarray(priorities, dtype=np.float32).reshape(idx.shape)
        self._ck(self._lib.ipp_ring_update_priorities(self._h, idx.size, _ptr(idx), _ptr(pr), 0))

    # -- sampling -----------------------------------------------------------------------------------
    def sample_indices(self, n: int, alpha: float = -1.0, beta: float = 0.0, uniforms=None, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """``alpha < 0``: uniform; else prioritised (replay_buffers.py:121-132).  ``uniforms`` (n,) fp64 in [0, 1) = the
        caller's ``np.random.random_sample`` stream (parity); None -> device Philox."""
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64).reshape(n)
        idx = np.empty(n, np.int64)
        w = np.empty(n, np.float32)
        self._ck(self._lib.ipp_ring_sample(self._h, n, float(alpha), float(beta), _ptr(u), int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(idx), _ptr(w), 0))
        return idx, w

    def gather(self, indices=None, n: Optional[int] = None, shifts=None, with_policy: bool = True):
        """Rows ``indices`` (None: the last draw of ``n`` samples) as one batch; ``shifts`` (n, 2) int {dy, dx} = the crop
        offsets of the reference's ReplicationPad2d + RandomCrop augmentation relative to the centre."""
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int64)
        n = idx.size if idx is not None else int(n)
        sh = None if shifts is None else np.ascontiguousarray(shifts, dtype=np.int8).reshape(n, 2)
        obs = np.empty((n,) + self.obs_shape, np.float32)
        pol = np.empty((n, self.policy_slots), np.float32) if with_policy else None
        msk = np.empty((n, self.policy_slots), np.uint8) if with_policy else None
        val = np.empty(n, np.float32)
        rew = np.empty(n, np.float32)
        self._ck(self._lib.ipp_ring_gather(self._h, n, _ptr(idx), _ptr(sh), _ptr(obs), _ptr(pol), _ptr(msk), _ptr(val), _ptr(rew), 0))
        return obs, pol, msk, val, rew

    def gather_device(self, n: int, obs_ptr: int = 0, policy_ptr: int = 0, mask_ptr: int = 0, values_ptr: int = 0, rewards_ptr: int = 0,
                      indices_ptr: int = 0, shifts_ptr: int = 0) -> None:
        """Asynchronous gather into device buffers (raw addresses), e.g. torch tensors of the training step."""
        self._ck(self._lib.ipp_ring_gather(self._h, n, indices_ptr or None, shifts_ptr or None, obs_ptr or None, policy_ptr or None,
                                           mask_ptr or None, values_ptr or None, rewards_ptr or None, 1))


class ReplayBuffer:
    """``ReplayBuffer`` of the reference (replay_buffers.py:15-80) over an :class:`ExperienceRing` instead of a file list.
    ``window_size`` of the reference = the ring's capacity; ``num_augmented_samples`` extra randomly shifted copies per
    drawn sample (``augment_random_crop``: pad 4, replicate)."""

    PAD = 4  # nn.ReplicationPad2d(4), replay_buffers.py:71

    def __init__(self, ring: ExperienceRing, batch_size: int = 32, num_augmented_samples: int = 0, rng: Optional[np.random.RandomState] = None):
        self.ring = ring
        self.batch_size = batch_size
        self.num_augmented_samples = num_augmented_samples
        self.rng = rng if rng is not None else np.random

    @property
    def sample_size(self) -> int:
        return max(1, int(self.batch_size / (self.num_augmented_samples + 1)))

    def step(self):
        pass

    def update(self, indices: np.ndarray, priorities: np.ndarray):
        pass

    def __len__(self):
        return len(self.ring)

    def _gather_augmented(self, idx: np.ndarray):
        """Originals first, then ``num_augmented_samples`` blocks of shifted copies — the np.vstack / np.tile order of
        ``augment_random_crop`` (replay_buffers.py:58-77).  One crop offset per block, as torchvision's RandomCrop draws
        one offset for the whole batched tensor."""
        k = self.num_augmented_samples
        if k <= 0:
            return self.ring.gather(idx)
        all_idx = np.tile(idx, k + 1)
        shifts = np.zeros((k + 1, 2), np.int64)
        shifts[1:] = self.rng.randint(0, 2 * self.PAD + 1, size=(k, 2)) - self.PAD
        return self.ring.gather(all_idx, shifts=np.repeat(shifts, idx.size, axis=0))

    def sample(self):
        raise NotImplementedError("Replay buffer does not implement 'sample()' method!")


class ExperienceReplayBuffer(ReplayBuffer):
    """Uniform sampling (replay_buffers.py:83-101)."""

    def sample(self):
        u = self.rng.random_sample(self.sample_size)
        idx, _ = self.ring.sample_indices(self.sample_size, alpha=-1.0, uniforms=u)
        states, policies, msk, values, rewards = self._gather_augmented(idx)
        return states, policies, values, rewards, msk.astype(bool), idx, np.ones(len(states))


class PrioritizedExperienceReplayBuffer(ReplayBuffer):
    """Proportional prioritised replay (replay_buffers.py:104-141): P(i) = p_i^alpha / sum, importance weights
    (P(i) N)^-beta / max, beta annealed to 1 over ``total_steps`` calls of ``step()``."""

    def __init__(self, ring: ExperienceRing, batch_size: int = 32, alpha: float = 0.75, beta0: float = 0.5, num_epochs: int = 3,
                 rng: Optional[np.random.RandomState] = None):
        super().__init__(ring, batch_size, 0, rng)
        self.alpha = alpha
        self.beta0 = beta0
        self.beta = beta0
        self.ring.reset_priorities()  # np.ones(N) / N, :115
        self.total_steps = max(1, (len(self.ring) // self.sample_size) * num_epochs)

    @property
    def priorities(self) -> np.ndarray:
        return self.ring.priorities()

    def step(self):
        self.beta = np.minimum(self.beta + (1 - self.beta0) / self.total_steps, 1)

    def sample(self):
        u = self.rng.random_sample(self.sample_size)
        idx, weights = self.ring.sample_indices(self.sample_size, alpha=self.alpha, beta=float(self.beta), uniforms=u)
        states, policies, msk, values, rewards = self.ring.gather(idx)
        return states, policies, values, rewards, msk.astype(bool), idx, weights

    def update(self, indices: np.ndarray, priorities: np.ndarray):
        self.ring.update_priorities(indices, priorities)
