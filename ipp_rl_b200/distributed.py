"""Env-batch sharding across GPUs (one process per GPU, ``torch.distributed``).

Envs are independent (the reference deep-copies one Mapping per repetition / episode,
experiments/experiments.py:181-185, planning/mcts_zero/episode_generators.py:53), so the hot path
shards with NO collective: rank r owns the contiguous env slice ``shard_bounds(total, world, r)``
with its own maps in its own HBM.  An env's RNG stream is keyed by its GLOBAL id
(``ipp_config.env_id_offset``), so results do not depend on the split.  The only exchange is the
one a trainer's experience buffer needs: an all-gather of the per-env rewards (4 B per env per step).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from .engine import BatchedEngine, EngineConfig


def shard_bounds(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(first_env, count) of rank's contiguous slice; the remainder goes to the lowest ranks."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(int(total), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def all_shard_counts(total: int, world_size: int) -> List[int]:
    return [shard_bounds(total, world_size, r)[1] for r in range(world_size)]


def sharded_config(cfg: EngineConfig, total_envs: int, world_size: int, rank: int, device: Optional[int] = None) -> EngineConfig:
    """Config of this rank's engine: its slice size, its global env-id offset, its device."""
    first, count = shard_bounds(total_envs, world_size, rank)
    if count < 1:
        raise ValueError(f"rank {rank} of {world_size} gets no env out of {total_envs}")
    kw = dict(cfg.__dict__)
    kw.update(batch=count, env_id_offset=cfg.env_id_offset + first)
    if device is not None:
        kw["device"] = device
    return EngineConfig(**kw)


_scratch = {}


def _scratch_tensor(key, shape, dtype, device):
    """Reused staging tensors of the uneven-split path (a trainer calls the gathers every step)."""
    import torch

    t = _scratch.get(key)
    if t is None or t.shape != shape or t.dtype != dtype or t.device != device:
        t = torch.zeros(shape, dtype=dtype, device=device)
        _scratch[key] = t
    return t


def gather_rows(local, total_envs: int, group=None, out=None):
    """All-gather per-env rows of every rank into one (total_envs, ...) tensor ordered by global env id.  ``local`` is
    this rank's torch tensor whose first dimension is its env slice (CUDA with the nccl backend, CPU with gloo); ``out``
    (optional) receives the result — with it and an even split the call allocates nothing."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    counts = all_shard_counts(total_envs, world)
    mine = counts[dist.get_rank(group)]
    if local.shape[0] != mine:
        raise ValueError(f"local rows: {local.shape[0]}, this rank owns {mine} envs")
    tail = tuple(local.shape[1:])
    if out is None:
        out = torch.empty((total_envs,) + tail, dtype=local.dtype, device=local.device)
    elif tuple(out.shape) != (total_envs,) + tail or out.dtype != local.dtype:
        raise ValueError("out must be (total_envs, ...) with the dtype of the local rows")
    if len(set(counts)) == 1:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # uneven split: pad every shard to the largest one (reused staging), gather, copy the slices into place
    width = max(counts)
    padded = _scratch_tensor(("pad", tail), (width,) + tail, local.dtype, local.device)
    padded[:mine] = local
    wide = _scratch_tensor(("wide", tail, world), (world * width,) + tail, local.dtype, local.device)
    dist.all_gather_into_tensor(wide, padded, group=group)
    first = 0
    for r, c in enumerate(counts):
        out[first : first + c] = wide[r * width : r * width + c]
        first += c
    return out


def gather_rewards(local, total_envs: int, group=None):
    """The one exchange of the step path: per-env rewards (4 B per env per step) of all ranks, ordered by global env id."""
    if local.dim() != 1:
        raise ValueError("rewards must be 1-D (one per env of this rank)")
    return gather_rows(local, total_envs, group)


def gather_experience(local: dict, total_envs: int, group=None) -> dict:
    """Experience rows of all ranks for the learner's ring (the reference sums what its worker processes
    pickled to disk, mcts_zero_mission.py:504-521): every tensor of ``local`` ({"obs": (n, C, Y, X), "values": (n,), ...},
    first dimension = this rank's env slice) is all-gathered in global env order."""
    return {k: gather_rows(v, total_envs, group) for k, v in local.items()}


def gather_rows_to(local, total_envs: int, dst: int = 0, group=None):
    """Gather per-env rows to ONE rank (the learner): what the trainer's experience buffer needs without
    making every rank receive every other rank's rows.  Returns the (total_envs, ...) tensor on ``dst``, None elsewhere."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = all_shard_counts(total_envs, world)
    if local.shape[0] != counts[rank]:
        raise ValueError(f"local rows: {local.shape[0]}, this rank owns {counts[rank]} envs")
    tail = tuple(local.shape[1:])
    width = max(counts)
    src = local.contiguous()
    if counts[rank] != width:
        src = torch.zeros((width,) + tail, dtype=local.dtype, device=local.device)
        src[: counts[rank]] = local
    if rank == dst:
        parts = [torch.empty((width,) + tail, dtype=local.dtype, device=local.device) for _ in range(world)]
        dist.gather(src, parts, dst=dst, group=group)
        return torch.cat([p[:c] for p, c in zip(parts, counts)])
    dist.gather(src, None, dst=dst, group=group)
    return None


def all_reduce_policy(policy: np.ndarray, group=None) -> np.ndarray:
    """Root-parallel MCTS: sum of the visit-count policies of every rank's trees (the reference sums the policies of its
    worker processes, planning/mcts_zero/mcts_zero_mission.py:516-521).  Identity when torch.distributed is not
    initialised; NCCL (through a device tensor) or gloo otherwise."""
    try:
        import torch
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return policy
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return policy
    t = torch.from_numpy(np.ascontiguousarray(policy, dtype=np.float64))
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


class ShardedEngine:
    """This rank's slice of a ``total_envs`` batch.  ``step`` takes / returns LOCAL arrays; ``global_slice``
    tells which rows of a global action array belong here."""

    def __init__(self, cfg: EngineConfig, total_envs: int, world_size: int, rank: int, device: Optional[int] = None):
        self.total_envs, self.world_size, self.rank = total_envs, world_size, rank
        self.first, self.count = shard_bounds(total_envs, world_size, rank)
        self.engine = BatchedEngine(sharded_config(cfg, total_envs, world_size, rank, device))

    @property
    def global_slice(self) -> slice:
        return slice(self.first, self.first + self.count)

    def step(self, global_actions: np.ndarray, **kw) -> np.ndarray:
        return self.engine.step(np.asarray(global_actions)[self.global_slice], **kw)

    def close(self):
        self.engine.close()
