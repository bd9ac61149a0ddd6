"""ipp_rl_b200 — B200-native batched engine for the per-step hot path of dmar-bonn/ipp-rl.

``BatchedEngine`` (engine.py) is the batched twin; ``mapping`` / ``sensors`` / ``simulations`` /
``planning`` mirror the reference's factory / Mapping / GridMap surface for B=1 drop-in use.
All compute runs in ``csrc/libipp_b200.so`` (hand-written sm_100a CUDA behind a C ABI).
"""
from ._capi import (  # noqa: F401
    FLAG_ADAPTIVE,
    LAYOUT_MV,
    LAYOUT_PLANES,
    LAYOUT_SPLIT,
    LAYOUT_SUPER,
    LAYOUT_TILED,
    REWARD_GAUSS_ENTROPY,
    REWARD_TRACE,
    IppError,
    IppLibraryError,
)
from .engine import BatchedEngine, EngineConfig  # noqa: F401

__all__ = ["BatchedEngine", "EngineConfig", "IppError", "IppLibraryError"]
