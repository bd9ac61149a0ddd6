from .mcts import BatchedMCTS, LeafBatch  # noqa: F401
