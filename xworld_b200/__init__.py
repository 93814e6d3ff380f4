"""xworld_b200 -- B200-native batched XWorld2D simulator (hot path of PaddlePaddle/XWorld's
games/xworld behind the reference's SimulatorInterface / py_simulator API).

Layout: csrc/ (CUDA kernels + the C ABI of include/xworld_b200.h), simulator.py (the reference's
Python API names), catalog.py (icon set / atlas), sharding.py (multi-GPU env partition + counters).
"""
from .simulator import Simulator, decode_game_over_code  # noqa: F401
from .catalog import Catalog  # noqa: F401

__all__ = ["Simulator", "Catalog", "decode_game_over_code"]
