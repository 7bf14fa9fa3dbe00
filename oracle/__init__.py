"""CPU oracle for the Touch-GS rasterizer hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``touch-gs_b200/``,
``bench.py`` outside its ``cpu_baseline`` / ``--impl reference`` legs) may import
this package.  See ``oracle/gs_oracle.py`` for the status line: **parity
unpinned** (the reference tree vendors no rasterizer and no golden vectors).
"""
from .gs_oracle import *  # noqa: F401,F403
from . import train_oracle  # noqa: F401,E402  (train-step pieces: SURVEY §8f N1)
