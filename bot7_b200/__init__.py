"""bot7_b200 -- B200-native surrogate-fit-and-acquisition path of bot7 behind bot7's own API.

Sub-modules mirror the reference namespaces: bot7.grids, bot7.models, bot7.scores, bot7.samplers,
bot7.bots (reference init.lua:28-40).  Everything numerical runs in libbot7_b200.so (hand-written
sm_100a CUDA); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from . import grids, samplers, scores, models, bots, parallel  # noqa: F401

__all__ = ["grids", "samplers", "scores", "models", "bots", "parallel"]
