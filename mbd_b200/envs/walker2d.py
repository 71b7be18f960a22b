"""Walker2d — mirrors /root/reference/mbd/envs/walker2d.py (positional backend, n_frames=20): the hopper surface with
7 links, 6 motors and the height target 1.1 (walker2d.py:56-61)."""
from __future__ import annotations

from .hopper import Hopper


class Walker2d(Hopper):
    _asset, _z_target, _n_frames_ref = "walker2d.xml", 1.1, 20   # walker2d.py:14,19,60
