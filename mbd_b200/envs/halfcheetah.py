"""HalfCheetah — Brax's stock `halfcheetah` as the reference instantiates it
(`brax_envs.get_environment("halfcheetah", backend="positional")`, /root/reference/mbd/envs/__init__.py:30-31).
Neither the env nor its MJCF is in the reference tree: the env below restates brax/envs/half_cheetah.py **[brax-recalled]**.

Restated behaviour: positional backend => `opt.timestep = 0.003125`, `n_frames = 16` (env dt 0.05), actuator gears
[120, 90, 60, 120, 100, 100]; reset: `q = init_q + U(-0.1, 0.1)`, `qd = 0.1 * N(0, 1)`; reward = forward velocity of the
torso `(x' - x) / dt` - 0.1 * |action|^2."""
from __future__ import annotations

import numpy as np

from ..model import blob as blob_mod
from ..model import mjcf
from .ant import Ant
from .base import PipelineEnv, brax_asset


class HalfCheetah(Ant):
    reward_kind = blob_mod.REWARD_ANT   # same formula: forward velocity + RW1 - RW2 |a|^2 with RW1 = 0, RW2 = 0.1

    def __init__(self, xml_path: str | None = None, ctrl_cost_weight: float = 0.1, reset_noise_scale: float = 0.1):
        sys = mjcf.load(xml_path or brax_asset("half_cheetah.xml"))
        sys.dt = 0.003125
        sys.act_gear = np.array([120.0, 90.0, 60.0, 120.0, 100.0, 100.0])[: sys.act_size()]
        self._reset_noise_scale = float(reset_noise_scale)
        self._reward_params = (np.float32(0.003125 * 16), 0.0, float(ctrl_cost_weight), 0.0)
        PipelineEnv.__init__(self, sys=sys, backend="positional", n_frames=16)

    def _get_obs(self, pipeline_state) -> np.ndarray:
        return np.concatenate([pipeline_state.q[1:], pipeline_state.qd], axis=-1)   # exclude the root x position
