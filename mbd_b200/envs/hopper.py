"""Hopper — mirrors /root/reference/mbd/envs/hopper.py (positional backend, n_frames=20, the XML's own timestep).
4 links, planar root (slide x, slide z, hinge y on the torso), 3 motors.  The MJCF comes from `brax_asset`."""
from __future__ import annotations

import numpy as np

from .. import prng
from ..model import blob as blob_mod
from ..model import mjcf
from .base import PipelineEnv, PipelineState, State, brax_asset


class Hopper(PipelineEnv):
    reward_kind = blob_mod.REWARD_HOPPER
    _asset, _z_target, _n_frames_ref = "hopper.xml", 1.0, 20   # hopper.py:13,18,64

    def __init__(self, xml_path: str | None = None):
        sys = mjcf.load(xml_path or brax_asset(self._asset))
        self._reset_noise_scale = 5e-3                                  # hopper.py:16
        super().__init__(sys=sys, backend="positional", n_frames=self._n_frames_ref)

    def _pack_kwargs(self):
        return dict(reward_params=(self._z_target, 0.0, 0.0, 0.0))

    def reset(self, rng) -> State:
        """hopper.py:20-34"""
        rng, rng1, rng2 = prng.split(np.asarray(rng, dtype=np.uint32), 3)
        low, hi = -self._reset_noise_scale, self._reset_noise_scale
        qpos = self.sys.init_q.astype(np.float32) + prng.uniform(rng1, (self.sys.q_size(),), minval=low, maxval=hi)
        qvel = prng.uniform(rng2, (self.sys.qd_size(),), minval=low, maxval=hi)
        pipeline_state = self.pipeline_init(qpos, qvel)
        obs = self._get_obs(pipeline_state)
        return State(pipeline_state, obs, np.float32(0.0), np.float32(0.0), {})

    def step(self, state: State, action) -> State:
        """hopper.py:36-47"""
        raw, reward = self._gpu_step(state.pipeline_state.raw, action)
        pipeline_state = self._make_pipeline_state(raw)
        obs = self._get_obs(pipeline_state)
        return state.replace(pipeline_state=pipeline_state, obs=obs, reward=np.float32(reward), done=np.float32(0.0))

    def _get_obs(self, pipeline_state: PipelineState) -> np.ndarray:
        """hopper.py:49-55"""
        position = pipeline_state.q.copy()
        position[1] = pipeline_state.x.pos[0, 2]
        velocity = np.clip(pipeline_state.qd, -10, 10)
        return np.concatenate((position, velocity))

    def _get_reward(self, pipeline_state: PipelineState):
        """hopper.py:57-65 (host view of what the kernel accumulates per env step)"""
        x = pipeline_state.x.pos
        return np.float32(x[0, 0] - np.clip(np.abs(x[0, 2] - np.float32(self._z_target)), -1.0, 1.0) * 0.5)
