"""HumanoidStandup — mirrors /root/reference/mbd/envs/humanoidstandup.py (positional, n_frames=7).
Contact-heavy: 15 plane contacts (capsule end caps on torso / thighs / forearms, head and feet
spheres) against the single floor plane."""
from __future__ import annotations

import numpy as np

from .. import prng
from ..model import blob as blob_mod
from .base import PipelineEnv, PipelineState, State, load_system


class HumanoidStandup(PipelineEnv):
    reward_kind = blob_mod.REWARD_HUMANOIDSTANDUP

    def __init__(self):
        sys = load_system("humanoidstandup")
        super().__init__(sys=sys, backend="positional", n_frames=7)  # humanoidstandup.py:17

    def reset(self, rng) -> State:
        """humanoidstandup.py:19-38."""
        rng, rng1, rng2 = prng.split(np.asarray(rng, dtype=np.uint32), 3)
        qpos = self.sys.init_q.astype(np.float32) + prng.uniform(rng1, (self.sys.q_size(),), minval=-0.01, maxval=0.01)
        qvel = prng.uniform(rng2, (self.sys.qd_size(),), minval=-0.01, maxval=0.01)
        pipeline_state = self.pipeline_init(qpos, qvel)
        obs = self._get_obs(pipeline_state, np.zeros(self.sys.act_size(), np.float32))
        zero = np.float32(0.0)
        return State(pipeline_state, obs, zero, zero, {"reward_linup": zero, "reward_quadctrl": zero})

    def step(self, state: State, action) -> State:
        """humanoidstandup.py:40-48."""
        raw, reward = self._gpu_step(state.pipeline_state.raw, action)
        pipeline_state = self._make_pipeline_state(raw)
        obs = self._get_obs(pipeline_state, action)
        return state.replace(pipeline_state=pipeline_state, obs=obs, reward=np.float32(reward))

    def _get_obs(self, pipeline_state: PipelineState, action) -> np.ndarray:
        return np.concatenate([pipeline_state.q, pipeline_state.qd], axis=-1)

    def _get_reward(self, pipeline_state: PipelineState):
        x = pipeline_state.x.pos
        return np.float32(1.5 - np.clip(np.abs(x[0, 2] - 1.3), -2.0, 1.0) - np.abs(x[0, 0]) * 0.1 - np.abs(x[0, 1]) * 0.1)
