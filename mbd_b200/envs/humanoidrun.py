"""HumanoidRun — mirrors /root/reference/mbd/envs/humanoidrun.py (positional backend, n_frames=7)."""
from __future__ import annotations

import numpy as np

from .. import prng
from ..model import blob as blob_mod
from .base import PipelineEnv, PipelineState, State, load_system


class HumanoidRun(PipelineEnv):
    reward_kind = blob_mod.REWARD_HUMANOIDRUN

    def __init__(self):
        sys = load_system("humanoidrun")
        super().__init__(sys=sys, backend="positional", n_frames=7)  # humanoidrun.py:17

    def reset(self, rng) -> State:
        """humanoidrun.py:19-32: init_q + U(-0.01, 0.01) noise on qpos and qvel."""
        rng, rng1, rng2 = prng.split(np.asarray(rng, dtype=np.uint32), 3)
        low, hi = -0.01, 0.01
        qpos = self.sys.init_q.astype(np.float32) + prng.uniform(rng1, (self.sys.q_size(),), minval=-0.01, maxval=0.01)
        qvel = prng.uniform(rng2, (self.sys.qd_size(),), minval=low, maxval=hi)
        pipeline_state = self.pipeline_init(qpos, qvel)
        obs = self._get_obs(pipeline_state, np.zeros(self.sys.act_size(), np.float32))
        return State(pipeline_state, obs, np.float32(0.0), np.float32(0.0), {})

    def step(self, state: State, action) -> State:
        """humanoidrun.py:34-41."""
        raw, reward = self._gpu_step(state.pipeline_state.raw, action)  # reward: the kernel's (bit-exact with the planner)
        pipeline_state = self._make_pipeline_state(raw)
        obs = self._get_obs(pipeline_state, action)
        reward = np.float32(reward)
        return state.replace(pipeline_state=pipeline_state, obs=obs, reward=reward)

    def _get_obs(self, pipeline_state: PipelineState, action) -> np.ndarray:
        return np.concatenate([pipeline_state.q, pipeline_state.qd], axis=-1)

    def _get_reward(self, pipeline_state: PipelineState):
        """humanoidrun.py:46-51 (host view of what the kernel accumulates per env step)."""
        x = pipeline_state.x.pos
        return np.float32(x[0, 0] * 1.0 - np.clip(np.abs(x[0, 2] - 1.3), -1.0, 1.0) * 1.0 - np.abs(x[0, 1]) * 0.1)
