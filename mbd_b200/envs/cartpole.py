"""Cartpole — mirrors /root/reference/mbd/envs/cartpole.py (positional: dt 0.005, n_frames 4; cart on a slide joint,
pole on a hinge, one motor on the slider)."""
from __future__ import annotations

import numpy as np

from .. import prng
from ..model import blob as blob_mod
from .base import PipelineEnv, PipelineState, State, load_system


class Cartpole(PipelineEnv):
    reward_kind = blob_mod.REWARD_CARTPOLE

    def __init__(self, backend="positional", **kwargs):
        sys = load_system("cartpole")
        n_frames = 2
        if backend in ["spring", "positional"]:   # cartpole.py:17-19
            sys.dt = 0.005
            n_frames = 4
        kwargs["n_frames"] = kwargs.get("n_frames", n_frames)
        super().__init__(sys=sys, backend=backend, **kwargs)

    def reset(self, rng) -> State:
        """cartpole.py:25-38: the pole starts hanging down (q[1] = pi)"""
        rng, rng1, rng2 = prng.split(np.asarray(rng, dtype=np.uint32), 3)
        q = (self.sys.init_q.astype(np.float32) + prng.uniform(rng1, (self.sys.q_size(),), minval=-0.01, maxval=0.01)
             + np.array([0.0, np.pi], dtype=np.float32))
        qd = prng.uniform(rng2, (self.sys.qd_size(),), minval=-0.01, maxval=0.01)
        pipeline_state = self.pipeline_init(q, qd)
        return State(pipeline_state, self._get_obs(pipeline_state), np.float32(0.0), np.float32(0.0), {})

    def step(self, state: State, action) -> State:
        """cartpole.py:40-49; reward cos(q[1]) - |qd[0]| comes from the kernel (same expression as the planner path)"""
        raw, reward = self._gpu_step(state.pipeline_state.raw, action)
        pipeline_state = self._make_pipeline_state(raw)
        return state.replace(pipeline_state=pipeline_state, obs=self._get_obs(pipeline_state), reward=np.float32(reward),
                             done=np.float32(0.0))

    @property
    def action_size(self):
        return 1

    def _get_obs(self, pipeline_state: PipelineState) -> np.ndarray:
        return np.concatenate([pipeline_state.q, pipeline_state.qd])
