"""GenericPositionalEnv — any MJCF in the supported subset (free root, stacked hinges, sphere / capsule geoms
against the z = 0 floor, motor actuators, Brax custom numerics with spring_inertia_scale = 1) run through the
same compiler, oracle-checked kernels and planner as the reference's humanoid envs.  This is how a user brings a
model that lives outside the reference tree (e.g. the MJCF files inside their Brax install)."""
from __future__ import annotations

import numpy as np

from .. import prng
from ..model import blob as blob_mod
from ..model import mjcf
from .base import PipelineEnv, PipelineState, State


class GenericPositionalEnv(PipelineEnv):
    def __init__(self, xml_path: str, n_frames: int = 5, reward_kind: int = blob_mod.REWARD_HUMANOIDRUN, reset_noise: float = 0.01):
        self.reward_kind = reward_kind
        self._reset_noise = float(reset_noise)
        super().__init__(sys=mjcf.load(xml_path), backend="positional", n_frames=n_frames)

    def reset(self, rng) -> State:
        rng, rng1, rng2 = prng.split(np.asarray(rng, dtype=np.uint32), 3)
        s = self._reset_noise
        qpos = self.sys.init_q.astype(np.float32) + prng.uniform(rng1, (self.sys.q_size(),), minval=-s, maxval=s)
        qvel = prng.uniform(rng2, (self.sys.qd_size(),), minval=-s, maxval=s)
        ps = self.pipeline_init(qpos, qvel)
        return State(ps, self._get_obs(ps), np.float32(0.0), np.float32(0.0), {})

    def step(self, state: State, action) -> State:
        raw, reward = self._gpu_step(state.pipeline_state.raw, action)
        ps = self._make_pipeline_state(raw)
        return state.replace(pipeline_state=ps, obs=self._get_obs(ps), reward=np.float32(reward))

    def _get_obs(self, pipeline_state: PipelineState) -> np.ndarray:
        return np.concatenate([pipeline_state.q, pipeline_state.qd], axis=-1)
