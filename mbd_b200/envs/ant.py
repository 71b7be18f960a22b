"""Ant — Brax's stock `ant` as the reference instantiates it (`brax_envs.get_environment("ant", backend="positional")`,
/root/reference/mbd/envs/__init__.py:30-31).  Neither the env nor its MJCF is part of the reference tree: the env below
restates brax/envs/ant.py **[brax-recalled]**; the MJCF comes from `brax_asset` (a Brax install, `MBD_BRAX_ASSETS`, or this
repo's restatement of the public Gym model, mbd_b200/assets/ant.xml).

Restated behaviour: positional backend => `opt.timestep = 0.005`, `n_frames = 10` (env dt 0.05), every actuator gear 200;
reset: `q = init_q + U(-0.1, 0.1)`, `qd = 0.1 * N(0, 1)`; reward = forward velocity of the torso
`(x' - x) / dt` + healthy reward 1.0 (paid unconditionally: `terminate_when_unhealthy=True`) - 0.5 * |action|^2, contact
cost off; termination does not stop a planner rollout (`rollout_us` never resets)."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops, prng
from ..model import blob as blob_mod
from ..model import mjcf
from .base import PipelineEnv, PipelineState, State, brax_asset


class Ant(PipelineEnv):
    reward_kind = blob_mod.REWARD_ANT

    def __init__(self, xml_path: str | None = None, ctrl_cost_weight: float = 0.5, healthy_reward: float = 1.0,
                 reset_noise_scale: float = 0.1):
        sys = mjcf.load(xml_path or brax_asset("ant.xml"))
        sys.dt = 0.005                                   # ant.py: positional => opt.timestep 0.005, n_frames 10
        sys.act_gear = np.full_like(sys.act_gear, 200.0)  # ant.py: positional => gear 200 on every actuator
        self._reset_noise_scale = float(reset_noise_scale)
        self._reward_params = (np.float32(0.005 * 10), float(healthy_reward), float(ctrl_cost_weight), 0.0)
        super().__init__(sys=sys, backend="positional", n_frames=10)

    def _pack_kwargs(self):
        return dict(reward_params=self._reward_params)

    def reset(self, rng) -> State:
        rng, rng1, rng2 = prng.split(np.asarray(rng, dtype=np.uint32), 3)
        s = self._reset_noise_scale
        q = self.sys.init_q.astype(np.float32) + prng.uniform(rng1, (self.sys.q_size(),), minval=-s, maxval=s)
        # qd = scale * jax.random.normal: drawn by the sampling kernel (same threefry + erfinv as the hot path)
        nv = self.sys.qd_size()
        qd = ops.sample(rng2, nv, 0, 1, nv, s, torch.zeros(nv, device="cuda"))[0].cpu().numpy()
        ps = self.pipeline_init(q, qd)
        return State(ps, self._get_obs(ps), np.float32(0.0), np.float32(0.0), {})

    def step(self, state: State, action) -> State:
        raw, reward = self._gpu_step(state.pipeline_state.raw, action)
        ps = self._make_pipeline_state(raw)
        return state.replace(pipeline_state=ps, obs=self._get_obs(ps), reward=np.float32(reward))

    def _get_obs(self, pipeline_state: PipelineState) -> np.ndarray:
        return np.concatenate([pipeline_state.q[2:], pipeline_state.qd], axis=-1)   # exclude_current_positions_from_observation
