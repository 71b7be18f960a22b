"""Ant — Brax's stock `ant` as the reference instantiates it (`brax_envs.get_environment("ant", backend="positional")`,
/root/reference/mbd/envs/__init__.py:30-31).  Neither the env nor its MJCF is part of the reference tree: the env below
restates brax/envs/ant.py **[brax-recalled]** and the MJCF is looked up in the user's Brax install (or `MBD_BRAX_ASSETS`).

Restated behaviour: positional backend => `opt.timestep = 0.005`, `n_frames = 10` (env dt 0.05), every actuator gear 200;
reset: `q = init_q + U(-0.1, 0.1)`, `qd = 0.1 * N(0, 1)`; reward = forward velocity of the torso
`(x' - x) / dt` + healthy reward 1.0 (paid unconditionally: `terminate_when_unhealthy=True`) - 0.5 * |action|^2, contact
cost off; termination does not stop a planner rollout (`rollout_us` never resets)."""
from __future__ import annotations

import importlib.util
import os

import numpy as np
import torch

from .. import ops, prng
from ..model import blob as blob_mod
from ..model import mjcf
from .base import PipelineEnv, PipelineState, State


def find_brax_asset(name: str) -> str | None:
    """`<MBD_BRAX_ASSETS>/<name>` or `<site-packages>/brax/envs/assets/<name>`; None when neither exists."""
    cands = []
    if os.environ.get("MBD_BRAX_ASSETS"):
        cands.append(os.path.join(os.environ["MBD_BRAX_ASSETS"], name))
    try:
        spec = importlib.util.find_spec("brax")
    except (ImportError, ValueError):
        spec = None
    if spec is not None and spec.submodule_search_locations:
        cands.append(os.path.join(list(spec.submodule_search_locations)[0], "envs", "assets", name))
    for c in cands:
        if os.path.exists(c):
            return c
    return None


class Ant(PipelineEnv):
    reward_kind = blob_mod.REWARD_ANT

    def __init__(self, xml_path: str | None = None, ctrl_cost_weight: float = 0.5, healthy_reward: float = 1.0,
                 reset_noise_scale: float = 0.1):
        path = xml_path or find_brax_asset("ant.xml")
        if path is None:
            raise NotImplementedError("environment 'ant' needs Brax's ant.xml: it lives inside the Brax wheel "
                                      "(brax/envs/assets/ant.xml), not in the reference tree — install Brax, set "
                                      "MBD_BRAX_ASSETS to a directory that holds it, or pass xml_path")
        sys = mjcf.load(path)
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
