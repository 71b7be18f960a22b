"""HumanoidTrack — mirrors /root/reference/mbd/envs/humanoidtrack.py (positional, n_frames=5).

The 5 `*_ref` bodies (links 11-15) are cosmetic: world-parented slide bodies without
contacts, never read by the reward or eval_xref_logpd, and the env overwrites their x.pos with
the demo trajectory every step (humanoidtrack.py:66-74).  They are dynamically decoupled from
the humanoid, so the kernel simulates links 0-10 only; their x.pos on this surface is filled
from xref exactly as the reference does.
"""
from __future__ import annotations

import os

import numpy as np

from ..model import blob as blob_mod
from .base import ASSET_DIR, PipelineEnv, PipelineState, State, load_system


class HumanoidTrack(PipelineEnv):
    reward_kind = blob_mod.REWARD_HUMANOIDTRACK

    def __init__(self, mode="jog"):
        sys = load_system("humanoidtrack")
        self.H = 50  # traj time 1.5s (humanoidtrack.py:17)
        body_names = ["torso", "left_thigh", "right_thigh", "left_shin", "right_shin"]
        self.track_body_names = body_names
        self.track_body_idx = np.array([sys.link_names.index(n) for n in body_names])
        self.ref_body_names = [n + "_ref" for n in body_names]
        self.ref_body_idx = np.array([sys.link_names.index(n) for n in self.ref_body_names])
        demos = np.load(os.path.join(ASSET_DIR, "demos.npz"))
        self.xref = demos["jog_xref"].astype(np.float32)  # (5, 50, 3), built as humanoidtrack.py:33-44
        self.rew_xref = 1.0
        self.sim_links = [l for l in range(sys.num_links()) if l not in set(self.ref_body_idx.tolist())]
        self.track_links = tuple(int(i) for i in self.track_body_idx)
        super().__init__(sys=sys, backend="positional", n_frames=5)  # humanoidtrack.py:46

    def reset(self, rng) -> State:
        """humanoidtrack.py:48-61 — deterministic (rng unused)."""
        qpos = self.sys.init_q.astype(np.float32)
        qvel = np.zeros(self.sys.qd_size(), np.float32)
        pipeline_state = self.pipeline_init(qpos, qvel)
        obs = self._get_obs(pipeline_state)
        zero = np.float32(0.0)
        return State(pipeline_state, obs, zero, zero, {"reward_linup": zero, "reward_quadctrl": zero})

    def step(self, state: State, action) -> State:
        """humanoidtrack.py:63-82: reward from the PRE-step state, done counts time."""
        raw, reward = self._gpu_step(state.pipeline_state.raw, action)  # kernel reward = _get_reward(pre-step state)
        pipeline_state = self._make_pipeline_state(raw)
        t = min(int(np.int32(state.done)), self.xref.shape[1] - 1)  # clamped gather like XLA
        pos = pipeline_state.x.pos.copy()
        for i, idx in enumerate(self.ref_body_idx):
            pos[idx] = self.xref[i, t]
        pipeline_state = pipeline_state.replace(x=type(pipeline_state.x)(pos, pipeline_state.x.rot))
        obs = self._get_obs(pipeline_state)
        reward = np.float32(reward)
        return state.replace(pipeline_state=pipeline_state, obs=obs, reward=reward, done=state.done + 1)

    def _get_obs(self, pipeline_state: PipelineState) -> np.ndarray:
        return np.concatenate([pipeline_state.q, pipeline_state.qd], axis=-1)

    def _get_reward(self, state) -> np.float32:
        ps = state.pipeline_state
        return np.float32(1.0 + (-np.abs(ps.xd.vel[0, 0] - 1.6) - np.abs(ps.x.pos[0, 2] - 1.3) - np.abs(ps.x.pos[0, 1]) * 0.1))

    def eval_xref_logpd(self, xs_pos) -> np.float32:
        """humanoidtrack.py:98-106 on a stacked x.pos trajectory [H, L, 3] (host view; the planner
        uses the fused in-kernel reduction)."""
        xs = np.asarray(xs_pos)[:, self.track_body_idx].transpose(1, 0, 2)
        err = xs - self.xref[:, : xs.shape[1]]
        return np.float32(0.0 - ((np.clip(np.linalg.norm(err, axis=-1), 0.0, 0.5) / 0.5) ** 2).mean())
