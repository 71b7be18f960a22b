"""Env surface shared by the Brax-backed envs: State / PipelineState containers and the
PipelineEnv-like base class (mirrors `brax.envs.base.PipelineEnv, State` as used at
/root/reference/mbd/envs/humanoidrun.py:3,12-41).  Single-state `reset/step` run the same CUDA
rollout kernel as the planner (n=1, H=1); arrays on this surface are host numpy arrays.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Any, Dict, Optional

import numpy as np
import torch

from .. import ops
from ..model import blob as blob_mod
from ..model import kinematics, mjcf, system_io

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets")


def load_system(name: str) -> mjcf.System:
    """Loads the compiled model; falls back to compiling the reference XML when
    MBD_REFERENCE_ASSETS points at the reference's mbd/assets directory."""
    path = os.path.join(ASSET_DIR, name + ".json")
    if os.path.exists(path):
        return system_io.load(path)
    ref = os.environ.get("MBD_REFERENCE_ASSETS", "/root/reference/mbd/assets")
    return mjcf.load(os.path.join(ref, name + ".xml"))


def brax_asset(name: str) -> str:
    """Path of an MJCF the reference takes from INSIDE the Brax wheel (`epath.resource_path("brax") /
    "envs/assets/<name>"`, /root/reference/mbd/envs/hopper.py:13, walker2d.py:14; ant / halfcheetah through
    `brax.envs`).  Lookup order: `$MBD_BRAX_ASSETS/<name>`, an installed Brax, then this repo's own restatement
    of the public model under mbd_b200/assets/ (provenance in each file's header; unpinned like all Brax behaviour)."""
    import importlib.util
    cands = []
    if os.environ.get("MBD_BRAX_ASSETS"):
        cands.append(os.path.join(os.environ["MBD_BRAX_ASSETS"], name))
    try:
        spec = importlib.util.find_spec("brax")
    except (ImportError, ValueError):
        spec = None
    if spec is not None and spec.submodule_search_locations:
        cands.append(os.path.join(list(spec.submodule_search_locations)[0], "envs", "assets", name))
    cands.append(os.path.join(ASSET_DIR, name))
    for c in cands:
        if os.path.exists(c):
            return c
    raise FileNotFoundError(name)


@dataclasses.dataclass
class Transform:
    pos: np.ndarray
    rot: np.ndarray

    def replace(self, **kw):     # brax.base.Transform.replace (scripts/vis_diffusion.py:92-96 shifts x.pos for pushT with it)
        return dataclasses.replace(self, **kw)


@dataclasses.dataclass
class Motion:
    ang: np.ndarray
    vel: np.ndarray

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)


@dataclasses.dataclass
class PipelineState:
    """brax.positional.base.State fields the reference reads (x.pos, xd.vel, q, qd) + the raw
    [L,13] COM-frame state the kernels consume."""
    q: np.ndarray
    qd: np.ndarray
    x: Transform
    xd: Motion
    x_i: Transform
    xd_i: Motion
    raw: np.ndarray

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)


@dataclasses.dataclass
class State:
    pipeline_state: Any
    obs: np.ndarray
    reward: Any
    done: Any
    metrics: Dict[str, Any] = dataclasses.field(default_factory=dict)
    info: Dict[str, Any] = dataclasses.field(default_factory=dict)

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)


class PipelineEnv:
    """Positional-backend PipelineEnv: `sys`, `dt`, `action_size`, `observation_size`,
    `pipeline_init`, `pipeline_step` (brax/envs/base.py, restated)."""

    kind = "xpbd"
    reward_kind = blob_mod.REWARD_HUMANOIDRUN
    sim_links = None          # subset of links that are simulated (None = all)
    track_links = ()          # links whose positions feed eval_xref_logpd

    def __init__(self, sys: mjcf.System, backend: str = "positional", n_frames: int = 1):
        if backend != "positional":
            raise NotImplementedError("only the positional backend is implemented (SURVEY F4)")
        self.sys = sys
        self.backend = backend
        self._n_frames = n_frames
        self._links = list(range(sys.num_links())) if self.sim_links is None else list(self.sim_links)
        self.blob = blob_mod.pack(sys, n_frames, self.reward_kind, links=self._links, track_links=tuple(self.track_links),
                                  **self._pack_kwargs())
        self._models: Dict[int, ops.Model] = {}
        # world poses of links that are NOT simulated (cosmetic bodies) stay at their init_q pose
        self._static_x = kinematics.forward(sys, sys.init_q, np.zeros(sys.qd_size()))

    def _pack_kwargs(self):
        """extra keyword arguments for blob.pack (reward parameters)"""
        return {}

    # ---- brax PipelineEnv API ---------------------------------------------------------------
    @property
    def dt(self) -> float:
        return self.sys.dt * self._n_frames

    @property
    def action_size(self) -> int:
        return self.sys.act_size()

    @property
    def observation_size(self) -> int:
        return self.sys.q_size() + self.sys.qd_size()

    def device_model(self, device: Optional[torch.device] = None) -> ops.Model:
        idx = torch.cuda.current_device() if device is None else torch.device(device).index
        if idx not in self._models:
            self._models[idx] = ops.Model(self.blob, torch.device("cuda", idx))
        return self._models[idx]

    def _make_pipeline_state(self, raw: np.ndarray) -> PipelineState:
        """raw [Lsim,13] -> PipelineState with world-frame x, xd and joint coordinates q, qd."""
        pos, rot, ang, vel = kinematics.to_world(self.sys, raw, self._links)
        L = self.sys.num_links()
        xpos, xrot = self._static_x[0].astype(np.float32), self._static_x[1].astype(np.float32)
        xang, xvel = np.zeros((L, 3), np.float32), np.zeros((L, 3), np.float32)
        xipos, xirot = xpos.copy(), xrot.copy()
        xiang, xivel = xang.copy(), xvel.copy()
        for i, l in enumerate(self._links):
            xpos[l], xrot[l], xang[l], xvel[l] = pos[i], rot[i], ang[i], vel[i]
            xipos[l], xirot[l], xiang[l], xivel[l] = raw[i, 0:3], raw[i, 3:7], raw[i, 7:10], raw[i, 10:13]
        q, qd = kinematics.inverse(self.sys, xpos, xrot, xang, xvel, self._links)
        return PipelineState(q=q, qd=qd, x=Transform(xpos, xrot), xd=Motion(xang, xvel), x_i=Transform(xipos, xirot),
                             xd_i=Motion(xiang, xivel), raw=np.asarray(raw, dtype=np.float32))

    def pipeline_init(self, q, qd) -> PipelineState:
        raw = kinematics.pipeline_init(self.sys, q, qd, links=self._links)
        return self._make_pipeline_state(raw)

    def pipeline_step(self, pipeline_state: PipelineState, action) -> PipelineState:
        """n_frames positional steps with one action: the CUDA rollout kernel with n=1, H=1."""
        raw, _ = self._gpu_step(pipeline_state.raw, action)
        return self._make_pipeline_state(raw)

    def _gpu_step(self, raw, action):
        m = self.device_model()
        dev = m.device
        st = torch.as_tensor(np.ascontiguousarray(raw, dtype=np.float32), device=dev)
        u = torch.as_tensor(np.ascontiguousarray(action, dtype=np.float32).reshape(1, 1, -1), device=dev)
        out = ops.rollout(m, st, u, want_final=True)
        return out["final"][0].cpu().numpy(), float(out["rews"][0].item())
