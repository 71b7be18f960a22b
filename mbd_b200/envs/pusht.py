"""PushT — mirrors /root/reference/mbd/envs/pushT.py (class surface, reset chain, reward, obs), the reference's one env on
Brax's `generalized` backend.  The physics runs in csrc/pusht.cuh (one sample per thread): the planar reduced-coordinate pipeline
restated in include/mbd_pusht.h — Brax itself is an un-vendored dependency of the reference, the restatement is UNPINNED
(oracle/pusht_oracle.c lists what is recalled and what is an own choice).  Model data: assets/pusht.json, derived from the
reference's pushT.xml by scripts/make_pusht_asset.py."""
from __future__ import annotations

import dataclasses
import json
import os

import numpy as np
import torch

from .. import ops, prng
from .base import ASSET_DIR, Motion, State, Transform

# include/mbd_pusht.h
PT = dict(DT=0, NSUB=1, ITERS=2, GEAR0=3, GEAR1=4, MP=5, IMP=6, RP=7, DPX=8, DPY=9, MS=10, IMS=11, IS=12, IIS=13, CX=14, CY=15,
          DSX=16, DSY=17, DSTH=18, LIM0=19, BOX0=27, MU=35, DMIN=36, DMAX=37, WIDTH=38, MID=39, KB=40, KK=41, TOL=42, NPARAM=43)


@dataclasses.dataclass
class PushTPipelineState:
    """the fields of brax.generalized.base.State the reference reads: q, qd (pushT.py:47-58) and x.pos / x.rot for rendering"""
    q: np.ndarray
    qd: np.ndarray
    x: Transform
    xd: Motion

    @property
    def raw(self) -> np.ndarray:
        return np.concatenate([self.q, self.qd]).astype(np.float32)

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)


@dataclasses.dataclass
class _Geom:
    body: int            # link index, -1 = world
    type: str
    pos: np.ndarray
    quat: np.ndarray
    size: np.ndarray
    rgba: np.ndarray


class PushTSystem:
    """what the planner / renderer read from `env.sys`: dt, sizes, init_q, link names and geoms"""

    def __init__(self, model: dict):
        self.model = model
        self.dt = float(model["timestep"])
        self.link_names = list(model["bodies"].keys())      # pusher, slider, goal
        self.init_q = np.zeros(8, dtype=np.float64)          # MuJoCo qpos0 of slide / hinge joints
        self.link_parents = [-1, -1, -1]
        t = model["table"]
        self.geoms = [_Geom(-1, "plane", np.float32(t["pos"]), np.float32([1, 0, 0, 0]), np.float32(t["size"]), np.float32(t["rgba"]))]
        for l, name in enumerate(self.link_names):
            for g in model["bodies"][name]["geoms"]:
                self.geoms.append(_Geom(l, g["type"], np.float32(g["pos"]), np.float32([1, 0, 0, 0]), np.float32(g["size"]), np.float32(g["rgba"])))

    def q_size(self) -> int:
        return 8

    def qd_size(self) -> int:
        return 8

    def act_size(self) -> int:
        return 2

    def num_links(self) -> int:
        return 3

    def replace(self, **kw):
        out = PushTSystem(self.model)
        for k, v in kw.items():
            setattr(out, k, v)
        return out


def pack_params(model: dict, n_frames: int) -> np.ndarray:
    """assets/pusht.json -> the MBD_PT_* table (float32).  Python-double constants are rounded once."""
    b = model["bodies"]
    pusher, slider = b["pusher"], b["slider"]
    sphere = pusher["geoms"][0]
    assert sphere["type"] == "sphere" and [j["type"] for j in pusher["joints"]] == ["slide", "slide"]
    assert [j["type"] for j in slider["joints"]] == ["slide", "slide", "hinge"] and len(slider["geoms"]) == 2
    assert not any(g["collides"] for g in b["goal"]["geoms"]), "the goal is a ghost (contype = conaffinity = 0)"
    assert model["solimp"][4] == 2.0, "the impedance curve is implemented for power = 2 (MuJoCo default)"
    P = np.zeros(PT["NPARAM"], dtype=np.float64)
    P[PT["DT"]], P[PT["NSUB"]], P[PT["ITERS"]] = model["timestep"], n_frames, model["iterations"]
    gear = {a["joint"]: a["gear"] for a in model["actuators"]}
    P[PT["GEAR0"]], P[PT["GEAR1"]] = gear[pusher["joints"][0]["name"]], gear[pusher["joints"][1]["name"]]
    P[PT["MP"]], P[PT["IMP"]], P[PT["RP"]] = pusher["mass"], 1.0 / pusher["mass"], sphere["size"][0]
    P[PT["DPX"]], P[PT["DPY"]] = pusher["joints"][0]["damping"], pusher["joints"][1]["damping"]
    P[PT["MS"]], P[PT["IMS"]], P[PT["IS"]], P[PT["IIS"]] = slider["mass"], 1.0 / slider["mass"], slider["izz"], 1.0 / slider["izz"]
    P[PT["CX"]], P[PT["CY"]] = slider["com"][0], slider["com"][1]
    P[PT["DSX"]], P[PT["DSY"]], P[PT["DSTH"]] = (j["damping"] for j in slider["joints"])
    lim = [j for j in pusher["joints"] + slider["joints"][:2]]
    for k, j in enumerate(lim):
        P[PT["LIM0"] + 2 * k: PT["LIM0"] + 2 * k + 2] = j["range"] if j["limited"] else (-np.inf, np.inf)
    for k, g in enumerate(slider["geoms"]):
        assert g["type"] == "box" and g["pos"][2] == 0.0
        P[PT["BOX0"] + 4 * k: PT["BOX0"] + 4 * k + 4] = [g["pos"][0], g["pos"][1], g["size"][0], g["size"][1]]
    P[PT["MU"]] = max(sphere["friction"], max(g["friction"] for g in slider["geoms"]))   # MuJoCo / Brax: max of the pair
    tc, dr = model["solref"]
    dmin, dmax, width, mid, _ = model["solimp"]
    P[PT["DMIN"]], P[PT["DMAX"]], P[PT["WIDTH"]], P[PT["MID"]] = dmin, dmax, width, mid
    P[PT["KB"]] = 2.0 / (dmax * tc)
    P[PT["KK"]] = 1.0 / (dmax * dmax * tc * tc * dr * dr)
    P[PT["TOL"]] = 1e-6     # solver tolerance (own choice, like the solver): well below the softness of the constraints themselves
    return P.astype(np.float32)


class PushT:
    kind = "pusht"

    def __init__(self, backend: str = "generalized"):
        if backend != "generalized":
            raise NotImplementedError("pushT runs on the generalized backend (pushT.py:16)")
        with open(os.path.join(ASSET_DIR, "pusht.json")) as f:
            self.model = json.load(f)
        self.sys = PushTSystem(self.model)
        self.backend = backend
        self._n_frames = 5                                    # pushT.py:20
        self.params = pack_params(self.model, self._n_frames)
        self._dev = {}

    # ---- brax PipelineEnv surface --------------------------------------------------------------
    @property
    def dt(self) -> float:
        return self.sys.dt * self._n_frames

    @property
    def action_size(self) -> int:
        return 2                                              # pushT.py:69-70

    @property
    def observation_size(self) -> int:
        return 16                                             # pushT.py:73-74

    def device_params(self):
        idx = torch.cuda.current_device()
        if idx not in self._dev:
            self._dev[idx] = torch.as_tensor(self.params, device=torch.device("cuda", idx))
        return self._dev[idx]

    def pipeline_init(self, q, qd) -> PushTPipelineState:
        q = np.asarray(q, dtype=np.float32).copy()
        qd = np.asarray(qd, dtype=np.float32).copy()
        pos = np.zeros((3, 3), np.float32)
        rot = np.zeros((3, 4), np.float32)
        pos[0, :2] = q[0:2]; rot[0] = [1, 0, 0, 0]
        for l, o in ((1, 2), (2, 5)):
            pos[l, :2] = q[o:o + 2]
            rot[l] = [np.cos(q[o + 2] / 2), 0.0, 0.0, np.sin(q[o + 2] / 2)]
        ang = np.zeros((3, 3), np.float32); vel = np.zeros((3, 3), np.float32)
        vel[0, :2] = qd[0:2]
        for l, o in ((1, 2), (2, 5)):
            vel[l, :2] = qd[o:o + 2]; ang[l, 2] = qd[o + 2]
        return PushTPipelineState(q=q, qd=qd, x=Transform(pos, rot), xd=Motion(ang, vel))

    def reset(self, rng) -> State:
        """pushT.py:22-38"""
        rng, rng_goal_xy = prng.split(np.asarray(rng, dtype=np.uint32))
        q = self.sys.init_q.astype(np.float32)
        q[:2] = np.float32([0.1, -0.15])
        q[5:] = (prng.uniform(rng_goal_xy, (3,), minval=-1.0, maxval=1.0) * np.float32([0.2, 0.2, np.pi / 4])
                 + np.float32([-0.4, 0.4, np.pi])).astype(np.float32)
        ps = self.pipeline_init(q, np.zeros(8, np.float32))
        return State(ps, self._get_obs(ps), self._get_reward(ps), self._get_done(ps), {})

    def step(self, state: State, action) -> State:
        """pushT.py:40-47: n_frames physics steps with one action (the CUDA kernel with n = 1, H = 1)"""
        P = self.device_params()
        x0 = torch.as_tensor(state.pipeline_state.raw, device=P.device)
        u = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(1, 1, 2), device=P.device)
        out = ops.pusht_rollout(P, x0, u, want_final=True)
        fin = out["final"][0].cpu().numpy()
        ps = self.pipeline_init(fin[:8], fin[8:])
        return state.replace(pipeline_state=ps, obs=self._get_obs(ps), reward=np.float32(out["rews"][0].item()), done=self._get_done(ps))

    def _get_obs(self, pipeline_state) -> np.ndarray:
        return np.concatenate([pipeline_state.q, pipeline_state.qd], axis=-1)      # pushT.py:49-50

    def _get_reward(self, pipeline_state) -> np.float32:
        """pushT.py:52-62 (host copy for reset / tests; rollouts evaluate it in the kernel)"""
        q = np.asarray(pipeline_state.q, dtype=np.float32)
        d_pusher2slider = np.maximum(np.float32(np.linalg.norm(q[0:2] - q[2:4])) - np.float32(0.2), np.float32(0.0))
        return np.float32(1.0) - (np.float32(np.linalg.norm(q[5:7] - q[2:4])) + np.abs(q[7] - q[4]) / np.float32(np.pi) + d_pusher2slider)

    def _get_done(self, pipeline_state) -> np.float32:
        return np.float32(self._get_reward(pipeline_state) > 0.95)                 # pushT.py:64-66


def main():
    """pushT.py:77-98: 50 uniformly random actions from the seed-1 reset, written as a Brax-visualizer page"""
    import mbd_b200
    from ..io import brax_json
    env = PushT()
    rng = prng.PRNGKey(1)
    state = env.reset(rng)
    rollout = [state.pipeline_state]
    for _ in range(50):
        rng, rng_act = prng.split(rng)
        act = prng.uniform(rng_act, (env.action_size,), minval=-1.0, maxval=1.0)
        state = env.step(state, act)
        rollout.append(state.pipeline_state)
    path = f"{mbd_b200.__path__[0]}/../results/pushT"
    os.makedirs(path, exist_ok=True)
    with open(f"{path}/vis.html", "w") as f:
        f.write(brax_json.render(env.sys, rollout, env.dt))


if __name__ == "__main__":
    main()
