"""Car2d — mirrors /root/reference/mbd/envs/car2d.py (self-contained kinematic car, RK4,
11 disc obstacles, demo path).  Dynamics run in csrc/mbd_b200.cu::k_car2d."""
from __future__ import annotations

import dataclasses
import os

import numpy as np
import torch

from .. import ops
from .base import ASSET_DIR


@dataclasses.dataclass
class State:
    pipeline_state: np.ndarray
    obs: np.ndarray
    reward: float
    done: float

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)


class Car2d:
    kind = "car2d"

    def __init__(self):
        self.dt = 0.1
        self.H = 50
        r_obs = 0.3
        self.obs_center = np.array(  # car2d.py:48-62
            [[-r_obs * 3, r_obs * 2], [-r_obs * 2, r_obs * 2], [-r_obs * 1, r_obs * 2], [0.0, r_obs * 2], [0.0, r_obs * 1],
             [0.0, 0.0], [0.0, -r_obs * 1], [-r_obs * 3, -r_obs * 2], [-r_obs * 2, -r_obs * 2], [-r_obs * 1, -r_obs * 2],
             [0.0, -r_obs * 2]], dtype=np.float32)
        self.obs_radius = r_obs
        self.x0 = np.array([-0.5, 0.0, np.pi * 3 / 2], dtype=np.float32)
        self.xg = np.array([0.5, 0.0, 0.0], dtype=np.float32)
        self.xref = np.load(os.path.join(ASSET_DIR, "demos.npz"))["car2d_xref"].astype(np.float32)
        xref_diff = np.diff(self.xref, axis=0)
        theta = np.arctan2(xref_diff[:, 0], xref_diff[:, 1])
        self.thetaref = np.append(theta, theta[-1])
        # rew_xref = vmap(get_reward)(xref).mean()  (car2d.py:71)
        d = np.linalg.norm(self.xref[:, :2] - self.xg[:2], axis=-1).astype(np.float32)
        self.rew_xref = float(np.mean(np.float32(1.0) - (np.clip(d, 0.0, 0.2) / np.float32(0.2)) ** 2, dtype=np.float32))
        # kernel parameters: Python-double constants rounded once, as weak-typed JAX scalars are
        self.params = np.concatenate([self.obs_center.reshape(-1), np.float32([r_obs, self.dt, self.dt / 2, self.dt / 6])]).astype(np.float32)
        self._dev = {}

    def device_params(self):
        idx = torch.cuda.current_device()
        if idx not in self._dev:
            d = torch.device("cuda", idx)
            self._dev[idx] = (torch.as_tensor(self.params, device=d), torch.as_tensor(self.xref, device=d))
        return self._dev[idx]

    def reset(self, rng):
        return State(self.x0, self.x0, 0.0, 0.0)  # car2d.py:73-75

    def step(self, state: State, action) -> State:
        """car2d.py:77-86 through the CUDA kernel (n=1, H=1)."""
        params, _ = self.device_params()
        dev = params.device
        x0 = torch.as_tensor(np.asarray(state.pipeline_state, dtype=np.float32), device=dev)
        u = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(1, 1, 2), device=dev)
        out = ops.car2d_rollout(params, x0, u, want_traj=True)
        q = out["traj"][0, 0].cpu().numpy()
        return state.replace(pipeline_state=q, obs=q, reward=float(out["rews"][0].item()), done=0.0)

    def get_reward(self, q):
        d = np.float32(np.linalg.norm(np.asarray(q, dtype=np.float32)[:2] - self.xg[:2]))
        return np.float32(1.0) - (np.clip(d, 0.0, 0.2) / np.float32(0.2)) ** 2

    def eval_xref_logpd(self, xs):
        xs = np.asarray(xs, dtype=np.float32)
        err = xs[:, :2] - self.xref[: xs.shape[0], :2]
        return np.float32(0.0 - ((np.clip(np.linalg.norm(err, axis=-1), 0.0, 0.5) / 0.5) ** 2).mean())

    @property
    def action_size(self):
        return 2

    @property
    def observation_size(self):
        return 3

    def render(self, ax, xs):
        import matplotlib.pyplot as plt
        for i in range(self.obs_center.shape[0]):
            ax.add_artist(plt.Circle(self.obs_center[i, :], self.obs_radius, color="k", fill=True, alpha=0.5))
        ax.scatter(xs[:, 0], xs[:, 1], c=range(xs.shape[0]), cmap="Reds")
        ax.plot(xs[:, 0], xs[:, 1], "r-", label="Car path")
        ax.set_xlabel("x"); ax.set_ylabel("y"); ax.set_xlim(-2, 2); ax.set_ylim(-2, 2)
        ax.set_aspect("equal"); ax.grid(True)
