"""Rollout helpers — mirror /root/reference/mbd/utils.py:6-20 (`eval_us`, `rollout_us`).

The reference scans `env.step` over the horizon under jit; here the whole horizon is ONE call
of the CUDA rollout kernel (n=1) when `step_env` is the bound `step` of one of this package's
envs; any other callable is stepped in a plain Python loop.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _env_of(step_env):
    return getattr(step_env, "__self__", None)


def rollout_us(step_env, state, us):
    """-> (rews [H], pipeline_states: list of H pipeline states)."""
    env = _env_of(step_env)
    us = np.asarray(us.detach().cpu().numpy() if isinstance(us, torch.Tensor) else us, dtype=np.float32)
    if env is not None and getattr(env, "kind", None) == "car2d":
        params, _ = env.device_params()
        dev = params.device
        out = ops.car2d_rollout(params, torch.as_tensor(np.asarray(state.pipeline_state, np.float32), device=dev),
                                torch.as_tensor(us[None], device=dev), want_rewss=True, want_traj=True)
        return out["rewss"][0].cpu().numpy(), list(out["traj"][0].cpu().numpy())
    if env is not None and getattr(env, "kind", None) in ("xpbd", "pusht"):
        rews, states = [], []
        st = state
        # per-step states are needed by callers (rendering); one launch per step keeps them exact
        for t in range(us.shape[0]):
            st = env.step(st, us[t])
            rews.append(st.reward)
            states.append(st.pipeline_state)
        return np.asarray(rews, dtype=np.float32), states
    rews, states = [], []
    for t in range(us.shape[0]):
        state = step_env(state, us[t])
        rews.append(state.reward)
        states.append(state.pipeline_state)
    return np.asarray(rews), states


def eval_us(step_env, state, us):
    """-> rews [H]; a single fused launch for this package's envs."""
    env = _env_of(step_env)
    us_np = np.asarray(us.detach().cpu().numpy() if isinstance(us, torch.Tensor) else us, dtype=np.float32)
    if env is not None and getattr(env, "kind", None) == "xpbd":
        m = env.device_model()
        out = ops.rollout(m, torch.as_tensor(state.pipeline_state.raw, device=m.device), torch.as_tensor(us_np[None], device=m.device),
                          want_rewss=True)
        return out["rewss"][0].cpu().numpy()
    return rollout_us(step_env, state, us)[0]


def trajectory_arrays(env, pipeline_states):
    """The per-step world poses the reference hands to `brax.io.html.render` (utils.py:23-34 `render_us`,
    scripts/vis_diffusion.py:115-143): x.pos [T,L,3], x.rot [T,L,4] (w,x,y,z), plus q [T,nq], qd [T,nv]."""
    pos = np.stack([np.asarray(ps.x.pos, dtype=np.float32) for ps in pipeline_states])
    rot = np.stack([np.asarray(ps.x.rot, dtype=np.float32) for ps in pipeline_states])
    q = np.stack([np.asarray(ps.q, dtype=np.float32) for ps in pipeline_states])
    qd = np.stack([np.asarray(ps.qd, dtype=np.float32) for ps in pipeline_states])
    return dict(pos=pos, rot=rot, q=q, qd=qd, dt=np.float32(env.dt), link_names=np.asarray(list(env.sys.link_names)))


def rollout_states(step_env, state, us):
    """the `rollout` list of mbd.utils.render_us (utils.py:23-34) / vis_diffusion.py:115-121: the pipeline state BEFORE each of
    the H steps (initial state first)"""
    us = np.asarray(us.detach().cpu().numpy() if isinstance(us, torch.Tensor) else us, dtype=np.float32)
    rollout = []
    for i in range(us.shape[0]):
        rollout.append(state.pipeline_state)
        state = step_env(state, us[i])
    return rollout


def render_us(step_env, sys, state, us, dt=None):
    """Mirror of mbd.utils.render_us: steps `us` from `state` and returns the Brax-visualizer HTML page of the rollout
    (`brax.io.html.render(sys, rollout)` in the reference; produced here by mbd_b200.io.brax_json without Brax).  For an env
    without world poses (car2d) the list of states is returned."""
    env = _env_of(step_env)
    rollout = rollout_states(step_env, state, us)
    if env is not None and getattr(env, "kind", None) in ("xpbd", "pusht"):
        from .io import brax_json
        return brax_json.render(sys, rollout, env.dt if dt is None else dt)
    return rollout
