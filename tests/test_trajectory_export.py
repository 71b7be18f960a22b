"""render_us / trajectory_arrays (the rollout export the CLI writes instead of the reference's Brax HTML page,
mbd/utils.py:23-34, mbd_planner.py:168-178).  The env's GPU stepper is replaced by the CPU oracle here (test
infrastructure) so that the host-side logic runs without a GPU."""
import numpy as np

import mbd_b200
from mbd_b200 import prng
from mbd_b200.utils import render_us
from oracle import oracle as orc


def test_render_us_returns_world_trajectory(monkeypatch):
    env = mbd_b200.envs.get_env("humanoidrun")

    def cpu_step(raw, action):
        out = orc.xpbd_rollout(env.blob, raw, np.asarray(action, np.float32)[None, None], want_final=True)
        return out["final"][0], out["rews"][0]

    monkeypatch.setattr(env, "_gpu_step", cpu_step)
    state = env.reset(prng.split(prng.PRNGKey(0))[1])
    us = np.clip(np.random.default_rng(0).normal(size=(6, env.action_size)) * 0.5, -1, 1).astype(np.float32)
    tr = render_us(env.step, env.sys, state, us)
    L = env.sys.num_links()
    assert tr["pos"].shape == (6, L, 3) and tr["rot"].shape == (6, L, 4)
    assert tr["q"].shape == (6, env.sys.q_size()) and tr["qd"].shape == (6, env.sys.qd_size())
    assert list(tr["link_names"]) == list(env.sys.link_names) and np.isclose(tr["dt"], 0.042)
    # first entry is the initial state (the reference appends before stepping); quaternions stay unit; the torso moves
    assert np.array_equal(tr["pos"][0], np.asarray(state.pipeline_state.x.pos, np.float32))
    assert np.allclose(np.linalg.norm(tr["rot"], axis=-1), 1.0, atol=1e-5)
    assert not np.allclose(tr["pos"][5, 0], tr["pos"][0, 0])
    # joint coordinates are consistent with the world poses: at reset forward kinematics of (q, qd) reproduces them; later
    # the soft XPBD position constraints (joint_scale_pos = 0.5) leave a joint separation that accumulates down the chain
    from mbd_b200.model import kinematics
    for t, tol in ((0, 1e-5), (5, 0.2)):
        pos = kinematics.forward(env.sys, tr["q"][t].astype(np.float64), tr["qd"][t].astype(np.float64))[0]
        assert np.allclose(pos, tr["pos"][t], atol=tol)
