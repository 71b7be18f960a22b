"""Ant (Brax's stock env, restated; BASELINE config 3).  Brax's ant.xml is not available here, so the tests run the env
class on the repo's synthetic quadruped fixture (same topology: free root, 4 x (hip + ankle), capsule feet)."""
import os

import numpy as np
import pytest

import mbd_b200
from mbd_b200 import prng
from mbd_b200.envs.ant import Ant
from oracle import oracle as orc

FIXTURE = os.path.join(os.path.dirname(__file__), "fixtures", "quadruped.xml")


def _state(env, seed=3):
    """reset without the GPU: the qd noise is drawn by the oracle's normal() instead of the sampling kernel"""
    rng, r1, r2 = prng.split(prng.PRNGKey(seed), 3)
    q = env.sys.init_q.astype(np.float32) + prng.uniform(r1, (env.sys.q_size(),), minval=-0.1, maxval=0.1)
    qd = np.float32(0.1) * orc.normal(r2, (env.sys.qd_size(),))
    return env.pipeline_init(q, qd).raw


def test_registry_resolves_ant_from_the_repo_asset(monkeypatch):
    """without a Brax install the env is built from mbd_b200/assets/ant.xml (this repo's restatement of the public model)"""
    monkeypatch.delenv("MBD_BRAX_ASSETS", raising=False)
    env = mbd_b200.envs.get_env("ant")
    assert env.sys.link_types == "f11111111" and env.action_size == 8


def test_positional_overrides_and_reward_parameters():
    env = Ant(xml_path=FIXTURE)
    assert env.sys.dt == 0.005 and env._n_frames == 10 and np.isclose(env.dt, 0.05)
    assert np.all(env.sys.act_gear == 200.0)
    f = env.blob.view(np.float32)
    from mbd_b200.model import blob as B
    assert env.blob.view(np.int32)[B.H_REWARD] == B.REWARD_ANT
    assert f[B.H_RW0] == np.float32(0.05) and f[B.H_RW0 + 1] == 1.0 and f[B.H_RW0 + 2] == 0.5


def test_oracle_ant_reward_formula():
    env = Ant(xml_path=FIXTURE)
    st = _state(env)
    H = 6
    us = np.clip(np.random.default_rng(0).normal(size=(1, H, env.action_size)) * 0.6, -1, 1).astype(np.float32)
    out = orc.xpbd_rollout(env.blob, st, us, want_rewss=True)
    # independent restatement: step by step, root origin from the final states (the fixture's torso COM is its origin)
    x_prev = np.float32(st[0, 0])
    raw = st
    for t in range(H):
        raw = orc.xpbd_rollout(env.blob, raw, us[:, t:t + 1], want_final=True)["final"][0]
        x = np.float32(raw[0, 0])
        ss = np.float32(0.0)
        for k in range(env.action_size):
            ss = np.float32(ss + np.float32(us[0, t, k] * us[0, t, k]))
        r = np.float32(np.float32(np.float32(np.float32(x - x_prev) / np.float32(0.05)) + np.float32(1.0)) - np.float32(np.float32(0.5) * ss))
        assert r == out["rewss"][0, t]
        x_prev = x


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [1, 2])
def test_ant_kernels_match_oracle_bit_exact(variant):
    import torch
    from mbd_b200 import ops
    env = Ant(xml_path=FIXTURE)
    st = _state(env)
    us = np.clip(np.random.default_rng(5).normal(size=(70, 20, env.action_size)) * 0.7, -1, 1).astype(np.float32)
    ref = orc.xpbd_rollout(env.blob, st, us, want_rewss=True, want_final=True)
    ops.set_kernel_variant(variant)
    try:
        m = env.device_model(torch.device("cuda:0"))
        out = ops.rollout(m, torch.as_tensor(st, device="cuda:0"), torch.as_tensor(us, device="cuda:0"), want_rewss=True, want_final=True)
    finally:
        ops.set_kernel_variant(0)
    assert np.array_equal(out["final"].cpu().numpy().view(np.uint32), ref["final"].view(np.uint32))
    assert np.array_equal(out["rewss"].cpu().numpy().view(np.uint32), ref["rewss"].view(np.uint32))
