"""pushT (/root/reference/mbd/envs/pushT.py): parameter table, reset chain and reward of the env, physical
known-answer tests of the restated planar generalized pipeline (oracle/pusht_oracle.c — parity with Brax is UNPINNED, these pin
the mechanics), and — on the GPU — kernel == oracle bit for bit, the diffusion step, a short solve."""
import os
import re

import numpy as np
import pytest

import mbd_b200
from mbd_b200 import prng
from mbd_b200.envs import pusht
from oracle import oracle as orc
from oracle import planner as opl
from tests.conftest import assert_bit_exact

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def env():
    return mbd_b200.envs.get_env("pushT")


def _x0(env, q=None, qd=None):
    st = env.reset(prng.split(prng.PRNGKey(0))[1]).pipeline_state.raw.copy()
    if q is not None:
        for k, v in q.items():
            st[k] = v
    if qd is not None:
        for k, v in qd.items():
            st[8 + k] = v
    return st


def _roll(env, x0, us, **kw):
    us = np.asarray(us, np.float32)
    return orc.pusht_rollout(env.params, x0, us[None] if us.ndim == 2 else us, want_final=True, want_traj=True, want_rewss=True, **kw)


def test_parameter_indices_match_the_header():
    """envs/pusht.py::PT mirrors the enum of include/mbd_pusht.h (evaluated here with a tiny interpreter of the enum)"""
    src = open(os.path.join(ROOT, "include", "mbd_pusht.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define (MBD_PT_\w+) (\d+)", src)}
    body = re.sub(r"/\*.*?\*/", "", src[src.index("enum {"):src.index("};")], flags=re.S).replace("enum {", "")
    val, names = -1, {}
    for item in [t.strip() for t in body.split(",") if t.strip()]:
        if "=" in item:
            name, expr = [t.strip() for t in item.split("=")]
            val = int(eval(expr, {}, {**defs, **names}))   # noqa: S307 - arithmetic on the constants above
        else:
            name, val = item, val + 1
        names[name] = val
    for k, v in pusht.PT.items():
        assert names["MBD_PT_" + k] == v, k
    assert pusht.PT["NPARAM"] == len(mbd_b200.envs.get_env("pushT").params)


def test_model_facts_from_the_reference_xml(env):
    """what scripts/make_pusht_asset.py derived from mbd/assets/pushT.xml"""
    P, I = env.params, pusht.PT
    assert P[I["DT"]] == np.float32(0.01) and P[I["NSUB"]] == 5 and env.dt == pytest.approx(0.05)   # pushT.py:20
    assert P[I["GEAR0"]] == 30 and P[I["GEAR1"]] == 30 and P[I["MP"]] == 1.0 and P[I["RP"]] == np.float32(0.05)
    assert P[I["MS"]] == np.float32(0.1) and P[I["CX"]] == np.float32(-0.05) and P[I["CY"]] == 0.0       # two 0.05 kg boxes, the second at x = -0.1
    assert P[I["IS"]] == pytest.approx(2 * (0.05 / 12 * (0.3 ** 2 + 0.1 ** 2) + 0.05 * 0.05 ** 2), rel=1e-6)
    assert list(P[I["DSX"]:I["DSTH"] + 1]) == [3.0, 3.0, np.float32(0.03)]
    assert list(P[I["LIM0"]:I["LIM0"] + 8]) == [-1, 1] * 4
    assert list(P[I["BOX0"]:I["BOX0"] + 8]) == [0, 0, np.float32(0.15), np.float32(0.05), np.float32(-0.1), 0, np.float32(0.05), np.float32(0.15)]
    assert P[I["MU"]] == 1.0 and env.action_size == 2 and env.observation_size == 16
    assert env.sys.link_names == ["pusher", "slider", "goal"]


def test_reset_is_the_reference_chain(env):
    """pushT.py:22-38: split, pusher at (0.1, -0.15), goal = uniform(-1, 1) * (0.2, 0.2, pi/4) + (-0.4, 0.4, pi)"""
    rng = prng.split(prng.PRNGKey(3))[1]
    st = env.reset(rng)
    _, k = prng.split(rng)
    u = prng.uniform(k, (3,), minval=-1.0, maxval=1.0)
    q = st.pipeline_state.q
    assert q[0] == np.float32(0.1) and q[1] == np.float32(-0.15) and not q[2:5].any()
    np.testing.assert_array_equal(q[5:], (u * np.float32([0.2, 0.2, np.pi / 4]) + np.float32([-0.4, 0.4, np.pi])).astype(np.float32))
    assert not st.pipeline_state.qd.any() and st.obs.shape == (16,)
    assert -0.6 <= q[5] <= -0.2 and 0.2 <= q[6] <= 0.6 and np.pi * 0.75 <= q[7] <= np.pi * 1.25
    # reward of the reset state with the reference expression (pushT.py:52-62)
    r = 1.0 - (np.linalg.norm(q[5:7] - q[2:4]) + abs(q[7] - q[4]) / np.pi + max(np.linalg.norm(q[0:2] - q[2:4]) - 0.2, 0.0))
    assert st.reward == pytest.approx(r, abs=1e-6) and st.done == 0.0


def test_free_pusher_is_a_unit_mass_under_the_motor(env):
    """no contact, no damping on the pusher: v_n = n dt F / m and x_n = x_0 + dt^2 F/m n(n+1)/2 (semi-implicit Euler)"""
    x0 = _x0(env, q={0: 0.5, 1: 0.5})
    out = _roll(env, x0, np.tile([[0.5, -1.0]], (4, 1)))
    n = 4 * 5
    F = np.array([30 * 0.5, -30.0])
    fin = out["final"][0]
    np.testing.assert_allclose(fin[8:10], n * 0.01 * F, rtol=1e-5)
    np.testing.assert_allclose(fin[0:2] - [0.5, 0.5], 0.01 ** 2 * F * n * (n + 1) / 2, rtol=1e-5)
    assert not fin[2:5].any() and not fin[10:13].any()                      # the slider is never touched
    np.testing.assert_array_equal(fin[5:8], x0[5:8])                          # the goal never moves
    # ctrl is clipped to [-1, 1] (motor ctrlrange)
    assert_bit_exact(_roll(env, x0, np.tile([[7.0, -9.0]], (4, 1)))["final"], _roll(env, x0, np.tile([[1.0, -1.0]], (4, 1)))["final"])


def test_joint_damping_is_integrated_implicitly(env):
    """a spinning, drifting slider without forces: v+ = v m / (m + dt d) per step for the hinge; the COM offset couples
    x, y and theta through the centrifugal term, so check the hinge (whose row decouples when x, y rates are zero) and decay"""
    x0 = _x0(env, q={0: 0.9, 1: 0.9}, qd={4: 2.0})
    out = _roll(env, x0, np.zeros((1, 2)))
    w = out["final"][0][12]
    assert 0.0 < w < 2.0
    out2 = _roll(env, x0, np.zeros((40, 2)))
    assert abs(out2["final"][0][12]) < abs(w) and np.isfinite(out2["final"]).all()
    # pure translation of the slider (no spin): x decouples, v+ = v m / (m + dt d) exactly
    x1 = _x0(env, q={0: 0.9, 1: 0.9}, qd={2: 1.0})
    v = _roll(env, x1, np.zeros((1, 2)))["final"][0][10]
    assert v == pytest.approx((0.1 / (0.1 + 0.01 * 3.0)) ** 5, rel=1e-5)


def test_pushing_moves_the_slider_and_conserves_momentum(env):
    """the T is the bar x in [-0.15, 0.15], |y| <= 0.05 plus the crossbar x in [-0.15, -0.05], |y| <= 0.15 (COM at (-0.05, 0)).
    The pusher is driven into the crossbar's left face along the COM line: the slider is carried along without turning, the soft
    contact lets them interpenetrate by millimetres only, and without joint damping the contact impulse is equal and opposite"""
    x0 = _x0(env, q={0: -0.21, 1: 0.0})                     # 0.01 left of the face x = -0.15 (radius 0.05)
    tr = _roll(env, x0, np.tile([[1.0, 0.0]], (4, 1)))["traj"][0]
    assert tr[-1, 2] > 0.2 and abs(tr[-1, 4]) < 1e-3 and abs(tr[-1, 3]) < 1e-4      # pushed in +x, no rotation, no y drift
    gap = (tr[:, 2] - 0.15) - (tr[:, 0] + 0.05)
    assert -0.012 < gap.min() and gap[-1] < 0.0, "in contact, penetration of millimetres (solimp 0.9-0.95, F = 30 N)"
    assert tr[-1, 8] == pytest.approx(tr[-1, 10], rel=0.01), "moving together"
    P = env.params.copy()
    for k in ("DSX", "DSY", "DSTH"):
        P[pusht.PT[k]] = 0.0
    x1 = _x0(env, q={0: -0.21, 1: 0.0}, qd={0: 1.0})        # a flying pusher, no motor force, no damping anywhere
    o = orc.pusht_rollout(P, x1, np.zeros((1, 6, 2), np.float32), want_final=True)["final"][0]
    assert 1.0 * o[8] + 0.1 * o[10] == pytest.approx(1.0, rel=1e-4), "linear momentum"
    assert o[10] > o[8] > 0.5, "the light slider is kicked ahead of the heavy pusher"
    # off-centre push: angular momentum about the origin is conserved too (L = sum m (x vy - y vx) + I w)
    x2 = _x0(env, q={0: 0.05, 1: -0.12}, qd={1: 1.0})       # hits the bar from below at x = 0.05, 0.10 right of the COM
    P[pusht.PT["MU"]] = 0.0     # (friction on the pusher, which cannot rotate, is balanced by its slide joints: an external torque)
    f = orc.pusht_rollout(P, x2, np.zeros((1, 4, 2), np.float32), want_final=True)["final"][0]
    c, s_ = np.cos(f[4]), np.sin(f[4])
    com = np.array([f[2] - 0.05 * c, f[3] - 0.05 * s_]); vcom = np.array([f[10] + f[12] * 0.05 * s_, f[11] - f[12] * 0.05 * c])
    L1 = 1.0 * (f[0] * f[9] - f[1] * f[8]) + 0.1 * (com[0] * vcom[1] - com[1] * vcom[0]) + env.params[pusht.PT["IS"]] * f[12]
    assert L1 == pytest.approx(1.0 * 0.05 * 1.0, rel=2e-3) and f[12] > 1.0, "the bar spins up, total angular momentum stays"
    assert 1.0 * f[9] + 0.1 * vcom[1] == pytest.approx(1.0, rel=1e-4)


def test_friction_drags_the_slider_sideways(env):
    """mu = 1: a pusher sliding along the bar while pressing on it drags the bar with it; with mu = 0 it cannot"""
    x0 = _x0(env, q={0: 0.05, 1: -0.099}, qd={0: 0.5})      # touching the bar from below (1 mm in), moving in +x
    us = np.tile([[0.3, 0.6]], (6, 1))
    a = _roll(env, x0, us)["final"][0]
    P = env.params.copy(); P[pusht.PT["MU"]] = 0.0
    b = orc.pusht_rollout(P, x0, np.asarray(us, np.float32)[None], want_final=True)["final"][0]
    assert a[2] > 0.002 and abs(b[2]) < 0.2 * a[2]


def test_joint_limit_stops_the_pusher(env):
    x0 = _x0(env, q={0: 0.97, 1: 0.9})
    tr = _roll(env, x0, np.tile([[1.0, 0.0]], (30, 1)))["traj"][0]
    assert tr[:, 0].max() < 1.05 and abs(tr[-1, 8]) < 0.5    # held near the +1 limit by the soft constraint against F = 30


def test_oracle_is_deterministic_and_sample_independent(env):
    rng = np.random.default_rng(0)
    Y = np.clip(rng.normal(size=(9, 12, 2)).astype(np.float32), -1, 1)
    x0 = _x0(env, q={0: -0.05, 1: -0.16})
    a = orc.pusht_rollout(env.params, x0, Y, want_final=True, want_rewss=True)
    b = orc.pusht_rollout(env.params, x0, Y[4:5], want_final=True, want_rewss=True, nthreads=1)
    assert_bit_exact(a["final"][4:5], b["final"]); assert_bit_exact(a["rewss"][4:5], b["rewss"])
    np.testing.assert_allclose(a["rews"], a["rewss"].mean(axis=1), rtol=1e-6)


def test_trajectory_export_document(env):
    """the Brax-visualizer JSON document of a pushT rollout (mbd_planner.py:171-178 / vis_diffusion.py:27-112): three links, the
    sphere and the four boxes with the colours of the XML, the table under "world", x.pos / x.rot stacked over time"""
    from mbd_b200.io import brax_json
    st = env.reset(prng.split(prng.PRNGKey(0))[1])
    tr = _roll(env, st.pipeline_state.raw, np.tile([[0.0, 1.0]], (3, 1)))["traj"][0]
    states = [st.pipeline_state] + [env.pipeline_init(r[:8], r[8:]) for r in tr]
    doc = brax_json.to_dict(env.sys, states, env.dt)
    assert doc["link_names"] == ["pusher", "slider", "goal"] and doc["opt"]["timestep"] == pytest.approx(0.05)
    assert [g["name"] for g in doc["geoms"]["slider"]] == ["Box", "Box"] and doc["geoms"]["pusher"][0]["name"] == "Sphere"
    assert doc["geoms"]["pusher"][0]["rgba"] == [0, 1, 0, 1] and doc["geoms"]["world"][0]["name"] == "Plane"
    pos = np.asarray(doc["states"]["x"]["pos"]); rot = np.asarray(doc["states"]["x"]["rot"])
    assert pos.shape == (4, 3, 3) and rot.shape == (4, 3, 4)
    np.testing.assert_allclose(pos[:, 0, 1], [-0.15] + list(tr[:, 1]), atol=1e-6)
    np.testing.assert_allclose(rot[0, 2], [np.cos(states[0].q[7] / 2), 0, 0, np.sin(states[0].q[7] / 2)], atol=1e-6)
    # vis_diffusion.py:92-96 lifts every frame a little with x.replace(pos=...)
    lifted = states[1].replace(x=states[1].x.replace(pos=states[1].x.pos + np.float32([0, 0, 0.01])))
    assert lifted.x.pos[0, 2] == np.float32(0.01) and "<html>" in brax_json.render(env.sys, states, env.dt)


def test_oracle_matches_the_committed_fixture(env):
    """tests/golden/pusht_oracle.npz (scripts/make_golden.py pusht): pins the oracle against accidental change"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "pusht_oracle.npz"))
    assert_bit_exact(env.params, g["params"], "parameter table")
    o = orc.pusht_rollout(env.params, g["x0"], g["Y0s"], want_rewss=True, want_final=True)
    assert_bit_exact(o["rews"], g["rews"]); assert_bit_exact(o["rewss"], g["rewss"]); assert_bit_exact(o["final"], g["final"])
    s_ = orc.pusht_rollout(env.params, g["x1"], g["script"][None], want_traj=True, want_rewss=True)
    assert_bit_exact(s_["traj"][0], g["script_traj"]); assert_bit_exact(s_["rewss"][0], g["script_rewss"])
    tr = g["script_traj"]
    assert tr[:, 2].max() > 0.9 and np.abs(tr[:, 4]).max() > 1.0, "the scripted push drives the slider to its limit and spins it"


# ---- GPU -----------------------------------------------------------------------------------------------------------
def _T(a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")


@pytest.mark.gpu
def test_kernel_equals_oracle_bit_for_bit(env):
    from mbd_b200 import ops
    rng = np.random.default_rng(5)
    for n, H, start in [(1, 1, {}), (77, 40, {}), (130, 7, {0: -0.05, 1: -0.16}), (64, 25, {0: 0.2, 1: 0.0})]:
        x0 = _x0(env, q=start, qd={1: 0.3} if start else None)
        Y = (rng.normal(size=(n, H, 2)) * 1.2).astype(np.float32)          # includes |u| > 1
        ref = orc.pusht_rollout(env.params, x0, Y, want_rewss=True, want_final=True, want_traj=True)
        out = ops.pusht_rollout(env.device_params(), _T(x0), _T(Y), want_rewss=True, want_final=True, want_traj=True)
        for k in ("rews", "rewss", "final", "traj"):
            assert_bit_exact(out[k].cpu().numpy(), ref[k], f"pushT {k} n={n} H={H}")
        assert np.abs(ref["traj"][:, :, 2:5]).max() > 0 or not start, "the contact cases really push the slider"


@pytest.mark.gpu
def test_env_step_equals_rollout(env):
    from mbd_b200 import ops
    st = env.reset(prng.split(prng.PRNGKey(1))[1])
    us = np.clip(np.random.default_rng(2).normal(size=(6, 2)), -1, 1).astype(np.float32) * np.float32([0.2, 1.0])
    out = ops.pusht_rollout(env.device_params(), _T(st.pipeline_state.raw), _T(us[None]), want_rewss=True, want_traj=True)
    s = st
    for t in range(6):
        s = env.step(s, us[t])
        assert_bit_exact(s.pipeline_state.raw, out["traj"][0, t].cpu().numpy(), f"state after step {t}")
        assert_bit_exact(np.float32(s.reward), out["rewss"][0, t].cpu().numpy(), f"reward of step {t}")
    assert s.obs.shape == (16,) and s.pipeline_state.x.pos.shape == (3, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("P", [1, 2])
def test_diffusion_step_vs_oracle(env, P):
    """one reverse step at the reference's recommended pushT shape (Hsample 40, temp 0.2; mbd_planner.py:52-62), one rank and two
    emulated ranks"""
    import torch
    from mbd_b200.planners import engine as eng
    Nn, H, temp, Nd, i = 512, 40, 0.2, 30, 21
    st = env.reset(prng.split(prng.PRNGKey(0))[1]).pipeline_state.raw
    _, alphas, alphas_bar, sigmas = opl.make_schedule(1e-4, 1e-2, Nd)
    keys = eng.key_chain(np.uint32([5, 6]), Nd)
    Ybar_i = (np.random.default_rng(3).normal(size=2 * H) * 0.3).astype(np.float32)
    ref = opl.reverse_once(opl.OracleEnv("pusht", 2, params=env.params, x0=st), keys[i], Nn, H, float(sigmas[i]), Ybar_i, temp,
                           alphas, alphas_bar, i)
    engines = ([eng.DiffusionEngine(env, Nn, H, temp, False, st, Ndiffuse=Nd)] if P == 1 else
               eng.DiffusionEngine.make_emulated_ranks(env, Nn, H, temp, False, st, P, Ndiffuse=Nd))
    for e in engines:
        e.load_schedule(keys, sigmas, alphas, alphas_bar); e.set_step(i)
        e.Ybars[i].copy_(torch.as_tensor(Ybar_i, device="cuda:0"))
    if P == 1:
        engines[0].step()
    else:
        eng.DiffusionEngine.step_emulated_ranks(engines)
    torch.cuda.synchronize()
    assert_bit_exact(np.concatenate([e.Y0s.cpu().numpy() for e in engines]), ref["Y0s"], "sampled actions")
    assert_bit_exact(np.concatenate([e.rews_local.cpu().numpy() for e in engines]), ref["rews"], "per-sample returns")
    scale = max(float(np.abs(ref["Ybar_im1"]).max()), 1e-6)
    for e in engines:
        e.check_exchange()
        assert np.abs(e.Ybars[i - 1].cpu().numpy() - ref["Ybar_im1"]).max() / scale < 1e-4
        assert abs(float(e.rew_hist[i].item()) - float(ref["rew_mean"])) < 1e-4 * max(abs(float(ref["rew_mean"])), 1e-6) + 1e-6


@pytest.mark.gpu
def test_run_diffusion_pushT_short(tmp_path, monkeypatch, capsys):
    """the reference CLI surface on pushT: recommended parameters are applied (temp 0.2, Ndiffuse 200, Hsample 40) and the solve
    does not end below the reward of the zero plan"""
    from mbd_b200.planners.mbd_planner import Args, run_diffusion
    args = Args(env_name="pushT", Nsample=512, not_render=True)
    rew = run_diffusion(args)
    assert args.Hsample == 40 and args.temp_sample == pytest.approx(0.2) and args.Ndiffuse == 200   # mbd_planner.py:52-62
    env = mbd_b200.envs.get_env("pushT")
    st = env.reset(prng.split(prng.PRNGKey(args.seed))[1]).pipeline_state.raw
    zero = orc.pusht_rollout(env.params, st, np.zeros((1, 40, 2), np.float32))["rews"][0]
    assert np.isfinite(rew) and rew >= zero - 1e-3


@pytest.mark.gpu
def test_kernel_matches_the_committed_fixture(env):
    """the CUDA path against tests/golden/pusht_oracle.npz (no oracle involved on the GPU box)"""
    from mbd_b200 import ops
    g = np.load(os.path.join(ROOT, "tests", "golden", "pusht_oracle.npz"))
    o = ops.pusht_rollout(_T(g["params"]), _T(g["x0"]), _T(g["Y0s"]), want_rewss=True, want_final=True)
    for k in ("rews", "rewss", "final"):
        assert_bit_exact(o[k].cpu().numpy(), g[k], f"pushT golden {k}")
    s_ = ops.pusht_rollout(_T(g["params"]), _T(g["x1"]), _T(g["script"][None]), want_traj=True, want_rewss=True)
    assert_bit_exact(s_["traj"][0].cpu().numpy(), g["script_traj"]); assert_bit_exact(s_["rewss"][0].cpu().numpy(), g["script_rewss"])
