"""Physical sanity of the oracle's positional (XPBD) step on hand-built 1-2 link systems, plus the
regression fixture of the humanoid.  These are the substitute for the golden vectors the
reference does not have (Brax is un-vendored: parity unpinned, see oracle/mbd_oracle.c header)."""
import os
import tempfile

import numpy as np
import pytest

from mbd_b200.model import blob, kinematics, mjcf

HDR = """<mujoco><compiler angle="degree" inertiafromgeom="true"/>
<option timestep="0.005"/>
<custom><numeric data="1" name="spring_inertia_scale"/><numeric data="0" name="spring_mass_scale"/>
<numeric data="{ang_damp}" name="constraint_ang_damping"/><numeric data="0.5" name="joint_scale_pos"/>
<numeric data="0.2" name="joint_scale_ang"/></custom><worldbody>{floor}"""
FLOOR = '<geom name="floor" type="plane" size="5 5 .1" conaffinity="1" contype="0" friction="1 .1 .1"/>'


def _load(xml):
    with tempfile.NamedTemporaryFile("w", suffix=".xml", delete=False) as f:
        f.write(xml)
    try:
        return mjcf.load(f.name)
    finally:
        os.unlink(f.name)


def _ball(z=1.0, floor=True, contype=1):
    return _load(HDR.format(ang_damp=0, floor=FLOOR if floor else "") +
                 f'<body name="b" pos="0 0 {z}"><joint type="free" name="root"/>'
                 f'<geom type="sphere" size="0.1" contype="{contype}" conaffinity="0"/></body></worldbody></mujoco>')


def _pendulum(act=False):
    s = HDR.format(ang_damp=0, floor="") + """<body name="base" pos="0 0 2"><joint type="free" name="root"/>
      <geom type="sphere" size="0.2" contype="0" conaffinity="0"/>
      <body name="arm" pos="0 0 -0.3"><joint type="hinge" name="h" axis="0 1 0" pos="0 0 0.15" range="-90 90" limited="true"/>
        <geom type="capsule" fromto="0 0 0.1 0 0 -0.4" size="0.05" contype="0" conaffinity="0"/></body></body></worldbody>"""
    if act:
        s += '<actuator><motor joint="h" gear="10" ctrllimited="true" ctrlrange="-1 1"/></actuator>'
    return _load(s + "</mujoco>")


def _roll(orc, sys, nsub, q=None, qd=None, act=None, reward=blob.REWARD_HUMANOIDRUN):
    b = blob.pack(sys, 1, reward)
    q = sys.init_q if q is None else q
    qd = np.zeros(sys.qd_size()) if qd is None else qd
    st = kinematics.pipeline_init(sys, q, qd)
    nu = max(sys.act_size(), 1)
    Y = np.zeros((1, 1, nu), np.float32) if act is None else np.float32(act).reshape(1, 1, nu)
    if sys.act_size() == 0:
        b = b.copy(); b.view(np.int32)[blob.H_NU] = 1
    return st, orc.xpbd_rollout(b, st, Y, want_final=True, nsub_override=nsub)["final"][0]


def test_free_fall_is_semi_implicit_euler(orc):
    sys = _ball(z=5.0, floor=False)
    n, dt, g = 100, 0.005, -9.81
    st, fin = _roll(orc, sys, n)
    # v_k = k g dt ; z_n = z0 + g dt^2 n(n+1)/2 ; velocity is re-derived from positions (project_xd)
    assert np.isclose(fin[0, 2], 5.0 + g * dt * dt * n * (n + 1) / 2, rtol=0, atol=3e-3)
    assert np.isclose(fin[0, 12], g * dt * n, rtol=2e-3)
    assert np.allclose(fin[0, 3:7], [1, 0, 0, 0], atol=1e-6) and np.allclose(fin[0, 0:2], 0, atol=1e-6)


def test_ball_comes_to_rest_on_the_plane(orc):
    sys = _ball(z=0.3)
    st, fin = _roll(orc, sys, 600)
    assert abs(fin[0, 2] - 0.1) < 5e-3, fin[0]          # sphere radius 0.1 resting on z = 0
    assert np.abs(fin[0, 10:13]).max() < 0.05


def test_friction_stops_sliding_ball(orc):
    sys = _ball(z=0.1)
    qd = np.zeros(6); qd[0] = 1.0                        # 1 m/s along x, resting height
    st, fin = _roll(orc, sys, 400, qd=qd)
    assert abs(fin[0, 10]) < 0.05 and 0.0 < fin[0, 0] < 0.5   # decelerated by mu = 1 contact friction
    st, fin_nofloor = _roll(orc, _ball(z=0.1, contype=0), 400, qd=qd)
    assert abs(fin_nofloor[0, 10] - 1.0) < 1e-3          # without a colliding geom nothing brakes it


def test_pendulum_joint_holds_and_momentum_is_conserved(orc):
    sys = _pendulum()
    q = sys.init_q.copy(); q[7] = 0.8
    st, fin = _roll(orc, sys, 200, q=q)
    # anchor distance stays small (XPBD translation constraint)
    from mbd_b200.model.mjcf import rotate
    a_c = fin[1, 0:3] + rotate(sys.joint_pos[1] - sys.com[1], fin[1, 3:7].astype(np.float64))
    anchor_p = sys.link_pos[1] + rotate(sys.joint_pos[1], sys.link_rot[1]) - sys.com[0]
    a_p = fin[0, 0:3] + rotate(anchor_p, fin[0, 3:7].astype(np.float64))
    assert np.linalg.norm(a_c - a_p) < 2e-2
    # joint corrections are internal: total linear momentum = (sum m) g t exactly along z, 0 along x,y
    p = (sys.mass[:, None] * fin[:, 10:13]).sum(0)
    assert np.allclose(p[:2], 0, atol=2e-3)
    assert np.isclose(p[2], sys.mass.sum() * -9.81 * 200 * 0.005, rtol=5e-3)


def test_actuator_torque_direction_and_limit(orc):
    sys = _pendulum(act=True)
    b = blob.pack(sys, 1, blob.REWARD_HUMANOIDRUN)
    st = kinematics.pipeline_init(sys, sys.init_q, np.zeros(sys.qd_size()))
    out = {}
    for u in (1.0, -1.0, 5.0):
        fin = orc.xpbd_rollout(b, st, np.float32([[[u]]]), want_final=True, nsub_override=20)["final"][0]
        xp, xr, xa, xv = (fin[:, 0:3], fin[:, 3:7], fin[:, 7:10], fin[:, 10:13])
        x, r, a, v = kinematics.to_world(sys, fin)
        out[u] = kinematics.inverse(sys, x, r, a, v)[0][7]
    assert out[1.0] > 0.01 and out[-1.0] < -0.01        # positive control -> positive joint angle
    assert out[5.0] == out[1.0]                          # ctrl_range clip (actuator.to_tau)
    # joint limit: drive hard into the +90 deg stop for a long time
    fin = orc.xpbd_rollout(b, st, np.float32([[[1.0]]]), want_final=True, nsub_override=1500)["final"][0]
    x, r, a, v = kinematics.to_world(sys, fin)
    assert kinematics.inverse(sys, x, r, a, v)[0][7] < np.pi / 2 + 0.15


def test_capsule_rests_on_its_two_end_caps(orc):
    """MJX plane_capsule = one sphere contact per end cap (humanoidstandup's torso/thigh/forearm geoms)."""
    sys = _load(HDR.format(ang_damp=0, floor=FLOOR) +
                '<body name="b" pos="0 0 0.3"><joint type="free" name="root"/>'
                '<geom type="capsule" fromto="-0.2 0 0 0.2 0 0" size="0.05" contype="1" conaffinity="0"/></body></worldbody></mujoco>')
    assert len(sys.contacts) == 2 and {round(float(c["pos"][0]), 3) for c in sys.contacts} == {-0.2, 0.2}
    st, fin = _roll(orc, sys, 800)
    assert abs(fin[0, 2] - 0.05) < 5e-3 and np.abs(fin[0, 10:13]).max() < 0.05
    assert abs(fin[0, 3]) > 0.999          # stays flat: both caps carry it


def test_humanoidstandup_oracle_is_stable(orc):
    import mbd_b200
    from mbd_b200 import prng
    env = mbd_b200.envs.get_env("humanoidstandup")
    assert len(env.sys.contacts) == 15
    st = env.reset(prng.split(prng.PRNGKey(0))[1]).pipeline_state.raw
    out = orc.xpbd_rollout(env.blob, st, np.zeros((1, 100, 17), np.float32), want_final=True, want_rewss=True)
    fin = out["final"][0]
    assert np.isfinite(fin).all() and 0.03 < fin[0, 2] < 0.2 and np.abs(fin[:, 10:13]).max() < 0.2   # lies on the floor
    # reward = 1.5 - clip(|z-1.3|,-2,1) - 0.1|x| - 0.1|y| with the torso on the ground: ~0.5
    assert abs(out["rewss"][0, -1] - 0.5) < 0.05


def test_humanoid_regression_fixture(orc, humanoidrun_setup):
    """Oracle output pinned by a committed fixture (scripts/make_golden.py)."""
    env, b, st = humanoidrun_setup
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "humanoidrun_oracle.npz"))
    out = orc.xpbd_rollout(b, st, g["Y0s"], want_rewss=True, want_final=True)
    assert np.array_equal(out["rews"].view(np.uint32), g["rews"].view(np.uint32))
    assert np.array_equal(out["final"].view(np.uint32), g["final"].view(np.uint32))
    assert np.isfinite(out["final"]).all()
