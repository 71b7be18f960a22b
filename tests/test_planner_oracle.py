"""Planner arithmetic of the oracle (oracle/planner.py) pinned by the in-repo reference facts:
schedule constants, the Ybar_im1 == Ybar identity (App. D), the demo branch, and
the car2d full-solve fixtures (BASELINE config 1)."""
import os

import numpy as np

import mbd_b200
from mbd_b200.planners import engine as eng
from oracle import planner as opl

G = os.path.join(os.path.dirname(__file__), "golden")


def test_schedule_kats():
    for N, exp in ((100, 0.6305), (200, 0.7981), (300, 0.8839)):
        s = opl.make_schedule(1e-4, 1e-2, N)[3]
        assert abs(float(s[-1]) - exp) < 5e-5
        assert f"{s[-1]:.2e}" == f"{exp:.2e}"          # the line the reference prints (mbd_planner.py:93)
        # product-side schedule is the same arithmetic
        assert np.array_equal(s, eng.make_schedule(1e-4, 1e-2, N)[3])


def test_update_identity_appendix_d():
    _, alphas, alphas_bar, _ = opl.make_schedule(1e-4, 1e-2, 100)
    rng = np.random.default_rng(0)
    Ybar_i = rng.uniform(-1, 1, 100).astype(np.float32)
    Ybar = rng.uniform(-1, 1, 100).astype(np.float32)
    for i in (99, 50, 1):
        out = opl.update(Ybar_i, Ybar, alphas, alphas_bar, i)
        assert np.allclose(out, Ybar, rtol=0, atol=2e-6)   # Ybar_im1 == Ybar up to a few ulp
        c = eng.update_coef(alphas, alphas_bar, i)
        Yi = Ybar_i * c[0]
        lit = (c[3] * (Yi + c[2] * (c[1] * (-Yi + c[0] * Ybar)))) / c[4]
        assert np.array_equal(out, lit.astype(np.float32))


def test_stats_softmax_and_demo_branch():
    rng = np.random.default_rng(1)
    rews = rng.normal(size=256).astype(np.float32)
    Y = rng.uniform(-1, 1, (256, 20)).astype(np.float32)
    Ybar, mean, w = opl.reverse_once_stats(rews, Y, 0.1)
    assert abs(w.sum() - 1) < 1e-5 and w.argmax() == rews.argmax()
    assert np.isclose(mean, rews.mean())
    # std guard: constant rewards -> uniform weights -> plain mean
    Ybar0, _, w0 = opl.reverse_once_stats(np.zeros(256, np.float32), Y, 0.1)
    assert np.allclose(w0, 1 / 256) and np.allclose(Ybar0, Y.mean(0), atol=1e-6)
    # demo branch: a sample whose demo log-density dominates gets its logit replaced
    logpd = np.full(256, -1.0, np.float32); logpd[7] = 0.0
    _, _, wd = opl.reverse_once_stats(rews - 5.0, Y, 0.1, logpd=logpd, rew_xref=1.0)
    assert wd.argmax() == 7


def test_car2d_full_solve_fixture():
    """BASELINE config 1: car2d Nsample=64 Hsample=40, one full diffusion solve on the CPU."""
    g = np.load(os.path.join(G, "car2d_oracle.npz"))
    car = mbd_b200.envs.get_env("car2d")
    env = opl.OracleEnv("car2d", 2, params=car.params, x0=car.x0)
    rf, Yi, rews = opl.run_diffusion(env, 0, 64, 40, 100, 0.1)
    assert Yi.shape == (99, 80)
    assert np.float32(rf) == g["rew_final"] and np.array_equal(Yi[-1], g["Yi_last"]) and np.array_equal(rews, g["rews"])


def test_car2d_dynamics_properties(orc):
    car = mbd_b200.envs.get_env("car2d")
    # zero action: the car does not move, reward of the start pose
    out = orc.car2d_rollout(car.params, car.x0, np.zeros((1, 5, 2), np.float32), want_traj=True, want_rewss=True)
    assert np.allclose(out["traj"][0], car.x0) and np.allclose(out["rewss"], 0.0)
    # full throttle straight ahead from heading 3pi/2: x' = 3 sin(theta) = -3 -> moves in -x
    u = np.zeros((1, 3, 2), np.float32); u[..., 1] = 1.0
    tr = orc.car2d_rollout(car.params, car.x0, u, want_traj=True)["traj"][0]
    assert np.allclose(tr[:, 0], -0.5 - 0.3 * np.arange(1, 4), atol=1e-5) and np.allclose(tr[:, 1], 0, atol=1e-5)
    # driving into an obstacle freezes the state (car2d.py:82-83): start next to the disc at (0,0)
    x0 = np.float32([-0.35, 0.0, np.pi / 2])
    u = np.zeros((1, 4, 2), np.float32); u[..., 1] = 1.0
    tr = orc.car2d_rollout(car.params, x0, u, want_traj=True)["traj"][0]
    assert np.allclose(tr, x0)
    # action clip at +-1
    a = orc.car2d_rollout(car.params, car.x0, np.full((1, 2, 2), 5.0, np.float32), want_traj=True)["traj"]
    b = orc.car2d_rollout(car.params, car.x0, np.full((1, 2, 2), 1.0, np.float32), want_traj=True)["traj"]
    assert np.array_equal(a, b)
    assert abs(car.rew_xref - 0.18002363) < 1e-6
