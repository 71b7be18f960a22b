"""Accuracy of the shared fp32 math specification (include/mbd_fp32.h) against float64."""
import numpy as np
from scipy.special import erfinv


def _ulp_err(got, ref):
    ref32 = ref.astype(np.float32)
    ulp = np.spacing(np.abs(ref32)).astype(np.float64)
    return np.max(np.abs(got.astype(np.float64) - ref) / ulp)


def test_atan2(orc):
    rng = np.random.default_rng(0)
    y = rng.normal(size=100000).astype(np.float32)
    x = rng.normal(size=100000).astype(np.float32)
    assert _ulp_err(orc.fmap("atan2", y, x), np.arctan2(y.astype(np.float64), x.astype(np.float64))) <= 4.0
    # axes, signs and zeros
    yy = np.float32([0, 0, 1, -1, 0.0, 1e-30, -1e-30, 3, -3])
    xx = np.float32([1, -1, 0, 0, 0.0, 1, -1, 3, -3])
    got = orc.fmap("atan2", yy, xx)
    ref = np.arctan2(yy.astype(np.float64), xx.astype(np.float64))
    assert np.allclose(got, ref, atol=3e-7)


def test_sincos(orc):
    a = np.random.default_rng(1).uniform(-40, 40, size=100000).astype(np.float32)
    assert np.max(np.abs(orc.fmap("sin", a) - np.sin(a.astype(np.float64)))) < 2.5e-7
    assert np.max(np.abs(orc.fmap("cos", a) - np.cos(a.astype(np.float64)))) < 2.5e-7


def test_log_exp(orc):
    rng = np.random.default_rng(2)
    l = np.exp(rng.uniform(-17, 5, size=100000)).astype(np.float32)
    assert _ulp_err(orc.fmap("log", l), np.log(l.astype(np.float64))) <= 3.0 or \
        np.max(np.abs(orc.fmap("log", l) - np.log(l.astype(np.float64)))) < 1e-6
    e = rng.uniform(-86.9, 0, size=100000).astype(np.float32)
    assert _ulp_err(orc.fmap("exp", e), np.exp(e.astype(np.float64))) <= 3.0
    assert orc.fmap("exp", np.float32([-100.0, 0.0]))[0] == 0.0
    assert orc.fmap("exp", np.float32([0.0]))[0] == 1.0


def test_erfinv(orc):
    u = np.random.default_rng(3).uniform(-1, 1, size=100000).astype(np.float32)
    got, ref = orc.fmap("erfinv", u), erfinv(u.astype(np.float64))
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-6
    edge = np.float32([-0.99999994, 0.99999994, 0.0])
    assert np.isfinite(orc.fmap("erfinv", edge)).all()
