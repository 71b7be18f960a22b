import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def humanoidrun_setup():
    """(sys, blob, state_init) of the north-star env with the seed-0 reset of the planner."""
    import mbd_b200
    from mbd_b200 import prng
    env = mbd_b200.envs.get_env("humanoidrun")
    rng, rng_reset = prng.split(prng.PRNGKey(0))
    st = env.reset(rng_reset)
    return env, env.blob, st.pipeline_state.raw


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_exact(a, b, what=""):
    a, b = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    same = bits(a) == bits(b)
    if not same.all():
        idx = np.argwhere(~same)[0]
        raise AssertionError(f"{what}: {np.count_nonzero(~same)} of {same.size} words differ; first at {tuple(idx)}: "
                             f"{a[tuple(idx)]!r} vs {b[tuple(idx)]!r}")
