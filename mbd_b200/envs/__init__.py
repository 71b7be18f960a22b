"""Env registry — mirrors /root/reference/mbd/envs/__init__.py:13-33 (same names, same ValueError)."""
from .ant import Ant  # noqa: F401
from .car2d import Car2d
from .generic import GenericPositionalEnv  # noqa: F401
from .humanoidrun import HumanoidRun
from .humanoidstandup import HumanoidStandup
from .humanoidtrack import HumanoidTrack

_NOT_VENDORED = {
    "hopper": "its MJCF lives inside the Brax wheel (hopper.py:13), not in the reference tree",
    "walker2d": "its MJCF lives inside the Brax wheel (walker2d.py:14)",
    "halfcheetah": "env and MJCF are Brax's stock `halfcheetah`, not in the reference tree",
    "pushT": "uses Brax's `generalized` backend (pushT.py:16), outside the positional hot path",
    "cartpole": "slide joints are not enabled in this round (SURVEY 8f.3)",
}


def get_env(env_name: str):
    if env_name == "humanoidrun":
        return HumanoidRun()
    elif env_name == "humanoidstandup":
        return HumanoidStandup()
    elif env_name == "humanoidtrack":
        return HumanoidTrack()
    elif env_name == "car2d":
        return Car2d()
    elif env_name == "ant":
        return Ant()   # raises NotImplementedError with instructions when Brax's ant.xml cannot be found
    elif env_name in _NOT_VENDORED:
        raise NotImplementedError(f"environment {env_name!r} is recognised but not available: {_NOT_VENDORED[env_name]}")
    else:
        raise ValueError(f"Unknown environment: {env_name}")
