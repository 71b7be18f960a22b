"""Env registry — mirrors /root/reference/mbd/envs/__init__.py:13-33 (same names, same ValueError)."""
from .ant import Ant
from .car2d import Car2d
from .cartpole import Cartpole
from .generic import GenericPositionalEnv  # noqa: F401
from .halfcheetah import HalfCheetah
from .hopper import Hopper
from .humanoidrun import HumanoidRun
from .humanoidstandup import HumanoidStandup
from .humanoidtrack import HumanoidTrack
from .pusht import PushT
from .walker2d import Walker2d

_NOT_AVAILABLE = {}   # every env name of the reference's registry is built (pushT since round 2: envs/pusht.py)


def get_env(env_name: str):
    if env_name == "pushT":
        return PushT()
    elif env_name == "hopper":
        return Hopper()
    elif env_name == "humanoidstandup":
        return HumanoidStandup()
    elif env_name == "humanoidrun":
        return HumanoidRun()
    elif env_name == "humanoidtrack":
        return HumanoidTrack()
    elif env_name == "walker2d":
        return Walker2d()
    elif env_name == "cartpole":
        return Cartpole()
    elif env_name == "car2d":
        return Car2d()
    elif env_name == "ant":
        return Ant()
    elif env_name == "halfcheetah":
        return HalfCheetah()
    elif env_name in _NOT_AVAILABLE:
        raise NotImplementedError(f"environment {env_name!r} is recognised but not available: {_NOT_AVAILABLE[env_name]}")
    else:
        raise ValueError(f"Unknown environment: {env_name}")
