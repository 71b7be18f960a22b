"""mbd_b200 — B200-native implementation of the Model-Based Diffusion reverse-step hot path.

Mirrors the reference package layout for the path in scope: `mbd_b200.envs.get_env`,
`mbd_b200.utils.rollout_us`, `mbd_b200.planners.mbd_planner.{Args, run_diffusion}`.
"""
from . import envs, planners, utils  # noqa: F401

__all__ = ["envs", "planners", "utils"]
