"""Path-integral baselines (MPPI / CMA-ES / CEM) — drop-in for
/root/reference/mbd/planners/path_integral.py: same `Args`, same recommended-parameter table, same
`run_path_integral(args) -> rew`.  The sample -> rollout -> softmax skeleton reuses the MBD kernels
unchanged; only the update rules differ:
    mppi   mu = einsum(w, Y0s)                                      (path_integral.py:33-36)
    cma-es mu as mppi; sigma = sqrt(einsum(w, (Y0s-mu_0t)^2)).mean() * sigma, floored at 1e-3 (:39-45)
    cem    mu = mean of the 10 best samples, idx = argsort(w)[::-1][:10]  (:48-52, bit-exact index work)
One deviation: the shared statistics kernel keeps MBD's guard `std < 1e-4 -> 1` (mbd_planner.py:112),
which path_integral.py:121 lacks (there a zero std turns every weight into NaN).
Single-GPU (the reference is single-device; mppi alone would shard like MBD).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

import mbd_b200
from mbd_b200 import ops, prng
from mbd_b200.planners.engine import DiffusionEngine

try:
    from tqdm import tqdm
except Exception:  # noqa: BLE001
    tqdm = None


@dataclass
class Args:
    # exp
    seed: int = 0
    disable_recommended_params: bool = False
    update_method: str = "mppi"  # mppi, cma-es, cem
    # env
    env_name: str = (
        "ant"  # "humanoidstandup", "ant", "halfcheetah", "hopper", "walker2d"
    )
    # diffusion
    Nsample: int = 2048  # number of samples
    Hsample: int = 50  # horizon
    Nrefine: int = 100  # number of repeat steps
    temp_sample: float = 0.1  # temperature for sampling


TEMP_RECOMMEND = {"ant": 0.1, "halfcheetah": 0.4, "hopper": 0.1, "humanoidstandup": 0.1, "humanoidrun": 0.1, "walker2d": 0.1,
                  "pushT": 0.2}
NREFINE_RECOMMEND = {"pushT": 200, "humanoidrun": 300}
NSAMPLE_RECOMMEND = {"humanoidrun": 8192}
HSAMPLE_RECOMMEND = {"pushT": 40}


def apply_recommended_params(args: Args) -> Args:
    """path_integral.py:86-91."""
    if not args.disable_recommended_params:
        args.temp_sample = TEMP_RECOMMEND.get(args.env_name, args.temp_sample)
        args.Nrefine = NREFINE_RECOMMEND.get(args.env_name, args.Nrefine)
        args.Nsample = NSAMPLE_RECOMMEND.get(args.env_name, args.Nsample)
        args.Hsample = HSAMPLE_RECOMMEND.get(args.env_name, args.Hsample)
        print(f"override temp_sample to {args.temp_sample}")
    return args


class PathIntegralEngine(DiffusionEngine):
    """update_once (path_integral.py:111-127) on the device."""

    def __init__(self, env, Nsample, Hsample, temp_sample, state_init, update_method: str):
        super().__init__(env, Nsample, Hsample, temp_sample, False, state_init)
        if self.P != 1:
            raise NotImplementedError("path-integral baselines are single-GPU")
        if update_method not in ("mppi", "cma-es", "cem"):
            raise KeyError(update_method)
        self.update_method = update_method
        self.zeros = torch.zeros(self.HNu, device=self.device)
        self.sq = torch.empty(self.HNu, device=self.device)

    def update_once(self, key, mu_0t: torch.Tensor, sigma: float, out: torch.Tensor):
        """returns (mu_0tm1 [device], sigma [host float], rews.mean() [device scalar])"""
        self.rollout_phase(key, sigma, mu_0t)                       # eps*sigma + mu_0t, clip, eval_us
        ops.softmax_weights(self.rews_local, None, 0, self.n_local, self.temp, 0.0, self.weights, self.scalars, self.logp_scratch)
        if self.update_method == "cem":
            order = torch.sort(self.weights, stable=True).indices   # jnp.argsort (stable, ascending)
            idx = order.flip(0)[:10]                                # [::-1][:10]
            out.copy_(self.Y0s[idx].mean(dim=0))
            return out, sigma, self.scalars[0]
        ops.weighted_sum(self.weights, self.Y0s, self.HNu, self.run_scratch, self.partial)
        ops.update(self.partial, 1, self.HNu, self.zeros, [1.0, 1.0, 1.0, 1.0, 1.0], out)   # out = einsum(w, Y0s) exactly
        if self.update_method == "cma-es":
            ops.weighted_sqerr_sum(self.weights, self.Y0s, mu_0t, self.HNu, self.run_scratch, self.sq)
            sigma = float(torch.sqrt(self.sq).mean().item()) * sigma
            sigma = max(sigma, 1e-3)
        return out, sigma, self.scalars[0]


def run_path_integral(args: Args, log_every: int = 10, return_trajectory: bool = False):
    rng = prng.PRNGKey(seed=args.seed)
    apply_recommended_params(args)
    env = mbd_b200.envs.get_env(args.env_name)
    Nu = env.action_size
    rng, rng_reset = prng.split(rng)  # NOTE: rng_reset should never be changed.
    state_init = env.reset(rng_reset)
    eng = PathIntegralEngine(env, args.Nsample, args.Hsample, args.temp_sample, state_init, args.update_method)
    HNu = args.Hsample * Nu
    mus = torch.zeros((args.Nrefine, HNu), device=eng.device)
    rews = torch.zeros(args.Nrefine, device=eng.device)
    rng_exp, rng = prng.split(rng)
    r = rng_exp
    sigma = 1.0
    steps = range(args.Nrefine - 1, 0, -1)
    pbar = tqdm(steps, desc="Path Integrating") if tqdm is not None else None
    for n_done, t in enumerate(pbar if pbar is not None else steps):
        r, k = prng.split(r)
        _, sigma, rew = eng.update_once(k, mus[t], sigma, mus[t - 1])
        rews[t].copy_(rew, non_blocking=True)
        if pbar is not None and (n_done % log_every == log_every - 1 or t == 1):
            pbar.set_postfix({"rew": f"{rews[t].item():.2e}"})
    mu_0ts = mus[: args.Nrefine - 1].flip(0).reshape(args.Nrefine - 1, args.Hsample, Nu)
    from mbd_b200.planners.mbd_planner import final_reward
    rew_final = final_reward(env, eng, mu_0ts[-1])
    if return_trajectory:
        return rew_final, mu_0ts
    return rew_final


if __name__ == "__main__":
    import tyro

    rew = run_path_integral(args=tyro.cli(Args))
    print(f"rew: {rew:.2e}")
