"""MBD planner — drop-in for /root/reference/mbd/planners/mbd_planner.py.

Same `Args` fields, same recommended-parameter override, same `run_diffusion(args) -> rew_final`
signature, stdout lines and `results/{env}/mu_0ts.npy` artefact; the jitted `reverse_once` is
replaced by `DiffusionEngine.reverse_once` (hand-written sm_100a CUDA behind the C ABI).
Under torchrun (WORLD_SIZE > 1) the Nsample axis is sharded over ranks.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

import mbd_b200
from mbd_b200 import ops, prng
from mbd_b200.planners.engine import DiffusionEngine, key_chain, make_schedule

try:  # tqdm is cosmetic
    from tqdm import tqdm
except Exception:  # noqa: BLE001
    tqdm = None


## load config
@dataclass
class Args:
    # exp
    seed: int = 0
    disable_recommended_params: bool = False
    not_render: bool = False
    # env
    env_name: str = (
        "ant"  # "humanoidstandup", "ant", "halfcheetah", "hopper", "walker2d", "car2d"
    )
    # diffusion
    Nsample: int = 2048  # number of samples
    Hsample: int = 50  # horizon
    Ndiffuse: int = 100  # number of diffusion steps
    temp_sample: float = 0.1  # temperature for sampling
    beta0: float = 1e-4  # initial beta
    betaT: float = 1e-2  # final beta
    enable_demo: bool = False


# recommended parameters (mbd_planner.py:45-63)
TEMP_RECOMMEND = {"ant": 0.1, "halfcheetah": 0.4, "hopper": 0.1, "humanoidstandup": 0.1, "humanoidrun": 0.1, "walker2d": 0.1,
                  "pushT": 0.2}
NDIFFUSE_RECOMMEND = {"pushT": 200, "humanoidrun": 300}
NSAMPLE_RECOMMEND = {"humanoidrun": 8192}
HSAMPLE_RECOMMEND = {"pushT": 40}


def apply_recommended_params(args: Args) -> Args:
    """mbd_planner.py:64-69 — mutates args in place exactly like the reference."""
    if not args.disable_recommended_params:
        args.temp_sample = TEMP_RECOMMEND.get(args.env_name, args.temp_sample)
        args.Ndiffuse = NDIFFUSE_RECOMMEND.get(args.env_name, args.Ndiffuse)
        args.Nsample = NSAMPLE_RECOMMEND.get(args.env_name, args.Nsample)
        args.Hsample = HSAMPLE_RECOMMEND.get(args.env_name, args.Hsample)
        if _is_main():   # one line per job, as in the single-process reference (every rank of a torchrun job runs this)
            print(f"override temp_sample to {args.temp_sample}")
    return args


def _is_main() -> bool:
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


def run_diffusion(args: Args, log_every: int = 10, return_trajectory: bool = False):
    rng = prng.PRNGKey(seed=args.seed)

    ## setup env
    apply_recommended_params(args)
    env = mbd_b200.envs.get_env(args.env_name)
    Nu = env.action_size

    rng, rng_reset = prng.split(rng)  # NOTE: rng_reset should never be changed.
    state_init = env.reset(rng_reset)

    ## run diffusion
    betas, alphas, alphas_bar, sigmas = make_schedule(args.beta0, args.betaT, args.Ndiffuse)
    if _is_main():
        print(f"init sigma = {sigmas[-1]:.2e}")

    engine = DiffusionEngine(env, args.Nsample, args.Hsample, args.temp_sample, args.enable_demo, state_init, Ndiffuse=args.Ndiffuse)
    HNu = args.Hsample * Nu
    # Everything the loop of mbd_planner.py:138-148 feeds into reverse_once is uploaded ONCE: the Y0s_rng chain
    # (rng, Y0s_rng = split(rng) per step, :103), sigmas[i] and the schedule scalars.  engine.Ybars row N-1 = YN = 0 and
    # row i-1 receives Ybar_{i-1}; one step is captured in a CUDA graph and replayed, nothing is copied to the host inside
    # the loop.
    rng_exp, rng = prng.split(rng)
    engine.load_schedule(key_chain(rng_exp, args.Ndiffuse), sigmas, alphas, alphas_bar)
    engine.set_step(args.Ndiffuse - 1)
    if os.environ.get("MBD_GRAPH", "1") != "0":
        engine.capture()
    steps = range(args.Ndiffuse - 1, 0, -1)
    pbar = tqdm(steps, desc="Diffusing") if (tqdm is not None and _is_main()) else None
    for n_done, i in enumerate(pbar if pbar is not None else steps):
        engine.step()
        if pbar is not None and (n_done % log_every == log_every - 1 or i == 1):
            # the reference formats rew every step (a device->host sync each step, mbd_planner.py:147);
            # here the sync is paid every `log_every` steps only
            pbar.set_postfix({"rew": f"{engine.rew_hist[i].item():.2e}"})
            engine.check_exchange()
    engine.check_exchange()
    Ybars = engine.Ybars
    Yi = Ybars[: args.Ndiffuse - 1].flip(0).reshape(args.Ndiffuse - 1, args.Hsample, Nu)  # jnp.array(Ybars) order

    if not args.not_render and _is_main():
        path = f"{mbd_b200.__path__[0]}/../results/{args.env_name}"
        if not os.path.exists(path):
            os.makedirs(path)
        np.save(f"{path}/mu_0ts.npy", Yi.cpu().numpy())
        if args.env_name == "car2d":
            _render_car2d(env, state_init, Yi[-1].cpu().numpy(), args, path)
        elif env.kind in ("xpbd", "pusht"):
            # mbd_planner.py:168-178: rollout.html = brax.io.html.render(sys with opt.timestep = env.dt, rollout).  The same
            # page (and the JSON document inside it, which vis_diffusion.py / brax.io.html.render_from_json consume) is written
            # by mbd_b200.io.brax_json; rollout_states.npz keeps the plain arrays
            from ..io import brax_json
            from ..utils import rollout_states, trajectory_arrays
            rollout = rollout_states(env.step, state_init, Yi[-1].cpu().numpy())
            with open(f"{path}/rollout.html", "w") as f:
                f.write(brax_json.render(env.sys, rollout, env.dt))
            with open(f"{path}/rollout.json", "w") as f:
                f.write(brax_json.dumps(env.sys, rollout, env.dt))
            np.savez(f"{path}/rollout_states.npz", **trajectory_arrays(env, rollout))
    rew_final = final_reward(env, engine, Yi[-1])
    if return_trajectory:
        return rew_final, Yi
    return rew_final


def final_reward(env, engine: DiffusionEngine, us: torch.Tensor) -> float:
    """rollout_us(state_init, Yi[-1])[0].mean()  (mbd_planner.py:179-180) — one n=1 launch."""
    us = us.reshape(1, engine.H, engine.Nu).contiguous()
    if env.kind == "xpbd":
        out = ops.rollout(engine.model, engine.state_init, us)
    elif env.kind == "pusht":
        out = ops.pusht_rollout(engine.params_car, engine.state_init, us)
    else:
        out = ops.car2d_rollout(engine.params_car, engine.state_init, us)
    return float(out["rews"][0].item())


def _render_car2d(env, state_init, us, args, path):
    try:
        import matplotlib
        matplotlib.use("Agg")
        from matplotlib import pyplot as plt
    except Exception:  # noqa: BLE001
        return
    fig, ax = plt.subplots(1, 1, figsize=(3, 3))
    xs = [np.asarray(state_init.pipeline_state)]
    state = state_init
    for t in range(us.shape[0]):
        state = env.step(state, us[t])
        xs.append(np.asarray(state.pipeline_state))
    env.render(ax, np.stack(xs))
    if args.enable_demo:
        ax.plot(env.xref[:, 0], env.xref[:, 1], "g--", label="RRT path")
    ax.legend()
    plt.savefig(f"{path}/rollout.png")


def _maybe_init_distributed():
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")


if __name__ == "__main__":
    import tyro

    _maybe_init_distributed()
    rew_final = run_diffusion(args=tyro.cli(Args))
    if _is_main():
        print(f"final reward = {rew_final:.2e}")
