"""Where the wall clock of a full humanoidrun solve goes (bench.py's e2e_solve): phase timers around run_diffusion's steps."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mbd_b200
from mbd_b200 import prng
from mbd_b200.planners import engine as eng
from mbd_b200.planners.mbd_planner import Args, run_diffusion, final_reward

def once(tag):
    T = {}
    def tick(name, t0):
        torch.cuda.synchronize(); T[name] = (time.perf_counter() - t0) * 1e3; return time.perf_counter()
    t = time.perf_counter()
    env = mbd_b200.envs.get_env("humanoidrun"); t = tick("get_env", t)
    rng, rr = prng.split(prng.PRNGKey(0)); st = env.reset(rr); t = tick("reset", t)
    _, alphas, alphas_bar, sigmas = eng.make_schedule(1e-4, 1e-2, 300)
    e = eng.DiffusionEngine(env, 8192, 50, 0.1, False, st, Ndiffuse=300); t = tick("engine_init", t)
    rng_exp, rng = prng.split(rng)
    e.load_schedule(eng.key_chain(rng_exp, 300), sigmas, alphas, alphas_bar); e.set_step(299); t = tick("load_schedule", t)
    e.capture(); t = tick("capture", t)
    for _ in range(299): e.step()
    t = tick("299 steps", t)
    e.check_exchange(); t = tick("check_exchange", t)
    Yi = e.Ybars[:299].flip(0).reshape(299, 50, 17); t = tick("flip", t)
    rf = final_reward(env, e, Yi[-1]); t = tick("final_reward", t)
    print(tag, {k: round(v, 2) for k, v in T.items()}, "total", round(sum(T.values()), 1), "ms; rew_final", rf)

once("first"); once("second"); once("third")
import gc; gc.disable(); once("gc-off"); gc.enable()
t0 = time.perf_counter(); run_diffusion(Args(env_name="humanoidrun", not_render=True), log_every=10**9); torch.cuda.synchronize()
print("run_diffusion wall", round((time.perf_counter() - t0) * 1e3, 1), "ms")
