"""Generates tests/golden/*.npz from the CPU oracle (the reference itself cannot run here:
no jax/brax).  The fixtures pin (a) the oracle against accidental change, (b) the CUDA path on
the GPU box where neither /root/reference nor a recompiled oracle may be assumed identical.
    python scripts/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import mbd_b200  # noqa: E402
from mbd_b200 import prng  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle import planner as opl  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
os.makedirs(G, exist_ok=True)

# ---- humanoidrun: 32 samples x 50 steps from the planner's own seed-0 chain --------------------------
env = mbd_b200.envs.get_env("humanoidrun")
rng, rng_reset = prng.split(prng.PRNGKey(0))
st = env.reset(rng_reset).pipeline_state.raw
rng_exp, rng = prng.split(rng)
_, k = prng.split(rng_exp)
sig = float(opl.make_schedule(1e-4, 1e-2, 300)[3][299])
Y0s = orc.sample_Y0s(k, 8192, 850, sig, np.zeros(850, np.float32), 0, 32).reshape(32, 50, 17)
out = orc.xpbd_rollout(env.blob, st, Y0s, want_rewss=True, want_final=True)
np.savez_compressed(os.path.join(G, "humanoidrun_oracle.npz"), state_init=st, key=k, sigma=np.float32(sig), Y0s=Y0s,
                    rews=out["rews"], rewss=out["rewss"], final=out["final"])
print("humanoidrun rews", out["rews"][:4])

# ---- humanoidtrack with demo -----------------------------------------------------------------------------
envt = mbd_b200.envs.get_env("humanoidtrack")
stt = envt.reset(None).pipeline_state.raw
Yt = orc.sample_Y0s(k, 1024, 850, 0.63, np.zeros(850, np.float32), 0, 16).reshape(16, 50, 17)
o = orc.xpbd_rollout(envt.blob, stt, Yt, xref=envt.xref, want_track=True)
np.savez_compressed(os.path.join(G, "humanoidtrack_oracle.npz"), state_init=stt, Y0s=Yt, rews=o["rews"], logpd=o["logpd"],
                    track=o["track"][:, ::10])
print("humanoidtrack rews", o["rews"][:4], "logpd", o["logpd"][:4])

# ---- car2d: BASELINE config 1 (Nsample=64, H=40) full solve + a demo solve ---------------------------------
car = mbd_b200.envs.get_env("car2d")
oenv = opl.OracleEnv("car2d", 2, params=car.params, x0=car.x0)
rf, Yi, rews = opl.run_diffusion(oenv, 0, 64, 40, 100, 0.1)
rfd, Yid, rewsd = opl.run_diffusion(oenv, 0, 512, 50, 100, 0.1, xref=car.xref, rew_xref=car.rew_xref)
# short demo chain (8 steps): long chains amplify rounding differences of the statistics through the
# collision freeze / reward clip discontinuities, so step-level parity is pinned on a short one
rfs, Yis, rewss_ = opl.run_diffusion(oenv, 0, 512, 50, 9, 0.1, xref=car.xref, rew_xref=car.rew_xref)
np.savez_compressed(os.path.join(G, "car2d_oracle.npz"), rew_final=np.float32(rf), Yi_last=Yi[-1], rews=rews,
                    rew_final_demo=np.float32(rfd), Yi_last_demo=Yid[-1], rews_demo=rewsd,
                    Yi_short_demo=Yis, rews_short_demo=rewss_)
print("car2d", rf, rfd)
