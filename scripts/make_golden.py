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

def pusht():
    """pushT: a scripted push (contact with both boxes of the T, friction, a joint limit) and 24 sampled rollouts"""
    pt = mbd_b200.envs.get_env("pushT")
    x0 = pt.reset(prng.split(prng.PRNGKey(0))[1]).pipeline_state.raw
    Y = np.clip(np.random.default_rng(12).normal(size=(24, 40, 2)).astype(np.float32) * 0.8 + np.float32([-0.2, 0.5]), -1.5, 1.5)
    o = orc.pusht_rollout(pt.params, x0, Y, want_rewss=True, want_final=True)
    x1 = x0.copy(); x1[0:2] = [-0.21, 0.0]
    script = np.float32([[1.0, 0.0]] * 14 + [[0.2, 1.0]] * 10 + [[-1.0, -0.3]] * 6)
    s_ = orc.pusht_rollout(pt.params, x1, script[None], want_traj=True, want_rewss=True)
    np.savez_compressed(os.path.join(G, "pusht_oracle.npz"), params=pt.params, x0=x0, Y0s=Y, rews=o["rews"], rewss=o["rewss"], final=o["final"],
                        x1=x1, script=script, script_traj=s_["traj"][0], script_rewss=s_["rewss"][0])
    print("pushT rews", o["rews"][:4], "scripted push: slider ends at", s_["traj"][0][-1][2:5])


if len(sys.argv) > 1 and sys.argv[1] == "pusht":
    pusht()
    sys.exit(0)

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
pusht()
