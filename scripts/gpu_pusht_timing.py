"""Where does the pushT kernel spend its time?  ms per launch of mbd_pusht_rollout (2048 x 40, sampled actions around zero) for
several solver sweep counts and for a contact-free start (pusher parked far from the T and from the walls)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mbd_b200
from mbd_b200 import ops, prng
from mbd_b200.envs import pusht

env = mbd_b200.envs.get_env("pushT")
x0 = env.reset(prng.split(prng.PRNGKey(0))[1]).pipeline_state.raw
rng = np.random.default_rng(0)
n, H = 2048, 40
Y = torch.as_tensor((rng.normal(size=(n, H, 2)) * 0.9).astype(np.float32), device="cuda:0")
Yz = torch.zeros_like(Y)


def t(params, x, Yt, label):
    P = torch.as_tensor(params, device="cuda:0"); X = torch.as_tensor(x, device="cuda:0")
    for _ in range(2):
        ops.pusht_rollout(P, X, Yt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        o = ops.pusht_rollout(P, X, Yt, want_final=True)
    e1.record(); torch.cuda.synchronize()
    f = o["final"].cpu().numpy()
    print(f"{label:58s} {e0.elapsed_time(e1) / 5:8.3f} ms   |q| max {np.abs(f[:, :5]).max():.2f}  at a wall: {(np.abs(f[:, :4]).max(1) > 0.999).mean():.2f}", flush=True)


for it in (100, 10, 1):
    P = env.params.copy(); P[pusht.PT["ITERS"]] = it
    t(P, x0, Y, f"random actions, {it} sweeps")
t(env.params, x0, Yz, "zero actions (nothing ever touches anything)")
P = env.params.copy(); P[pusht.PT["LIM0"]:pusht.PT["LIM0"] + 8] = [-1e9, 1e9] * 4
t(P, x0, Y, "random actions, no joint limits (contacts only)")
P = env.params.copy(); P[pusht.PT["RP"]] = 1e-6
t(P, x0, Y, "random actions, point pusher (limits only, ~no contacts)")
for it in (100, 1):
    P = env.params.copy(); P[pusht.PT["RP"]] = 1e-6; P[pusht.PT["ITERS"]] = it
    t(P, x0, Y, f"point pusher (limits only), {it} sweeps")
    P = env.params.copy(); P[pusht.PT["LIM0"]:pusht.PT["LIM0"] + 8] = [-1e9, 1e9] * 4; P[pusht.PT["ITERS"]] = it
    t(P, x0, Y, f"no joint limits (contacts only), {it} sweeps")
P = env.params.copy(); P[pusht.PT["TOL"]] = 1e-3
t(P, x0, Y, "random actions, tolerance 1e-3")
