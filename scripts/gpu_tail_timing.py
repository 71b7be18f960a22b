"""Pure kernel time of the two tail kernels at the 8-rank weak-scaling size (65,536 samples) WITHOUT NVLink and without
inter-GPU skew: 8 emulated ranks on one GPU (peer loads hit local memory).  The rank that arrives last at each rendezvous
does not wait, so min over ranks of the statistics / update time is the kernel's own cost."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mbd_b200
from mbd_b200 import ops, prng
from mbd_b200.planners import engine as eng

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_per = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
env = mbd_b200.envs.get_env("humanoidrun")
st = env.reset(prng.split(prng.PRNGKey(0))[1])
Nd = 40
_, alphas, alphas_bar, sigmas = eng.make_schedule(1e-4, 1e-2, Nd)
keys = eng.key_chain(np.uint32([1, 2]), Nd)
ranks = eng.DiffusionEngine.make_emulated_ranks(env, P * n_per, 50, 0.1, False, st, P, Ndiffuse=Nd) if P > 1 else \
    [eng.DiffusionEngine(env, n_per, 50, 0.1, False, st, Ndiffuse=Nd)]
for e in ranks:
    e.load_schedule(keys, sigmas, alphas, alphas_bar); e.set_step(Nd - 1)
    if P == 1: e.stream = torch.cuda.current_stream()
evs = [[ops.Event() for _ in range(4)] for _ in ranks]
res = []
for it in range(12):
    cur = torch.cuda.current_stream()
    for e in ranks: e.stream.wait_stream(cur)
    for e, ev in zip(ranks, evs):
        with torch.cuda.stream(e.stream):
            ops.step_launch_timed(e._plan_c, ev[0], ev[1], ev[2], ev[3])
    for e in ranks: cur.wait_stream(e.stream)
    torch.cuda.synchronize()
    if it >= 2:
        res.append([(ev[1].elapsed_ms(ev[2]), ev[2].elapsed_ms(ev[3])) for ev in evs])
res = np.array(res)      # [it, rank, 2]
print(f"P={P} N={P * n_per}: statistics kernel min over ranks {res[:, :, 0].min(1).mean() * 1e3:.1f} us, update kernel min over ranks "
      f"{res[:, :, 1].min(1).mean() * 1e3:.1f} us (mean over {len(res)} steps)")
