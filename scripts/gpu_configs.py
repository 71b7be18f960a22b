"""env-steps/s of one diffusion step (mean over 20 consecutive steps of the chain, captured-graph replay) for every BASELINE.json config and
every other positional env (1 GPU), with the
per-sample returns of the first 32 samples checked bit for bit against the CPU oracle."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mbd_b200
from mbd_b200 import prng
from mbd_b200.planners import engine as eng
from oracle import oracle as orc

CONFIGS = [("car2d", 64, 40, False, 100), ("car2d", 2048, 50, True, 100), ("hopper", 1024, 50, False, 100), ("ant", 4096, 50, False, 100),
           ("humanoidrun", 8192, 50, False, 300), ("humanoidtrack", 16384, 50, True, 100), ("humanoidtrack", 16384, 60, True, 100),
           ("humanoidstandup", 8192, 50, False, 100), ("walker2d", 2048, 50, False, 100), ("halfcheetah", 2048, 50, False, 100),
           ("cartpole", 2048, 50, False, 100), ("pushT", 2048, 40, False, 200)]
rows = []
for name, N, H, demo, Nd in CONFIGS:
    env = mbd_b200.envs.get_env(name)
    rng, rr = prng.split(prng.PRNGKey(0))
    st = env.reset(rr)
    _, alphas, alphas_bar, sigmas = eng.make_schedule(1e-4, 1e-2, Nd)
    e = eng.DiffusionEngine(env, N, H, 0.2 if name == "pushT" else 0.1, demo, st, Ndiffuse=Nd)
    Nu = env.action_size
    e.load_schedule(eng.key_chain(np.uint32([5, 7]), Nd), sigmas, alphas, alphas_bar)
    e.set_step(Nd - 1)
    e.capture()                          # the product path: one captured step (three launches), replayed
    K = 20
    for _ in range(3):
        e.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        e.step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    Y = e.Y0s[:32].cpu().numpy().reshape(32, H, Nu)
    if env.kind == "xpbd":
        ref = orc.xpbd_rollout(env.blob, st.pipeline_state.raw, Y, xref=env.xref if demo else None)
    elif env.kind == "pusht":
        ref = orc.pusht_rollout(env.params, st.pipeline_state.raw, Y)
    else:
        ref = orc.car2d_rollout(env.params, env.x0, Y, xref=env.xref if demo else None)
    ok = np.array_equal(e.rews_local[:32].cpu().numpy().view(np.uint32), ref["rews"].view(np.uint32))
    if demo:
        ok = ok and np.array_equal(e.logpd_local[:32].cpu().numpy().view(np.uint32), ref["logpd"].view(np.uint32))
    rows.append(dict(n_frames=(env._n_frames if env.kind in ('xpbd', 'pusht') else 1), links=(int(env.blob.view(np.int32)[1]) if env.kind == 'xpbd' else (3 if env.kind == 'pusht' else 1)), env=name, Nsample=N, Hsample=H, demo=demo, ms_per_step=ms, env_steps_per_s=N * H / ms * 1e3, oracle_bit_exact=bool(ok)))
    print(rows[-1])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/configs_r02.json", "w"), indent=1)
