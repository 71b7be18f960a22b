"""Shard-size sweep of the humanoidrun rollout kernel: every kernel variant at the per-GPU shard sizes of a strong-scaling run
(8192 samples over 8 / 4 / 2 / 1 GPUs) — the data behind the auto-selector in launch_rollout (csrc/mbd_b200.cu).
Each variant is checked bit for bit against variant 2."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mbd_b200
from mbd_b200 import ops, prng

env = mbd_b200.envs.get_env(sys.argv[1] if len(sys.argv) > 1 else "humanoidrun")
st = torch.as_tensor(env.reset(prng.split(prng.PRNGKey(0))[1]).pipeline_state.raw, device="cuda:0")
m = ops.Model(env.blob)
key = np.uint32([1, 2]); H = 50; HNu = H * env.action_size
rows = []
for n in (256, 512, 1024, 2048, 4096, 8192):
    Y0s = torch.empty((n, HNu), device="cuda:0"); rews = torch.empty(n, device="cuda:0"); Yb = torch.zeros(HNu, device="cuda:0")
    ops.set_kernel_variant(2)
    ops.sample_rollout(m, st, key, n, 0, n, H, 0.88, Yb, Y0s, rews); torch.cuda.synchronize()
    ref = rews.cpu().numpy().copy()
    for v in (0, 1, 2, 3, 5, 6, 8, 9):
        try:
            ops.set_kernel_variant(v)
        except Exception:
            continue
        for _ in range(2): ops.sample_rollout(m, st, key, n, 0, n, H, 0.88, Yb, Y0s, rews)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.sample_rollout(m, st, key, n, 0, n, H, 0.88, Yb, Y0s, rews)
        e1.record(); torch.cuda.synchronize()
        rows.append(dict(n=n, variant=v, ms=e0.elapsed_time(e1) / 5, bit_identical=bool(np.array_equal(rews.cpu().numpy(), ref))))
        print(rows[-1], flush=True)
ops.set_kernel_variant(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/shard_sweep_r02.json", "w"), indent=1)
