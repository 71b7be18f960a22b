"""A few launches of the fused sampling + rollout kernel for ncu:  python scripts/gpu_ncu_target.py <env> <nsample> [variant] [H]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mbd_b200
from mbd_b200 import ops, prng
name, n = sys.argv[1], int(sys.argv[2])
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
H = int(sys.argv[4]) if len(sys.argv) > 4 else 50
env = mbd_b200.envs.get_env(name)
st = torch.as_tensor(env.reset(prng.split(prng.PRNGKey(0))[1]).pipeline_state.raw, device="cuda:0")
m = ops.Model(env.blob)
HNu = H * env.action_size
Y0s = torch.empty((n, HNu), device="cuda:0"); rews = torch.empty(n, device="cuda:0"); Yb = torch.zeros(HNu, device="cuda:0")
ops.set_kernel_variant(variant)
for _ in range(3):
    ops.sample_rollout(m, st, np.uint32([1, 2]), n, 0, n, H, 0.88, Yb, Y0s, rews)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ops.sample_rollout(m, st, np.uint32([1, 2]), n, 0, n, H, 0.88, Yb, Y0s, rews)
e1.record(); torch.cuda.synchronize()
print(f"{name} n={n} variant={variant}: {e0.elapsed_time(e1) / 3:.4f} ms per launch")
