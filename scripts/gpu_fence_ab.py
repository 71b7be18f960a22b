"""A/B on ONE box: the rollout kernel with and without round 1's __threadfence_block() before every bar.arrive of the named edge
barriers (-DMBD_NAMED_FENCE).  Usage: python scripts/gpu_fence_ab.py build   (here, needs nvcc: writes mbd_b200/_C/libmbd_b200_fence.so)
       python scripts/gpu_fence_ab.py run fence|nofence   (on the GPU box; prints ms per launch, alternate the two a few times)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mbd_b200 import build as b
mode = sys.argv[1]
alt = os.path.join(os.path.dirname(b.OUT), "libmbd_b200_fence.so")
if mode == "build":
    b.OUT = alt
    b.NVCC_FLAGS = b.NVCC_FLAGS + ["-DMBD_NAMED_FENCE"]
    b.build(force=True)
    print("built", alt)
    sys.exit(0)
if sys.argv[2] == "fence":
    b.OUT = alt
    b.is_stale = lambda: False
import numpy as np, torch
import mbd_b200
from mbd_b200 import ops, prng
env = mbd_b200.envs.get_env("humanoidrun")
st = torch.as_tensor(env.reset(prng.split(prng.PRNGKey(0))[1]).pipeline_state.raw, device="cuda:0")
m = ops.Model(env.blob)
n, H = 8192, 50
Y0s = torch.empty((n, 850), device="cuda:0"); rews = torch.empty(n, device="cuda:0"); Yb = torch.zeros(850, device="cuda:0")
for v in (9, 3):
    nn = n if v == 9 else 4096
    ops.set_kernel_variant(v)
    for _ in range(5):
        ops.sample_rollout(m, st, np.uint32([1, 2]), nn, 0, nn, H, 0.88, Yb, Y0s[:nn], rews[:nn])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        ops.sample_rollout(m, st, np.uint32([1, 2]), nn, 0, nn, H, 0.88, Yb, Y0s[:nn], rews[:nn])
    e1.record(); torch.cuda.synchronize()
    print(f"{sys.argv[2]:8s} variant {v} n={nn}: {e0.elapsed_time(e1) / 30:.4f} ms per launch", flush=True)
