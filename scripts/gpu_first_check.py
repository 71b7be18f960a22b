"""First GPU bring-up check: CUDA path vs CPU oracle (bit-exact expectations)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mbd_b200 import ops
from mbd_b200.model import system_io, blob, kinematics
from oracle import oracle as orc

A = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mbd_b200", "assets")
dev = torch.device("cuda:0")
s = system_io.load(os.path.join(A, "humanoidrun.json"))
b = blob.pack(s, 7, blob.REWARD_HUMANOIDRUN)
st = kinematics.pipeline_init(s, s.init_q, np.zeros(s.qd_size()))
m = ops.Model(b)
rng = np.random.default_rng(0)
for (n, H, nsub) in [(8, 1, 1), (8, 1, 7), (37, 5, 0), (256, 50, 0)]:
    Y = np.clip(rng.normal(size=(n, H, 17)) * 0.88, -1, 1).astype(np.float32)
    ref = orc.xpbd_rollout(b, st, Y, want_rewss=True, want_final=True, nsub_override=nsub)
    out = ops.rollout(m, torch.tensor(st, device=dev), torch.tensor(Y, device=dev), want_rewss=True, want_final=True, nsub_override=nsub)
    torch.cuda.synchronize()
    f = out["final"].cpu().numpy(); r = out["rews"].cpu().numpy(); rs = out["rewss"].cpu().numpy()
    print(f"n={n} H={H} nsub={nsub}: final bit-exact={np.array_equal(f.view(np.uint32), ref['final'].view(np.uint32))} "
          f"maxabs={np.abs(f-ref['final']).max():.3e} rews bit-exact={np.array_equal(r.view(np.uint32), ref['rews'].view(np.uint32))} "
          f"rewss maxabs={np.abs(rs-ref['rewss']).max():.3e}")
# sampling
key = orc.split(orc.prng_key(0))[0]
Ybar = rng.normal(size=(50 * 17)).astype(np.float32) * 0.1
got = ops.sample(key, 512, 128, 64, 850, 0.7, torch.tensor(Ybar, device=dev)).cpu().numpy()
ref = orc.sample_Y0s(key, 512, 850, 0.7, Ybar, 128, 192)
print("sample bit-exact", np.array_equal(got.view(np.uint32), ref.view(np.uint32)), np.abs(got - ref).max())
# timing of the fused kernel at the headline size
n, H = 8192, 50
Y0s = torch.empty((n, H * 17), device=dev); rews = torch.empty(n, device=dev)
Yb = torch.zeros(H * 17, device=dev); sti = torch.tensor(st, device=dev)
for _ in range(2):
    ops.sample_rollout(m, sti, key, n, 0, n, H, 0.88, Yb, Y0s, rews)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.sample_rollout(m, sti, key, n, 0, n, H, 0.88, Yb, Y0s, rews)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"fused sample+rollout n={n} H={H}: {ms:.3f} ms -> {n*H/ms*1e3/1e6:.2f} M env-steps/s; rews mean {rews.mean().item():.4f}")
refY = orc.sample_Y0s(key, n, 850, 0.88, np.zeros(850, np.float32), 0, 64)
print("fused Y0s bit-exact", np.array_equal(Y0s[:64].cpu().numpy().view(np.uint32), refY.view(np.uint32)))
ref = orc.xpbd_rollout(b, st, refY.reshape(64, 50, 17))
print("fused rews bit-exact", np.array_equal(rews[:64].cpu().numpy().view(np.uint32), ref["rews"].view(np.uint32)))
