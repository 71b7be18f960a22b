"""Traces phase boundaries (clock64 per SM) of co-resident CTAs of the v2 kernel to see whether the two CTAs of an
SM run their heavy phases (A, C) at the same time.  Instrumented build (-DMBD_PROFILE_PHASES)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mbd_b200 import build as b
b.OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmbd_prof.so")
b.NVCC_FLAGS = b.NVCC_FLAGS + ["-DMBD_PROFILE_PHASES"]
b.build(force=True)
b.is_stale = lambda: False
import mbd_b200
from mbd_b200 import ops, prng, _lib
env = mbd_b200.envs.get_env("humanoidrun")
rng, rr = prng.split(prng.PRNGKey(0))
st = torch.as_tensor(env.reset(rr).pipeline_state.raw, device="cuda:0")
m = env.device_model(); L = _lib.lib(); key = np.uint32([1, 2]); n = 8192
Y0s = torch.empty((n, 850), device="cuda:0"); rews = torch.empty(n, device="cuda:0"); Yb = torch.zeros(850, device="cuda:0")
ops.set_kernel_variant(2)
L.mbd_trace_arm(ctypes.c_uint(70000))      # disarmed warm-up
ops.sample_rollout(m, st, key, n, 0, n, 1, 0.88, Yb[:17], Y0s[:, :17].contiguous(), rews); torch.cuda.synchronize()
L.mbd_trace_arm(ctypes.c_uint(0))
# short rollout: H=3 env steps = 21 physics steps; each traced warp logs 8 events per step
Y3 = torch.empty((n, 51), device="cuda:0")
ops.sample_rollout(m, st, key, n, 0, n, 3, 0.88, Yb[:51], Y3, rews); torch.cuda.synchronize()
buf = np.zeros((65536, 2), np.uint64); cnt = ctypes.c_uint(0)
L.mbd_trace_read(buf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cnt))
k = min(int(cnt.value), 60000)
tag, t = buf[:k, 0], buf[:k, 1].astype(np.int64)
blk = (tag >> np.uint64(32)).astype(np.int64); smid = ((tag >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
link = ((tag >> np.uint64(8)) & np.uint64(0xFF)).astype(np.int64); ph = (tag & np.uint64(0xFF)).astype(np.int64)
print("events", k)
# pick an SM that hosts two CTAs
sms = {}
for b_, s_ in zip(blk, smid): sms.setdefault(int(s_), set()).add(int(b_))
two = [s_ for s_, bs in sms.items() if len(bs) == 2][:2]
for s_ in two:
    bs = sorted(sms[s_])
    print(f"SM {s_}: blocks {bs}")
    t0 = t[(smid == s_)].min()
    for b_ in bs:
        sel = (smid == s_) & (blk == b_) & (link == 4)
        order = np.argsort(t[sel])
        ev = list(zip((t[sel][order] - t0).tolist(), ph[sel][order].tolist()))
        # print phase start of A (event 7 = end of previous step) for the first 12 physics steps: time of phase-6 (D end) events
        d_end = [e[0] for e in ev if e[1] == 6][:14]
        c_end = [e[0] for e in ev if e[1] == 4][:14]
        a_end = [e[0] for e in ev if e[1] == 0][:14]
        print(f"  block {b_} shin: A-end {a_end}\n               C-end {c_end}\n               D-end {d_end}")
