"""Two-group CTA: which warps (= SM sub-partition schedulers, warp id % 4) each sample group's links occupy."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mbd_b200
from mbd_b200 import ops, prng, _lib
env = mbd_b200.envs.get_env("humanoidrun")
rng, rr = prng.split(prng.PRNGKey(0))
st = torch.as_tensor(env.reset(rr).pipeline_state.raw, device="cuda:0")
key = np.uint32([1, 2]); n = 8192
Y0s = torch.empty((n, 850), device="cuda:0"); rews = torch.empty(n, device="cuda:0"); Yb = torch.zeros(850, device="cuda:0")
MAPS = {
    "A interleave (each scheduler serves one group)": lambda w: (w & 1, w >> 1),
    "B quad-swap (both groups on every scheduler)": lambda w: ((w ^ (w >> 2)) & 1, w >> 1),
    "C blocks (group 1 above group 0)": lambda w: (w // 11, w % 11),
    "D interleave, group 1 reversed slots": lambda w: (w & 1, (w >> 1) if (w & 1) == 0 else 10 - (w >> 1)),
}
ORDERS = {
    "default": None,
    "balA": [0, 1, 9, 2, 7, 8, 10, 3, 4, 5, 6],
    "balB": [8, 9, 0, 10, 7, 2, 1, 3, 4, 5, 6],
    "balC": [10, 9, 2, 1, 8, 0, 7, 3, 5, 4, 6],
    "shins-top": [0, 8, 10, 7, 9, 1, 2, 3, 5, 4, 6],
}
ref = None
ops.set_kernel_variant(6)
for mname, mp in MAPS.items():
    for oname, order in ORDERS.items():
        m = ops.Model(env.blob)
        tab = (ctypes.c_int * 22)(*[(mp(w)[0] << 4) | mp(w)[1] for w in range(22)])
        _lib.check(_lib.lib().mbd_model_set_group_map(m.handle, tab, 22), "set map")
        if order is not None:
            arr = (ctypes.c_int * 11)(*order)
            _lib.check(_lib.lib().mbd_model_set_warp_order(m.handle, arr, 11), "set order")
        for _ in range(2): ops.sample_rollout(m, st, key, n, 0, n, 50, 0.88, Yb, Y0s, rews)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.sample_rollout(m, st, key, n, 0, n, 50, 0.88, Yb, Y0s, rews)
        e1.record(); torch.cuda.synchronize()
        r = rews.cpu().numpy(); ref = r if ref is None else ref
        print(f"{mname:48s} {oname:10s}: {e0.elapsed_time(e1)/5:.3f} ms same={np.array_equal(r, ref)}", flush=True)
