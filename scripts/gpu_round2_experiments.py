"""Round-2 experiments (NEXT.md): group de-phasing by a start offset, shared schedulers, neighbourhood barriers.
    python scripts/gpu_round2_experiments.py            # drives one subprocess per configuration (a deadlock only loses that one)
Every configuration is checked bit for bit against kernel variant 2."""
import ctypes, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

MAPS = {"A": lambda w: (w & 1, w >> 1), "B": lambda w: ((w ^ (w >> 2)) & 1, w >> 1)}


def child(variant, mp, stagger):
    import numpy as np, torch
    import mbd_b200
    from mbd_b200 import ops, prng, _lib
    env = mbd_b200.envs.get_env("humanoidrun")
    st = torch.as_tensor(env.reset(prng.split(prng.PRNGKey(0))[1]).pipeline_state.raw, device="cuda:0")
    key = np.uint32([1, 2]); n = 8192
    Y0s = torch.empty((n, 850), device="cuda:0"); rews = torch.empty(n, device="cuda:0"); Yb = torch.zeros(850, device="cuda:0")
    m = ops.Model(env.blob)
    ops.set_kernel_variant(2)
    ops.sample_rollout(m, st, key, n, 0, n, 50, 0.88, Yb, Y0s, rews); torch.cuda.synchronize()
    ref = rews.cpu().numpy().copy()
    tab = (ctypes.c_int * 22)(*[(MAPS[mp](w)[0] << 4) | MAPS[mp](w)[1] for w in range(22)])
    _lib.check(_lib.lib().mbd_model_set_group_map(m.handle, tab, 22), "map")
    _lib.check(_lib.lib().mbd_set_group_stagger(stagger), "stagger")
    ops.set_kernel_variant(variant)
    for _ in range(2): ops.sample_rollout(m, st, key, n, 0, n, 50, 0.88, Yb, Y0s, rews)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.sample_rollout(m, st, key, n, 0, n, 50, 0.88, Yb, Y0s, rews)
    e1.record(); torch.cuda.synchronize()
    print(f"variant={variant} map={mp} stagger={stagger:5d}: {e0.elapsed_time(e1) / 5:.3f} ms  bit-identical={np.array_equal(rews.cpu().numpy(), ref)}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 4:
        child(int(sys.argv[1]), sys.argv[2], int(sys.argv[3]))
    else:
        cases = [(v, mp, st) for v in (6, 10) for mp in ("A", "B") for st in (0, 1500, 3000, 4500)] + [(v, "A", 0) for v in (8, 9, 11)]
        for variant, mp, stagger in cases:
            if True:
                if True:
                    try:
                        subprocess.run([sys.executable, os.path.abspath(__file__), str(variant), mp, str(stagger)], timeout=40, check=False)
                    except subprocess.TimeoutExpired:
                        print(f"variant={variant} map={mp} stagger={stagger}: TIMEOUT (deadlock?)", flush=True)
