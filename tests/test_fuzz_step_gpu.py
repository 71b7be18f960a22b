"""Shape sweep of the diffusion step (`mbd_step_launch`: rollouts | statistics + softmax in one cluster | weighted mean + update)
on car2d (cheap rollouts, BASELINE config 1) against the numpy planner oracle, single rank and emulated ranks: sample counts
from one per rank to several thousand that are NOT multiples of the 64-sample runs / 1024-thread blocks the tail kernels tile
by, horizons from 1 to 50, with and without the demo blend (mbd_planner.py:110-133).  Where every rank holds whole 64-sample
runs the sharded result must equal the single-rank one bit for bit; otherwise it must stay within the 1e-4 of north_star."""
import numpy as np
import pytest
import torch

import mbd_b200
from mbd_b200.planners import engine as eng
from oracle import planner as opl
from tests.conftest import assert_bit_exact

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ND = 40


def _cases():
    rng = np.random.default_rng(7)
    out = []
    per_rank = [1, 2, 3, 31, 33, 63, 64, 65, 100, 128, 192, 257, 640, 1000, 1024, 1031]
    for k, nl in enumerate(per_rank):
        P = [1, 2, 4, 8, 3][k % 5] if nl != 1 else 8
        H = int([1, 2, 7, 50, 23, 13][k % 6])
        out.append((nl * P, H, P, bool(k % 2), float([0.1, 0.5, 0.05][k % 3]), int(rng.integers(2, ND))))
    out.append((8 * 4096, 50, 8, True, 0.1, 17))   # 32768 returns through the 8-CTA statistics cluster, whole runs on every rank
    out.append((5 * 64, 50, 5, False, 0.1, 9))     # a rank count that is not a power of two, whole runs: still bit-identical
    return out


def N(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("Nn,H,P,demo,temp,i", _cases())
def test_step_shapes_vs_oracle_and_single_rank(orc, Nn, H, P, demo, temp, i):
    car = mbd_b200.envs.get_env("car2d")
    _, alphas, alphas_bar, sigmas = opl.make_schedule(1e-4, 1e-2, ND)
    keys = eng.key_chain(np.uint32([Nn, H]), ND)
    Ybar_i = (np.random.default_rng(Nn + H).normal(size=2 * H) * 0.3).astype(np.float32)
    oenv = opl.OracleEnv("car2d", 2, params=car.params, x0=car.x0)
    ref = opl.reverse_once(oenv, keys[i], Nn, H, float(sigmas[i]), Ybar_i, temp, alphas, alphas_bar, i,
                           xref=car.xref if demo else None, rew_xref=car.rew_xref)

    def run(engines):
        for e in engines:
            e.load_schedule(keys, sigmas, alphas, alphas_bar)
            e.set_step(i)
            e.Ybars[i].copy_(torch.as_tensor(Ybar_i, device=DEV))
        if len(engines) == 1:
            engines[0].step()
        else:
            eng.DiffusionEngine.step_emulated_ranks(engines)
        torch.cuda.synchronize()
        for e in engines:
            e.check_exchange()
            assert int(e.ctl[0].item()) == i - 1

    e1 = [eng.DiffusionEngine(car, Nn, H, temp, demo, car.reset(None), Ndiffuse=ND)]
    run(e1)
    groups = [e1] if P == 1 else [e1, eng.DiffusionEngine.make_emulated_ranks(car, Nn, H, temp, demo, car.reset(None), P, Ndiffuse=ND)]
    if P > 1:
        run(groups[1])
    for engines in groups:
        what = f"{len(engines)} rank(s)"
        assert_bit_exact(np.concatenate([N(e.Y0s) for e in engines]), ref["Y0s"], what + ": sampled actions")
        assert_bit_exact(np.concatenate([N(e.rews_local) for e in engines]), ref["rews"], what + ": per-sample returns")
        w = np.concatenate([N(e.weights) for e in engines])
        assert abs(float(w.astype(np.float64).sum()) - 1.0) < 1e-4
        assert np.abs(w - ref["weights"]).max() / float(ref["weights"].max()) < 2e-4, what + ": softmax weights"
        assert int(w.argmax()) == int(ref["weights"].argmax()) or Nn < 4
        scale = max(float(np.abs(ref["Ybar_im1"]).max()), 1e-6)
        for e in engines:
            assert np.abs(N(e.Ybars[i - 1]) - ref["Ybar_im1"]).max() / scale < 1e-4, what + f": Ybar_im1 on rank {e.rank}"
            assert abs(float(e.rew_hist[i].item()) - float(ref["rew_mean"])) <= 1e-4 * max(abs(float(ref["rew_mean"])), 1e-6) + 1e-6
    if P > 1:
        # every rank holds the same iterate, bit for bit (each one reduces all partials in the same fixed order) ...
        for e in groups[1][1:]:
            assert_bit_exact(N(e.Ybars[i - 1]), N(groups[1][0].Ybars[i - 1]), "ranks disagree on the iterate")
            assert_bit_exact(N(e.rews_all), N(groups[1][0].rews_all), "ranks disagree on the gathered returns")
        # ... and it is the single-rank iterate whenever every rank holds 64 * 2^k samples: the single-rank pairwise tree over the
        # 64-sample runs then contains the per-rank trees as subtrees (mbd_b200/planners/sharding.py)
        nl = Nn // P
        if nl % 64 == 0 and ((nl // 64) & (nl // 64 - 1)) == 0:
            assert_bit_exact(N(groups[1][0].Ybars[i - 1]), N(e1[0].Ybars[i - 1]), f"{P} ranks vs one rank")
            assert_bit_exact(np.concatenate([N(e.weights) for e in groups[1]]), N(e1[0].weights), f"{P} ranks vs one rank: weights")
