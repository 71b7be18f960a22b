"""Real multi-GPU check (needs >= 2 GPUs; skipped otherwise): a 2-rank NCCL run of the sharded
engine gives bit-identical Ybar_im1 to the single-GPU run."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MBD_ROOT"])
import mbd_b200
from mbd_b200 import prng
from mbd_b200.planners import engine as eng
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
demo = os.environ.get("MBD_TEST_DEMO", "0") == "1"   # BASELINE config 5: humanoidtrack + enable_demo, sample-sharded
env = mbd_b200.envs.get_env("humanoidtrack" if demo else "humanoidrun")
rng, rr = prng.split(prng.PRNGKey(0))
st = env.reset(rr)
_, alphas, alphas_bar, sigmas = eng.make_schedule(1e-4, 1e-2, 100)
e = eng.DiffusionEngine(env, 2048, 50, 0.1, demo, st)
Yb = torch.zeros(850, device="cuda")
key = np.uint32([3, 1])
outs = []
for i in (99, 98, 97):
    key2 = prng.split(key)[1]; key = prng.split(key)[0]
    out, rew = e.reverse_once(key2, float(sigmas[i]), Yb, eng.update_coef(alphas, alphas_bar, i))
    Yb = out.clone(); outs.append(out.cpu().numpy().copy()); outs.append(np.float32([rew.item()]))
assert e.P == 1 or e.exchange == os.environ.get("MBD_EXCHANGE", "p2p"), e.exchange
if e.sym is not None:
    assert int(e.xerr.item()) == 0
if e.rank == 0:
    np.save(os.environ["MBD_OUT"], np.concatenate(outs))
if e.P > 1:
    dist.destroy_process_group()
'''


def _port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("demo", ["0", "1"], ids=["humanoidrun", "humanoidtrack-demo"])
def test_two_rank_nccl_equals_single_gpu(tmp_path, demo):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, MBD_ROOT=ROOT, MBD_TEST_DEMO=demo)
    env1 = dict(env, MBD_OUT=str(tmp_path / "p1.npy"))
    subprocess.run([sys.executable, str(w)], check=True, env=env1, timeout=600)
    a = np.load(tmp_path / "p1.npy")
    for mode in ("p2p", "nccl"):   # fused NVLink peer-memory gather, and the NCCL all_gather fallback
        env2 = dict(env, MBD_OUT=str(tmp_path / f"p2_{mode}.npy"), MBD_EXCHANGE=mode)
        subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_port()), str(w)], check=True, env=env2, timeout=600)
        b = np.load(tmp_path / f"p2_{mode}.npy")
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), mode
