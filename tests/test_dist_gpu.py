"""Real multi-GPU check (needs >= 2 GPUs; skipped otherwise): a 2-rank NCCL run of the sharded
engine (exchange inside the tail kernels over NVLink peer memory) gives bit-identical iterates to the single-GPU run."""
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
Nd = 100
_, alphas, alphas_bar, sigmas = eng.make_schedule(1e-4, 1e-2, Nd)
e = eng.DiffusionEngine(env, 2048, 50, 0.1, demo, st, Ndiffuse=Nd)
e.load_schedule(eng.key_chain(np.uint32([3, 1]), Nd), sigmas, alphas, alphas_bar)
e.set_step(Nd - 1)
if os.environ.get("MBD_TEST_GRAPH", "1") == "1":
    e.capture()                      # one captured step, replayed: the cross-GPU rendezvous epochs live on the device
for _ in range(4):
    e.step()
torch.cuda.synchronize()
e.check_exchange()
assert e.P == 1 or e.exchange == "p2p", e.exchange
assert int(e.ctl[0].item()) == Nd - 5
if e.rank == 0:
    np.save(os.environ["MBD_OUT"], np.concatenate([e.Ybars[Nd - 5:Nd - 1].cpu().numpy().reshape(-1), e.rew_hist.cpu().numpy()]))
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
    for mode in ("graph", "direct"):   # one captured step replayed / three direct launches per step
        env2 = dict(env, MBD_OUT=str(tmp_path / f"p2_{mode}.npy"), MBD_TEST_GRAPH="1" if mode == "graph" else "0")
        subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_port()), str(w)], check=True, env=env2, timeout=600)
        b = np.load(tmp_path / f"p2_{mode}.npy")
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), mode
