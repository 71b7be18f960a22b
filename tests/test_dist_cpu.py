"""world_size-2 gloo test of the sample-sharding protocol (host logic) with the CPU oracle as the
compute backend: sharded == unsharded, bit for bit, for noise, returns, statistics and Ybar."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mbd_b200
    from mbd_b200.planners.sharding import ShardPlan, tree_sum_rows
    from oracle import oracle as orc
    from oracle import planner as opl
    car = mbd_b200.envs.get_env("car2d")
    N, H, Nu = 256, 20, 2
    key = np.uint32([11, 22])
    Ybar = np.linspace(-0.2, 0.2, H * Nu).astype(np.float32)
    plan = ShardPlan.from_env(N)
    assert (plan.P, plan.rank, plan.n_local, plan.n_begin) == (world, rank, N // world, rank * N // world)
    Y_loc = orc.sample_Y0s(key, N, H * Nu, 0.6, Ybar, plan.n_begin, plan.n_begin + plan.n_local)
    out = orc.car2d_rollout(car.params, car.x0, Y_loc.reshape(plan.n_local, H, Nu), xref=car.xref)
    rews_all, logpd_all = torch.empty(N), torch.empty(N)
    plan.all_gather(rews_all, torch.from_numpy(out["rews"]))
    plan.all_gather(logpd_all, torch.from_numpy(out["logpd"]))
    # every rank computes identical global statistics, then its local partial
    _, mean, w = opl.reverse_once_stats(rews_all.numpy(), np.zeros((N, 1), np.float32), 0.1, logpd=logpd_all.numpy(),
                                        rew_xref=car.rew_xref)
    w_loc = w[plan.n_begin:plan.n_begin + plan.n_local]
    partial = torch.from_numpy((w_loc[:, None].astype(np.float64) * Y_loc).sum(0).astype(np.float32))
    partials = torch.empty((world, H * Nu))
    plan.all_gather(partials, partial)
    q.put((rank, rews_all.numpy().copy(), logpd_all.numpy().copy(), tree_sum_rows(partials).numpy().copy(), float(mean)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_unsharded_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # reference: the unsharded computation in this process
    import mbd_b200
    from oracle import oracle as orc
    from oracle import planner as opl
    car = mbd_b200.envs.get_env("car2d")
    N, H, Nu = 256, 20, 2
    key = np.uint32([11, 22])
    Ybar = np.linspace(-0.2, 0.2, H * Nu).astype(np.float32)
    Y = orc.sample_Y0s(key, N, H * Nu, 0.6, Ybar)
    out = orc.car2d_rollout(car.params, car.x0, Y.reshape(N, H, Nu), xref=car.xref)
    for r in res:
        assert np.array_equal(r[1], out["rews"]) and np.array_equal(r[2], out["logpd"])
    assert np.array_equal(res[0][3], res[1][3]) and res[0][4] == res[1][4]
    _, mean, w = opl.reverse_once_stats(out["rews"], Y, 0.1, logpd=out["logpd"], rew_xref=car.rew_xref)
    full = (w[:, None].astype(np.float64) * Y).sum(0)
    assert np.allclose(res[0][3], full, rtol=1e-5, atol=1e-7)


def test_shard_plan_and_tree_sum():
    from mbd_b200.planners.sharding import ShardPlan, tree_sum_rows
    import pytest
    p = ShardPlan(8192, 8, 3)
    assert (p.n_local, p.n_begin) == (1024, 3072)
    with pytest.raises(ValueError):
        ShardPlan(100, 8, 0)
    rows = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
    assert torch.equal(tree_sum_rows(rows), rows.sum(0))
    a = torch.tensor([[1e8], [1.0], [-1e8], [1.0]])
    assert tree_sum_rows(a).item() == 0.0  # fp32 pairwise (1e8+1)+(-1e8+1) = 0; left-to-right would give 1
