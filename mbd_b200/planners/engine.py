"""Device-side engine of one reverse-diffusion step (`reverse_once`,
/root/reference/mbd/planners/mbd_planner.py:97-135), sample-sharded over ranks.

Per step and rank r of P (samples [r*N/P, (r+1)*N/P), noise addressed by GLOBAL index so the
result does not depend on P):
  1. mbd_sample_rollout      fused sampling + rollouts   -> Y0s_local, rews_local (+logpd_local)
  2. gather(rews[, logpd]) from all ranks                (skipped for P == 1; fused NVLink peer-memory
                                                           kernel mbd_peer_gather, NCCL all_gather as fallback)
  3. mbd_softmax_weights     global mean/std/demo/softmax -> weights_local
  4. mbd_weighted_sum        partial Ybar over local samples (deterministic order)
  5. gather(partial) from all ranks                      (same mechanism)
  6. mbd_update              tree-sum of rank partials + the literal update lines 130-133
All launches go to the current CUDA stream; nothing synchronises with the host.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from .. import ops
from .sharding import ShardPlan


def make_schedule(beta0: float, betaT: float, Ndiffuse: int):
    """mbd_planner.py:84-87 in float32."""
    betas = np.linspace(beta0, betaT, Ndiffuse, dtype=np.float32)
    alphas = (np.float32(1.0) - betas).astype(np.float32)
    alphas_bar = np.cumprod(alphas, dtype=np.float32)
    sigmas = np.sqrt(np.float32(1.0) - alphas_bar).astype(np.float32)
    return betas, alphas, alphas_bar, sigmas


def update_coef(alphas, alphas_bar, i: int):
    """The float32 scalars of mbd_planner.py:100,130-133 for step i."""
    one = np.float32(1.0)
    ab = np.float32(alphas_bar[i])
    return [np.sqrt(ab), one / (one - ab), one - ab, one / np.sqrt(np.float32(alphas[i])), np.sqrt(np.float32(alphas_bar[i - 1]))]


class DiffusionEngine:
    def __init__(self, env, Nsample: int, Hsample: int, temp_sample: float, enable_demo: bool, state_init,
                 device: Optional[torch.device] = None, group=None):
        self.env = env
        self.N, self.H, self.temp = int(Nsample), int(Hsample), float(temp_sample)
        self.enable_demo = bool(enable_demo)
        self.plan = ShardPlan.from_env(self.N, group)
        self.group, self.P, self.rank = group, self.plan.P, self.plan.rank
        self.n_local, self.n_begin = self.plan.n_local, self.plan.n_begin
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.Nu = env.action_size
        self.HNu = self.H * self.Nu
        d = self.device
        f = dict(device=d, dtype=torch.float32)
        self.Y0s = torch.empty((self.n_local, self.HNu), **f)
        # ---- exchange buffers.  P > 1 on CUDA: one peer-mapped symmetric buffer per rank
        #      [rews n_local | logpd n_local | partial HNu | 2 flag rows of 8 words], exchanged by the fused
        #      mbd_peer_gather kernel over NVLink; fallback (MBD_EXCHANGE=nccl, or no symmetric memory): NCCL.
        self.sym = None
        self.exchange = "none" if self.P == 1 else os.environ.get("MBD_EXCHANGE", "p2p")
        if self.P > 1 and self.exchange == "p2p":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                import torch.distributed as dist
                words = 2 * self.n_local + self.HNu + 16
                self.sym = symm_mem.empty(words, dtype=torch.float32, device=d)
                self.sym.zero_()
                self.sym_hdl = symm_mem.rendezvous(self.sym, dist.group.WORLD if group is None else group)
                self.peer_ptrs = [int(p) for p in self.sym_hdl.buffer_ptrs]
                self.off_rews, self.off_logpd, self.off_partial = 0, self.n_local, 2 * self.n_local
                self.off_flags = 2 * self.n_local + self.HNu
                self.epoch = 0
                self.xerr = torch.zeros(1, dtype=torch.int32, device=d)
                torch.cuda.synchronize()
                dist.barrier(group=group)
            except Exception as e:  # noqa: BLE001
                if os.environ.get("MBD_EXCHANGE") == "p2p":
                    raise
                self.sym, self.exchange = None, "nccl"
        if self.sym is not None:
            self.rews_local = self.sym[self.off_rews:self.off_rews + self.n_local]
            self.logpd_local = self.sym[self.off_logpd:self.off_logpd + self.n_local] if self.enable_demo else None
        else:
            self.rews_local = torch.empty(self.n_local, **f)
            self.logpd_local = torch.empty(self.n_local, **f) if self.enable_demo else None
        self.rews_all = self.rews_local if self.P == 1 else torch.empty(self.N, **f)
        self.logpd_all = None
        if self.enable_demo:
            self.logpd_all = self.logpd_local if self.P == 1 else torch.empty(self.N, **f)
        self.weights = torch.empty(self.n_local, **f)
        self.scalars = torch.zeros(4, **f)
        self.logp_scratch = torch.empty(self.N, **f)
        self.run_scratch = torch.empty(((self.n_local + ops.RUN - 1) // ops.RUN) * self.HNu, **f)
        self.partial = self.sym[self.off_partial:self.off_partial + self.HNu] if self.sym is not None else torch.empty(self.HNu, **f)
        self.partials = self.partial if self.P == 1 else torch.empty((self.P, self.HNu), **f)
        self.Ybar_out = torch.empty(self.HNu, **f)
        # own kernels per step: sample_rollout, softmax_weights, wsum_runs, update (+ wsum_tree and, with the fused NVLink
        # exchange, two k_peer_gather launches when sharded)
        self.launches_per_step = 4 if self.P == 1 else (7 if self.sym is not None else 5)
        self.launches_last_step = self.launches_per_step
        # ONE cooperative kernel per diffusion step (mbd_reverse_step; 1 GPU, no demo, Brax env) is available with
        # MBD_SINGLE_KERNEL=1.  It is bit-identical but measured 2 % SLOWER than the five launches (1.581 vs 1.551 ms
        # at 8192x50: every CTA recomputes the global statistics after the grid barrier, which costs more than the
        # four launch gaps it removes), so the separate launches stay the default.
        self.single_kernel = (self.P == 1 and not self.enable_demo and env.kind == "xpbd"
                              and os.environ.get("MBD_SINGLE_KERNEL", "0") == "1")
        if env.kind == "xpbd":
            self.model = env.device_model(d)
            raw = state_init.pipeline_state.raw if hasattr(state_init, "pipeline_state") else state_init
            self.state_init = torch.as_tensor(np.ascontiguousarray(raw, dtype=np.float32), device=d)
            self.xref = torch.as_tensor(env.xref, device=d).contiguous() if self.enable_demo else None
        elif env.kind == "car2d":
            self.params, xref = env.device_params()
            x0 = state_init.pipeline_state if hasattr(state_init, "pipeline_state") else state_init
            self.state_init = torch.as_tensor(np.ascontiguousarray(x0, dtype=np.float32), device=d)
            self.xref = xref if self.enable_demo else None
        else:
            raise ValueError(env.kind)
        self.rew_xref = float(getattr(env, "rew_xref", 0.0))

    # ---- pieces (also used one by one in tests) ------------------------------------------------
    def rollout_phase(self, key, sigma: float, Ybar_i: torch.Tensor):
        if self.env.kind == "xpbd":
            ops.sample_rollout(self.model, self.state_init, key, self.N, self.n_begin, self.n_local, self.H, float(sigma), Ybar_i,
                               self.Y0s, self.rews_local, xref=self.xref, logpd_out=self.logpd_local)
        else:
            ops.car2d_rollout(self.params, self.state_init, self.Y0s.view(self.n_local, self.H, 2), xref=self.xref, key=key,
                              n_total=self.N, n_begin=self.n_begin, sigma=float(sigma), Ybar=Ybar_i, rews_out=self.rews_local,
                              logpd_out=self.logpd_local)

    def gather_phase(self):
        if self.sym is not None:
            self.epoch += 1
            if self.enable_demo:   # rews and logpd are adjacent: one gather of 2*n_local words per rank, then unzip
                if not hasattr(self, "_both"):
                    self._both = torch.empty((self.P, 2 * self.n_local), device=self.device)
                both = self._both
                ops.peer_gather(self.peer_ptrs, self.P, self.rank, self.off_rews, 2 * self.n_local, self.off_flags, self.epoch, both,
                                self.xerr)
                self.rews_all.view(self.P, self.n_local).copy_(both[:, : self.n_local])
                self.logpd_all.view(self.P, self.n_local).copy_(both[:, self.n_local:])
            else:
                ops.peer_gather(self.peer_ptrs, self.P, self.rank, self.off_rews, self.n_local, self.off_flags, self.epoch,
                                self.rews_all, self.xerr)
        elif self.P > 1:
            self.plan.all_gather(self.rews_all, self.rews_local)
            if self.enable_demo:
                self.plan.all_gather(self.logpd_all, self.logpd_local)

    def reduce_phase(self, Ybar_i: torch.Tensor, coef, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        out = self.Ybar_out if out is None else out
        ops.softmax_weights(self.rews_all, self.logpd_all, self.n_begin, self.n_local, self.temp, self.rew_xref, self.weights,
                            self.scalars, self.logp_scratch)
        if self.P == 1:
            # one rank: the tree over the 64-sample runs IS the rank tree of mbd_update — skip the separate tree launch
            nruns = ops.weighted_sum_runs(self.weights, self.Y0s, self.HNu, self.run_scratch)
            ops.update(self.run_scratch, nruns, self.HNu, Ybar_i, coef, out)
            return out
        ops.weighted_sum(self.weights, self.Y0s, self.HNu, self.run_scratch, self.partial)
        if self.sym is not None:
            ops.peer_gather(self.peer_ptrs, self.P, self.rank, self.off_partial, self.HNu, self.off_flags + 8, self.epoch,
                            self.partials, self.xerr)
        elif self.P > 1:
            self.plan.all_gather(self.partials, self.partial)
        ops.update(self.partials, self.P, self.HNu, Ybar_i, coef, out)
        return out

    def check_exchange(self):
        """Raises if a fused peer gather ever timed out on its cross-GPU barrier (a peer died or diverged).
        Synchronises: call it outside the step loop."""
        if self.sym is not None and int(self.xerr.item()) != 0:
            raise ops.MbdError("mbd_peer_gather: cross-GPU barrier timed out (a peer rank stopped participating)")

    def reverse_once(self, key, sigma: float, Ybar_i: torch.Tensor, coef, out: Optional[torch.Tensor] = None):
        """One diffusion step.  Returns (Ybar_im1 [HNu] device tensor, rews.mean() device scalar view)."""
        if self.single_kernel:
            out_t = self.Ybar_out if out is None else out
            if ops.reverse_step(self.model, self.state_init, key, self.n_local, self.H, float(sigma), Ybar_i, self.temp, coef, self.Y0s,
                                self.rews_local, self.weights, self.scalars, self.run_scratch, out_t):
                self.launches_last_step = 1
                return out_t, self.scalars[0]
            self.single_kernel = False   # configuration not covered: stay on the separate launches
        self.launches_last_step = self.launches_per_step
        self.rollout_phase(key, sigma, Ybar_i)
        self.gather_phase()
        out = self.reduce_phase(Ybar_i, coef, out)
        return out, self.scalars[0]
