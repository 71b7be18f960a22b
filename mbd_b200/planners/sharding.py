"""Sample-axis sharding plan (host logic, backend agnostic: NCCL on GPUs, gloo in CPU tests).

Rank r of P owns global samples [r*N/P, (r+1)*N/P).  The noise of a sample is addressed by its
GLOBAL index (jax.random.normal over the full (Nsample,H,Nu) array), so the union of the shards
is exactly the unsharded sample set.  Exchange per diffusion step: all-gather of the per-sample
scalars (rews, and logpd with demos), all-gather of the per-rank partial weighted sums, which
are then combined in a fixed pairwise order — results are bit-identical for every P that keeps
N/P a multiple of 64*2^k (see DESIGN.md "determinism").
"""
from __future__ import annotations

import dataclasses

import torch
import torch.distributed as dist


@dataclasses.dataclass
class ShardPlan:
    N: int
    P: int = 1
    rank: int = 0
    group: object = None

    @classmethod
    def from_env(cls, N: int, group=None) -> "ShardPlan":
        if dist.is_available() and dist.is_initialized():
            return cls(N, dist.get_world_size(group), dist.get_rank(group), group)
        return cls(N, 1, 0, None)

    def __post_init__(self):
        if self.N % self.P != 0:
            raise ValueError(f"Nsample={self.N} must be divisible by the number of ranks ({self.P})")

    @property
    def n_local(self) -> int:
        return self.N // self.P

    @property
    def n_begin(self) -> int:
        return self.rank * self.n_local

    def all_gather(self, out_all: torch.Tensor, local: torch.Tensor) -> torch.Tensor:
        """out_all [P*len(local)] <- concat over ranks (rank order).  No-op alias when P == 1."""
        if self.P == 1:
            if out_all.data_ptr() != local.data_ptr():
                out_all.copy_(local)
            return out_all
        dist.all_gather_into_tensor(out_all.view(-1), local.view(-1), group=self.group)
        return out_all


def tree_sum_rows(rows: torch.Tensor) -> torch.Tensor:
    """Adjacent-pairwise tree sum over dim 0 — the order k_update uses for the rank partials."""
    r = rows
    while r.shape[0] > 1:
        if r.shape[0] % 2:
            # same fold as the device binary-counter stack for non power-of-two counts
            head, tail = r[:-1], r[-1:]
            r = torch.cat([head[0::2] + head[1::2], tail], dim=0)
        else:
            r = r[0::2] + r[1::2]
    return r[0]
