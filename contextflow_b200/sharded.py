"""Batch sharding of the log-density path over the GPUs of one box.

Every sample's (z, ldj, logp) depends only on that sample, its context row and its noise rows; weights are replicated.
One process per GPU evaluates a contiguous batch slice; the only exchange on the inference path is the final gather of
the (B/N, M) log-probabilities (NCCL all-gather over NVLink; `gloo` in the CPU tests).  No data-path collective."""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, near-equal slices: the first n % world ranks get one extra row."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedLogProb:
    """log_prob over a global batch: each rank scores its slice with `local_fn(x, ctx) -> (b, M)`, then all-gather."""

    def __init__(self, local_fn: Callable, mixtures: int, group: Optional[dist.ProcessGroup] = None):
        self.local_fn, self.M, self.group = local_fn, mixtures, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def local_slice(self, x, ctx):
        lo, hi = shard_bounds(x.shape[0], self.world, self.rank)
        return x[lo:hi], (None if ctx is None else ctx[lo:hi])

    def gather(self, local_logp: torch.Tensor, total: int) -> torch.Tensor:
        """(b_r, M) per rank -> (total, M) on every rank, rows in global batch order (ragged slices are padded)."""
        if self.world == 1:
            return local_logp
        width = -(-total // self.world)
        pad = torch.zeros((width, self.M), device=local_logp.device, dtype=local_logp.dtype)
        pad[: local_logp.shape[0]] = local_logp
        out = torch.empty((self.world * width, self.M), device=local_logp.device, dtype=local_logp.dtype)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        rows = []
        for r in range(self.world):
            lo, hi = shard_bounds(total, self.world, r)
            rows.append(out[r * width: r * width + (hi - lo)])
        return torch.cat(rows, 0)

    def log_prob(self, x, ctx=None):
        """x, ctx hold the GLOBAL batch on every rank (replicated loader); returns the global (B, M) log-prob."""
        xs, cs = self.local_slice(x, ctx)
        return self.gather(self.local_fn(xs, cs), x.shape[0])

    def log_prob_local(self, x_local, ctx_local, total: int):
        """Each rank already holds only its slice (sharded loader)."""
        return self.gather(self.local_fn(x_local, ctx_local), total)


class GradAllReduce:
    """Training-mode exchange: after `loss.backward()` on each rank's slice, one all-reduce(sum) of the trainable
    parameters' fp32 gradients, flattened into a single bucket (0.4-1.2 M floats for the BASELINE configs: one NCCL launch, sized for
    launch latency rather than link count on NVSwitch).  The reference's loss is a batch MEAN (experiment_ad.py:207), so a rank that
    averaged over b_r of the B global rows contributes with weight b_r / B."""

    def __init__(self, params, group: Optional[dist.ProcessGroup] = None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._flat = None

    def __call__(self, local_rows: int, total_rows: int):
        if self.world == 1:
            return
        # parameters that received no gradient (e.g. an embedding table in front of probsample) stay at grad=None: which ones those are is
        # structural, hence identical on every rank, and AdamW must skip them exactly as a single-process run does (no weight decay / moments)
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return
        sizes = [g.numel() for g in grads]
        n = sum(sizes)
        if self._flat is None or self._flat.numel() != n or self._flat.device != grads[0].device:
            self._flat = torch.empty(n, device=grads[0].device, dtype=torch.float32)
            self._views = [c.view_as(g) for c, g in zip(self._flat.split(sizes), grads)]
        # two fused multi-tensor launches around ONE collective instead of a copy per parameter
        torch._foreach_copy_(self._views, grads)
        self._flat.mul_(float(local_rows) / float(total_rows))
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
        torch._foreach_copy_(grads, self._views)


def sharded_actnorm_stats(x_local, group: Optional[dist.ProcessGroup] = None, local_stats: Optional[Callable] = None):
    """ActNorm's data-dependent initialisation (actnorm.py:28-35: t <- mean, logs <- log(unbiased std + 1e-8) over (B,H,W)) when the
    first batch is sharded over ranks: each rank computes its slice's (mean, logstd) with the same kernel as the
    single-process path, one all-gather of (mean, M2, count) per channel, merged with the pairwise-variance formula in float64 -- the
    result is the statistic of the GLOBAL batch, so a sharded run initialises exactly like the reference's single process."""
    if local_stats is None:
        from . import ops
        local_stats = ops.actnorm_stats
    mean_r, logstd_r = local_stats(x_local)
    n_r = x_local.numel() // x_local.shape[1]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mean_r, logstd_r
    std_r = (torch.exp(logstd_r.double()) - 1e-8).clamp_min(0.0) if n_r > 1 else torch.zeros_like(mean_r, dtype=torch.float64)
    pack = torch.stack([mean_r.double(), std_r * std_r * max(n_r - 1, 0), torch.full_like(std_r, float(n_r))], 0).contiguous()   # (3, D)
    world = dist.get_world_size(group)
    flat = torch.empty(world * pack.numel(), device=pack.device, dtype=pack.dtype)
    dist.all_gather_into_tensor(flat, pack.reshape(-1), group=group)
    allp = flat.view((world,) + tuple(pack.shape))
    n = allp[:, 2].sum(0)
    mean = (allp[:, 0] * allp[:, 2]).sum(0) / n
    m2 = (allp[:, 1] + allp[:, 2] * (allp[:, 0] - mean) ** 2).sum(0)
    std = torch.sqrt(m2 / (n - 1))
    return mean.float(), torch.log(std.float() + 1e-8)


def enable_sharded_actnorm_init(group: Optional[dist.ProcessGroup] = None):
    """Route every ActNorm's first-batch initialisation through `sharded_actnorm_stats` (all ranks must run the same model in lockstep)."""
    from .layers.actnorm import ActNorm
    ActNorm.stats_fn = staticmethod(lambda x: sharded_actnorm_stats(x, group))


def disable_sharded_actnorm_init():
    from .layers.actnorm import ActNorm
    ActNorm.stats_fn = None
