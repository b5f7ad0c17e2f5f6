"""Multi-GPU plumbing: chains shard over ranks with NO per-iteration communication; collectives appear only
at the end of a run, to gather cold-rung draws for the summaries of R/summarize_plot.R:46-153 (Gelman-Rubin, credible
intervals).  `torch.distributed` is the transport: NCCL over NVLink for device tensors, gloo for the CPU tests.

RNG streams are keyed by GLOBAL chain id (covidcurve_b200/csrc/cc_rng.cuh), so a run sharded over N ranks reproduces the
single-process run chain by chain.
"""
from __future__ import annotations

import numpy as np


def shard_chains(chains: int, world: int, rank: int):
    """Contiguous block of chains owned by `rank`: (chain_offset, n_local).  Blocks differ by at most one chain."""
    if chains < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError("need chains >= 0, world >= 1 and 0 <= rank < world")
    base, extra = divmod(chains, world)
    n_local = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, n_local


def shard_vectors(n: int, world: int, rank: int):
    """Contiguous block of parameter vectors owned by `rank` for the likelihood-only workload."""
    return shard_chains(n, world, rank)


def allgather_draws(local, chains: int, group=None):
    """All-gather draws[chain_local, ...] over ranks into draws[chains, ...] in global chain order.

    `local` is a torch tensor (CPU for gloo, CUDA for NCCL) whose first dimension is this rank's chain count; ragged
    shards (chains not divisible by the world size) are padded to the largest shard for the collective.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [shard_chains(chains, world, r)[1] for r in range(world)]
    if local.shape[0] != counts[rank]:
        raise ValueError("rank %d holds %d chains, expected %d" % (rank, local.shape[0], counts[rank]))
    width = max(counts)
    if min(counts) == width:
        # equal shards (the usual case: chains divisible by the GPU count): one collective straight into the result tensor -
        # contiguous chain blocks in rank order ARE the global chain order, so no re-assembly copy is needed
        out = torch.empty((chains,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = local
    if local.shape[0] < width:
        pad = torch.zeros((width - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def rhat_allreduce(local, chains: int, group=None):
    """Gelman-Rubin R-hat per parameter from per-chain moments only: each rank reduces its draws[chain_local, iteration,
    param] to (sum of chain means, sum of squared chain means, sum of chain variances) and one all-reduce of 3 x P numbers
    replaces the gather of every draw.  Same statistic as summaries.rhat_per_parameter on the gathered draws."""
    import torch
    import torch.distributed as dist
    n = local.shape[1]
    mean = local.mean(dim=1)
    var = local.var(dim=1, unbiased=True) if n > 1 else torch.zeros_like(mean)
    stats = torch.stack([mean.sum(dim=0), (mean * mean).sum(dim=0), var.sum(dim=0)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    C = float(chains)
    gmean = stats[0] / C
    W = stats[2] / C
    B = n / (C - 1) * (stats[1] - C * gmean * gmean)
    V = (1 - 1 / n) * W + B / n
    return torch.sqrt(V / W)


def rhat_torch(draws):
    """summaries.rhat_per_parameter for a torch tensor draws[chain, iteration, param] on any device (used on the gathered
    draws where they live, i.e. in HBM)."""
    import torch
    C, n = draws.shape[0], draws.shape[1]
    mean = draws.mean(dim=1)
    var = draws.var(dim=1, unbiased=True) if n > 1 else torch.zeros_like(mean)
    W = var.mean(dim=0)
    B = n / (C - 1) * ((mean - mean.mean(dim=0, keepdim=True)) ** 2).sum(dim=0)
    V = (1 - 1 / n) * W + B / n
    return torch.sqrt(V / W)


def gelman_allgather(local, chains: int, confidence: float = 0.95, autoburnin: bool = True, group=None):
    """coda::gelman.diag and drjacoby's rhat over ALL ranks' chains without moving a draw: each rank reduces its CUDA tensor
    draws[chain_local, iteration, param] to per-chain means and variances with the library's kernel (cc_chain_moments), the
    ranks all-gather those 2 x chains x params numbers (NCCL), and every rank finishes with cc_gelman_from_moments.
    Returns (point estimate, upper limit, drjacoby rhat) - the same values as summaries.device_gelman_diag on gathered draws."""
    import torch
    from . import _lib
    from .summaries import gelman_from_chain_moments
    n = local.shape[1]
    first = n // 2 if autoburnin else 0
    mean, var = _lib.chain_moments(local, first_iter=first)
    both = torch.stack([mean, var], dim=1)                     # [chain_local, 2, param]
    allm = allgather_draws(both, chains, group=group)
    return gelman_from_chain_moments(allm[:, 0, :].contiguous(), allm[:, 1, :].contiguous(), n - first, confidence)


def cred_intervals_allgather(local, chains: int, names, chain_offset: int = 0, group=None):
    """get_cred_intervals(by_chain = TRUE) for all chains of a sharded run: every rank summarises its own chains on its GPU
    (cc_summarize_draws) and only the 9-number rows travel (all-gather of [chain_local, param, field])."""
    import pandas as pd
    import torch
    from . import _lib
    tab = torch.from_numpy(_lib.summarize_draws(local, by_chain=True)).to(local.device)
    allt = allgather_draws(tab, chains, group=group).cpu().numpy()
    G, P = allt.shape[:2]
    cols = {"chain": np.repeat(np.arange(1, G + 1), P), "param": np.tile(np.asarray(list(names), dtype=object), G)}
    for f, c in enumerate(_lib.SUMM_FIELDS[:9]):
        cols[c] = allt[:, :, f].reshape(-1)
    return pd.DataFrame(cols)


def draws_tensor(mcmc, device):
    """Zero-copy torch view of the recorded draws of a covidcurve_b200._lib.Mcmc run that still lives on `device`:
    theta [chains_local, recorded rungs, rows_cap, d] (only the first `mcmc.rows` rows are valid)."""
    import torch

    theta_ptr, _, _, cap = mcmc.device_draws()
    shape = (mcmc.chains, mcmc.n_rec_rungs, cap, mcmc.d)

    class _Wrap:
        __cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (theta_ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(_Wrap(), device=device)
