"""Multi-process (world_size 2, gloo, CPU) tests of the sharding / gathering host logic of the N > 1 path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from covidcurve_b200 import dist as ccdist
from covidcurve_b200.summaries import rhat_per_parameter


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_draws(chain_ids, n_it=50, d=4):
    """Deterministic draws that depend only on the GLOBAL chain id (like the device RNG keyed by global chain id)."""
    out = np.empty((len(chain_ids), n_it, d))
    for i, c in enumerate(chain_ids):
        out[i] = np.random.default_rng(1000 + c).normal(loc=0.1 * c, size=(n_it, d))
    return out


def _worker(rank, world, port, chains, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        off, n_local = ccdist.shard_chains(chains, world, rank)
        local = torch.from_numpy(_fake_draws(range(off, off + n_local)))
        full = ccdist.allgather_draws(local, chains)
        rhat = ccdist.rhat_allreduce(local, chains)
        q.put((rank, full.numpy(), rhat.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("chains", [6, 5])  # even and ragged shards
def test_allgather_and_rhat_match_single_process(chains):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, chains, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _fake_draws(range(chains))
    for rank, full, rhat in res:
        assert np.array_equal(full, want), "rank %d gathered draws differ from the single-process run" % rank
        np.testing.assert_allclose(rhat, rhat_per_parameter(want), rtol=1e-10)


def _numpy_standins():
    """The three native calls behind the sharded summary helpers, replaced by numpy restatements (test-only: this suite runs without a
    GPU and checks the gather logic - shard order, ragged shards, the autoburnin window -, not the kernels)."""
    import covidcurve_b200._lib as L
    from covidcurve_b200 import summaries as S

    def chain_moments(draws, first_iter=0, device=0):
        _, m, v = S.chain_moments(draws.numpy()[:, first_iter:, :])
        return torch.from_numpy(m), torch.from_numpy(v)

    def gelman_from_moments(mean, var, n_used, device=0):
        xbar, s2 = np.asarray(mean), np.asarray(var)
        C, n = xbar.shape[0], float(n_used)
        W, B, muhat = s2.mean(0), n * xbar.var(0, ddof=1), xbar.mean(0)
        var_w, var_b = s2.var(0, ddof=1) / C, 2 * B ** 2 / (C - 1)
        cov = lambda a, b: ((a - a.mean(0)) * (b - b.mean(0))).sum(0) / (C - 1)
        cov_wb = (n / C) * (cov(s2, xbar ** 2) - 2 * muhat * cov(s2, xbar))
        V = (n - 1) * W / n + (1 + 1 / C) * B / n
        var_V = ((n - 1) ** 2 * var_w + (1 + 1 / C) ** 2 * var_b + 2 * (n - 1) * (1 + 1 / C) * cov_wb) / n ** 2
        df_adj = (2 * V ** 2 / var_V + 3) / (2 * V ** 2 / var_V + 1)
        R2f, R2r = (n - 1) / n, (1 + 1 / C) / n * B / W
        return np.stack([np.sqrt(((1 - 1 / n) * W + B / n) / W), np.sqrt(df_adj * (R2f + R2r)), W, B, 2 * W ** 2 / var_w, df_adj,
                         np.full_like(W, R2f), R2r], axis=1)

    def summarize_draws(draws, by_chain=True, device=0):
        x = draws.numpy()
        out = np.zeros((x.shape[0], x.shape[2], len(L.SUMM_FIELDS)))
        for c in range(x.shape[0]):
            for p in range(x.shape[2]):
                v = x[c, :, p]
                q = np.quantile(v, [0.025, 0.5, 0.975])
                out[c, p, :8] = [v.min(), q[0], q[1], v.mean(), q[2], v.max(), S.effective_size(v), S.geweke_z(v)]
        return out
    L.chain_moments, L.gelman_from_moments, L.summarize_draws = chain_moments, gelman_from_moments, summarize_draws


def _summary_worker(rank, world, port, chains, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _numpy_standins()
        off, n_local = ccdist.shard_chains(chains, world, rank)
        local = torch.from_numpy(_fake_draws(range(off, off + n_local), n_it=61))
        est, upper, rhat = ccdist.gelman_allgather(local, chains)
        table = ccdist.cred_intervals_allgather(local, chains, ["a", "b", "c", "d"])
        q.put((rank, est, upper, rhat, table[["chain", "param", "median", "ESS"]].to_numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("chains", [6, 5])
def test_sharded_summaries_match_single_process(chains):
    """gelman_allgather / cred_intervals_allgather over two ranks (even and ragged shards) against the single-process summaries of the
    same chains: coda's gelman.diag with its autoburnin window, drjacoby's R-hat, the per-chain table in global chain order."""
    from covidcurve_b200 import summaries as S
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_summary_worker, args=(r, world, port, chains, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _fake_draws(range(chains), n_it=61)
    e0, u0 = S.gelman_diag(want)                                   # autoburnin: iterations 30 .. 60
    r0 = S.rhat_per_parameter(want[:, 61 // 2:, :])
    for rank, est, upper, rhat, tab in res:
        np.testing.assert_allclose(est, e0, rtol=1e-10)
        np.testing.assert_allclose(upper, u0, rtol=1e-10)
        np.testing.assert_allclose(rhat, r0, rtol=1e-10)
        assert list(tab[:, 0]) == [c + 1 for c in range(chains) for _ in range(4)] and list(tab[:4, 1]) == ["a", "b", "c", "d"]
        np.testing.assert_allclose(tab[:, 2].astype(float).reshape(chains, 4), np.median(want, axis=1), rtol=1e-12)
        assert np.all(tab[:, 3].astype(float) > 0)


def test_shard_chains_partitions_exactly():
    for chains in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            blocks = [ccdist.shard_chains(chains, world, r) for r in range(world)]
            assert sum(n for _, n in blocks) == chains
            assert blocks[0][0] == 0
            for (o1, n1), (o2, _) in zip(blocks, blocks[1:]):
                assert o1 + n1 == o2
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1
    with pytest.raises(ValueError):
        ccdist.shard_chains(4, 2, 2)
