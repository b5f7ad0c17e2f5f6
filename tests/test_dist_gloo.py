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
