"""world_size-2 (and 3) gloo test of the multi-GPU host logic on CPU: index-range sharding of the seeded workload
and the final statistics gather.  The per-shard solver here is the CPU oracle (test infrastructure) standing in for
the per-GPU kernel; what is under test is that shards tile the batch exactly and the gathered statistics equal the
single-process ones."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, per_rank, q):
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    from oracle.pyoracle import OracleLib
    pkg = load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = pkg.problems.quadrotor(20)
    b0, b1 = pkg.sharding.shard_range(rank, world, per_rank=per_rank)
    x0, xref = pkg.workloads.quadrotor_hover_batch(b0, b1, mult=0.25)
    r = OracleLib().solve_batch(prob, x0, xref, dtype=np.float32, nthreads=2)
    vec = pkg.sharding.local_stats(r.iter, r.status, prob.max_iter)
    tot, tmax = pkg.sharding.gather_stats(vec, [10.0 + rank, 1.0], dist)
    if rank == 0:
        q.put((tot, tmax, b0, b1))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_statistics_equal_single_process(pkg, oracle, world):
    per_rank = 700
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + world + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    tot, tmax, b0, b1 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    prob = pkg.problems.quadrotor(20)
    x0, xref = pkg.workloads.quadrotor_hover_batch(0, per_rank * world, mult=0.25)
    r = oracle.solve_batch(prob, x0, xref, dtype=np.float32, nthreads=4)
    ref = pkg.sharding.local_stats(r.iter, r.status, prob.max_iter)
    assert np.array_equal(tot, ref)
    assert tmax[0] == 10.0 + world - 1 and (b0, b1) == (0, per_rank)


def test_shard_ranges_tile(pkg):
    S = pkg.sharding
    for world in (1, 2, 4, 8):
        edges = [S.shard_range(r, world, total=1000003) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == 1000003
        assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
        assert [S.shard_range(r, world, per_rank=1 << 20) for r in range(world)] == [(r << 20, (r + 1) << 20) for r in range(world)]
