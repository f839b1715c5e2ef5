"""Multi-process path: shard by path index, track per rank, gather on rank 0 (world_size 2, gloo, CPU).
The per-rank backend here is the oracle -- this tests the host-side sharding / gather logic, the
GPU library is exercised by the -m gpu tests and by bench.py under torchrun."""
import os
import socket

import numpy as np
import pytest

import hcb200
from hcb200 import sharding, workloads


def test_shard_ranges_partition_the_batch():
    for n, w in ((10, 3), (7, 8), (110592, 8), (0, 2), (5, 1)):
        r = [sharding.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    api = pyoracle.load()
    w = workloads.katsura8(1)
    lo, hi = sharding.shard_range(w.N, rank, world)
    part = w.slice(lo, hi)
    res = part.track(api, w.build(api))
    full = sharding.gather_results(res, dist)
    if rank == 0:
        np.savez(out, return_code=full.return_code, solution=full.solution, accepted=full.accepted_steps)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(tmp_path, oracle):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    w = workloads.katsura8(1)
    ref = w.track(oracle, w.build(oracle))
    assert (got["return_code"] == ref.return_code).all() and (got["accepted"] == ref.accepted_steps).all()
    assert np.array_equal(got["solution"], ref.solution)          # same code, same inputs: bit-identical, in path order
    c = sharding.class_counts(ref)
    assert c["success"] == 256 and c["nonsingular"] == 256 and c["paths"] == 256


def _worker_index_range(rank, world, port, out):
    """Each rank tracks its index range of the total-degree start system: no start matrix exists anywhere
    (hc_track_total_degree, first = lo).  Backend: the device code compiled for the host."""
    import sys
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_sim"))
    import pysim
    from hcb200 import capi, start_systems, systems
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    api = pysim.load()
    td = start_systems.total_degree(systems.katsura(6), 0.4 + 1.3j)
    H = api.homotopy(capi.H_STRAIGHT_LINE, api.system(td.F), api.system(td.G), gamma=td.gamma, G_params=td.scaling, F_params=[])
    lo, hi = sharding.shard_range(td.n_paths(), rank, world)
    res = capi.track_total_degree(H, td.degrees, first=lo, count=hi - lo)
    full = sharding.gather_results(res, dist)
    if rank == 0:
        np.savez(out, return_code=full.return_code, solution=full.solution, accepted=full.accepted_steps)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_the_total_degree_index_range(tmp_path, sim):
    import torch.multiprocessing as mp
    from helpers import straight_line
    from hcb200 import systems
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker_index_range, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    td, H = straight_line(sim, systems.katsura(6), 0.4 + 1.3j)
    ref = H.track_batch(td.start_solutions())
    assert (got["return_code"] == ref.return_code).all() and (got["accepted"] == ref.accepted_steps).all()
    assert np.array_equal(got["solution"], ref.solution)
    assert int((ref.return_code == 1).sum()) == 64
