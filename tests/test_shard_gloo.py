"""CPU: the N>1 path (contiguous node ranges + optional gather) with world_size 2 on gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from specfab_b200.shard import node_range


def test_node_range_partitions_exactly():
    for N in (0, 1, 7, 8, 1000, 10 ** 7 + 3):
        for world in (1, 2, 3, 8):
            spans = [node_range(N, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, N, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "oracle")); sys.path.insert(0, os.path.join(root, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from specfab_b200.shard import node_range, gather_rows
    import oracle_c as oc
    from util import random_states, random_ugrad
    L = 4
    n = oc.init(L)
    x = random_states(L, N, 1, True)      # every rank builds the same global field, then takes its range
    ug = random_ugrad(N, 2)
    lo, hi = node_range(N, rank, world)
    # stand-in for the GPU step on this rank's shard (CPU test: the oracle port does the arithmetic)
    loc = oc.step_batch(x[lo:hi], ug[lo:hi], dt=1e-2)
    full = gather_rows(torch.from_numpy(np.ascontiguousarray(loc.T)), N)      # library layout (n, n_local)
    if rank == 0:
        ref = oc.step_batch(x, ug, dt=1e-2)
        q.put(float(np.abs(full.numpy().T - ref).max()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_matches_single_rank():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert err == 0.0
