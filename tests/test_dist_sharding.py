"""Sharding / gathering logic of the multi-GPU launcher on CPU (gloo, world_size 2) and as a 1-process
fake world.  The solve itself is replaced by the oracle HERE (test only) because no GPU is present."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lbmpc_b200.dist import gather_results, reduce_stats, sample_initial_states, shard_range


def test_shard_ranges_partition_the_batch():
    for batch in (0, 1, 7, 1024, 1025, 65536, 10 ** 6):
        for world in (1, 2, 3, 4, 8):
            rs = [shard_range(batch, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1


def test_initial_states_do_not_depend_on_world_size():
    full = sample_initial_states(1000, seed=3)
    for world in (2, 4, 8):
        parts = [full[slice(*shard_range(1000, r, world))] for r in range(world)]
        assert np.array_equal(np.concatenate(parts), full)
    assert np.array_equal(full[0], [-0.35, -0.4, 0.0, 0.0])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "learning-based-mpc_b200"), os.path.join(root, "oracle")):
        sys.path.insert(0, p)
    import lbmpc_b200
    from oracle_py import OracleProblem
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mdl = lbmpc_b200.moore_greitzer_model("LBMPC")
    X0 = sample_initial_states(batch, seed=1)
    lo, hi = shard_range(batch, rank, world)
    r = OracleProblem("C", "LBMPC", mdl, 20).solve_batch(X0[lo:hi])     # stand-in for the GPU shard solve
    local = {k: torch.from_numpy(np.ascontiguousarray(r[k])) for k in ("uc", "theta", "obj", "iters", "status")}
    full = gather_results(local, batch)
    from lbmpc_b200.dist import gather_packed
    packed = gather_packed(local, batch, dst=0)          # the one-collective form bench.py uses
    if rank == 0:
        for k in full:
            assert packed[k].dtype == full[k].dtype and torch.equal(packed[k], full[k]), k
    else:
        assert packed == {}
    hosted = gather_packed(local, batch, dst=0, to_host=True)   # ... and its one-copy-to-the-host form
    if rank == 0:
        for k in full:
            assert hosted[k].dtype == full[k].numpy().dtype and np.array_equal(hosted[k], full[k].numpy()), k
    else:
        assert hosted == {}
    stats = reduce_stats(local["status"], local["iters"], local["obj"])
    if rank == 0:
        q.put(({k: v.numpy() for k, v in full.items()}, stats))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gloo_world2_gather_equals_single_rank(models):
    from oracle_py import OracleProblem
    batch, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, stats = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = OracleProblem("C", "LBMPC", models["LBMPC"], 20).solve_batch(sample_initial_states(batch, seed=1))
    for k in ("uc", "theta", "obj", "iters", "status"):
        assert np.array_equal(full[k], ref[k]), k
    assert stats["n_optimal"] == int((ref["status"] == 0).sum())
    assert stats["sum_iters"] == int(ref["iters"].sum()) and stats["max_iters"] == int(ref["iters"].max())
