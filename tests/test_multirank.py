"""world_size-2 gloo test of the multi-GPU host logic (shard by batch index, no collective on the solve path,
one final gather).  The per-rank 'solve' is the CPU oracle here; on GPUs it is the CUDA path (bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle as O
    from ros2_mpc_b200.sharding import solve_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    B = 11
    x0 = np.c_[rng.uniform(-1, 1, (B, 2)), rng.uniform(0, 6.28, B)]
    goal = x0 + np.c_[rng.uniform(-0.5, 0.5, (B, 2)), rng.uniform(-1, 1, B)]
    p = O.variant_params("B", N=8)
    calls = []

    def solve_fn(a):
        calls.append(a["x0"].shape[0])
        r = O.solve_batch(p, a["x0"], a["goal"], nthreads=1)
        return dict(X=r["X"], U=r["U"], status=r["status"])

    # a shared table whose length happens to equal the batch size must not be sliced (it is named in `shared`)
    table = np.arange(B, dtype=np.float64)

    def solve_fn_checked(a):
        assert a["table"].shape == (B,) and np.array_equal(a["table"], table)
        return solve_fn(a)

    out = solve_sharded(solve_fn_checked, dict(x0=x0, goal=goal, table=table), B, rank, world, dist=dist, shared=("table",))
    try:
        solve_sharded(solve_fn, dict(x0=x0, goal=goal, table=table[:3]), B, rank, world, dist=dist)
        raise AssertionError("a mis-shaped batched array must be rejected")
    except ValueError:
        pass
    dist.barrier()
    if rank == 0:
        full = O.solve_batch(p, x0, goal, nthreads=1)
        ok = (np.array_equal(out["X"], full["X"]) and np.array_equal(out["U"], full["U"])
              and np.array_equal(out["status"], full["status"]))
        q.put(("ok" if ok else "mismatch", calls))
    else:
        assert out is None
        q.put(("rank1", calls))
    dist.destroy_process_group()


def test_two_rank_sharded_solve_matches_single_rank():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    tags = dict(res)
    assert "ok" in tags, res
    assert sorted(tags["ok"] + tags["rank1"]) == [5, 6]  # 11 problems -> shards of 6 and 5


def test_contiguous_shards_cover_the_batch():
    """The C ABI (b200mpc_solve_batch_multi) and sharding.contiguous_shard use the same rule: sizes differ by at most one."""
    from ros2_mpc_b200.sharding import contiguous_shard
    for B, G in ((11, 2), (1048576, 8), (5, 8), (0, 3), (4096, 3)):
        sl = [contiguous_shard(B, g, G) for g in range(G)]
        assert sl[0][0] == 0 and sl[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
        sizes = [h - l for l, h in sl]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
