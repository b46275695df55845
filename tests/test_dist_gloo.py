"""The N > 1 path on CPU: two processes over gloo (the product's GPU kernels cannot run
here, so each rank integrates its shard with the CPU oracle -- test infrastructure --
while everything else is the code bench.py and a multi-GPU user run: rank context from
the torchrun environment, shard bounds, barrier, max-over-ranks timing, count sums,
gather of the sharded rows on rank 0)."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import _oracle as orc
    import _problems as pr
    from numbalsoda_b200._dist import Ranks
    R = Ranks(backend="gloo")
    assert (R.rank, R.world) == (rank, world)
    b0, b1 = R.shard(B)
    P = pr.robertson_sweep(B)  # the global ensemble; this rank owns rows [b0, b1)
    u0 = np.array([1.0, 0.0, 0.0])
    te = np.linspace(0, 1e5, 20)
    R.barrier()
    us, ok, cnt = orc.solve_batch("lsoda", orc.builtin_rhs("robertson"), u0, te, P[b0:b1], 1e-8, 1e-8)
    t_local = 0.25 * (rank + 1)  # a stand-in for the CUDA-event time of this rank
    t_max = R.max(t_local)
    n_ok = R.sum(int(ok.sum()))
    steps = R.sum(int(cnt[:, 0].sum()))
    full = R.gather_rows(us, B)
    R.barrier()
    if rank == 0:
        q.put((b0, b1, t_max, n_ok, steps, full))
    R.close()


@pytest.mark.parametrize("B", [37, 64])
def test_two_ranks_shard_reduce_and_gather(B):
    import torch.multiprocessing as mp
    sys.path.insert(0, HERE)
    import _oracle as orc
    import _problems as pr
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    [p.start() for p in procs]
    try:
        b0, b1, t_max, n_ok, steps, full = q.get(timeout=120)
    finally:
        [p.join(60) for p in procs]
        [p.kill() for p in procs if p.is_alive()]
    assert all(p.exitcode == 0 for p in procs)
    assert (b0, b1) == (0, (B + 1) // 2)
    assert t_max == 0.5  # the slowest rank's time
    ref, okr, cr = orc.solve_batch("lsoda", orc.builtin_rhs("robertson"), np.array([1.0, 0.0, 0.0]),
                                   np.linspace(0, 1e5, 20), pr.robertson_sweep(B), 1e-8, 1e-8)
    assert n_ok == int(okr.sum()) == B
    assert steps == int(cr[:, 0].sum())
    assert full.shape == ref.shape and np.array_equal(full, ref)  # sharded == unsharded


def test_single_process_context_is_a_no_op():
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    from numbalsoda_b200._dist import Ranks
    R = Ranks()
    assert (R.rank, R.world) == (0, 1) and R.shard(10) == (0, 10)
    assert R.max(1.5) == 1.5 and R.sum(3) == 3.0
    R.barrier()
    R.close()
