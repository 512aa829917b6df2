"""N > 1 host logic on CPU: the bk_comm callbacks (TorchComm) over gloo with world_size 2.
The C library calls these callbacks with raw pointers; here they are driven directly."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    from bigkrls_b200.dist import TorchComm
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = TorchComm(device="cpu")
        s = comm.struct
        assert (s.rank, s.world) == (rank, world)
        # allreduce_sum
        a = np.arange(5, dtype=np.float64) * (rank + 1)
        assert s.allreduce_sum(None, a.ctypes.data, a.size) == 0
        assert np.array_equal(a, np.arange(5) * 3.0)
        # in-place allgatherv with unequal segments (column blocks of an n x n matrix, n = 7)
        n = 7
        bounds = [n * r // world for r in range(world + 1)]
        counts = (C.c_int64 * world)(*[(bounds[r + 1] - bounds[r]) * n for r in range(world)])
        displs = (C.c_int64 * world)(*[bounds[r] * n for r in range(world)])
        full = np.zeros(n * n)
        full[displs[rank]:displs[rank] + counts[rank]] = rank + 1
        assert s.allgatherv(None, full.ctypes.data, counts, displs) == 0
        expect = np.concatenate([np.full(counts[r], r + 1.0) for r in range(world)])
        assert np.array_equal(full, expect)
        # broadcast
        b = np.full(4, float(rank))
        assert s.broadcast(None, b.ctypes.data, 4, 1) == 0
        assert np.array_equal(b, np.ones(4))
        # fold sharding of crossvalidate: folds k with (k-1) % world == rank, gathered
        mine = {k: {"R2_is": k * 1.0} for k in range(1, 6) if (k - 1) % world == rank}
        merged = {}
        for d in comm.gather_objects(mine):
            merged.update(d)
        assert sorted(merged) == [1, 2, 3, 4, 5]
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_torchcomm_callbacks_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
