"""torch.distributed plumbing for the bk_comm callbacks (one process per GPU, NCCL over NVLink).

The C library hands the callbacks raw DEVICE pointers on this rank's GPU; they are wrapped as
torch tensors (no copy) through the CUDA array interface and passed to the NCCL collectives.
On CPU-only boxes the same class runs over gloo with host pointers (tests, world_size 2)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._lib import ALLGATHERV_FN, ALLREDUCE_FN, BROADCAST_FN, Comm


class _DevView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _wrap(ptr, n, device):
    if device.type == "cuda":
        return torch.as_tensor(_DevView(ptr, n), device=device)
    buf = (C.c_double * int(n)).from_address(int(ptr))
    return torch.from_numpy(np.frombuffer(buf, dtype=np.float64, count=int(n)))


class TorchComm:
    """bk_comm implemented with torch.distributed (backend nccl on GPUs, gloo on CPU)."""

    def __init__(self, device=None, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device(device) if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available()
            else torch.device("cpu"))
        self._cbs = (ALLREDUCE_FN(self._allreduce), ALLGATHERV_FN(self._allgatherv),
                     BROADCAST_FN(self._broadcast))
        self.struct = Comm(self.rank, self.world, None, *self._cbs)

    def _sync(self):
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def _allreduce(self, user, ptr, n):
        try:
            t = _wrap(ptr, n, self.device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self._sync()
            return 0
        except Exception as e:  # noqa: BLE001 - must not unwind through C
            print("bk_comm.allreduce failed:", e, flush=True)
            return 1

    def _allgatherv(self, user, ptr, counts, displs):
        try:
            cs = [int(counts[r]) for r in range(self.world)]
            ds = [int(displs[r]) for r in range(self.world)]
            total = max(d + c for d, c in zip(ds, cs))
            full = _wrap(ptr, total, self.device)
            # in place: every rank broadcasts its own segment (segments may be unequal)
            for r in range(self.world):
                if cs[r] > 0:
                    dist.broadcast(full[ds[r]:ds[r] + cs[r]], src=dist.get_global_rank(self.group, r)
                                   if self.group is not None else r, group=self.group)
            self._sync()
            return 0
        except Exception as e:  # noqa: BLE001
            print("bk_comm.allgatherv failed:", e, flush=True)
            return 1

    def _broadcast(self, user, ptr, n, root):
        try:
            t = _wrap(ptr, n, self.device)
            src = dist.get_global_rank(self.group, root) if self.group is not None else root
            dist.broadcast(t, src=src, group=self.group)
            self._sync()
            return 0
        except Exception as e:  # noqa: BLE001
            print("bk_comm.broadcast failed:", e, flush=True)
            return 1

    def gather_objects(self, obj):
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def barrier(self):
        dist.barrier(group=self.group)
