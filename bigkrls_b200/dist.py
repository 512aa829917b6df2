"""torch.distributed plumbing for multi-GPU runs (one process per GPU).

Two layers:

* **native peer communicator** (GPUs, default): torch.distributed is only the *bootstrap* - it carries the
  64-byte CUDA IPC handles between the ranks once.  After that every exchange of the fit is a kernel of
  libbigkrls_b200.so storing into the other GPUs' HBM over NVLink 5 / NVSwitch, synchronised by system-scope
  flags (csrc/peer.cu): no Python callback, no host synchronisation on the data path.  This is what enables
  the distributed dense->band stage of the eigensolver and the distributed K X of the Krylov path.

* **callback communicator** (`bk_comm` vtable: allreduce / allgatherv / broadcast on raw pointers): the generic
  fallback, used with gloo on CPU-only boxes to test the host-side partitioning logic (tests/test_dist_cpu.py)
  and available for any other transport.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import ALLGATHERV_FN, ALLREDUCE_FN, BROADCAST_FN, EXCHANGE_FN, Comm, check


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
    """bk_comm over torch.distributed.  On GPUs (`native=True`, the default) the data path is the library's own
    peer-memory communicator; the callbacks remain as the generic path (gloo on CPU)."""

    def __init__(self, device=None, group=None, native=None, ctx=None):
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
        self.struct = Comm(self.rank, self.world, None, *self._cbs, None)
        self.peer = None
        self._exchange_cb = EXCHANGE_FN(self._exchange)
        if native is None:
            native = self.device.type == "cuda" and self.world > 1
        if native:
            self.ctx = ctx or _lib.default_context(self.device.index or 0)
            h = C.c_void_p()
            check(_lib.load().bk_peer_create(self.ctx.handle, self.rank, self.world, self._exchange_cb, None,
                                             C.byref(h)))
            self.peer = h
            self.struct.peer = h

    # ---- bootstrap: host all-gather of a few bytes (IPC handles) -----------------------------------------
    def _exchange(self, user, send, recv, nbytes):
        try:
            mine = C.string_at(send, int(nbytes))
            out = [None] * self.world
            dist.all_gather_object(out, mine, group=self.group)
            C.memmove(recv, b"".join(out), int(nbytes) * self.world)
            return 0
        except Exception as e:  # noqa: BLE001 - must not unwind through C
            print("bk_peer exchange failed:", e, flush=True)
            return 1

    def selftest(self):
        """Collective self-test of the peer collectives; returns the number of mismatching values."""
        bad = C.c_int()
        check(_lib.load().bk_peer_selftest(self.peer, C.byref(bad)))
        return bad.value

    def close(self):
        """Collective: unmap and free the symmetric heaps."""
        if self.peer is not None:
            _lib.load().bk_peer_destroy(self.peer)
            self.peer = None
            self.struct.peer = None

    # ---- generic callback path --------------------------------------------------------------------------
    def _sync(self):
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def _src(self, r):
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def _allreduce(self, user, ptr, n):
        try:
            t = _wrap(ptr, n, self.device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self._sync()
            return 0
        except Exception as e:  # noqa: BLE001
            print("bk_comm.allreduce failed:", e, flush=True)
            return 1

    def _allgatherv(self, user, ptr, counts, displs):
        try:
            cs = [int(counts[r]) for r in range(self.world)]
            ds = [int(displs[r]) for r in range(self.world)]
            total = max(d + c for d, c in zip(ds, cs))
            full = _wrap(ptr, total, self.device)
            if len(set(cs)) == 1 and all(ds[r] == r * cs[0] for r in range(self.world)):
                # equal, contiguous segments: ONE in-place all-gather
                dist.all_gather_into_tensor(full[:cs[0] * self.world], full[ds[self.rank]:ds[self.rank] + cs[0]].clone(),
                                            group=self.group)
            else:
                # ragged: pad to the longest segment, one all-gather, scatter back
                mx = max(cs)
                send = torch.zeros(mx, dtype=torch.float64, device=self.device)
                send[:cs[self.rank]] = full[ds[self.rank]:ds[self.rank] + cs[self.rank]]
                recv = torch.empty(mx * self.world, dtype=torch.float64, device=self.device)
                dist.all_gather_into_tensor(recv, send, group=self.group)
                for r in range(self.world):
                    if r != self.rank and cs[r] > 0:
                        full[ds[r]:ds[r] + cs[r]] = recv[r * mx:r * mx + cs[r]]
            self._sync()
            return 0
        except Exception as e:  # noqa: BLE001
            print("bk_comm.allgatherv failed:", e, flush=True)
            return 1

    def _broadcast(self, user, ptr, n, root):
        try:
            t = _wrap(ptr, n, self.device)
            dist.broadcast(t, src=self._src(root), group=self.group)
            self._sync()
            return 0
        except Exception as e:  # noqa: BLE001
            print("bk_comm.broadcast failed:", e, flush=True)
            return 1

    # ---- host-level helpers used by the Python API ---------------------------------------------------------
    def gather_objects(self, obj):
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def broadcast_object(self, obj, root=0):
        box = [obj]
        dist.broadcast_object_list(box, src=self._src(root), group=self.group)
        return box[0]

    def barrier(self):
        dist.barrier(group=self.group)

    def sub(self, k):
        """Communicator over the first k ranks (collective over ALL ranks of this one); ranks >= k get None.
        This is how `Ncores` selects the number of GPUs a fit uses."""
        k = max(1, min(int(k), self.world))
        if k == self.world:
            return self
        ranks = [self._src(r) for r in range(k)]
        g = dist.new_group(ranks=ranks)
        if self.rank >= k:
            return None
        if k == 1:
            return None if self.rank else _Single()
        return TorchComm(device=self.device, group=g, native=self.peer is not None,
                         ctx=getattr(self, "ctx", None))


class _Single:
    """Degenerate one-rank communicator (Ncores = 1 inside a multi-rank job)."""
    rank, world, peer, struct = 0, 1, None, None

    def gather_objects(self, obj):
        return [obj]

    def broadcast_object(self, obj, root=0):
        return obj

    def barrier(self):
        pass
