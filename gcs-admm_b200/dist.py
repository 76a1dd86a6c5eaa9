"""One-process-per-GPU driver of the vertex-partitioned ADMM (torch.distributed plumbing).

Per iteration (same order as the reference loop ``admm_solver_v3.py:655-733``):
    K1 on the owned vertices -> halo exchange of the cut half-edges' consensus copies (5 doubles
    each, ``all_to_all_single`` over NCCL/NVLink; gloo on CPU in the tests) -> fused edge kernel on
    all local edges -> all-reduce of the 6 partial sums -> control kernel (identical decision on
    every rank).
The compute backend is an object with ``vertex_update() / edge_update() / control() / xc / sums /
status()``; the product backend is ``CudaBackend`` (libgcsadmm.so).  Tests drive the same class
with a CPU stand-in to exercise the exchange logic under gloo.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["CudaBackend", "DistributedADMM"]


class _DevArray:
    """Wraps a raw device pointer for ``torch.as_tensor`` (CUDA array interface, zero copy)."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class CudaBackend:
    """libgcsadmm.so on the current CUDA device, kernels enqueued on torch's current stream."""

    def __init__(self, lp, device, perf=None, **params):
        """``perf``: None (exact x-update) or dict(inner_iters=K, tables=perf.local_tables(perf_tables(g), lp))."""
        from . import lib
        self.lp = lp
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.solver = lib.Solver(lp, device=device, **params)
        if perf is not None:
            self.solver.enable_perf(inner_iters=perf.get("inner_iters", 3), tables=perf["tables"])
        self.solver.set_stream(torch.cuda.current_stream().cuda_stream)
        nall = lp.he_off[-1] + lp.nH_ghost
        self.xc = torch.as_tensor(_DevArray(self.solver.xc_ptr(), (int(nall), 5)), device=self.device)
        self.sums = torch.as_tensor(_DevArray(self.solver.sums_ptr(), (8,)), device=self.device)

    def vertex_update(self):
        self.solver.vertex_update()

    def edge_update(self):
        self.solver.edge_update()

    def control(self):
        self.solver.control()

    def status(self):
        return self.solver.status()

    def history(self):
        return self.solver.history()

    def solution(self):
        return self.solver.solution()

    def close(self):
        self.solver.close()


class DistributedADMM:
    def __init__(self, lp, backend, group=None):
        self.lp, self.be, self.group = lp, backend, group
        dev = backend.xc.device
        self.send_idx = torch.as_tensor(lp.send_idx, dtype=torch.long, device=dev)
        self.in_splits = [int(x) * 5 for x in lp.send_counts]
        self.out_splits = [int(x) * 5 for x in lp.recv_counts]
        self.nH_own = int(lp.he_off[-1])
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.recv = torch.empty(int(lp.nH_ghost) * 5, dtype=torch.float64, device=dev)

    def exchange(self):
        """Owner -> mirror copy of every cut half-edge's 5 consensus scalars."""
        if self.world == 1:
            return
        xc = self.be.xc
        send = xc.index_select(0, self.send_idx).reshape(-1)
        dist.all_to_all_single(self.recv, send, self.out_splits, self.in_splits, group=self.group)
        xc[self.nH_own:].copy_(self.recv.view(-1, 5))

    def iterate(self, k=1):
        for _ in range(k):
            self.be.vertex_update()
            self.exchange()
            self.be.edge_update()
            if self.world > 1:
                dist.all_reduce(self.be.sums, group=self.group)
            self.be.control()

    def run(self, max_iters, check_every=8):
        done = 0
        st = self.be.status()
        while done < max_iters and not (st["converged"] or st["diverged"]):
            k = min(check_every, max_iters - done)
            self.iterate(k)
            done += k
            st = self.be.status()
        return st
