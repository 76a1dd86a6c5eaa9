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

__all__ = ["CudaBackend", "DistributedADMM", "PeerADMM", "peer_send_table"]


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
        self.sums = torch.as_tensor(_DevArray(self.solver.sums_ptr(), (10,)), device=self.device)

    supports_graph = True

    def use_stream(self, stream):
        """enqueue the library's kernels on this torch stream from now on (synchronises the previous one)"""
        self.solver.set_stream(stream.cuda_stream)

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
    def __init__(self, lp, backend, group=None, graph=False):
        """``graph``: after two eager iterations (NCCL warmed up) one whole iteration — K1, halo pack, all_to_all, unpack,
        K2-K4, all_reduce, K5 — is captured into a CUDA graph and replayed: one launch per ADMM iteration instead of ~10
        launches + 2 collective calls from Python (0.2 ms -> tens of microseconds of per-iteration overhead).  Call
        ``release_graph()`` before ``destroy_process_group()``: NCCL cannot tear a communicator down while a graph that
        captured its collectives is alive."""
        self.lp, self.be, self.group = lp, backend, group
        self._graph, self._eager_done = None, 0
        self._want_graph = bool(graph) and getattr(backend, "supports_graph", False)
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

    def _iterate_eager(self, k):
        for _ in range(k):
            self.be.vertex_update()
            self.exchange()
            self.be.edge_update()
            if self.world > 1:
                dist.all_reduce(self.be.sums, group=self.group)
            self.be.control()

    def _capture(self):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        self.be.use_stream(side)                 # before the capture starts: set_stream synchronises
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self._iterate_eager(1)
        cur.wait_stream(side)
        self.be.use_stream(cur)
        self._graph = g

    def release_graph(self):
        self._graph = None
        self._want_graph = False

    def iterate(self, k=1):
        if self._want_graph and self.world > 1:
            while k > 0 and self._eager_done < 2:
                self._iterate_eager(1)
                self._eager_done += 1
                k -= 1
            if k > 0 and self._graph is None:
                self._capture()
            for _ in range(k):
                self._graph.replay()
            return
        self._iterate_eager(k)

    def run(self, max_iters, check_every=8):
        done = 0
        st = self.be.status()
        while done < max_iters and not (st["converged"] or st["diverged"]):
            k = min(check_every, max_iters - done)
            self.iterate(k)
            done += k
            st = self.be.status()
        return st


def peer_send_table(rank, send_counts, recv_counts_all):
    """(send_rank, send_slot) of a rank's halo items, in the order of ``LocalProblem.send_idx`` (grouped by destination rank).
    ``recv_counts_all[q][p]`` = ghost slots rank q keeps for rank p; q's ghost range is ordered by source rank, and inside one
    source by global half-edge id — the order the source's send list has as well (``partition.split_graph``)."""
    send_rank, send_slot = [], []
    for q, n in enumerate(send_counts):
        base = int(sum(recv_counts_all[q][:rank]))
        send_rank += [q] * int(n)
        send_slot += list(range(base, base + int(n)))
    return np.asarray(send_rank, np.int32), np.asarray(send_slot, np.int32)


class PeerADMM:
    """The partitioned iteration with NO collective call inside it (``gcsadmm_peer_connect``): halos and residual sums travel as
    peer-memory stores over NVLink, ordered by flags; every rank replays its own CUDA graph.  ``torch.distributed`` is used once,
    to exchange the CUDA IPC handles and the partition sizes."""

    def __init__(self, lp, device, perf=None, group=None, **params):
        from . import lib
        self.lp, self.group = lp, group
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        self.solver = lib.Solver(lp, device=device, **params)
        if perf is not None:
            self.solver.enable_perf(inner_iters=perf.get("inner_iters", 1), tables=perf["tables"])
        mine = torch.from_numpy(self.solver.peer_export()).to(self.device)
        handles = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(handles, mine, group=group)
        meta = torch.tensor([int(lp.he_off[-1]), int(lp.nH_ghost)] + [int(x) for x in lp.recv_counts], dtype=torch.int64, device=self.device)
        metas = [torch.empty_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta, group=group)
        metas = [m.cpu().numpy() for m in metas]
        send_rank, send_slot = peer_send_table(rank, lp.send_counts, [m[2:] for m in metas])
        self.solver.peer_connect(rank, world, np.stack([h.cpu().numpy() for h in handles]), [m[0] for m in metas], [m[1] for m in metas],
                                 np.asarray(lp.send_idx, np.int32), send_rank, send_slot)
        dist.barrier(group=group)                 # nobody iterates before everybody is connected

    def iterate(self, k=1):
        self.solver.step(k)

    def run(self, max_iters):
        st = self.solver.run(max_iters)
        if self.solver.peer_error():
            raise RuntimeError("peer wait timed out: a rank fell out of step")
        return st

    def status(self):
        return self.solver.status()

    def history(self):
        return self.solver.history()

    def solution(self):
        return self.solver.solution()

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)            # peers may still be storing into this rank's buffers
        self.solver.close()
