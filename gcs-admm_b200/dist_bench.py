"""bench.py for N > 1: the same 100k-vertex grid, vertex-partitioned into strips across N GPUs
(strong scaling), one process per GPU, halo exchange + 6-double all-reduce over NCCL."""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def main(args):
    import utils  # noqa: F401
    from .dist import CudaBackend, DistributedADMM
    from .generator import grid_packed_graph
    from .partition import partition_vertices, split_graph
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    W = max(3, args.warmup)
    burn = max(W, args.grid + 10 if args.burn_in < 0 else args.burn_in)
    g = grid_packed_graph(args.grid)
    lp = split_graph(g, partition_vertices(g, world), world)[rank]
    tables = None
    if args.mode == "perf" or args.perf_report:
        from . import perf as perf_mod
        tables = perf_mod.local_tables(perf_mod.perf_tables(g), lp)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=torch.device("cuda", local_rank))

    def measure(mode, inner):
        """-> (ms per iteration [max over ranks], clocks, e2e seconds [max over ranks])"""
        pf = dict(inner_iters=inner, tables=tables) if mode == "perf" else None
        be = CudaBackend(lp, local_rank, perf=pf, max_it=max(1000, args.steps + burn + 8), eps_abs=0.0, eps_rel=0.0)
        drv = DistributedADMM(lp, be, graph=args.dist_graph)
        drv.iterate(burn)
        torch.cuda.synchronize()
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        sampler = None
        if rank == 0:
            import bench
            sampler = bench.ClockSampler(local_rank)
            sampler.start()
        dist.barrier()
        torch.cuda.synchronize()
        for i in range(args.steps):
            flush.zero_()                     # L2 eviction, outside the timed pair
            ev0[i].record()
            drv.iterate(1)
            ev1[i].record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
        t = torch.tensor([ms], dtype=torch.float64, device=be.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        clocks = sampler.finish() if sampler else None
        drv.release_graph()
        be.close()
        # end to end: local graph upload + K iterations + local solution download, wall clock, max over ranks
        dist.barrier()
        t0 = time.perf_counter()
        be2 = CudaBackend(lp, local_rank, perf=pf, max_it=max(1000, n_e2e + 8), eps_abs=0.0, eps_rel=0.0)
        drv2 = DistributedADMM(lp, be2, graph=args.dist_graph)
        drv2.iterate(n_e2e)
        be2.solution()
        be2.history()
        e2e = time.perf_counter() - t0
        t = torch.tensor([e2e], dtype=torch.float64, device=be2.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        drv2.release_graph()
        be2.close()
        return ms_max / args.steps, clocks, float(t.item())

    n_e2e = burn + args.steps
    per, clocks, e2e = measure(args.mode, args.inner)
    perf_rep = None
    if args.mode == "parity" and args.perf_report:
        perf_rep = {"what": "same partition and timing protocol with the inexact x-update (gcsadmm_enable_perf); see the 1-GPU line / DESIGN.md 5a"}
        for K in (3, 1):
            p_ms, p_clk, p_e2e = measure("perf", K)
            perf_rep[f"K={K}"] = {"value": 1e3 / p_ms, "unit": "it/s", "ms_per_step": p_ms, "e2e": n_e2e / p_e2e, "clocks": p_clk}
    if rank == 0:
        import bench
        k1b, k2b = bench.algorithmic_bytes(g)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        gs_bytes = sum(np.asarray(a).nbytes for a in (lp.poly_off, lp.polyA, lp.polyb, lp.he_off, lp.he_edge, lp.he_flags,
                                                      lp.edge_he_tail, lp.edge_he_head, lp.vtype, lp.cent))
        out_bytes = 8 * (9 * lp.nV + 5 * lp.nE)
        line = {"metric": bench.METRIC, "value": 1e3 / per, "unit": bench.UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": per, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"grid{args.grid}x{args.grid} 2-D GCS: {g.nV} vertices, {g.nE} directed edges, strips over {world} GPUs",
                           "mode": bench.MODE_TEXT[args.mode] + (f", K={args.inner}" if args.mode == "perf" else ""), "l2": "flushed (256 MiB) before every timed iteration", "burn_in_iterations": burn,
                           "halo_half_edges_rank0": int(lp.nH_ghost), "collectives": "all_to_all_single(halo) + all_reduce(8 doubles) per iteration, NCCL",
                           "cuda_graph": bool(args.dist_graph)},
                "clocks": clocks,
                "e2e": {"value": n_e2e / e2e, "unit": bench.UNIT, "h2d_bytes_per_step": gs_bytes / n_e2e,
                        "d2h_bytes_per_step": out_bytes / n_e2e, "note": "per rank, from a cold start: local graph upload + (burn_in + K) iterations + solution download; max over ranks"},
                "gpu_launches": 4 * args.steps,
                "roofline": {"bound": "hbm", "achieved": (k1b + k2b) / world / (per * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": (k1b + k2b) / world / (per * 1e-3) / 1e9 / peak, "traffic": None,
                             "note": "whole iteration, algorithmic bytes per GPU / max-over-ranks time"}}
        if perf_rep is not None:
            line["perf_mode"] = perf_rep
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
