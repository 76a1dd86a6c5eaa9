"""bench.py for N > 1: the same workload on N GPUs of one box, one process per GPU (strong scaling).
Grids are vertex-partitioned into strips (halo exchange of the cut half-edges + 6-double all-reduce per iteration);
a query batch is sharded rank::world with no communication at all (independent problems: replicas only)."""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def main(args):
    import utils  # noqa: F401
    from .dist import CudaBackend, DistributedADMM, PeerADMM
    from .generator import grid_packed_graph
    from .partition import partition_vertices, split_graph
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    import bench
    W = max(3, args.warmup)
    if args.workload.startswith("batch"):
        return batch_main(args, rank, world, local_rank, W)
    G = bench.WORKLOADS[args.workload]
    burn = max(W, G + 10 if args.burn_in < 0 else args.burn_in)
    g = grid_packed_graph(G)
    lp = split_graph(g, partition_vertices(g, world), world)[rank]
    from . import perf as perf_mod
    tables = perf_mod.local_tables(dict(zip(("cone_off", "cone"), perf_mod.cone_table(g)), kappa=1.0), lp)
    # headline mode: perf only if the parity gate passes in this run (rank 0 runs it on its GPU, everybody follows)
    flag = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", local_rank))
    gate = None
    if args.mode in ("auto", "perf") and not args.no_gate:
        if rank == 0:
            gate = bench.parity_gate()
            flag[0] = 1 if gate["passed"] else 0
        dist.broadcast(flag, 0)
    elif args.mode == "perf":          # --no-gate --mode perf: secondary workloads of a session whose gate has been run already
        flag[0] = 1
    headline = "perf" if int(flag.item()) == 1 else "parity"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=torch.device("cuda", local_rank))

    use_peer = args.dist == "peer"

    def measure(mode, inner):
        """-> (ms per iteration [max over ranks], clocks, e2e seconds [max over ranks], residual history)"""
        pf = dict(inner_iters=inner, tables=tables) if mode == "perf" else None
        kw = dict(max_it=max(1000, args.steps + burn + 8), eps_abs=0.0, eps_rel=0.0)
        if use_peer:
            drv = PeerADMM(lp, local_rank, perf=pf, **kw)
            be = drv
        else:
            be = CudaBackend(lp, local_rank, perf=pf, **kw)
            drv = DistributedADMM(lp, be, graph=args.dist_graph)
        drv.iterate(burn)
        torch.cuda.synchronize()
        sampler = None
        if rank == 0:
            sampler = bench.ClockSampler(local_rank)
            sampler.start()
        dist.barrier()
        torch.cuda.synchronize()
        ms = 0.0
        if use_peer:                          # CUDA events on the library's own stream, L2 flushed in-stream before every timed iteration,
            ms = float(drv.solver.time_window(args.steps, 256 << 20)[0].sum())     # the whole window enqueued ahead
        else:
            ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
            ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
            for i in range(args.steps):
                flush.zero_()                     # L2 eviction, outside the timed pair
                ev0[i].record()
                drv.iterate(1)
                ev1[i].record()
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
        dist.barrier()
        t = torch.tensor([ms], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        clocks = sampler.finish() if sampler else None
        hist = be.history()
        if not use_peer:
            drv.release_graph()
        be.close()
        # end to end: local graph upload + K iterations + local solution download, wall clock, max over ranks
        dist.barrier()
        t0 = time.perf_counter()
        kw2 = dict(max_it=max(1000, n_e2e + 8), eps_abs=0.0, eps_rel=0.0)
        if use_peer:
            drv2 = PeerADMM(lp, local_rank, perf=pf, **kw2)
            be2 = drv2
        else:
            be2 = CudaBackend(lp, local_rank, perf=pf, **kw2)
            drv2 = DistributedADMM(lp, be2, graph=args.dist_graph)
        drv2.iterate(n_e2e)
        be2.solution()
        be2.history()
        e2e = time.perf_counter() - t0
        t = torch.tensor([e2e], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if not use_peer:
            drv2.release_graph()
        be2.close()
        return ms_max / args.steps, clocks, float(t.item()), hist

    n_e2e = burn + args.steps
    per, clocks, e2e, hist = measure(headline, args.inner)
    other_rep = None
    if not args.no_other_mode:
        other = "parity" if headline == "perf" else "perf"
        o_ms, o_clk, o_e2e, _ = measure(other, args.inner)
        other_rep = {"what": "the other mode on the same partition with the same timing protocol", "mode": bench.MODE_TEXT[other],
                     "value": 1e3 / o_ms, "unit": "it/s", "ms_per_step": o_ms, "e2e": n_e2e / o_e2e, "clocks": o_clk}
    ttr = None
    if args.residual_budget > 0 and headline == "perf" and use_peer:
        ttr = time_to_residual(args, g, lp, rank, local_rank)
    check = None
    if rank == 0:
        # correctness of the partitioned run: the same iterations on ONE GPU (rank 0 replays them) give the same residual history
        from . import lib
        s1 = lib.Solver(g, device=local_rank, max_it=max(1000, args.steps + burn + 8), eps_abs=0.0, eps_rel=0.0)
        if headline == "perf":
            s1.enable_perf(inner_iters=args.inner, tables=perf_mod.perf_tables(g))
        s1.step(burn + args.steps)
        r1, p1, d1 = s1.history()
        s1.close()
        n = min(len(p1), len(hist[1]))
        scale = max(1.0, float(np.max(np.abs(p1[:n]))))
        check = {"iterations_compared": n - 1, "max_abs_diff_pri": float(np.max(np.abs(p1[:n] - hist[1][:n]))),
                 "max_abs_diff_dual": float(np.max(np.abs(d1[:n] - hist[2][:n]))), "max_abs_diff_rho": float(np.max(np.abs(r1[:n] - hist[0][:n]))),
                 "scale": scale, "what": "residual history of the N-GPU run vs a 1-GPU replay of the same iterations on rank 0"}
    if rank == 0:
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
                "config": {"workload": f"grid{G}x{G} 2-D GCS: {g.nV} vertices, {g.nE} directed edges, strips over {world} GPUs",
                           "mode": bench.MODE_TEXT[headline] + (f", K={args.inner}" if headline == "perf" else ""), "l2": "flushed (256 MiB memset in-stream before every timed iteration, outside its event pair; the window is enqueued ahead)", "burn_in_iterations": burn,
                           "halo_half_edges_rank0": int(lp.nH_ghost),
                           "exchange": ("peer memory: cut half-edges and residual sums stored straight into the neighbours' buffers over NVLink, flag-ordered; "
                                        "no collective call inside the iteration; 3 launches per iteration (K1 + halo push by its last block | halo wait + edges + sums | sums wait + control)") if use_peer else
                                       "all_to_all_single(halo) + all_reduce(8 doubles) per iteration, NCCL" + (", one CUDA graph per iteration" if args.dist_graph else "")},
                "clocks": clocks,
                "e2e": {"value": n_e2e / e2e, "unit": bench.UNIT, "h2d_bytes_per_step": gs_bytes / n_e2e,
                        "d2h_bytes_per_step": out_bytes / n_e2e, "note": "per rank, from a cold start: local graph upload + (burn_in + K) iterations + solution download; max over ranks"},
                "gpu_launches": ((3 if headline == "perf" else 4) if use_peer else 3) * args.steps,
                "roofline": {"bound": "hbm", "achieved": (k1b + k2b) / world / (per * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": (k1b + k2b) / world / (per * 1e-3) / 1e9 / peak, "traffic": None,
                             "note": "whole iteration, algorithmic bytes per GPU / max-over-ranks time"}}
        if gate is not None:
            line["parity_gate"] = gate
        if other_rep is not None:
            line[("parity" if headline == "perf" else "perf") + "_mode"] = other_rep
        line["consistency_vs_1gpu"] = check
        if ttr is not None:
            line["time_to_residual_1e-4"] = ttr
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()


def time_to_residual(args, g, lp, rank, local_rank, tol=1e-4):
    """BASELINE metric, second half, on N GPUs: the accelerated perf-mode configuration of bench.TTR (local frames, rho0 = 3,
    over-relaxed consensus step, duals started from the portal cost-to-go field) on the strips, peer-memory exchange, until
    max(pri, dual, inner) < tol on every rank (all ranks take the same control decisions) or --residual-cap iterations."""
    import bench
    from . import perf as perf_mod, warmstart
    from .dist import PeerADMM
    cfg = bench.TTR
    t_all = time.perf_counter()
    Tg = perf_mod.perf_tables(g, frames=cfg["frames"])
    tl = perf_mod.local_tables(Tg, lp)
    mu = warmstart.dual_start(g, Tg["edge_delta"], cfg["rho0"], field=cfg["warm"])[np.asarray(lp.global_he, dtype=np.int64)]
    t_host = time.perf_counter() - t_all
    cap = int(args.residual_cap)
    drv = PeerADMM(lp, local_rank, perf=dict(inner_iters=cfg["inner"], tables=tl), max_it=cap + 8, abs_stop=1, abs_tol=tol, check_every=256,
                   frac=cfg["window"] / cap, outer_alpha=cfg["outer_alpha"], rho0=cfg["rho0"])
    drv.solver.set_state(None, mu, None, cfg["rho0"], 0)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = drv.run(cap)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=torch.device("cuda", local_rank))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    x_v, z_v, y_v, z_e = drv.solution()
    part = torch.tensor([float(np.sum(np.linalg.norm(z_v[:, :2] - z_v[:, 2:], axis=1))),
                         1e-4 * float(np.sum(z_e[np.asarray(lp.edge_counted, dtype=bool), 4]))], dtype=torch.float64, device=torch.device("cuda", local_rank))
    dist.all_reduce(part)
    drv.close()
    return {"reached": bool(st["converged"]), "seconds": float(t.item()), "iterations": int(st["iterations"]), "pri_res": st["pri_res"], "dual_res": st["dual_res"],
            "inner_res": st["inner_res"], "pri_res_reference_definition": st["pri_res_ref"], "dual_res_reference_definition": st["dual_res_ref"],
            "rho": st["rho"], "tolerance": tol, "iteration_cap": cap, "outer_alpha": cfg["outer_alpha"],
            "relaxed_cost": float(part[0].item() + part[1].item()), "host_setup_seconds_rank0": t_host,
            "mode": f"perf K={cfg['inner']}, local coordinate frames, rho0 = {cfg['rho0']}, over-relaxed consensus step, duals started from the portal-graph "
                    f"cost-to-go field ({cfg['warm']}); strips over the GPUs, peer-memory exchange; max over ranks of the wall time of the run"}


def batch_main(args, rank, world, local_rank, W):
    """BASELINE config 4: independent queries sharded rank::world — no data-path collective (replicas only).  A step = one ADMM
    iteration of every query of the batch; value = batch iterations/s = 1 / (max over ranks of the time per iteration)."""
    import bench
    from . import lib
    g, desc = bench.build_workload(args.workload, rank, world)
    dev = torch.device("cuda", local_rank)
    gate, flag = None, torch.zeros(1, dtype=torch.int32, device=dev)
    if args.mode in ("auto", "perf") and not args.no_gate:
        if rank == 0:
            gate = bench.parity_gate()
            flag[0] = 1 if gate["passed"] else 0
        dist.broadcast(flag, 0)
    elif args.mode == "perf":
        flag[0] = 1
    headline = "perf" if int(flag.item()) == 1 else "parity"
    burn = max(W, 100 if args.burn_in < 0 else args.burn_in)
    s = lib.Solver(g, device=local_rank, max_it=max(1000, burn + args.steps + 8), eps_abs=0.0, eps_rel=0.0)
    if headline == "perf":
        s.enable_perf(inner_iters=args.inner)
    s.step(burn)
    sampler = None
    if rank == 0:
        sampler = bench.ClockSampler(local_rank)
        sampler.start()
    dist.barrier()
    tot = 0.0
    tot = float(s.time_window(args.steps, 256 << 20)[0].sum())
    dist.barrier()
    t = torch.tensor([tot], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per = float(t.item()) / args.steps
    clocks = sampler.finish() if sampler else None
    state = list(s.state()) + (list(s.perf_state()) if headline == "perf" else [])
    s.close()
    # end to end on every rank: upload of its share + warm state, K iterations, download; max over ranks
    dist.barrier()
    t0 = time.perf_counter()
    s2 = lib.Solver(g, device=local_rank, max_it=max(1000, state[4] + args.steps + 8), check_every=max(1, min(64, args.steps)), eps_abs=0.0, eps_rel=0.0)
    if headline == "perf":
        s2.enable_perf(inner_iters=args.inner)
    s2.set_state(state[0], state[1], state[2], state[3], state[4])
    if headline == "perf":
        s2.set_perf_state(state[5], state[6])
    s2.run(args.steps)
    s2.solution()
    e2e = time.perf_counter() - t0
    s2.close()
    t = torch.tensor([e2e], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nq = torch.tensor([g.prob_voff.shape[0] - 1], dtype=torch.float64, device=dev)
    dist.all_reduce(nq)
    if rank == 0:
        line = {"metric": bench.METRIC, "value": 1e3 / per, "unit": bench.UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": per, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc.split(":")[0] + f", {int(nq.item())} queries sharded rank::world over {world} GPUs, no communication",
                           "mode": bench.MODE_TEXT[headline] + (f", K={args.inner}" if headline == "perf" else ""), "l2": "flushed (256 MiB) before every timed iteration",
                           "burn_in_iterations": burn},
                "problem_iterations_per_second": nq.item() * 1e3 / per, "clocks": clocks,
                "e2e": {"value": args.steps / float(t.item()), "unit": bench.UNIT, "seconds": float(t.item()),
                        "note": "per rank: upload of its queries + warm state, K iterations, solution download; max over ranks"},
                "gpu_launches": 2 * args.steps}
        if gate is not None:
            line["parity_gate"] = gate
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
