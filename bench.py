"""bench.py — ADMM iterations/s of the full-vertex-split iteration on the 100k-vertex 2-D grid GCS.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--grid G] [--impl reference]

A "step" is one ADMM iteration (K1 vertex programs + fused edge/dual/residual kernel + control)
over the whole graph.  `value` = iterations/s with the graph resident in HBM (CUDA events on the
library's stream, L2 flushed between timed iterations); `e2e` = the same metric through the
host-buffer C-ABI call gcsadmm_solve_host (graph upload + K iterations + solution download inside
the timed region).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "admm_iterations_per_second"
UNIT = "it/s"


def algorithmic_bytes(g):
    """compulsory fp64 traffic of one ADMM iteration.
    K1: targets z (gather, H*5) + mu (H*5) read, xc (H*5) written, polytopes, CSR indices.
    K2-4: xc read (H*5), z read+written (2*E*5), mu read+written (2*H*5), edge->half-edge indices."""
    E, H, V = g.nE, 2 * g.nE, g.nV
    sum_m = int(g.poly_off[-1])
    k1 = 8 * (3 * 5 * H + 3 * sum_m) + 4 * (V + 1) + 4 * H + H + V + 16 * V
    k2 = 8 * (3 * 5 * H + 2 * 5 * E) + 8 * E
    return k1, k2


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons}


def cpu_baseline(G_sample, V_full, iters=10):
    """C oracle (port of the reference's algorithm, OpenMP over vertices) on a bounded sample:
    a G_sample x G_sample grid of the same family; cost per iteration is linear in |V|, so the
    figure is scaled to the full graph's vertex count."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import utils  # noqa: F401
    from c_oracle import COracle, lib as olib
    from gcs_admm_b200.generator import grid_packed_graph
    g = grid_packed_graph(G_sample)
    o = COracle(g)
    o.step(G_sample + 10)           # same burn-in rule as the GPU arm: the cold-start wave has reached every vertex
    t0 = time.perf_counter()
    o.step(iters)
    dt = (time.perf_counter() - t0) / iters
    its_sample = 1.0 / dt
    return {"value": its_sample * g.nV / V_full, "unit": UNIT, "cores": olib().gcso_num_threads(), "kind": "port",
            "sample": f"{iters} ADMM iterations (after a burn-in of {G_sample + 10}) of the C oracle on the {G_sample}x{G_sample} grid ({g.nV} vertices, "
                      f"{its_sample:.3f} it/s), scaled by |V| to the {V_full}-vertex workload"}


def run_reference(args):
    """--impl reference: the reference's own algorithm on the host cores.  The reference itself
    (pydrake + MOSEK) is not installable offline, so the oracle port stands in (kind='port')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import utils  # noqa: F401
    from c_oracle import COracle, lib as olib
    from gcs_admm_b200.generator import grid_packed_graph
    V_full = args.grid * args.grid + 2
    Gs = min(args.grid, 48)
    g = grid_packed_graph(Gs)
    o = COracle(g)
    o.step(Gs + 10)                 # same burn-in rule as the GPU arm (every vertex program live), untimed
    for _ in range(args.warmup):
        o.step(1)
    t0 = time.perf_counter()
    o.step(args.steps)
    dt = time.perf_counter() - t0
    val = args.steps / dt * g.nV / V_full
    sample = (f"each step = one ADMM iteration of the C oracle (after a burn-in of {Gs + 10}) on the {Gs}x{Gs} grid ({g.nV} vertices), "
              f"scaled by |V| to the {V_full}-vertex workload")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps * V_full / g.nV, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"grid{args.grid}x{args.grid} 2-D GCS ({V_full} vertices)", "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": olib().gcso_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_classic:
        # the reference's other CPU solver (classic_solver.py: one monolithic conic program), Drake-free restatement, on a bounded sample
        from gcs_admm_b200.classic import solve_classic
        from gcs_admm_b200.generator import grid_problem, packed_to_dicts
        off, A, b, _, _ = grid_problem(8)
        As, bs = packed_to_dicts(off, A, b)
        t0 = time.perf_counter()
        rc = solve_classic(As, bs, 2, round_solution=False)
        line["classic_solver"] = {"workload": f"grid8x8 ({len(As)} vertices)", "seconds": time.perf_counter() - t0, "status": rc["status"],
                                  "ip_iterations": rc["iterations"], "cost": rc["cost"], "kind": "port (gcs_admm_b200.classic, sparse interior point, 1 thread)",
                                  "note": "whole solve to optimality, not an iteration rate; 258 vertices take ~90 s, so it is not run at the metric's size"}
    print(json.dumps(line))


MODE_TEXT = {"parity": "parity (every vertex program solved to 1e-8 by the interior-point kernel)",
             "perf": "perf (inexact x-update: warm-started splitting iterations, gcsadmm_enable_perf)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--grid", type=int, default=316, help="G: the workload is the G x G grid GCS (316 -> 99 858 vertices)")
    ap.add_argument("--impl", type=str, default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-classic", action="store_true", help="reference arm: skip the classic-solver sample")
    ap.add_argument("--burn-in", type=int, default=-1, help="untimed iterations before the timed window (default: grid side + 10, so that the\n                    cold-start wave has reached every vertex and no vertex program is the trivial all-zero one)")
    ap.add_argument("--mode", type=str, default="parity", choices=["parity", "perf"],
                    help="parity: exact interior-point x-update (reference trajectory); perf: K closed-form splitting iterations per x-update")
    ap.add_argument("--inner", type=int, default=3, help="K of the perf mode")
    ap.add_argument("--no-perf-report", dest="perf_report", action="store_false",
                    help="parity runs also time the perf mode on the same graph and report it as a nested object; this switches that off")
    ap.add_argument("--perf-trace-iters", type=int, default=20000, help="perf report: residuals and wall time after this many K=1 iterations from a cold start")
    ap.add_argument("--dist-graph", action="store_true", help="N > 1: replay one captured CUDA graph per ADMM iteration (kernels + NCCL collectives)")
    ap.add_argument("--residual-run", type=int, default=0, help="also run up to this many iterations with the abs 1e-4 stop and report the time")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or args.gpus > 1:
        from gcs_admm_b200 import dist_bench  # multi-GPU path (vertex-partitioned, NCCL halo exchange)
        import gcs_admm_b200  # noqa: F401
        return dist_bench.main(args)

    import utils  # noqa: F401
    from gcs_admm_b200 import lib
    from gcs_admm_b200.generator import grid_packed_graph
    W = max(3, args.warmup)
    burn = max(W, args.grid + 10 if args.burn_in < 0 else args.burn_in)
    g = grid_packed_graph(args.grid)
    k1_bytes, k2_bytes = algorithmic_bytes(g)
    tables = None
    if args.mode == "perf" or args.perf_report:
        from gcs_admm_b200 import perf as perf_mod
        tables = perf_mod.perf_tables(g)

    def timed(mode, inner):
        """burn-in, then `steps` iterations timed one by one with CUDA events on the solver's stream, L2 flushed before each"""
        s = lib.Solver(g, device=0, max_it=max(1000, args.steps + burn + 8), eps_abs=0.0, eps_rel=0.0)
        if mode == "perf":
            s.enable_perf(inner_iters=inner, tables=tables)
        s.step(burn)
        st0 = s.status()
        sampler = ClockSampler(0)
        sampler.start()
        tot = k1 = ed = 0.0
        for _ in range(args.steps):          # L2 flushed before every timed iteration, outside the event pair
            s.flush_l2()
            a, b, c = s.time_steps(1, split=True)
            tot += a; k1 += b; ed += c
        clocks = sampler.finish()
        st = s.status()
        s.close()
        return dict(ms=tot / args.steps, k1_ms=k1 / args.steps, ed_ms=ed / args.steps, clocks=clocks, st0=st0, st=st)

    def end_to_end(mode, inner, n_e2e):
        """host buffers -> host buffers through the C-ABI, wall clock: graph (and table) upload + n iterations + download"""
        t0 = time.perf_counter()
        if mode == "perf":          # same sequence as gcsadmm_solve_host, plus the table upload
            s3 = lib.Solver(g, device=0, max_it=max(1000, n_e2e + 8), check_every=64, eps_abs=0.0, eps_rel=0.0).enable_perf(inner_iters=inner, tables=tables)
            s3.run(n_e2e)
            s3.solution(); s3.history()
            out = {"status": s3.status()}
            s3.close()
        else:
            out = lib.solve_host(g, device=0, max_iters=n_e2e, max_it=max(1000, n_e2e + 8), check_every=64,
                                 eps_abs=0.0, eps_rel=0.0)
        dt = time.perf_counter() - t0
        assert out["status"]["iterations"] == n_e2e
        return dt

    m = timed(args.mode, args.inner)
    ms, k1, ed, clocks, st0, st = m["ms"], m["k1_ms"] * args.steps, m["ed_ms"] * args.steps, m["clocks"], m["st0"], m["st"]
    value = 1e3 / ms
    gs_bytes = sum(a.nbytes for a in (g.poly_off, g.polyA, g.polyb, g.he_off, g.he_edge, g.he_flags, g.edge_he_tail,
                                      g.edge_he_head, g.vtype)) + 16 * g.nV
    out_bytes = 8 * (9 * g.nV + 5 * g.nE) + 3 * 8 * (burn + args.steps + 1)
    n_e2e = burn + args.steps
    e2e_s = end_to_end(args.mode, args.inner, n_e2e)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    k1_ms = m["k1_ms"]
    traffic = None          # dram__bytes_read.sum + dram__bytes_write.sum of one K1 launch, from the committed ncu capture
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_k1_traffic.json" if args.mode == "parity" else "r01_k1perf_traffic.json")))
        if tj.get("workload") == f"grid{args.grid}x{args.grid}" and (args.mode == "parity" or tj.get("inner_iters") == args.inner):
            traffic = tj["traffic_bytes_per_launch"]
    except Exception:
        pass
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": W,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"grid{args.grid}x{args.grid} 2-D GCS: {g.nV} vertices, {g.nE} directed edges, m=8 rows/region "
                               "(BASELINE.json metric config: 100k-vertex 2-D GCS)", "mode": MODE_TEXT[args.mode] + (f", K={args.inner}" if args.mode == "perf" else ""),
                   "l2": "flushed (256 MiB memset) before every timed iteration", "burn_in_iterations": burn, "inner_ipm_iters_per_vertex": (st["inner_iters"] - st0["inner_iters"]) / max(1, args.steps * g.nV),
                   "inner_tol": 1e-8, "warm_start_theta": 1e-3, "zero_tol": 1e-12,
                   "vertex_programs_skipped_as_zero_frac": (st["skipped"] - st0["skipped"]) / max(1, args.steps * g.nV)},
        "clocks": clocks,
        "e2e": {"value": n_e2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": gs_bytes / n_e2e, "d2h_bytes_per_step": out_bytes / n_e2e,
                "note": "gcsadmm_solve_host from a cold start: graph upload + (burn_in + K) iterations + solution/history download, wall clock; value = (burn_in + K) / time"},
        "gpu_launches": 4 * args.steps,
        "roofline": {"bound": "hbm", "kernel": "vertex_kernel (K1)" if args.mode == "parity" else "vertex_perf_kernel (K1, perf mode)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                     "traffic_note": "ncu --set full capture (profiles/r01_k1_grid316_ncu_summary.txt); includes the 2.7 KB/vertex warm-start records K1 reads and rewrites" if args.mode == "parity" else None,
                     "algorithmic_bytes_per_launch": k1_bytes, "kernel_ms": k1_ms,
                     "whole_iteration": {"bytes": k1_bytes + k2_bytes, "achieved": (k1_bytes + k2_bytes) / (ms * 1e-3) / 1e9,
                                         "frac": (k1_bytes + k2_bytes) / (ms * 1e-3) / 1e9 / peak},
                     "edge_kernel": {"bytes": k2_bytes, "ms": ed / args.steps, "achieved": k2_bytes / (ed / args.steps * 1e-3) / 1e9}},
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(min(args.grid, 48), g.nV)
    if args.mode == "parity" and args.perf_report:
        rep = {"what": "same graph, same timing protocol, x-update = K warm-started splitting iterations per ADMM iteration (gcsadmm_enable_perf) instead of an exact "
                       "interior-point solve; same fixed point (tests/test_gpu_perf.py), not the reference's trajectory, hence not the headline value"}
        for K in (3, 1):
            pm = timed("perf", K)
            pe = end_to_end("perf", K, n_e2e)
            rep[f"K={K}"] = {"value": 1e3 / pm["ms"], "unit": UNIT, "ms_per_step": pm["ms"], "kernel_ms": pm["k1_ms"], "edge_ms": pm["ed_ms"],
                             "roofline_frac_k1": k1_bytes / (pm["k1_ms"] * 1e-3) / 1e9 / peak,
                             "roofline_frac_iteration": (k1_bytes + k2_bytes) / (pm["ms"] * 1e-3) / 1e9 / peak,
                             "e2e": n_e2e / pe, "clocks": pm["clocks"]}
        if args.perf_trace_iters > 0:
            s4 = lib.Solver(g, device=0, max_it=10 * args.perf_trace_iters, abs_stop=1, abs_tol=1e-4, check_every=64).enable_perf(inner_iters=1, tables=tables)
            t0 = time.perf_counter()
            st4 = s4.run(args.perf_trace_iters)
            rep["residual_trace_K=1"] = {"iterations": st4["iterations"], "seconds": time.perf_counter() - t0, "pri_res": st4["pri_res"], "dual_res": st4["dual_res"],
                                         "rho": st4["rho"], "reached_1e-4": bool(st4["converged"])}
            s4.close()
        line["perf_mode"] = rep
    if args.residual_run:
        s2 = lib.Solver(g, device=0, max_it=args.residual_run, abs_stop=1, abs_tol=1e-4, check_every=64)
        t0 = time.perf_counter()
        st2 = s2.run(args.residual_run)
        line["time_to_residual_1e-4"] = {"seconds": time.perf_counter() - t0, "iterations": st2["iterations"], "reached": bool(st2["converged"]),
                                         "pri_res": st2["pri_res"], "dual_res": st2["dual_res"]}
        s2.close()
    print(json.dumps(line))


if __name__ == "__main__":
    main()
