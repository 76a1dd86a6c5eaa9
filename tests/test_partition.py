"""Vertex partitioning + halo exchange (multi-GPU plumbing) on the CPU.

The N>1 path is exercised with a CPU stand-in backend (tests/dist_helpers.py): in-process with a
simulated exchange, and as two real ranks over torch.distributed/gloo."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from gcs_admm_b200.generator import grid_packed_graph
from gcs_admm_b200.graph import pack_graph
from gcs_admm_b200.partition import partition_vertices, split_graph


def _graphs():
    As, bs, n, d, keys = load_golden("benchmark4")
    return {"benchmark4": pack_graph(As, bs), "grid6": grid_packed_graph(6)}


@pytest.mark.parametrize("name", ["benchmark4", "grid6"])
@pytest.mark.parametrize("R", [2, 3])
def test_split_maps(name, R):
    g = _graphs()[name]
    part = partition_vertices(g, R)
    lps = split_graph(g, part, R)
    assert sum(lp.nV for lp in lps) == g.nV
    assert sum(int(lp.edge_counted.sum()) for lp in lps) == g.nE           # every edge accounted exactly once
    assert sum(int(lp.he_off[-1]) for lp in lps) == g.H                     # every half-edge owned exactly once
    for lp in lps:
        assert int(lp.he_off[-1]) + lp.nH_ghost == 2 * lp.nE                # two slots per local edge
        assert lp.send_idx.shape[0] == lp.send_counts.sum() and lp.nH_ghost == lp.recv_counts.sum()
        assert np.all(lp.send_idx < lp.he_off[-1])
        # owned half-edges keep the reference order I_v_in + I_v_out of their vertex
        assert np.array_equal(g.he_edge[lp.global_he], lp.global_edges[lp.he_edge])


def _run_simulated(g, R, its):
    from dist_helpers import OracleBackend
    part = partition_vertices(g, R)
    lps = split_graph(g, part, R)
    bes = [OracleBackend(lp) for lp in lps]
    for _ in range(its):
        for be in bes:
            be.vertex_update()
        # exchange: rank r's send block for q lands in q's ghost range, in (sender rank, order) layout
        for q, (lq, bq) in enumerate(zip(lps, bes)):
            off = int(lq.he_off[-1])
            for r, (lr, br) in enumerate(zip(lps, bes)):
                cnt = int(lq.recv_counts[r])
                if cnt:
                    s0 = int(lr.send_counts[:q].sum())
                    bq.xc[off:off + cnt] = br.xc[torch.as_tensor(lr.send_idx[s0:s0 + cnt])]
                    off += cnt
        for be in bes:
            be.edge_update()
        tot = sum(be.sums.clone() for be in bes)
        for be in bes:
            be.sums[:] = tot
            be.control()
    return lps, bes


@pytest.mark.parametrize("name,R", [("benchmark4", 2), ("grid6", 3)])
def test_partitioned_iterates_equal_single_graph(name, R):
    from c_oracle import COracle
    g = _graphs()[name]
    its = 12
    lps, bes = _run_simulated(g, R, its)
    o = COracle(g)
    o.step(its)
    xc, mu, z = o.state()
    rho, pri, dual = o.history()
    for lp, be in zip(lps, bes):
        assert np.array_equal(be.z, z[lp.global_edges])                    # cut edges identical on both sides, bit for bit
        assert np.array_equal(be.xc.numpy()[:int(lp.he_off[-1])], xc[lp.global_he])
        assert np.allclose(be.mu * be.mu_scale, mu[lp.global_he], rtol=0, atol=1e-13)
        assert np.allclose(be.pri_seq, pri, rtol=1e-12, atol=0) and np.allclose(be.dual_seq, dual, rtol=1e-12, atol=0)
        assert be.rho_seq == list(rho)


def _gloo_worker(rank, world, port, name, its, out):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import utils  # noqa: F401
    import torch.distributed as dist
    from dist_helpers import OracleBackend
    from gcs_admm_b200.dist import DistributedADMM
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = _graphs()[name]
    lp = split_graph(g, partition_vertices(g, world), world)[rank]
    be = OracleBackend(lp)
    drv = DistributedADMM(lp, be)
    drv.run(its, check_every=4)
    np.savez(os.path.join(out, f"rank{rank}.npz"), z=be.z, ge=lp.global_edges, pri=np.array(be.pri_seq), it=be.it)
    dist.destroy_process_group()


def test_two_ranks_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    from c_oracle import COracle
    name, its = "grid6", 8
    port = 29500 + os.getpid() % 1000
    mp.spawn(_gloo_worker, args=(2, port, name, its, str(tmp_path)), nprocs=2, join=True)
    g = _graphs()[name]
    o = COracle(g)
    o.step(its)
    _, _, z = o.state()
    _, pri, _ = o.history()
    for r in range(2):
        d = np.load(tmp_path / f"rank{r}.npz")
        assert int(d["it"]) == its
        assert np.array_equal(d["z"], z[d["ge"]])
        assert np.allclose(d["pri"], pri, rtol=1e-12, atol=0)


def test_peer_send_table_matches_the_ghost_order():
    """peer mode: slot j of what rank r sends to rank q is the ghost slot q keeps for that half-edge"""
    from gcs_admm_b200.dist import peer_send_table
    from gcs_admm_b200.generator import grid_packed_graph
    from gcs_admm_b200.partition import partition_vertices, split_graph
    g = grid_packed_graph(12)
    for R in (2, 3, 4):
        lps = split_graph(g, partition_vertices(g, R), R)
        rc = [lp.recv_counts for lp in lps]
        for q, lq in enumerate(lps):
            nH = int(lq.he_off[-1])
            ghost_gid = -np.ones(lq.nH_ghost, dtype=np.int64)
            for arr, garr in ((lq.edge_he_tail, g.edge_he_tail), (lq.edge_he_head, g.edge_he_head)):
                m = arr >= nH
                ghost_gid[arr[m] - nH] = garr[lq.global_edges[m]]
            for r, lp in enumerate(lps):
                sr, ss = peer_send_table(r, lp.send_counts, rc)
                sel = sr == q
                assert np.array_equal(ghost_gid[ss[sel]], lp.global_he[lp.send_idx][sel])
