"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): the vertex-partitioned run over NCCL
must reproduce the single-GPU iterates."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["GCS_ROOT"])
import utils
from gcs_admm_b200.dist import CudaBackend, DistributedADMM
from gcs_admm_b200.generator import grid_packed_graph
from gcs_admm_b200.partition import partition_vertices, split_graph
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
g = grid_packed_graph(12)
lp = split_graph(g, partition_vertices(g, world), world)[rank]
pf = None
if int(os.environ.get("GCS_PERF", "0")):
    from gcs_admm_b200 import perf
    pf = dict(inner_iters=int(os.environ["GCS_PERF"]), tables=perf.local_tables(perf.perf_tables(g), lp))
if int(os.environ.get("GCS_PEER", "0")):
    from gcs_admm_b200.dist import PeerADMM
    be = drv = PeerADMM(lp, rank, perf=pf)
else:
    be = CudaBackend(lp, rank, perf=pf)
    drv = DistributedADMM(lp, be, graph=bool(int(os.environ.get("GCS_GRAPH", "0"))))
drv.iterate(20)
x_v, z_v, y_v, z_e = be.solution()
rho, pri, dual = be.history()
np.savez(os.path.join(os.environ["GCS_OUT"], f"rank{rank}.npz"), z=z_e, ge=lp.global_edges, pri=pri, dual=dual)
if hasattr(drv, "release_graph"):
    drv.release_graph()
be.close()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("perf_k,graph,peer", [(0, 0, 0), (2, 0, 0), (2, 1, 0), (0, 0, 1), (1, 0, 1)])
def test_two_gpus_match_one(tmp_path, perf_k, graph, peer):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from gcs_admm_b200.generator import grid_packed_graph
    from gcs_admm_b200.lib import Solver
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, GCS_ROOT=ROOT, GCS_OUT=str(tmp_path), GCS_PERF=str(perf_k), GCS_GRAPH=str(graph), GCS_PEER=str(peer))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", str(29533 + perf_k + graph + 7 * peer), str(script)], env=env, timeout=150)
    g = grid_packed_graph(12)
    s = Solver(g)
    if perf_k:
        s.enable_perf(inner_iters=perf_k)
    s.step(20)
    _, _, _, z = s.solution()
    rho, pri, dual = s.history()
    for r in range(2):
        d = np.load(tmp_path / f"rank{r}.npz")
        assert np.max(np.abs(d["z"] - z[d["ge"]])) < 1e-12
        assert np.allclose(d["pri"], pri, rtol=1e-10, atol=0) and np.allclose(d["dual"], dual, rtol=1e-10, atol=0)
