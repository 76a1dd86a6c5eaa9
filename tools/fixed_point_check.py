import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import utils
from conftest import load_golden
from gcs_admm_b200.graph import pack_graph
from gcs_admm_b200.lib import Solver
for name in ["benchmark1","benchmark2","benchmark3","benchmark4"]:
    As,bs,n,d,keys = load_golden(name)
    g = pack_graph(As,bs)
    s = Solver(g, max_it=20000, eps_abs=0.0, eps_rel=0.0, check_every=500)
    for k in range(4):
        s.step(5000)
        x_v,z_v,y_v,z_e = s.solution(); st = s.status()
        cost = float(np.sum(np.linalg.norm(z_v[:, :2]-z_v[:, 2:],axis=1)) + 1e-4*np.sum(z_e[:,4]))
        print(name, st['iterations'], 'pri %.2e dual %.2e rho %.3g'%(st['pri_res'], st['dual_res'], st['rho']), 'cost', cost, 'classic', float(d['classic_cost']), 'rel', abs(cost-float(d['classic_cost']))/float(d['classic_cost']))
    s.close()
