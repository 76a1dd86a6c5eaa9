#!/bin/bash
# GPU session 10: inner residual in the stop rule (tests), honest convergence sweep at G = 100, edge-kernel occupancy variants
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s10_smoke.log 2>&1
if ! grep -q "smoke ok" gpurun_out/s10_smoke.log; then tail -30 gpurun_out/s10_smoke.log; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/s10_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/s10_pytest.log | cut -c1-300
B="python bench.py --steps 20 --warmup 3 --no-gate --mode perf --no-cpu-baseline --no-other-mode --residual-budget 0"
run() { echo "== $1"; env $1 timeout 300 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('it/s %.0f  ms %.4f  k1 %.4f  edge %.4f  e2e %.1f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['edge_kernel']['ms'], d['e2e']['value']))"; }
run "GCS_EDGE_MINB=2"
run "GCS_EDGE_MINB=3 GCS_EDGE_BLOCKS_PER_SM=3"
run "GCS_EDGE_MINB=3 GCS_EDGE_BLOCKS_PER_SM=6"
run "GCS_EDGE_MINB=4 GCS_EDGE_BLOCKS_PER_SM=4"
run "GCS_EDGE_MINB=4 GCS_EDGE_BLOCKS_PER_SM=8"
run "GCS_EDGE_MINB=2 GCS_EDGE_BLOCKS_PER_SM=2"
T="timeout 80 python tools/time_to_residual.py --grid 100 --max-iters 600000 --trace 3 --budget 40"
for cfg in "--rho0 3" "--rho0 3 --outer-alpha 1.7" "--rho0 3 --warm dijkstra" "--rho0 3 --warm dijkstra --outer-alpha 1.7" \
           "--rho0 3 --warm dijkstra --outer-alpha 1.7 --inner 2" "--rho0 3 --warm dijkstra --outer-alpha 1.7 --inner 3" "--rho0 1 --warm dijkstra --outer-alpha 1.7" \
           "--rho0 10 --warm dijkstra --outer-alpha 1.7" "--rho0 3 --warm euclid --outer-alpha 1.7"; do
  echo "== $cfg"; $T $cfg 2>&1 | tail -3
done > gpurun_out/s10_conv_grid100.log 2>&1
cut -c1-330 gpurun_out/s10_conv_grid100.log
