#!/bin/bash
# GPU session 8: over-relaxed local-frames consensus + enqueue-ahead timing window (tests, bench), convergence sweep at G = 100
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s8_smoke.log 2>&1
if ! grep -q "smoke ok" gpurun_out/s8_smoke.log; then tail -30 gpurun_out/s8_smoke.log; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/s8_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/s8_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-gate --mode perf --no-cpu-baseline --no-other-mode --residual-budget 0"
timeout 300 $B > gpurun_out/s8_bench.json 2>gpurun_out/s8_bench.err; python -c "
import json
d=json.load(open('gpurun_out/s8_bench.json')); r=d['roofline']
print('it/s %.0f  ms %.4f  k1 %.4f  edge %.4f  e2e %.1f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['edge_kernel']['ms'], d['e2e']['value']))"
T="timeout 60 python tools/time_to_residual.py --grid 100 --max-iters 400000 --trace 4 --budget 30"
for cfg in "--rho0 3" "--rho0 3 --outer-alpha 1.7" "--rho0 10" "--rho0 10 --outer-alpha 1.7" "--rho0 3 --warm dijkstra" "--rho0 3 --warm dijkstra --outer-alpha 1.7" \
           "--window 400000 --adapt-every 100 --outer-alpha 1.7" "--window 400000 --adapt-every 100 --outer-alpha 1.7 --warm dijkstra" \
           "--rho0 3 --theta 0.1" "--rho0 3 --theta 10" "--rho0 3 --inner 2 --outer-alpha 1.7" "--window 400000 --adapt-every 1000 --outer-alpha 1.7" \
           "--rho0 30 --outer-alpha 1.7" "--rho0 1 --outer-alpha 1.7 --warm dijkstra"; do
  echo "== $cfg"; $T $cfg 2>&1 | tail -2
done > gpurun_out/s8_conv_grid100.log 2>&1
cut -c1-260 gpurun_out/s8_conv_grid100.log
