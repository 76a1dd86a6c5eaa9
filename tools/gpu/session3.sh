#!/bin/bash
# GPU session 3: local frames (tests + convergence at G=100 / G=316), edge-kernel grid sweep, ncu of the current K1
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke.log 2>&1
if ! grep -q "smoke ok" gpurun_out/s3_smoke.log; then tail -30 gpurun_out/s3_smoke.log; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/s3_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/s3_pytest.log
T="timeout 400 python tools/time_to_residual.py --trace 8"
for cfg in "--grid 100 --frames local" "--grid 100 --frames local --rho0 0.3" "--grid 100 --frames local --rho0 3" "--grid 100 --frames local --outer-alpha 1.6" \
           "--grid 100 --frames local --window 400000 --adapt-every 100" "--grid 316 --frames local --max-iters 1500000 --budget 300"; do
  echo "== $cfg"; $T $cfg 2>&1 | tail -4
done > gpurun_out/s3_conv.log 2>&1
cat gpurun_out/s3_conv.log | cut -c1-330
B="python bench.py --steps 4 --warmup 3 --no-gate --mode perf --no-cpu-baseline --no-other-mode --residual-budget 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vertex_perf_kernel -s 330 -c 1 -f -o gpurun_out/s3_k1perf $B > gpurun_out/s3_ncu2.log 2>&1
echo "ncu k1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_kernel -s 330 -c 1 -f -o gpurun_out/s3_edge $B > gpurun_out/s3_ncu3.log 2>&1
echo "ncu edge rc=$?"
