#!/bin/bash
# GPU session 2: updated kernels (tests + bench), then convergence experiments for the time-to-residual metric
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2_smoke.log 2>&1
if ! grep -q "smoke ok" gpurun_out/s2_smoke.log; then tail -30 gpurun_out/s2_smoke.log; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/s2_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/s2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --residual-budget 0 --no-cpu-baseline > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/s2_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['kernel_ms'], d['roofline']['edge_kernel'], d['e2e']['value'], d['roofline']['whole_iteration'])
PY
T="timeout 200 python tools/time_to_residual.py --grid 100 --max-iters 600000 --trace 6 --budget 60"
for cfg in "--window 100" "--window 600000 --adapt-every 1" "--window 600000 --adapt-every 200" "--window 600000 --adapt-every 2000" \
           "--window 100 --outer-alpha 1.8" "--window 600000 --adapt-every 200 --outer-alpha 1.8" "--window 100 --rho0 0.1" "--window 100 --rho0 0.1 --outer-alpha 1.8" \
           "--window 100 --rho0 0.03 --outer-alpha 1.8" "--window 600000 --adapt-every 200 --outer-alpha 1.8 --inner 2"; do
  echo "== $cfg"; $T $cfg 2>&1 | tail -1
done > gpurun_out/s2_conv_grid100.log 2>&1
cat gpurun_out/s2_conv_grid100.log | cut -c1-400
