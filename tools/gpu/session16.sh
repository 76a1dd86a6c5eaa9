#!/bin/bash
# GPU session 16: final build — full GPU test suite (incl. the grid-vs-classic fixtures), one bench line (edge kernel with index prefetch)
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s16_smoke.log 2>&1
if ! grep -q "smoke ok" gpurun_out/s16_smoke.log; then tail -30 gpurun_out/s16_smoke.log; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py --durations=5 > gpurun_out/s16_pytest.log 2>&1
echo "pytest rc=$?"; tail -14 gpurun_out/s16_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-mode --residual-budget 0 > gpurun_out/s16_bench.json 2>gpurun_out/s16_bench.err; python -c "
import json
d=json.load(open('gpurun_out/s16_bench.json')); r=d['roofline']
print('it/s %.0f  ms %.4f  k1 %.4f  edge %.4f  e2e %.1f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['edge_kernel']['ms'], d['e2e']['value']), d['parity_gate']['passed'], r['traffic'])"
