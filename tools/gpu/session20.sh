#!/bin/bash
# GPU session 20: class table transposed (coalesced lanes in the core product): perf-mode tests + bench line
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s20_smoke.log 2>&1
tail -1 gpurun_out/s20_smoke.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_perf.py -m gpu -q -x > gpurun_out/s20_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/s20_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-mode --residual-budget 0 > gpurun_out/s20_bench.json 2>gpurun_out/s20_bench.err; python -c "
import json
d=json.load(open('gpurun_out/s20_bench.json')); r=d['roofline']
print('it/s %.0f  ms %.4f  k1 %.4f  edge %.4f  e2e %.1f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['edge_kernel']['ms'], d['e2e']['value']), d['parity_gate']['passed'], r['whole_iteration']['frac'], r['frac'])"
