#!/bin/bash
# GPU session 17: BASELINE config 3 (10k-vertex grid) as a bench line with time to residual; config 4 on one GPU
set -u
mkdir -p gpurun_out
timeout 400 python bench.py --workload grid100 --residual-budget 40 > gpurun_out/s17_bench_grid100.json 2> gpurun_out/s17_bench_grid100.err
echo "grid100 rc=$?"; tail -2 gpurun_out/s17_bench_grid100.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/s17_bench_grid100.json')); r=d['roofline']
print('it/s %.0f  ms %.4f  k1 %.4f  edge %.4f  e2e %.1f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['edge_kernel']['ms'], d['e2e']['value']), d['parity_gate']['passed'])
t=d.get('time_to_residual_1e-4'); print({k: t[k] for k in t if k not in ('mode','reference_definition_note','definition','certificate')}); print(d.get('cpu_baseline'))"
timeout 300 python bench.py --workload batch4096 --no-gate --mode perf --no-cpu-baseline > gpurun_out/s17_bench_batch4096_1gpu.json 2> gpurun_out/s17_bench_batch4096_1gpu.err
echo "batch rc=$?"; tail -2 gpurun_out/s17_bench_batch4096_1gpu.err | cut -c1-300; cut -c1-600 gpurun_out/s17_bench_batch4096_1gpu.json
