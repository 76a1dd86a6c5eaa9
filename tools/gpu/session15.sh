#!/bin/bash
# GPU session 15: the driver's round-end sequence on the final build: smoke, reference arm, default bench line
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s15_smoke.log 2>&1
tail -2 gpurun_out/s15_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/s15_bench_ref.json 2> gpurun_out/s15_bench_ref.err
echo "ref rc=$?"; cut -c1-400 gpurun_out/s15_bench_ref.json
timeout 1500 python bench.py > gpurun_out/s15_bench_default.json 2> gpurun_out/s15_bench_default.err
echo "bench rc=$?"; tail -3 gpurun_out/s15_bench_default.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/s15_bench_default.json')); r=d['roofline']
print('it/s %.0f  ms %.4f  k1 %.4f  edge %.4f  e2e %.1f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['edge_kernel']['ms'], d['e2e']['value']), d['parity_gate']['passed'])
t=d.get('time_to_residual_1e-4'); print({k: t[k] for k in t if k not in ('mode','reference_definition_note','definition')}); print(d.get('cpu_baseline'))"
