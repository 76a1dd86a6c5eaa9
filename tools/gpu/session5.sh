#!/bin/bash
# GPU session 5: merged-phase K1 + per-edge kernel: tests, bench, ncu
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s5_smoke.log 2>&1
if ! grep -q "smoke ok" gpurun_out/s5_smoke.log; then tail -30 gpurun_out/s5_smoke.log; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/s5_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/s5_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-gate --mode perf --no-cpu-baseline --no-other-mode --residual-budget 0"
timeout 300 $B > gpurun_out/s5_bench.json 2>gpurun_out/s5_bench.err; python -c "
import json
d=json.load(open('gpurun_out/s5_bench.json')); r=d['roofline']
print('it/s %.0f  ms %.4f  k1 %.4f  edge %.4f  e2e %.1f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['edge_kernel']['ms'], d['e2e']['value']), d['e2e'])"
B4="python bench.py --steps 4 --warmup 3 --no-gate --mode perf --no-cpu-baseline --no-other-mode --residual-budget 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vertex_perf_kernel -s 330 -c 1 -f -o gpurun_out/s5_k1perf $B4 > gpurun_out/s5_ncu2.log 2>&1
echo "ncu k1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_frames_kernel -s 330 -c 1 -f -o gpurun_out/s5_edge $B4 > gpurun_out/s5_ncu3.log 2>&1
echo "ncu edge rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/s5_launches_perf_grid316.csv $B4 > gpurun_out/s5_ncu1.log 2>&1
