#!/bin/bash
# GPU session 12: quality of the iterate at the 1e-4 stop (G = 100, run on to 1e-5), the default bench.py run (headline line incl.
# time to residual + certificate at G = 316), launch list + edge-kernel capture of the final build
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s12_smoke.log 2>&1
if ! grep -q "smoke ok" gpurun_out/s12_smoke.log; then tail -30 gpurun_out/s12_smoke.log; exit 1; fi
timeout 600 python -m pytest tests/test_gpu_perf.py -m gpu -q -x > gpurun_out/s12_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/s12_pytest.log | cut -c1-300
T="timeout 120 python tools/time_to_residual.py --grid 100 --max-iters 1500000 --trace 30 --budget 70 --tol 1e-5"
for cfg in "--rho0 3 --warm dijkstra --outer-alpha 1.7" "--rho0 3 --outer-alpha 1.7"; do
  echo "== $cfg"; $T $cfg 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    if 'it' in d: print(d['it'], '%.3e %.3e %.3e' % (d['pri'], d['dual'], d['inner']), d.get('cost'))
    else: print({k: d[k] for k in d if k in ('reached', 'iterations', 'seconds', 'cost', 'straight_line', 'warm')})"
done > gpurun_out/s12_quality_grid100.log 2>&1
cat gpurun_out/s12_quality_grid100.log
timeout 1500 python bench.py > gpurun_out/s12_bench_default.json 2> gpurun_out/s12_bench_default.err
echo "bench rc=$?"; tail -3 gpurun_out/s12_bench_default.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/s12_bench_default.json')); r=d['roofline']
print('it/s %.0f  ms %.4f  k1 %.4f  edge %.4f  e2e %.1f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['edge_kernel']['ms'], d['e2e']['value']), d['parity_gate']['passed'])
print(d.get('time_to_residual_1e-4')); print(d.get('cpu_baseline'))"
B4="python bench.py --steps 4 --warmup 3 --no-gate --mode perf --no-cpu-baseline --no-other-mode --residual-budget 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/s12_launches_perf_grid316.csv $B4 > gpurun_out/s12_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_frames_kernel -s 330 -c 1 -f -o gpurun_out/s12_edge $B4 > gpurun_out/s12_ncu3.log 2>&1
echo "ncu rc=$?"
