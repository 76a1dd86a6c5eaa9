#!/bin/bash
# GPU session 9: persistent double-buffered perf K1 (tests, bench, ncu), time to residual 1e-4 at G = 316 with the accelerators of session 8
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s9_smoke.log 2>&1
if ! grep -q "smoke ok" gpurun_out/s9_smoke.log; then tail -30 gpurun_out/s9_smoke.log; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_perf.py tests/test_gpu_solve.py -m gpu -q -x > gpurun_out/s9_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -8 gpurun_out/s9_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
B="python bench.py --steps 20 --warmup 3 --no-gate --mode perf --no-cpu-baseline --no-other-mode --residual-budget 0"
run() { echo "== $1"; env $1 timeout 300 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('it/s %.0f  ms %.4f  k1 %.4f  edge %.4f  e2e %.1f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['edge_kernel']['ms'], d['e2e']['value']))"; }
run "GCS_PERF_BLOCKS_PER_SM=4"
run "GCS_PERF_BLOCKS_PER_SM=3"
run "GCS_PERF_BLOCKS_PER_SM=5"
run "GCS_PERF_BLOCKS_PER_SM=4 GCS_TILE_BLOCKS=32"
timeout 300 $B > gpurun_out/s9_bench.json 2>gpurun_out/s9_bench.err
T="timeout 600 python tools/time_to_residual.py --grid 316 --max-iters 2000000 --trace 10"
$T --rho0 3 --warm dijkstra --outer-alpha 1.7 --budget 420 > gpurun_out/s9_tt_g316_fixed.jsonl 2>&1
tail -3 gpurun_out/s9_tt_g316_fixed.jsonl | cut -c1-300
$T --window 2000000 --adapt-every 100 --warm dijkstra --outer-alpha 1.7 --budget 300 > gpurun_out/s9_tt_g316_adapt.jsonl 2>&1
tail -3 gpurun_out/s9_tt_g316_adapt.jsonl | cut -c1-300
B4="python bench.py --steps 4 --warmup 3 --no-gate --mode perf --no-cpu-baseline --no-other-mode --residual-budget 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vertex_perf_kernel -s 330 -c 1 -f -o gpurun_out/s9_k1perf $B4 > gpurun_out/s9_ncu2.log 2>&1
echo "ncu k1 rc=$?"
