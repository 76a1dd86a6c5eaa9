#!/bin/bash
# GPU session 13: stop test on the reference's residual definitions (global coordinates) in local frames: tests, G = 100, G = 316
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s13_smoke.log 2>&1
if ! grep -q "smoke ok" gpurun_out/s13_smoke.log; then tail -30 gpurun_out/s13_smoke.log; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_perf.py tests/test_gpu_solve.py -m gpu -q -x > gpurun_out/s13_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/s13_pytest.log | cut -c1-300
show() { python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    if 'it' in d: print(d['it'], d['s'], 'pri %.3e dual %.3e inner %.3e | ref %.3e %.3e' % (d['pri'], d['dual'], d['inner'], d['pri_ref'], d['dual_ref']), 'cost', d.get('cost'))
    else: print({k: d[k] for k in d if k in ('reached', 'iterations', 'seconds', 'cost', 'straight_line', 'warm')})"; }
timeout 120 python tools/time_to_residual.py --grid 100 --max-iters 600000 --trace 10 --budget 60 --rho0 3 --warm dijkstra --outer-alpha 1.7 2>&1 | show
timeout 700 python tools/time_to_residual.py --grid 316 --max-iters 2000000 --trace 20 --budget 520 --rho0 3 --warm dijkstra --outer-alpha 1.7 > gpurun_out/s13_tt_g316.jsonl 2>&1
show < gpurun_out/s13_tt_g316.jsonl
