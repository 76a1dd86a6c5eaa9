#!/bin/bash
# GPU session 14a: full GPU test suite on the final build; exact-mode K1 capture (the headline kernel of round 1, two revisions newer than its last profile)
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s14_smoke.log 2>&1
if ! grep -q "smoke ok" gpurun_out/s14_smoke.log; then tail -30 gpurun_out/s14_smoke.log; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/s14_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/s14_pytest.log | cut -c1-300
BP="python bench.py --steps 3 --warmup 3 --no-gate --mode parity --no-cpu-baseline --no-other-mode --residual-budget 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vertex_kernel -s 336 -c 1 -f -o gpurun_out/s14_k1exact $BP > gpurun_out/s14_ncu_exact.log 2>&1
echo "ncu exact rc=$?"
timeout 120 python tools/time_to_residual.py --grid 100 --max-iters 600000 --trace 4 --budget 60 --rho0 3 --warm dijkstra --outer-alpha 1.7 2>&1 | tail -1 | cut -c1-500
