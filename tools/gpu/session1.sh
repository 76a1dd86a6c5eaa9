#!/bin/bash
# GPU session 1 of round 2: smoke, GPU tests, bench, launch list, ncu captures of the two kernels of a perf-mode iteration
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_gpu.txt 2>&1
nproc >> gpurun_out/s1_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s1_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/s1_smoke.log
if ! grep -q "smoke ok" gpurun_out/s1_smoke.log; then tail -30 gpurun_out/s1_smoke.log; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/s1_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/s1_pytest.log
tail -25 gpurun_out/s1_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 --residual-budget 150 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/s1_bench.json; tail -5 gpurun_out/s1_bench.err
B="python bench.py --steps 4 --warmup 3 --no-gate --mode perf --no-cpu-baseline --no-other-mode --residual-budget 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/s1_launches_perf_grid316.csv $B > gpurun_out/s1_ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vertex_perf_kernel -s 330 -c 1 -f -o gpurun_out/s1_k1perf $B > gpurun_out/s1_ncu2.log 2>&1
echo "ncu k1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_kernel -s 330 -c 1 -f -o gpurun_out/s1_edge $B > gpurun_out/s1_ncu3.log 2>&1
echo "ncu edge rc=$?"
ls -la gpurun_out | tail -20
