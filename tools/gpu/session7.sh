#!/bin/bash
# GPU session 7 (8 GPUs): the three multi-GPU configurations of BASELINE.json
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$1', 'it/s %.0f ms %.4f' % (d['value'], d['ms_per_step']), d['config']['mode'][:12], d.get('consistency_vs_1gpu'), {k: (v['value'], v['ms_per_step']) for k, v in d.items() if k.endswith('_mode')}, d.get('problem_iterations_per_second'), d['e2e']['value'])"; }
timeout 700 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s7_bench_8gpu_grid316.json 2> gpurun_out/s7_bench_8gpu_grid316.err
echo "grid316 rc=$?"; tail -2 gpurun_out/s7_bench_8gpu_grid316.err; show gpurun_out/s7_bench_8gpu_grid316.json
timeout 500 $TR bench.py --gpus 8 --steps 20 --warmup 3 --workload batch4096 --no-gate --mode perf > gpurun_out/s7_bench_8gpu_batch4096.json 2> gpurun_out/s7_bench_8gpu_batch4096.err
echo "batch4096 rc=$?"; tail -2 gpurun_out/s7_bench_8gpu_batch4096.err; show gpurun_out/s7_bench_8gpu_batch4096.json
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 3 --workload grid1000 --burn-in 50 --no-gate --mode perf --no-other-mode > gpurun_out/s7_bench_8gpu_grid1000.json 2> gpurun_out/s7_bench_8gpu_grid1000.err
echo "grid1000 rc=$?"; tail -2 gpurun_out/s7_bench_8gpu_grid1000.err; show gpurun_out/s7_bench_8gpu_grid1000.json
