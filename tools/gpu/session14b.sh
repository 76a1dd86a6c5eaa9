#!/bin/bash
# GPU session 14b (N GPUs, N = $1, default 8): the multi-GPU configurations of BASELINE.json on the final build, incl. time to residual 1e-4
set -u
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$1', 'it/s %.0f ms %.4f' % (d['value'], d['ms_per_step']), d['config']['mode'][:12], d.get('consistency_vs_1gpu'), {k: (v['value'], v['ms_per_step']) for k, v in d.items() if k.endswith('_mode')}, d.get('problem_iterations_per_second'), d['e2e']['value'])
t=d.get('time_to_residual_1e-4')
if t: print({k: t[k] for k in ('reached','seconds','iterations','pri_res','dual_res','inner_res','pri_res_reference_definition','relaxed_cost')})"; }
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --residual-budget 1 --residual-cap 1000000 > gpurun_out/s14_bench_${N}gpu_grid316.json 2> gpurun_out/s14_bench_${N}gpu_grid316.err
echo "grid316 rc=$?"; tail -2 gpurun_out/s14_bench_${N}gpu_grid316.err | cut -c1-300; show gpurun_out/s14_bench_${N}gpu_grid316.json
if [ "$N" = "8" ]; then
timeout 500 $TR bench.py --gpus $N --steps 20 --warmup 3 --workload batch4096 --no-gate --mode perf > gpurun_out/s14_bench_${N}gpu_batch4096.json 2> gpurun_out/s14_bench_${N}gpu_batch4096.err
echo "batch4096 rc=$?"; tail -2 gpurun_out/s14_bench_${N}gpu_batch4096.err | cut -c1-300; show gpurun_out/s14_bench_${N}gpu_batch4096.json
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 --workload grid1000 --burn-in 50 --no-gate --mode perf --no-other-mode > gpurun_out/s14_bench_${N}gpu_grid1000.json 2> gpurun_out/s14_bench_${N}gpu_grid1000.err
echo "grid1000 rc=$?"; tail -2 gpurun_out/s14_bench_${N}gpu_grid1000.err | cut -c1-300; show gpurun_out/s14_bench_${N}gpu_grid1000.json
fi
