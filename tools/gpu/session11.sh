#!/bin/bash
# GPU session 11 (2 GPUs): fused peer kernels (multi-GPU parity tests), bench at N = 2 incl. time to residual 1e-4
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/s11_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -8 gpurun_out/s11_pytest.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 1; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --residual-budget 1 --residual-cap 1100000 > gpurun_out/s11_bench_2gpu.json 2> gpurun_out/s11_bench_2gpu.err
echo "bench rc=$?"; tail -3 gpurun_out/s11_bench_2gpu.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/s11_bench_2gpu.json').read().strip().splitlines()[-1])
print('it/s %.0f ms %.4f' % (d['value'], d['ms_per_step']), d.get('consistency_vs_1gpu'), d.get('time_to_residual_1e-4'), {k: (v['value'], v['ms_per_step']) for k, v in d.items() if k.endswith('_mode')})"
