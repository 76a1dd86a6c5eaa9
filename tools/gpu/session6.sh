#!/bin/bash
# GPU session 6 (2 GPUs): multi-GPU parity tests (NCCL and peer-memory paths), bench at N = 2 with both exchanges
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/s6_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/s6_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$1', 'it/s %.0f ms %.4f' % (d['value'], d['ms_per_step']), d['config']['mode'][:12], d.get('consistency_vs_1gpu'), {k: (v['value'], v['ms_per_step']) for k, v in d.items() if k.endswith('_mode')})"; }
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s6_bench_2gpu_peer.json 2> gpurun_out/s6_bench_2gpu_peer.err
echo "peer rc=$?"; tail -3 gpurun_out/s6_bench_2gpu_peer.err; show gpurun_out/s6_bench_2gpu_peer.json
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --dist nccl --no-gate --mode perf > gpurun_out/s6_bench_2gpu_nccl.json 2> gpurun_out/s6_bench_2gpu_nccl.err
echo "nccl rc=$?"; tail -3 gpurun_out/s6_bench_2gpu_nccl.err; show gpurun_out/s6_bench_2gpu_nccl.json
