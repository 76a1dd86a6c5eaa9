#!/bin/bash
# GPU session 4: A/B of tile sizes (perf K1) and edge-kernel variants on the bench workload; tests of the restructured P0
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4_smoke.log 2>&1
if ! grep -q "smoke ok" gpurun_out/s4_smoke.log; then tail -30 gpurun_out/s4_smoke.log; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_perf.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/s4_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/s4_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-gate --mode perf --no-cpu-baseline --no-other-mode --residual-budget 0"
run() { echo "== $1"; env $1 timeout 300 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('it/s %.0f  ms %.4f  k1 %.4f  edge %.4f  e2e %.1f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['edge_kernel']['ms'], d['e2e']['value']))"; }
run "GCS_TILE_BLOCKS=64"
run "GCS_TILE_BLOCKS=32"
run "GCS_TILE_BLOCKS=48"
run "GCS_TILE_BLOCKS=16"
run "GCS_TILE_BLOCKS=64 GCS_EDGE_BLOCKS_PER_SM=6"
run "GCS_TILE_BLOCKS=64 GCS_EDGE_BLOCKS_PER_SM=12"
run "GCS_TILE_BLOCKS=64 GCS_EDGE_KERNEL=per_edge GCS_EDGE_BLOCKS_PER_SM=4"
run "GCS_TILE_BLOCKS=64 GCS_EDGE_KERNEL=per_edge GCS_EDGE_BLOCKS_PER_SM=8"
run "GCS_TILE_BLOCKS=64 GCS_EDGE_KERNEL=per_edge GCS_EDGE_BLOCKS_PER_SM=16"
