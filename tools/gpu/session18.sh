#!/bin/bash
# GPU session 18: warp-cooperative edge kernel (opt-in): the GPU tests with it, A/B against the per-edge kernel
set -u
mkdir -p gpurun_out
GCS_EDGE_KERNEL=coop timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s18_smoke.log 2>&1
tail -1 gpurun_out/s18_smoke.log | cut -c1-200
GCS_EDGE_KERNEL=coop timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/s18_pytest_coop.log 2>&1
echo "pytest(coop) rc=$?"; tail -6 gpurun_out/s18_pytest_coop.log | cut -c1-300
B="python bench.py --steps 20 --warmup 3 --no-gate --mode perf --no-cpu-baseline --no-other-mode --residual-budget 0"
run() { echo "== $1"; env $1 timeout 300 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('it/s %.0f  ms %.4f  k1 %.4f  edge %.4f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['edge_kernel']['ms']))"; }
run "GCS_EDGE_KERNEL=per_edge"
run "GCS_EDGE_KERNEL=coop"
run "GCS_EDGE_KERNEL=coop GCS_EDGE_BLOCKS_PER_SM=3"
run "GCS_EDGE_KERNEL=coop GCS_EDGE_MINB=3 GCS_EDGE_BLOCKS_PER_SM=3"
