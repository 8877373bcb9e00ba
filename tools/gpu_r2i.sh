#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -x -p no:cacheprovider 2>&1 | tail -3
timeout 600 python tools/gemm_bench.py 30 2>&1 | tail -14
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
for bn in 0 512; do
  TAPCLIP_GEMM_BN=$bn $B > gpurun_out/bench_bn$bn.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_bn$bn.log').read().strip().splitlines()[-1])
print('BN=$bn ms/step=%.3f fwd=%.3f frac=%.4f fwdfrac=%.4f' % (d['ms_per_step'], d['forward']['ms_per_step'], d['roofline']['step_frac_of_peak'], d['forward']['step_frac_of_peak']))
for s in d['roofline']['top_shapes'][:8]: print('   ', s)
PY
done
exit 0
