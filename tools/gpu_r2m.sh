#!/bin/bash
for d in 0 64; do echo "== TAPCLIP_GEMM_DEBUG=$d"; TAPCLIP_GEMM_DEBUG=$d timeout 600 python tools/gemm_bench.py 30 2>&1 | tail -11 | cut -c1-100; done
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
for d in 0 64 0 64; do
  TAPCLIP_GEMM_DEBUG=$d $B > gpurun_out/bench_d$d.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_d$d.log').read().strip().splitlines()[-1])
print('DEBUG=$d ms/step=%.3f fwd=%.3f' % (d['ms_per_step'], d['forward']['ms_per_step']))
PY
done
exit 0
