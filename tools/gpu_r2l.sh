#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
for h in 1 0 1 0; do
  TAPCLIP_FUSE_HEAD=$h $B > gpurun_out/bench_h$h.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_h$h.log').read().strip().splitlines()[-1])
print('FUSE_HEAD=$h ms/step=%.3f e2e=%.3f fwd=%.3f launches=%d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['forward']['ms_per_step'], d['gpu_launches']))
PY
done
exit 0
