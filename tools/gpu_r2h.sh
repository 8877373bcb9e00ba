#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
for rep in 1 2; do
for f in 2 1 0; do
  TAPCLIP_FUSE_LN=$f $B > gpurun_out/bench_f${f}_$rep.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_f${f}_$rep.log').read().strip().splitlines()[-1])
print('FUSE_LN=$f rep$rep ms/step=%.3f e2e=%.3f fwd=%.3f frac=%.4f fwdfrac=%.4f clocks=%s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['forward']['ms_per_step'], d['roofline']['step_frac_of_peak'], d['forward']['step_frac_of_peak'], d['clocks']))
PY
done
done
TAPCLIP_FUSE_LN=0 TAPCLIP_GEMM_BN=512 $B > gpurun_out/bench_f0_2cta.log 2>&1
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_f0_2cta.log').read().strip().splitlines()[-1])
print('FUSE_LN=0 2CTA ms/step=%.3f fwd=%.3f' % (d['ms_per_step'], d['forward']['ms_per_step']))
for s in d['roofline']['top_shapes'][:5]: print('   ', s)
PY
exit 0
