#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; r=$?; echo "== $name exit $r"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=8 run api python -m pytest tests/test_gpu_api.py tests/test_gpu_gemm.py -m gpu -q --tb=short -x -p no:cacheprovider
TAILN=8 run parity python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -p no:cacheprovider -k "golden or layernorm_folded"
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
for f in 0 0; do
  TAPCLIP_FUSE_LN=$f $B > gpurun_out/bench_f$f.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_f$f.log').read().strip().splitlines()[-1])
print('FUSE_LN=$f ms/step=%.3f e2e=%.3f fwd=%.3f frac=%.4f fwdfrac=%.4f fwd512=%.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['forward']['ms_per_step'], d['roofline']['step_frac_of_peak'], d['forward']['step_frac_of_peak'], d['forward']['at_batch_512']['step_frac_of_peak']))
PY
done
python tools/one_step.py 5 > gpurun_out/one_step_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/one_step.py 5 > gpurun_out/ncu_launches.log 2>&1
echo "== ncu launch list exit $?"
exit 0
