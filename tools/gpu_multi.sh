#!/bin/bash
# multi-GPU evidence: bench lines of the class-sharded workloads at N GPUs (N = $1), kept under gpurun_out/ for profiles/
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -$N
for wl in eval_c3 train_c5 fwd_c4 train_c2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl > gpurun_out/r02_bench_${wl}_${N}gpu.json 2>gpurun_out/bench_${wl}_${N}gpu.err
  echo "== $wl ${N}gpu rc=$?"; tail -1 gpurun_out/r02_bench_${wl}_${N}gpu.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('   ms/step=%.3f img/s=%.0f e2e=%.3f per-GPU frac=%.4f' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['step_frac_of_peak']))
except Exception as e: print('   parse error', e)
"
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" gpurun_out/bench_${wl}_${N}gpu.err | tail -3 | cut -c1-200
done
exit 0
