#!/bin/bash
# 2-GPU call: NCCL vs fused-gather parity, then short multi-GPU bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -x -p no:cacheprovider 2>&1 | tail -25
for wl in train_c2 train_c5 eval_c3; do
for fg in 1 0; do
  TAPCLIP_FUSED_GATHER=$fg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --workload $wl > gpurun_out/bench_${wl}_2gpu_fg$fg.log 2>gpurun_out/bench_${wl}_2gpu_fg$fg.err
  echo "== $wl 2gpu fused_gather=$fg rc=$?"; tail -1 gpurun_out/bench_${wl}_2gpu_fg$fg.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('   ms/step=%.3f img/s=%.0f e2e=%.3f frac=%.4f' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['step_frac_of_peak']))
except Exception as e: print('   parse error', e)
"
  tail -3 gpurun_out/bench_${wl}_2gpu_fg$fg.err | cut -c1-200
done
done
exit 0
