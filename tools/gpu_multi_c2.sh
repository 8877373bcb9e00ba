#!/bin/bash
# final-code confirmation at N GPUs: the C2 bench line (and, at N=2, the multi-GPU tests)
N=${1:-8}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --workload train_c2 > gpurun_out/r02b_bench_train_c2_${N}gpu.json 2>gpurun_out/bench_train_c2_${N}gpu.err
echo "== train_c2 ${N}gpu rc=$?"; tail -1 gpurun_out/r02b_bench_train_c2_${N}gpu.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('   ms/step=%.3f img/s=%.0f e2e=%.3f per-GPU frac=%.4f' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['step_frac_of_peak']))"
grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" gpurun_out/bench_train_c2_${N}gpu.err | tail -3 | cut -c1-200
exit 0
