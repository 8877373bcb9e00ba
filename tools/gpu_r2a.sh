#!/bin/bash
# round-2 call A: new GEMM epilogues + scheduler: tests, then bench A/B over the env switches
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; r=$?; echo "== $name exit $r"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=15 run gemm_sched1 python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -x -p no:cacheprovider
TAPCLIP_GEMM_SCHED=0 TAILN=15 run gemm_sched0 python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -x -p no:cacheprovider
TAILN=25 run parity python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider
TAILN=15 run api python -m pytest tests/test_gpu_api.py tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
TAILN=1 run bench_default $B
TAPCLIP_FUSE_LN=0 TAPCLIP_GEMM_SCHED=0 TAILN=1 run bench_f0s0 $B
TAPCLIP_FUSE_LN=2 TAPCLIP_GEMM_SCHED=0 TAILN=1 run bench_f2s0 $B
TAPCLIP_FUSE_LN=0 TAPCLIP_GEMM_SCHED=1 TAILN=1 run bench_f0s1 $B
TAPCLIP_TEXT_PRIO=1 TAILN=1 run bench_prio $B
TAPCLIP_NO_OVERLAP=1 TAILN=1 run bench_noov $B
TAILN=6 run breakdown python tools/step_breakdown.py
TAPCLIP_FUSE_LN=0 TAILN=6 run breakdown_f0 python tools/step_breakdown.py
exit 0
