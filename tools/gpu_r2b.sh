#!/bin/bash
# round-2 call B: tuned RESID/FOLD epilogues (isolated + in the step), new API tests, full-size parity
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; r=$?; echo "== $name exit $r"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=8 run gemm python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -x -p no:cacheprovider
TAILN=45 run fused_ln_bench python tools/fused_ln_bench.py
TAILN=12 run api python -m pytest tests/test_gpu_api.py -m gpu -q --tb=short -p no:cacheprovider
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
TAILN=1 run bench_default $B
TAPCLIP_FUSE_LN=0 TAILN=1 run bench_f0 $B
TAPCLIP_FUSE_LN=1 TAILN=1 run bench_f1 $B
TAILN=6 run breakdown python tools/step_breakdown.py
TAPCLIP_PARITY_REPORT=gpurun_out/parity_fullsize.txt TAILN=30 run fullsize python -m pytest tests/test_gpu_parity_fullsize.py -m gpu -q --tb=short -s -p no:cacheprovider
exit 0
