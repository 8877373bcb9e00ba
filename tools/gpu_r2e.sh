#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; r=$?; echo "== $name exit $r"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=12 run gemm python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -x -p no:cacheprovider
TAILN=12 run api python -m pytest tests/test_gpu_api.py -m gpu -q --tb=short -p no:cacheprovider
TAILN=45 run fused_ln_bench python tools/fused_ln_bench.py
B="python bench.py --steps 10 --warmup 3"
TAILN=1 run bench_default $B --no-cpu-baseline
TAPCLIP_FUSE_LN=0 TAILN=1 run bench_f0 $B --no-cpu-baseline
TAPCLIP_FUSE_LN=1 TAILN=1 run bench_f1 $B --no-cpu-baseline
TAILN=1 run bench_default2 $B --no-cpu-baseline
TAPCLIP_FUSE_LN=0 TAILN=1 run bench_f0_2 $B --no-cpu-baseline
python tools/one_gemm.py > gpurun_out/one_gemm_plain.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -c 12 -o gpurun_out/gemm_epi2 python tools/one_gemm.py > gpurun_out/ncu_gemm_epi.log 2>&1
echo "== ncu exit $?"; tail -3 gpurun_out/ncu_gemm_epi.log
exit 0
