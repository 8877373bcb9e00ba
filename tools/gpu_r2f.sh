#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; r=$?; echo "== $name exit $r"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=6 run gemm python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -x -p no:cacheprovider -k "residual"
TAILN=45 run fused_ln_bench python tools/fused_ln_bench.py
python tools/one_gemm.py > gpurun_out/one_gemm_plain.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -c 6 -o gpurun_out/gemm_epi3 python tools/one_gemm.py > gpurun_out/ncu_gemm_epi.log 2>&1
echo "== ncu exit $?"; tail -3 gpurun_out/ncu_gemm_epi.log
exit 0
