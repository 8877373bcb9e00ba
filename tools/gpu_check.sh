#!/bin/bash
# Runs every GPU test file in its own process (a CUDA fault in one does not poison the rest) with a
# timeout, and collects logs under gpurun_out/.  Usage (through gpurun): bash tools/gpu_check.sh [files...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
files="$@"
[ -z "$files" ] && files="tests/test_gpu_gemm.py tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_preprocess.py"
rc=0
for f in $files; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --tb=short -s -p no:cacheprovider > gpurun_out/$name.log 2>&1
  r=$?
  echo "== $f exit $r"; tail -n 25 gpurun_out/$name.log
  [ $r -ne 0 ] && rc=$r
done
exit $rc
