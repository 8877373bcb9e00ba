#!/bin/bash
# round-2 call D: direct-layout RESID epilogue with shifted 16-bit copy, FOLD stats prefetch, new bench.py
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; r=$?; echo "== $name exit $r"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=12 run gemm python -m pytest tests/test_gpu_gemm.py -m gpu -q --tb=short -x -p no:cacheprovider
TAILN=12 run api python -m pytest tests/test_gpu_api.py tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider
TAILN=30 run parity python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -s -p no:cacheprovider -k "layernorm_folded or golden"
grep "\[parity\]" gpurun_out/parity.log | grep -E "FUSE_LN|vitb16_c1" | cut -c1-200
TAILN=45 run fused_ln_bench python tools/fused_ln_bench.py
TAPCLIP_PARITY_REPORT=gpurun_out/parity_fullsize.txt TAILN=12 run fullsize python -m pytest tests/test_gpu_parity_fullsize.py -m gpu -q --tb=short -s -p no:cacheprovider
TAPCLIP_FUSE_LN=0 TAPCLIP_PARITY_REPORT=gpurun_out/parity_fullsize_f0.txt TAILN=12 run fullsize_f0 python -m pytest tests/test_gpu_parity_fullsize.py -m gpu -q --tb=short -s -p no:cacheprovider
B="python bench.py --steps 10 --warmup 3"
TAILN=1 run bench_default $B --quick-cpu
TAPCLIP_FUSE_LN=0 TAILN=1 run bench_f0 $B --no-cpu-baseline
TAPCLIP_FUSE_LN=1 TAILN=1 run bench_f1 $B --no-cpu-baseline
exit 0
