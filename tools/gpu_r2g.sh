#!/bin/bash
mkdir -p gpurun_out
for d in 0 4 8 16 32 12 28 60; do
  echo "== TAPCLIP_GEMM_DEBUG=$d"
  TAPCLIP_GEMM_DEBUG=$d timeout 600 python tools/fused_ln_bench.py 2>&1 | grep -E "image tower|resid" | head -3
done
exit 0
