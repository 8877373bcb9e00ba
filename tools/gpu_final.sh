#!/bin/bash
# round-2 evidence run (1 GPU): full GPU test suite, bench lines of every workload, ncu launch list + full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
export TAPCLIP_PARITY_REPORT=gpurun_out/r02b_parity_fullsize.txt
rm -f $TAPCLIP_PARITY_REPORT
timeout 1500 python -m pytest tests -m gpu -q --tb=short -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu.log
grep "\[parity\]" gpurun_out/pytest_gpu.log > gpurun_out/r02b_parity_report.txt
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -3 gpurun_out/smoke.log
SECONDS=0; python bench.py > gpurun_out/r02b_bench_train_c2.json 2> gpurun_out/bench_train_c2.err; echo "== bench train_c2 exit $? in ${SECONDS}s"; cut -c1-400 gpurun_out/r02b_bench_train_c2.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02b_bench_reference.json 2>/dev/null; echo "== reference arm exit $?"
for wl in fwd_c1 fwd_b128 eval_c3 train_c5 fwd_c4; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_$wl.json 2> gpurun_out/bench_$wl.err
  echo "== bench $wl exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r02b_bench_$wl.json').read().strip().splitlines()[-1]); print('   ms/step=%.3f img/s=%.0f e2e=%.3f frac=%.4f' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['step_frac_of_peak']))"
done
# ncu: launch list of 4 train steps, then full captures of the image-tower GEMMs and the vision attention of one step
python tools/one_step.py 5 > gpurun_out/one_step_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches.csv python tools/one_step.py 5 > gpurun_out/ncu_launches.log 2>&1
echo "== ncu launch list exit $?"
TAPCLIP_NO_OVERLAP=1 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -s 387 -c 8 -o gpurun_out/r02b_gemm python tools/one_step.py 4 > gpurun_out/ncu_gemm.log 2>&1
echo "== ncu gemm exit $?"
TAPCLIP_NO_OVERLAP=1 ncu --set full --import-source on --clock-control none -k regex:attn_fwd_tc2 -s 37 -c 3 -o gpurun_out/r02b_attn python tools/one_step.py 4 > gpurun_out/ncu_attn.log 2>&1
echo "== ncu attn exit $?"

du -sh gpurun_out; ls -la gpurun_out/*.ncu-rep
exit 0
