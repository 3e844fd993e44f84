#!/bin/bash
# round 2 full validation on one B200: whole GPU suite, smoke, both bench arms with the default arguments, then the evidence
# files of profiles/ (launch list, timelines, micro-benchmarks)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=1200 --timeout-method=thread -s --durations=12 > gpurun_out/r2_tests_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_full.log
grep -vE "^\s*$|UserWarning|_warn|return float" gpurun_out/r2_tests_full.log | grep -v "^input_blocks\|^output_blocks\|^middle" | tail -30
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/r2_smoke.log
( time timeout 1200 python bench.py ) > gpurun_out/r2_bench_full.log 2>&1
tail -6 gpurun_out/r2_bench_full.log | cut -c1-3000
( time timeout 1200 python bench.py --impl reference ) > gpurun_out/r2_bench_reference.log 2>&1
tail -6 gpurun_out/r2_bench_reference.log | cut -c1-1500
if [ "$1" != "quick" ]; then
  timeout 600 python tools/r2_timeline.py gpurun_out/r2_timeline_train_step.csv > gpurun_out/r2_timeline_train_step_summary.txt 2>&1
  timeout 600 python tools/r2_timeline.py gpurun_out/r2_timeline_ddim.csv ddim > gpurun_out/r2_timeline_ddim_summary.txt 2>&1
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_step.csv python tools/r2_profile_step.py train > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/r2_launches_train_step.csv > gpurun_out/r2_launches_train_step_summary.txt 2>&1
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_ddim_2steps_b512.csv python tools/r2_profile_step.py ddim > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/r2_launches_ddim_2steps_b512.csv > gpurun_out/r2_launches_ddim_summary.txt 2>&1
  timeout 600 python tools/gpu_gn_bench.py > gpurun_out/r2_gn_bench_final.log 2>&1
  timeout 600 python tools/gpu_igemm_bench.py fwd stats > gpurun_out/r2_igemm_bench_final_stats.log 2>&1
  timeout 600 python tools/r2_gnload_bench.py > gpurun_out/r2_gnload_bench.log 2>&1
  timeout 600 python tools/r2_bn_sweep.py > gpurun_out/r2_bn_sweep.log 2>&1
  timeout 600 python tools/r2_wgrad_mc.py > gpurun_out/r2_wgrad_isolated.log 2>&1
  head -3 gpurun_out/r2_timeline_train_step_summary.txt | tail -2
fi
