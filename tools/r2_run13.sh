#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_step.csv python tools/r2_profile_step.py train > gpurun_out/r2_prof_train.log 2>&1
tail -2 gpurun_out/r2_prof_train.log; wc -l gpurun_out/r2_launches_train_step.csv
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_ddim_step.csv python tools/r2_profile_step.py ddim > gpurun_out/r2_prof_ddim.log 2>&1
tail -2 gpurun_out/r2_prof_ddim.log; wc -l gpurun_out/r2_launches_ddim_step.csv
