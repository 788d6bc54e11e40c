#!/bin/bash
# bf16 profile captures (run under gpurun): launch list of one SR step at 1024x1024 in bf16 + full captures of the dominant
# conv launch (level-0 3x3 up conv fprop), its dgrad and its wgrad.  The kernel indices follow scripts/ncu_round2.sh
# (same launch order; only the kernel names carry the _bf16 suffix).
export DIP_PROF_PREC=bf16 DIP_PROF_SR=1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bf16_launches.csv python scripts/profile_step.py 2 1024 1024 > /dev/null 2>&1
DIP_NO_GRAPH=1 DIP_NO_SIDE=1 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel_bf16 -s 18 -c 1 -o gpurun_out/bf16_conv_l0up python scripts/profile_step.py 1 1024 1024 > gpurun_out/bf16_ncu.log 2>&1
DIP_NO_GRAPH=1 DIP_NO_SIDE=1 ncu --set full --clock-control none -k regex:tc_conv_kernel_bf16 -s 20 -c 2 -o gpurun_out/bf16_conv_l0up_dgrad python scripts/profile_step.py 1 1024 1024 >> gpurun_out/bf16_ncu.log 2>&1
DIP_NO_GRAPH=1 DIP_NO_SIDE=1 ncu --set full --clock-control none -k regex:tc_wgrad_kernel_bf16 -s 1 -c 1 -o gpurun_out/bf16_wgrad_l0up python scripts/profile_step.py 1 1024 1024 >> gpurun_out/bf16_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
