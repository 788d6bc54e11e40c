#!/bin/bash
# Round-2 profile captures (run under gpurun): launch list of one step + full capture of the dominant conv launch.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python scripts/profile_step.py 2 > /dev/null 2>&1
DIP_NO_GRAPH=1 DIP_NO_SIDE=1 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 18 -c 1 -o gpurun_out/r2_conv_l0up python scripts/profile_step.py 1 > gpurun_out/r2_ncu_conv.log 2>&1
DIP_NO_GRAPH=1 DIP_NO_SIDE=1 ncu --set full --clock-control none -k regex:tc_conv_kernel -s 20 -c 2 -o gpurun_out/r2_conv_l0up_dgrad python scripts/profile_step.py 1 >> gpurun_out/r2_ncu_conv.log 2>&1
DIP_NO_GRAPH=1 DIP_NO_SIDE=1 ncu --set full --clock-control none -k regex:tc_wgrad_kernel -s 1 -c 1 -o gpurun_out/r2_wgrad_l0up python scripts/profile_step.py 1 >> gpurun_out/r2_ncu_conv.log 2>&1
ls -la gpurun_out/*.ncu-rep
