#!/bin/bash
# A/B of the wgrad K-block width (DIP_WGRAD_KP) in both tensor-core modes + the wgrad / engine tests with 64-pixel blocks in tf32
mkdir -p gpurun_out
( for rep in 1 2; do
  for kp in 32 64; do
    DIP_WGRAD_KP=$kp DIP_PROF_TIME=1 timeout 120 python scripts/profile_step.py 600 512 512 2>&1 | grep config | sed "s/^/kp=$kp /"
  done; done
  for kp in 32 64; do
    DIP_WGRAD_KP=$kp DIP_PROF_PREC=bf16 DIP_PROF_SR=1 DIP_PROF_TIME=1 timeout 120 python scripts/profile_step.py 200 1024 1024 2>&1 | grep config | sed "s/^/kp=$kp /"
    DIP_WGRAD_KP=$kp DIP_PROF_SR=1 DIP_PROF_TIME=1 timeout 120 python scripts/profile_step.py 200 1024 1024 2>&1 | grep config | sed "s/^/kp=$kp /"
  done ) | tee gpurun_out/wgrad_kp_ab.txt
DIP_WGRAD_KP=64 timeout -s KILL 300 python -m pytest tests/test_conv_ops_gpu.py tests/test_engine_gpu.py tests/test_baseline_shapes_gpu.py -q -p no:cacheprovider -k "wgrad or forward_backward or one_step" 2>&1 | tail -4
