#!/bin/bash
# A/B of run-time knobs on the device runner (it/s, graph replay, 512x512 denoise). usage: scripts/ab_env.sh VAR v1 v2 ...
VAR=$1; shift
for v in "$@"; do
  for rep in 1 2; do
    env $VAR=$v DIP_PROF_TIME=1 timeout 120 python scripts/profile_step.py 600 2>&1 | grep config | sed "s|^|[$VAR=$v] |"
  done
done
