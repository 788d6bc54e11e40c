#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python scripts/time_variants.py 2>&1 | tail -4
  DIP_PROF_TIME=1 DIP_PROF_CS=128 DIP_PROF_MODE=nearest DIP_PROF_MASK=1 timeout 120 python scripts/profile_step.py 400 512 512 2>&1 | tail -2
  DIP_PROF_TIME=1 DIP_PROF_CS=128 DIP_PROF_MODE=nearest DIP_PROF_MASK=1 DIP_PROF_PREC=bf16 timeout 120 python scripts/profile_step.py 400 512 512 2>&1 | tail -2
) | tee gpurun_out/variants_timings.txt
