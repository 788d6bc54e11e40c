#!/bin/bash
# compute-sanitizer over the bf16 kernels and the per-scale-width / avg-downsampling step
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  DIP_SAN_PREC=2 DIP_SAN_NARROW=1 timeout -s KILL 500 compute-sanitizer --tool $tool python scripts/sanitize_ops.py > gpurun_out/san_bf16_$tool.txt 2>&1
  tail -4 gpurun_out/san_bf16_$tool.txt
done
DIP_SAN_NARROW=1 timeout -s KILL 400 compute-sanitizer --tool memcheck python scripts/sanitize_ops.py > gpurun_out/san_tf32_narrow_memcheck.txt 2>&1; tail -3 gpurun_out/san_tf32_narrow_memcheck.txt
