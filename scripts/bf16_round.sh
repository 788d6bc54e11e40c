#!/bin/bash
# One GPU call: per-scale widths (snail) + bf16 parity tests (verbose), the SR full-budget runs, regression of the rest.
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_variants.py -q -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/variants.log; tail -12 gpurun_out/variants.log
timeout -s KILL 600 python -m pytest tests/test_bf16_gpu.py -q -s -p no:cacheprovider > gpurun_out/bf16_tests_full.log 2>&1
grep -n "^\[bf16\|passed\|failed\|^FAILED\|Error" gpurun_out/bf16_tests_full.log | cut -c1-400 | tail -60
timeout -s KILL 600 python -m pytest "tests/test_full_budget_gpu.py::test_sr_zebra_2000_iterations_vs_reference_runs" -q -s -p no:cacheprovider > gpurun_out/bf16_sr_full.log 2>&1
grep -n "tail50_engine\|diff_vs_ref_mean\|ref_spread\|precision\|passed\|failed\|^FAILED\|Error" gpurun_out/bf16_sr_full.log | cut -c1-300 | tail -40
timeout -s KILL 600 python -m pytest tests/test_conv_ops_gpu.py tests/test_engine_gpu.py tests/test_baseline_shapes_gpu.py tests/test_downsampler.py tests/test_noise_and_guards_gpu.py -q -p no:cacheprovider 2>&1 | tail -8
timeout -s KILL 600 python -m pytest tests/test_notebooks_gpu.py -q -p no:cacheprovider 2>&1 | tail -8
