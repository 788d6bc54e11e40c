#!/bin/bash
# One GPU call: bf16 parity tests (verbose), the default bench line, bf16 ncu captures.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_bf16_gpu.py -q -s -p no:cacheprovider > gpurun_out/bf16_tests_full.log 2>&1
grep -n "^\[bf16\|passed\|failed\|^FAILED\|^E  " gpurun_out/bf16_tests_full.log | cut -c1-420 | tail -40
timeout -s KILL 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_default.json') if l.startswith('{')][-1])
    print('value', d['value'], 'e2e', d['e2e']['value'], 'img', d['image_run']['it_per_s'], 'roof', d['roofline']['frac'], 'cpu', d.get('cpu_baseline',{}).get('value'))
    print(json.dumps(d.get('config3_sr_x4_1024'), indent=1)[:3000])
except Exception as e: print('bench parse failed', e)
PY
timeout -s KILL 900 bash scripts/ncu_bf16.sh > gpurun_out/ncu_bf16_script.log 2>&1; tail -5 gpurun_out/ncu_bf16_script.log
