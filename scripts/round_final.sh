#!/bin/bash
# Final check of the round: the whole GPU suite, smoke(), the default bench line.
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/final_gpu_tests.log 2>&1; tail -15 gpurun_out/final_gpu_tests.log | cut -c1-300
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | cut -c1-400
timeout -s KILL 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 400 gpurun_out/bench_final.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_final.json') if l.startswith('{')][-1])
    print('value', d['value'], 'e2e', d['e2e']['value'], 'img', d['image_run']['it_per_s'], 'roof', d['roofline']['frac'], 'cpu', d.get('cpu_baseline',{}).get('value'))
    c=d['config3_sr_x4_1024']
    print({p:(c[p]['it_per_s'], c[p].get('e2e_it_per_s')) for p in ('tf32','bf16')})
except Exception as e: print('bench parse failed', e)
PY
