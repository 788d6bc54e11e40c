#!/bin/bash
# Runs bench.py under a list of environment settings (experiment builds / runtime switches) and prints one summary line each.
# usage: scripts/bench_matrix.sh "VAR=1" "DIP_LIB=path" ...   ("-" = defaults)
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  if [ "$cfg" = "-" ]; then envs=""; else envs="$cfg"; fi
  out=$(env $envs python bench.py --steps ${STEPS:-40} --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$out" | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-60s value %.1f it/s  e2e %.1f  dom %.1f TF/s (%.3f)  conv_all %.1f  wgrad %.1f  clk %s' % ('$cfg', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline_all_conv']['achieved'], d['roofline_wgrad']['achieved'], d['clocks']['sm_mhz']))"
done
