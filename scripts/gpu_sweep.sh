#!/bin/bash
# quick A/B runs of the headline bench with tuning knobs (GPU box): each argument is a space-separated list of VAR=value settings
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('  Mrays/s %.1f  ms/frame %.2f  e2e %.1f | closest %.2f shadow %.2f shade %.2f other %.2f | frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel_ms_per_frame']['trace_closest'], r['kernel_ms_per_frame']['trace_shadow'], r['kernel_ms_per_frame']['shade'], r['kernel_ms_per_frame']['other'], r['frac']))
    elif 'rror' in l: print(l.strip())
"
done
