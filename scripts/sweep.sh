#!/bin/bash
# usage: scripts/sweep.sh VAR v1 v2 ... -- [bench args]   (prints ms_per_step / frac per value)
var=$1; shift
vals=()
while [ "$1" != "--" ] && [ $# -gt 0 ]; do vals+=("$1"); shift; done
shift
for v in "${vals[@]}"; do
  env $var=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$var=$v', 'ms %.3f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'])"
done
