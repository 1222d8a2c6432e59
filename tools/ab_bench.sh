#!/bin/bash
# A/B of one environment switch on the same box: tools/ab_bench.sh VAR "v1 v2 v1 v2" [bench args]
VAR=$1; VALS=$2; shift 2
for v in $VALS; do
  env $VAR=$v python bench.py --steps 30 --warmup 5 "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('$VAR=$v  ms/step %.3f  e2e %.3f  conv frac %.3f  deep %.3f ms  large %.3f ms  launches %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], r['frac'], r['deep_layers']['ms_per_step'], r['large_layers']['ms_per_step'], d['gpu_launches']))
for s in r['secondary']: print('    %-50s %.3f ms  frac %.3f' % (s['stage'][:50], s['ms_per_step'], s['frac']))
"
done
