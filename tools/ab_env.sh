#!/bin/bash
# A/B of several environment settings on the same box: tools/ab_env.sh "A=1,B=0 A=0,B=0 ..." [bench args]
# (each configuration is a comma-separated list of VAR=value; "-" = no change); errors go to gpurun_out/ab_env.err
CFGS=$1; shift
mkdir -p gpurun_out
for c in $CFGS; do
  envs=$(echo "$c" | tr ',' ' ')
  [ "$c" = "-" ] && envs=""
  env $envs python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" 2>>gpurun_out/ab_env.err | python -c "
import sys, json
lines = sys.stdin.read().strip().splitlines()
if not lines:
    print('$c  FAILED (see gpurun_out/ab_env.err)'); sys.exit(0)
d = json.loads(lines[-1]); r = d['roofline']
print('$c  ms/step %.3f  e2e %.3f  conv frac %.3f  deep %.3f ms  large %.3f ms  launches %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], r['frac'], r['deep_layers']['ms_per_step'], r['large_layers']['ms_per_step'], d['gpu_launches']))
for k, f in r['families'].items(): print('    %-20s %.3f ms  %s' % (k, f['ms_per_step'], ('%.0f TF' % f['tflops']) if f['tflops'] else ''))
for s in r['secondary']: print('    %-50s %.3f ms  frac %.3f' % (s['stage'][:50], s['ms_per_step'], s['frac']))
"
done
