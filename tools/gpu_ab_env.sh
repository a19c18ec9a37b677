#!/bin/bash
# A/B of an environment switch on ONE box, alternating runs.  Usage: tools/gpu_ab_env.sh VAR "cfg1 cfg2 ..." [reps]
var=$1; cfgs=${2:-toc3d_fast}; reps=${3:-2}
mkdir -p gpurun_out
for rep in $(seq $reps); do for cfg in $cfgs; do for ab in 0 1; do
  if [ $ab = 1 ]; then export $var=1; else unset $var; fi
  timeout 600 python bench.py --config $cfg --no-cpu-baseline --no-other-configs --no-batch4 --no-roofline --steps 20 > gpurun_out/ab_$cfg.$ab.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_$cfg.$ab.json')); print('$cfg', '$var=$ab', '%.2f samples/s %.3f ms' % (d['value'], d['ms_per_step']), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done; done
