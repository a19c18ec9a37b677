#!/bin/bash
# A/B on one box: analytic pad keys vs written pad constants, three configs, alternating.
mkdir -p gpurun_out
for rep in 1 2; do
for cfg in toc3d_faster_1600 eva_vit_l_1600 toc3d_fast; do
for ab in 0 1; do
  if [ $ab = 1 ]; then export TOC3D_NO_ANALYTIC_PADS=1; else unset TOC3D_NO_ANALYTIC_PADS; fi
  timeout 600 python bench.py --config $cfg --no-cpu-baseline --no-other-configs --no-batch4 --no-roofline --steps 20 > gpurun_out/ab_$cfg.$ab.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_$cfg.$ab.json')); print('$cfg', 'no_analytic=$ab', '%.2f samples/s %.3f ms' % (d['value'], d['ms_per_step']), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done; done
