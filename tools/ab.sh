#!/bin/bash
# A/B of raster-scorer variants on one GPU: prints kernel ms / roofline fraction / e2e per configuration
# usage: tools_ab.sh "layout variant l2fetch" ...
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  export UAM_RASTER_LAYOUT=$1 UAM_INT_VARIANT=$2
  if [ "$3" != "0" ]; then export UAM_L2_FETCH_GRANULARITY=$3; else unset UAM_L2_FETCH_GRANULARITY; fi
  python bench.py --steps 5 --warmup 3 --no-cpu $EXTRA 2>gpurun_out/ab_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('layout=$1 variant=$2 l2fetch=$3', 'kernel_ms=%.3f frac=%.3f value=%.3e e2e_ms=%.3f samples/s=%.3e launches=%d' % (d['roofline']['kernel_ms'], d['roofline']['frac'], d['value'], d['e2e']['ms_per_step'], d['samples_per_s'], d['gpu_launches']))
" || tail -5 gpurun_out/ab_err.log
done
