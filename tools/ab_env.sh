#!/bin/bash
# A/B of environment knobs on one GPU: tools/ab_env.sh "UAM_BIN_SHIFT=6 UAM_BIN_CHUNK=8192" "UAM_BIN_SHIFT=7" ...
# prints ms per step / scoring kernel ms / e2e ms for each setting (bench.py without the CPU legs and the other configs)
mkdir -p gpurun_out
for cfg in "$@"; do
  env $cfg python bench.py --steps ${STEPS:-50} --warmup 5 --no-cpu --no-configs --e2e-steps 10 2>gpurun_out/ab_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$cfg', 'step_ms=%.4f kernel_ms=%.4f e2e_ms=%.3f value=%.4e' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['value']))
" || tail -5 gpurun_out/ab_err.log
done
