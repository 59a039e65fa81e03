#!/bin/bash
# Run on the GPU box: quick C2 bench over (lanes, trajectories-per-group) and print one line each.
mkdir -p gpurun_out
SPAN=${SPAN:-0.25}
for T in ${TS:-1 2}; do
  for L in ${LANES:-8}; do
    NYXB_COOP_T=$T python bench.py --steps 2 --warmup 1 --span-days $SPAN --no-cpu-baseline --lanes $L 2> gpurun_out/sweep_err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('T', $T, 'lanes', $L, 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'], 'kern_ms %.1f' % d['roofline']['kernel_ms'], d['clocks']['sm_mhz'])
" | tee -a gpurun_out/sweep.log
  done
done
