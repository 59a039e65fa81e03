#!/bin/bash
# Run on the GPU box: sweep library variants x schedules over the quick C2 bench.
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  env "$@" python bench.py --steps 2 --warmup 1 --span-days ${SPAN:-0.25} --no-cpu-baseline --lanes ${LANES:-8} 2> gpurun_out/sweep_err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$label', 'value %.3e' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'kern_ms %.1f' % d['roofline']['kernel_ms'])
" | tee -a gpurun_out/sweep.log
}
run default X=1
run rounds NYXB_COOP_SCHED=rounds
for so in nyx_b200/csrc/variants/libnyxb_*.so; do
  run $so NYXB_LIBRARY=$so
  run $so+rounds NYXB_LIBRARY=$so NYXB_COOP_SCHED=rounds
done
