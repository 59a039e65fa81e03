#!/bin/bash
# Run on the GPU box: per-thread column-walk kernel (lanes = 1) at the C2 size over prefetch-distance / block-size variants.
mkdir -p gpurun_out
run() {
  label=$1; shift
  env "$@" timeout 120 python bench.py --lanes 1 --steps 2 --warmup 1 --no-cpu-baseline --no-strict 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label', 'value %.3e' % d['value'], 'ms %.1f' % d['ms_per_step'])" | tee -a gpurun_out/${TAG:-r01m}_sweep.log
}
run pf1_b64 X=1
run pf1_b32 NYXB_K1_BLOCK=32
for so in nyx_b200/csrc/variants/libnyxb_*.so; do
  [ -f "$so" ] || continue
  run "$(basename $so)_b64" NYXB_LIBRARY=$so
  run "$(basename $so)_b32" NYXB_LIBRARY=$so NYXB_K1_BLOCK=32
done
