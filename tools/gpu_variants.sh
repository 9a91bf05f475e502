#!/bin/bash
# A/B of tuning builds on the GPU box:  tools/gpu_variants.sh <tag> name1 name2 ...   (variants/libuavsim_<name>.so)
# per variant: the 64x64 parity tests, then the driver's bench command without extras (window + episode numbers)
tag=$1; shift
for n in "$@"; do
  export UAVSIM_LIB=$PWD/variants/libuavsim_$n.so
  if [ "$NOTEST" != "1" ]; then
    timeout 600 python -m pytest tests/test_gpu_fast_step.py tests/test_gpu_bounds.py -x -q 2>&1 | tail -2 > gpurun_out/${tag}_${n}_test.log
  fi
  python bench.py --steps 20 --warmup 5 --no-extras --e2e-steps 2 > gpurun_out/${tag}_${n}_bench.log 2> gpurun_out/${tag}_${n}_bench.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/${tag}_${n}_bench.log').read().strip().splitlines()[-1])
    t=open('gpurun_out/${tag}_${n}_test.log').read().strip().splitlines()[-1] if "$NOTEST"!="1" else ""
    print('%-12s window %.4f ms  episode %.4f ms  frac %.3f   %s' % ('$n', d['ms_per_step'], d['episode']['ms_per_step'], d['roofline']['frac'], t))
except Exception as e:
    print('$n', 'FAILED', e)
P
done
