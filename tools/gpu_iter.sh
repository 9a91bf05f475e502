#!/bin/bash
# one optimisation iteration on the GPU box: parity tests of the 64x64 kernels, the driver's bench command, one ncu capture
tag=${1:-iter}
timeout 600 python -m pytest tests/test_gpu_fast_step.py -x -q 2>&1 | tail -5 > gpurun_out/${tag}_test.log
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err
if [ "$2" != "noncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:uavsim_step_${3:-fast} -s 10 -c 1 -o gpurun_out/prof_${tag} -f python bench.py --steps 20 --warmup 5 --no-extras --e2e-steps 2 > gpurun_out/ncu_${tag}.log 2>&1
fi
cat gpurun_out/${tag}_test.log
python - <<P
import json
d=json.loads(open('gpurun_out/${tag}_bench.log').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'episode', d['episode']['ms_per_step'], 'frac', d['roofline']['frac'])
P
