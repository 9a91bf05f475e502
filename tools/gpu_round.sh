#!/bin/bash
# Round evidence on one B200: GPU tests, the driver's bench command, the launch list, full captures of the dominant kernel.
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_gputests.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.log 2> gpurun_out/r2_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uavsim_step_fast -s 10 -c 1 -o gpurun_out/prof_r2_fast_dense -f python bench.py --steps 20 --warmup 5 --no-extras --e2e-steps 2 > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uavsim_step_fast -s 100 -c 1 -o gpurun_out/prof_r2_fast_s100 -f python bench.py --steps 110 --warmup 3 --no-extras --e2e-steps 2 > gpurun_out/r2_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uavsim_step_tile -s 10 -c 1 -o gpurun_out/prof_r2_tile_dense -f python bench.py --steps 20 --warmup 5 --no-extras --e2e-steps 2 --step-path 3 > gpurun_out/r2_ncu_c.log 2>&1
tail -3 gpurun_out/r2_gputests.log
