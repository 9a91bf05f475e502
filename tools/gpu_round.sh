#!/bin/bash
# Round evidence on one B200: GPU tests, the driver's bench command (both arms), the launch list, full ncu captures of
# the step kernels, the no-pair ablation.  Outputs land in gpurun_out/; tools/make_profiles.sh turns them into profiles/.
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_gputests.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.log 2> gpurun_out/r2_bench_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench.err
python bench.py --gpus 1 --steps 200 --warmup 3 --no-extras --trace-every 25 --e2e-steps 2 > gpurun_out/r2_bench_trace.log 2>&1
UAVSIM_LIB=variants/libuavsim_nopairs.so python bench.py --steps 20 --warmup 5 --no-extras --e2e-steps 2 > gpurun_out/r2_bench_nopairs.log 2>&1
python bench.py --steps 20 --warmup 5 --no-extras --e2e-steps 2 --step-path 3 > gpurun_out/r2_bench_tile.log 2>&1
python bench.py --steps 20 --warmup 5 --no-extras --e2e-steps 2 --step-path 1 > gpurun_out/r2_bench_generic.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uavsim_step_fast -s 10 -c 1 -o gpurun_out/prof_r2_fast_dense -f python bench.py --steps 20 --warmup 5 --no-extras --e2e-steps 2 > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uavsim_step_fast -s 100 -c 1 -o gpurun_out/prof_r2_fast_s100 -f python bench.py --steps 110 --warmup 3 --no-extras --e2e-steps 2 > gpurun_out/r2_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uavsim_step_tile -s 10 -c 1 -o gpurun_out/prof_r2_tile_dense -f python bench.py --steps 20 --warmup 5 --no-extras --e2e-steps 2 --step-path 3 > gpurun_out/r2_ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uavsim_step_small -s 10 -c 1 -o gpurun_out/prof_r2_small -f python bench.py --workload default4096 --no-extras --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/r2_ncu_d.log 2>&1
UAVSIM_LIB=variants/libuavsim_nopairs.so ncu --set full --clock-control none -k regex:uavsim_step_fast -s 10 -c 1 -o gpurun_out/prof_r2_fast_nopairs -f python bench.py --steps 20 --warmup 5 --no-extras --e2e-steps 2 > gpurun_out/r2_ncu_e.log 2>&1
tail -3 gpurun_out/r2_gputests.log
python tools/generic_timing.py > gpurun_out/r2_generic_timing.log 2>&1
PYTHONPATH=. python tools/pmi_hidden_timing.py > gpurun_out/r2_pmi_hidden_timing.log 2>&1
python tools/pmi_small_timing.py > gpurun_out/r2_pmi_small_timing.log 2>&1
python tools/e2e_timing.py > gpurun_out/r2_e2e_timing.log 2>&1
ncu --set full --clock-control none -k regex:pmi_tc -s 3 -c 1 -o gpurun_out/prof_r2_pmi_tc -f python bench.py --workload swarm64_pmi --steps 8 --warmup 3 --no-extras --e2e-steps 2 > gpurun_out/r2_ncu_pmi.log 2>&1
