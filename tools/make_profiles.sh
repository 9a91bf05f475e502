#!/bin/bash
# Turns the ncu reports / logs that tools/gpu_round.sh left in gpurun_out/ into the tracked text summaries of profiles/.
set -e
cd "$(dirname "$0")/.."
G=gpurun_out P=profiles
cp $G/r2_launches.csv $P/r2_launches.csv
python tools/ncu_summary.py $G/prof_r2_fast_dense.ncu-rep > $P/r2_step_fast_kernel_ncu.txt
python tools/ncu_summary.py $G/prof_r2_fast_s100.ncu-rep > $P/r2_step_fast_kernel_step100_ncu.txt
python tools/ncu_summary.py $G/prof_r2_fast_nopairs.ncu-rep > $P/r2_step_fast_kernel_nopairs_ncu.txt
python tools/ncu_summary.py $G/prof_r2_tile_dense.ncu-rep > $P/r2_step_tile_kernel_ncu.txt
python tools/ncu_summary.py $G/prof_r2_small.ncu-rep > $P/r2_step_small_kernel_ncu.txt
python tools/ncu_summary.py $G/prof_r2_pmi_tc.ncu-rep > $P/r2_pmi_tc_kernel_ncu.txt
cp $P/r2_step_fast_kernel_ncu.txt $P/r2_step_kernel_ncu.txt   # the name VERDICT r1 asked for: the default 64x64 step kernel
for f in r2_generic_timing r2_pmi_hidden_timing r2_pmi_small_timing r2_e2e_timing; do cp $G/$f.log $P/$f.txt; done
python tools/ncu_segments.py $G/prof_r2_fast_dense.ncu-rep uavsim_step_fast_kernelILi64ELi64ELb0 --warps 2 --min-share 0.3 > $P/r2_step_fast_kernel_segments.txt
python tools/ncu_segments.py $G/prof_r2_tile_dense.ncu-rep uavsim_step_tile_kernelILb0 --warps 4 --min-share 0.3 > $P/r2_step_tile_kernel_segments.txt
python tools/ncu_segments.py $G/prof_r2_small.ncu-rep uavsim_step_small_kernelILb0 --envs 1366 --warps 2 --min-share 0.5 > $P/r2_step_small_kernel_segments.txt
python tools/ncu_lines.py $G/prof_r2_fast_dense.ncu-rep uavsim_step_fast_kernelILi64ELi64ELb0 --top 25 > $P/r2_step_fast_kernel_lines.txt
python tools/sass_lines.py uavsim_step_fast_kernelILi64ELi64ELb0 --top 25 --listing $P/r2_step_fast_kernel.sass > $P/r2_step_fast_kernel_sass_lines.txt
python tools/sass_lines.py uavsim_step_small_kernelILb0 --file step_small_kernel.cuh --top 15 --listing $P/r2_step_small_kernel.sass > /dev/null
python - <<'P'
import csv, io, json, subprocess
txt = subprocess.run(["ncu", "-i", "gpurun_out/prof_r2_fast_dense.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
def mb(k):
    v = float(d[k]); return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u[k]]
tb = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
out = {"swarm64": tb, "swarm64_self": tb,
       "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of %s, ncu --set full, launch 10 after a reset (the driver's window; "
               "profiles/r2_step_fast_kernel_ncu.txt); algorithmic bytes per launch 721682432" % d["Kernel Name"].split("(")[0]}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(out)
for n in ("r2_bench", "r2_bench_ref", "r2_bench_trace", "r2_bench_nopairs", "r2_bench_tile"):
    try:
        line = open("gpurun_out/%s.log" % n).read().strip().splitlines()[-1]
        json.dump(json.loads(line), open("profiles/%s.json" % n, "w"), indent=1)
    except Exception as e:
        print("skip", n, e)
P
ls -la $P
