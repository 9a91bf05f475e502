#!/usr/bin/env python
"""Plain-text summary of one kernel launch of an ncu report: the counters DESIGN.md and the verdicts quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--launch 0] > profiles/<name>_ncu.txt
"""
import argparse, csv, io, subprocess
ap = argparse.ArgumentParser()
ap.add_argument("report"); ap.add_argument("--launch", type=int, default=0)
a = ap.parse_args()
txt = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, val = rows[0], rows[1], rows[2 + a.launch]
d = {h: (v, u) for h, u, v in zip(hdr, units, val)}
def g(k):
    return d.get(k, ("n/a", ""))
print("kernel        :", g("Kernel Name")[0])
print("grid x block  :", g("launch__grid_size")[0], "x", g("launch__block_size")[0], "  registers/thread", g("launch__registers_per_thread")[0],
      "  smem/CTA", g("launch__shared_mem_per_block_allocated")[0], g("launch__shared_mem_per_block_allocated")[1])
print("resident CTAs/SM limits: regs %s smem %s warps %s" % (g("launch__occupancy_limit_registers")[0], g("launch__occupancy_limit_shared_mem")[0], g("launch__occupancy_limit_warps")[0]))
sel = [("duration", "gpu__time_duration.sum"), ("SM clock", "sm__cycles_elapsed.avg.per_second"),
       ("warp instructions executed", "smsp__inst_executed.sum"), ("active lanes per instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
       ("issue slots busy", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("warps active (of 64/SM)", "sm__warps_active.avg.pct_of_peak_sustained_active"),
       ("DRAM read", "dram__bytes_read.sum"), ("DRAM write", "dram__bytes_write.sum"), ("DRAM throughput", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
       ("pipe ALU", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), ("pipe FMA", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
       ("pipe FP64", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), ("pipe XU", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
       ("pipe LSU", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), ("pipe tensor (HMMA)", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"),
       ("tensor cycles active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), ("pipe TMA", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active"),
       ("shared-memory bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
       ("shared-memory wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
       ("local-memory loads (spills)", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum.pct_of_peak_sustained_elapsed")]
for name, k in sel:
    v, u = g(k)
    print("%-30s: %s %s" % (name, v, u))
print("stalls per issued instruction (warps waiting, by reason):")
for k in sorted(d):
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "selected" not in k:
        v = float(d[k][0] or 0)
        if v >= 0.05:
            print("  %-28s %.2f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
