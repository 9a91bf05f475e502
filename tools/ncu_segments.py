#!/usr/bin/env python
"""Straight-line segments of a kernel with their execution count per environment and warp, from an ncu report.

    python tools/ncu_segments.py gpurun_out/prof.ncu-rep <kernel substring> [--envs 65536] [--warps 4]

Consecutive SASS instructions with the same executed count form a segment (a loop body, a branch arm, a phase); the
table gives its length, how often one warp runs it per environment, its share of the launch, the source lines it
mostly comes from and its opcode mix.  This is the instruction budget DESIGN.md quotes.
"""
import argparse, collections, csv, io, os, re, subprocess, tempfile
ap = argparse.ArgumentParser()
ap.add_argument("report"); ap.add_argument("kernel")
ap.add_argument("--so", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "marl_uavs_targets_tracking_b200", "csrc", "libuavsim.so"))
ap.add_argument("--envs", type=int, default=65536); ap.add_argument("--warps", type=int, default=4)
ap.add_argument("--min-share", type=float, default=0.3)
a = ap.parse_args()
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(a.so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, ln in enumerate(dis) if ln.startswith(".text.") and a.kernel in ln)
line, cur = {}, None
for ln in dis[start + 1:]:
    if ln.startswith("//--------------------- "):
        break
    m = re.match(r'\s*//## File ".*/([^/"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        line[int(m.group(1), 16)] = cur
txt = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
iA, iS, iN, iT = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
base = int(rows[hi + 1][iA], 16)
data = {int(r[iA], 16) - base: (r[iS], int(r[iN]), int(r[iT])) for r in rows[hi + 1:] if len(r) == len(h)}
tot = sum(v[1] for v in data.values())
print("warp-instructions per launch %d = %.0f per environment = %.0f per warp and environment; %.1f active lanes"
      % (tot, tot / a.envs, tot / a.envs / a.warps, sum(v[2] for v in data.values()) / tot))
offs = sorted(data)
i = 0
print("%6s %5s %9s %7s  %s" % ("offset", "instr", "runs/warp", "share", "source lines | opcodes"))
while i < len(offs):
    j, n0 = i, data[offs[i]][1]
    while j < len(offs) and abs(data[offs[j]][1] - n0) <= 0.02 * max(n0, 1):
        j += 1
    cnt = j - i
    if 100.0 * n0 * cnt / tot >= a.min_share:
        op = lambda s_: (s_.split()[1] if s_.startswith("@") else s_.split()[0]).split(".")[0]
        ops = collections.Counter(op(data[o][0]) for o in offs[i:j])
        lines = collections.Counter(line.get(o) for o in offs[i:j])
        print("%6x %5d %9.2f %6.2f%%  %s | %s" % (offs[i], cnt, n0 / a.envs / a.warps, 100.0 * n0 * cnt / tot,
              " ".join("%s:%d" % (l[0].replace("step_", "").replace("_kernel.cuh", "").replace(".cuh", ""), l[1]) for l, _ in lines.most_common(3) if l),
              " ".join("%s%d" % kv for kv in ops.most_common(6))))
    i = j
