#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel of libuavsim.so (needs -lineinfo).

    python tools/sass_lines.py <kernel name substring> [--so path] [--top N] [--listing out.sass]

Used to keep the instruction budget of the step kernels reproducible: the table it prints (and the listing it can
write) are what profiles/*_sass_lines.txt hold.
"""
import argparse, collections, os, re, subprocess, sys, tempfile

ap = argparse.ArgumentParser()
ap.add_argument("kernel")
ap.add_argument("--so", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "marl_uavs_targets_tracking_b200", "csrc", "libuavsim.so"))
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--listing")
ap.add_argument("--file", default="step_fast_kernel.cuh", help="only break down lines of this source file")
a = ap.parse_args()
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(a.so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = None
for i, ln in enumerate(dis):
    if ln.startswith(".text.") and a.kernel in ln:
        start = i
        break
if start is None:
    sys.exit("kernel not found")
end = len(dis)
for i in range(start + 1, len(dis)):
    if dis[i].startswith("//--------------------- "):
        end = i
        break
body = dis[start:end]
if a.listing:
    open(a.listing, "w").write("\n".join(body) + "\n")
cur = ("?", 0)
per_line, per_file, ops = collections.Counter(), collections.Counter(), collections.Counter()
total = 0
stack = []
for ln in body:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        # inlined-at chains: attribute to the outermost line inside the file of interest
        chain = re.findall(r'inlined at "([^"]+)", line (\d+)', ln)
        for f, l in chain:
            if os.path.basename(f) == a.file:
                cur = (os.path.basename(f), int(l))
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
    if m:
        total += 1
        per_line[cur] += 1
        per_file[cur[0]] += 1
        ops[m.group(2).split(".")[0]] += 1
print("kernel", a.kernel, "instructions", total)
print("by file:", dict(per_file))
print("top opcodes:", ops.most_common(25))
print("top lines:")
for (f, l), c in per_line.most_common(a.top):
    print("  %-28s %5d  %d" % (f, l, c))
