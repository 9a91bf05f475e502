#!/usr/bin/env python
"""Executed warp-instructions per source line of a kernel, from an ncu report + the shipped .so.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel substring> [--launch 0] [--top 30]

ncu's CSV export of the source page lists per-SASS-instruction counters but no line numbers; nvdisasm -g on the same
libuavsim.so gives the line of every SASS offset.  The two are joined on the instruction offset inside the kernel.
Lines inside inlined helpers are attributed to the outermost line of --file (default step_fast_kernel.cuh).
"""
import argparse, collections, csv, io, os, re, subprocess, sys, tempfile

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("kernel")
ap.add_argument("--so", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "marl_uavs_targets_tracking_b200", "csrc", "libuavsim.so"))
ap.add_argument("--launch", type=int, default=0)
ap.add_argument("--top", type=int, default=30)
ap.add_argument("--file", default="step_fast_kernel.cuh")
ap.add_argument("--inner", action="store_true", help="attribute to the innermost line instead")
a = ap.parse_args()

# ---- nvdisasm: offset -> (file, line)
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(a.so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, ln in enumerate(dis) if ln.startswith(".text.") and a.kernel in ln)
# helpers (inline PTX wrappers, intrinsics headers, fast_math) carry their own line: attribute them to the most recent
# line of the kernel body instead (nvdisasm gives no inlined-at chain)
main_src = open(os.path.join(os.path.dirname(os.path.abspath(a.so)), a.file)).read().splitlines()
body_lo = next((i + 1 for i, l in enumerate(main_src) if "__device__ __noinline__ void" in l or "__global__" in l), 1)
MARKERS_FAST = [("exact path (fp64 row, whole-environment fallbacks)", "__device__ __noinline__ void fast_agent_exact"),
           ("checked twins (fp64 decisions inside a guard band)", "__device__ __noinline__ void sf_targets_checked"),
           ("kernel prologue", "__global__ void"),
           ("loop top: waits", "for (int64_t k = blockIdx.x"), ("phase 0a targets", "// ---- phase 0a"), ("phase 0b UAVs", "// ---- phase 0b"),
           ("phase 1 setup / guards", "// ---- phase 1"), ("targets: prefilter + walk", "// -- targets:"),
           ("UAV prefilter (one radius)", "// -- UAV partners."), ("UAV walk: communication + duplicate + neighbours", "// -- ONE walk over the candidate slots"),
           ("UAV walk epilogue (own entry, means)", "// own entry of the own slot"),
           ("exact call / masks / coverage", "if (exact) {"),
           ("boundary, normalise", "float raw, ttn, bpn, dupn;"), ("phase 2 cooperative reward, staging", "// ---- phase 2"),
           ("output stores", "asm volatile(\"fence.proxy.async.shared::cta;\""), ("epilogue", "// outputs complete before the CTA retires")]
MARKERS_TILE = [("pair fix-up (fp64)", "__device__ __noinline__ void tile_fix"), ("exact path (fp64 row)", "__device__ __noinline__ void tile_agent_exact"),
           ("kernel prologue", "__global__ void"), ("loop top: waits", "for (int64_t k = blockIdx.x"),
           ("phase 0a targets", "// ---- phase 0a"), ("phase 0b UAVs", "// ---- phase 0b"), ("phase 0 tail: radius, barrier, next loads", "if (!(rabs == rabs))"),
           ("phase 1 setup: bands, rows", "// ================= phase 1"), ("target tiles", "// ---- targets:"),
           ("UAV tiles: distances + communication", "// ---- UAV partners"), ("UAV tiles: duplicate / neighbours", "// duplicate tracking / neighbours on the new-new"),
           ("pair-phase epilogue (sums, coverage)", "// ---- hand the sums"), ("finish the UAVs", "// ================= finish the UAVs"),
           ("cooperative reward", "// ================= cooperative reward"),
           ("output stores", "asm volatile(\"fence.proxy.async.shared::cta;\""), ("epilogue", "// outputs complete before the CTA retires")]
MARKERS = MARKERS_TILE if "tile" in a.file else MARKERS_FAST
regions = []
for name, key in MARKERS:
    ln_ = next((i + 1 for i, l in enumerate(main_src) if key in l), None)
    if ln_:
        regions.append((ln_, name))
regions.sort()
def region_of(line):
    r = "helpers / other"
    for lo, name in regions:
        if line >= lo:
            r = name
    return r
off2line, cur, last_body = {}, ("?", 0), ("?", 0)
for ln in dis[start + 1:]:
    if ln.startswith("//--------------------- "):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        if cur[0] == a.file and cur[1] >= body_lo:
            last_body = cur
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        off2line[int(m.group(1), 16)] = (cur if a.inner else last_body, m.group(2).strip(), last_body)

# ---- ncu: per-SASS counters of the chosen launch
txt = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, curb = [], None
for row in csv.reader(io.StringIO(txt)):
    if row and row[0] == "Kernel Name":
        curb = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(curb)
    elif curb is not None and row and row[0] == "Address":
        curb["hdr"] = row
    elif curb is not None and curb["hdr"] and len(row) == len(curb["hdr"]):
        curb["rows"].append(row)
blocks = [b for b in blocks if a.kernel.replace("ILi", "<(int)").split("<")[0].split("_Z")[-1][-18:] in b["name"] or True]
b = blocks[a.launch]
h = b["hdr"]
iA, iS, iN, iT, iP = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
base = int(b["rows"][0][iA], 16)
per = collections.defaultdict(lambda: [0, 0, 0])
reg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
missing = 0
for r in b["rows"]:
    off = int(r[iA], 16) - base
    n, t, s = int(r[iN]), int(r[iT]), int(r[iP] or 0)
    ent = off2line.get(off, (("?", 0), "", ("?", 0)))
    key = ent[0]
    if off not in off2line:
        missing += 1
    for k, v in enumerate((n, t, s)):
        per[key][k] += v
        reg[region_of(ent[2][1])][k] += v
        tot[k] += v
src = {}
print("kernel: %s" % b["name"])
print("warp-instructions executed: %d   thread-instructions: %d (%.1f active lanes)   stall samples: %d   unmatched SASS: %d"
      % (tot[0], tot[1], tot[1] / max(tot[0], 1), tot[2], missing))
print("by region of the kernel (share of executed warp-instructions, active lanes, share of stall samples):")
for name, (n, t, s_) in sorted(reg.items(), key=lambda kv: -kv[1][0]):
    print("  %6.2f%%  %5.1f lanes  %6.2f%% samples  %s" % (100.0 * n / tot[0], t / max(n, 1), 100.0 * s_ / max(tot[2], 1), name))
print("%7s %7s %6s %7s  %s" % ("inst%", "cum%", "lanes", "smpl%", "file:line  source"))
cum = 0.0
for (f, l), (n, t, s) in sorted(per.items(), key=lambda kv: -kv[1][0])[: a.top]:
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(a.so)), f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:84] if 0 < l <= len(src[f]) else ""
    cum += 100.0 * n / tot[0]
    print("%6.2f%% %6.1f%% %6.1f %6.2f%%  %s:%d  %s" % (100.0 * n / tot[0], cum, t / max(n, 1), 100.0 * s / max(tot[2], 1), f, l, text))
