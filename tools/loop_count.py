#!/usr/bin/env python
"""Instruction count and opcode mix of the UAV walk loop (the backward branch whose body holds two MUFU.EX2) of
uavsim_step_fast_kernel<64,64,false> in a built library:  python tools/loop_count.py variants/libuavsim_x.so [--list]"""
import collections, re, subprocess, sys
so = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", "-fun", "_Z23uavsim_step_fast_kernelILi64ELi64ELb0EEv7KParams13UavSimBuffersPK8ActEntryllidiPd", so],
                     capture_output=True, text=True).stdout
ins = []
for ln in out.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA.*`\(\.L_x_\d+\)|BRA\s+0x([0-9a-f]+)", t)
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)*0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            body = ins[addr[tgt]:i + 1]
            if sum("MUFU.EX2" in x for _, x in body) == 2 and (best is None or len(body) < len(best)):
                best = body
print(so, "total kernel instructions", len(ins))
if best:
    ops = collections.Counter((x.split()[1] if x.startswith("@") else x.split()[0]).split(".")[0] for _, x in best)
    print(" walk loop: %d instructions at 0x%x:" % (len(best), best[0][0]), " ".join("%s%d" % kv for kv in ops.most_common()))
    if "--list" in sys.argv:
        for a, x in best:
            print("   %04x  %s" % (a, x))
