#!/usr/bin/env python
"""Generic step kernel at training batch sizes: ms per step of 65 536 environments of 10x10 (MAAC-G, random policy)
over 100 steps of the device loop, and the fraction of the HBM roofline.  python tools/generic_timing.py [envs [step path]]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from marl_uavs_targets_tracking_b200 import BatchedEnvironment, default_config
E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = m = 10
cfg = default_config("MAAC-G", n, m)
env = BatchedEnvironment(n, m, 2000, 2000, 12, n_envs=E, device="cuda:0", seed=3)
path = int(sys.argv[2]) if len(sys.argv) > 2 else 1
env.set_step_path(path)
env.reset(cfg)
env.run_random_policy(cfg, None, 5, 0, 10)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
env.run_random_policy(cfg, None, 5, 10, 100)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 100
peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6552.3) if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")) else 6552.3
byt = E * (124 * n + 48 * m + 4)
print("step path %d: %d envs %dx%d: %.5f ms per step (device loop: action draw + step), %.3f of the HBM roofline" % (path, E, n, m, ms, byt / (ms * 1e-3) / 1e9 / peak))
