#!/usr/bin/env python
"""BASELINE configs[2] (10x10, 16 384 environments, MAAC-R, hidden 128): ms per step on the tensor (2) and CUDA-core (1)
PMI kernels, and of the step kernel alone (MAAC, no PMI launch).  python tools/pmi_small_timing.py"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from marl_uavs_targets_tracking_b200 import BatchedEnvironment, PMINetwork, default_config
n = m = 10
E = 16384
torch.manual_seed(1)
pmi = PMINetwork(hidden_dim=128).eval()
for method, path in (("MAAC-R", 2), ("MAAC-R", 1), ("MAAC", 0)):
    cfg = default_config(method, n, m)
    env = BatchedEnvironment(n, m, 2000, 2000, 12, n_envs=E, device="cuda:0", seed=3)
    if path:
        env.set_pmi_path(path)
    env.reset(cfg)
    net = pmi if method == "MAAC-R" else None
    env.run_random_policy(cfg, net, 1, 0, 10)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    env.run_random_policy(cfg, net, 1, 10, 100)
    b.record()
    torch.cuda.synchronize()
    print("%s pmi path %d: %.4f ms per step (device loop)" % (method, path, a.elapsed_time(b) / 100))
    env.close()
