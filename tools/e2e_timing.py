#!/usr/bin/env python
"""Host-buffer step of the headline workload (64x64, 65 536 envs): ms per step of uavsim_step_host and of the queued
variant (step_host_async / _wait, one step ahead) for several chunk counts.  python tools/e2e_timing.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from marl_uavs_targets_tracking_b200 import BatchedEnvironment, default_config
n = m = 64
E = 65536
cfg = default_config("MAAC-G", n, m)
env = BatchedEnvironment(n, m, 2000, 2000, 12, n_envs=E, device="cuda:0", seed=3)
env.reset(cfg)
NH = 6
h_act = torch.stack([env.random_actions(2, t).cpu() for t in range(NH)]).pin_memory()
outs = [(torch.empty((E, n, 12), dtype=torch.float32).pin_memory(), torch.empty((4, E, n), dtype=torch.float32).pin_memory(),
         torch.empty((E,), dtype=torch.int32).pin_memory()) for _ in range(2)]
K = 20
for chunks in (1, 2, 4, 8, 16):
    for i in range(2):
        env.step_host(cfg, None, h_act[i % NH], *outs[0], chunks=chunks)
    t0 = time.perf_counter()
    for i in range(K):
        env.step_host(cfg, None, h_act[i % NH], *outs[0], chunks=chunks)
    ts = (time.perf_counter() - t0) / K * 1e3
    t0 = time.perf_counter()
    tk = env.step_host_async(cfg, None, h_act[0], *outs[0], chunks=chunks)
    for i in range(1, K):
        tk2 = env.step_host_async(cfg, None, h_act[i % NH], *outs[i & 1], chunks=chunks)
        env.step_host_wait(tk)
        tk = tk2
    env.step_host_wait(tk)
    tp = (time.perf_counter() - t0) / K * 1e3
    print("chunks %2d: step_host %.3f ms per step, queued one ahead %.3f ms" % (chunks, ts, tp))
