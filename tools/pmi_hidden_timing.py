# PYTHONPATH=. python tools/pmi_hidden_timing.py -- ms per MAAC-R step at 64 x 64 x 16 384 for hidden 64 / 128 on the tensor (2) and CUDA-core (1) PMI kernels
import torch, time
from marl_uavs_targets_tracking_b200 import BatchedEnvironment, PMINetwork, default_config
n=m=64; E=16384
cfg=default_config("MAAC-R", n, m)
for hidden in (64, 128):
    torch.manual_seed(1)
    pmi=PMINetwork(hidden_dim=hidden).eval()
    for path in (2, 1):
        env=BatchedEnvironment(n,m,2000,2000,12,n_envs=E,device="cuda:0",seed=3)
        env.set_pmi_path(path)
        env.reset(cfg)
        for t in range(5):
            env.random_actions(1,t); env.step_device(cfg,pmi)
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(5,15):
            env.random_actions(1,t); env.step_device(cfg,pmi)
        e1.record(); torch.cuda.synchronize()
        print("hidden", hidden, "path", path, "ms/step", e0.elapsed_time(e1)/10)
        env.close()
