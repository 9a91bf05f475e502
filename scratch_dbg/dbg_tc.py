import sys, numpy as np, torch
sys.path.insert(0, "tests")
from test_gpu_pmi_tensor import _env, _pmi
from marl_uavs_targets_tracking_b200 import default_config
n, m, E = 10, 10, int(sys.argv[1]) if len(sys.argv) > 1 else 300
cfg = default_config("MAAC-R", n, m)
pmi = _pmi()
envs = {}
for path in (1, 2):
    e = _env(n, m, cfg, E, seed=21); e.set_pmi_path(path); e.reset(cfg); envs[path] = e
for t in range(12):
    envs[1].random_actions(5, t); envs[2].random_actions(5, t)
    r = {}
    for path in (1, 2):
        _, rew4, _ = envs[path].step_device(cfg, pmi)
        r[path] = rew4[0].double().cpu().numpy()
    d = np.abs(r[1] - r[2])
    bad = np.argwhere(d > 2e-6)
    nb = envs[2]._nbr_bits.cpu().numpy() if hasattr(envs[2], "_nbr_bits") else None
    print("step", t, "max", d.max(), "nbad", len(bad), "of", d.size, "bad envs", sorted(set(bad[:, 0].tolist()))[:40])
    if len(bad): print("   first bad (env,uav,err):", [(int(a), int(b), float(d[a, b])) for a, b in bad[:12]])
