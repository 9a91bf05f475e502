"""MAAC-R reward through both PMI kernels: fp32 CUDA cores (path 1) and tcgen05 tensor cores with split fp16 hi + lo
operands (path 2; hidden sizes 128 = the shipped YAML files and 64 = the default of the reference's PMINetwork
constructor), against the CPU oracle (src/agent/uav.py:262-291, src/models/PMINet.py:41-72)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import golden_config, golden_pmi_module, max_scaled_err, oracle_params_from_config, oracle_pmi_from_module

pytestmark = pytest.mark.gpu
TOL_PMI = 1e-5      # contract
TOL_TC = 2e-6       # what the split-operand path should achieve (fp32-class products)


def _env(n, m, cfg, E, **kw):
    from marl_uavs_targets_tracking_b200 import BatchedEnvironment
    e = cfg["environment"]
    return BatchedEnvironment(n, m, e["x_max"], e["y_max"], e["na"], n_envs=E, device="cuda:0", **kw)


def _pmi(seed=1, hidden=128):
    from marl_uavs_targets_tracking_b200 import PMINetwork
    torch.manual_seed(seed)
    pmi = PMINetwork(hidden_dim=hidden)
    for bn in (pmi.bn_comm, pmi.bn_obs, pmi.bn_boundary_state, pmi.bn1):
        bn.running_mean.normal_(0, 0.3)
        bn.running_var.uniform_(0.5, 1.5)
    return pmi.eval()


@pytest.mark.parametrize("hidden", [128, 64])
@pytest.mark.parametrize("n,m,E,T", [(10, 10, 300, 40), (64, 64, 9, 12), (5, 3, 1, 10), (32, 32, 40, 20), (90, 4, 3, 6)])
def test_tensor_path_matches_oracle_and_cuda_core_path(oracle, n, m, E, T, hidden):
    from marl_uavs_targets_tracking_b200 import default_config
    cfg = default_config("MAAC-R", n, m)
    pmi = _pmi(hidden=hidden)
    envs = {}
    for path in (1, 2):
        e = _env(n, m, cfg, E, seed=21)
        e.set_pmi_path(path)
        e.reset(cfg)
        envs[path] = e
    P = oracle_params_from_config(cfg, n, m)
    opmi = oracle_pmi_from_module(pmi)
    st = {k: np.ascontiguousarray(v.cpu().numpy()) for k, v in envs[1].get_state().items()}
    worst = {1: 0.0, 2: 0.0, "1v2": 0.0}
    for t in range(T):
        a = envs[1].random_actions(5, t).cpu().numpy().copy()
        envs[2].random_actions(5, t)
        ref = oracle.step_batch(P, 2, float(cfg["cooperative"]), opmi, st, a, nthreads=8)
        r = {}
        for path in (1, 2):
            _, rew4, _ = envs[path].step_device(cfg, pmi)
            r[path] = rew4[0].double().cpu().numpy()
            worst[path] = max(worst[path], max_scaled_err(r[path], ref["rew4"][0]))
        worst["1v2"] = max(worst["1v2"], float(np.abs(r[1] - r[2]).max()))
    print(n, m, E, {k: "%.1e" % v for k, v in worst.items()})
    assert worst[1] <= TOL_TC and worst[2] <= TOL_TC, worst
    s1, s2 = envs[1].episode_stats(), envs[2].episode_stats()
    assert abs(s1["rewards"] - s2["rewards"]) <= 1e-6 * max(1.0, abs(s1["rewards"])) + 1e-4
    for e in envs.values():
        e.close()


def test_tensor_path_on_reference_golden():
    g = load_golden("s64_pmi_s42")
    cfg = golden_config(g)
    pmi = golden_pmi_module(g)
    n = m = 64
    T = g["actions"].shape[0]
    env = _env(n, m, cfg, 2, seed=0)
    env.set_pmi_path(2)
    rep = lambda a: np.broadcast_to(np.asarray(a), (2,) + np.asarray(a).shape).copy()  # noqa: E731
    env.set_state(cfg, *(rep(g[k + "0"]) for k in ("ux", "uy", "uh", "ua", "tx", "ty", "th")))
    worst = 0.0
    for t in range(T):
        _, rew, _ = env.step(cfg, pmi, torch.as_tensor(rep(g["actions"][t]), device="cuda:0"))
        worst = max(worst, max_scaled_err(rew["rewards"][0].double().cpu().numpy(), g["rewards"][t]))
    print("golden s64_pmi tensor path", worst)
    assert worst <= TOL_TC
    env.close()


def test_tensor_path_unsupported_sizes_fall_back_or_fail_loudly():
    from marl_uavs_targets_tracking_b200 import PMINetwork, UavSimError, default_config
    cfg = default_config("MAAC-R", 10, 10)
    env = _env(10, 10, cfg, 8)
    env.reset(cfg)
    torch.manual_seed(0)
    small = PMINetwork(hidden_dim=32).eval()
    env.random_actions(1, 0)
    env.step_device(cfg, small)            # automatic: hidden 32 -> CUDA-core kernel
    with pytest.raises(UavSimError):
        env.set_pmi_path(2)                # tensor path demanded but hidden is neither 64 nor 128
    env.close()


def test_tensor_path_serves_swarms_the_cuda_core_kernel_cannot_hold(oracle):
    """n = 128 (UAVSIM_MAX_UAV): 16 256 ordered pairs per environment exceed the CUDA-core kernel's 8 192-row pair
    buffer; the tensor path takes the shape, and asking for the CUDA-core path fails loudly instead of falling back."""
    from marl_uavs_targets_tracking_b200 import UavSimError, default_config
    n, m, E, T = 128, 5, 3, 4
    cfg = default_config("MAAC-R", n, m)
    pmi = _pmi()
    env = _env(n, m, cfg, E, seed=21)
    env.reset(cfg)
    P = oracle_params_from_config(cfg, n, m)
    opmi = oracle_pmi_from_module(pmi)
    st = {k: np.ascontiguousarray(v.cpu().numpy()) for k, v in env.get_state().items()}
    worst = 0.0
    for t in range(T):
        a = env.random_actions(5, t).cpu().numpy().copy()
        ref = oracle.step_batch(P, 2, float(cfg["cooperative"]), opmi, st, a, nthreads=8)
        _, rew4, _ = env.step_device(cfg, pmi)
        worst = max(worst, max_scaled_err(rew4[0].double().cpu().numpy(), ref["rew4"][0]))
    print("n=128 tensor path", worst)
    assert worst <= TOL_TC
    env.set_pmi_path(1)
    with pytest.raises(UavSimError):
        env.step_device(cfg, pmi)
    env.close()


def test_configs2_batch_sampled_against_the_oracle(oracle):
    """BASELINE.json configs[2] at full size (default scenario, 16 384 environments, MAAC-R, hidden 128): 64 environments
    spread over the batch replayed in the CPU oracle for 80 steps; the reward goes through the tensor-core kernel."""
    from marl_uavs_targets_tracking_b200 import default_config
    n = m = 10
    E, T = 16384, 80
    cfg = default_config("MAAC-R", n, m)
    pmi = _pmi(seed=4)
    env = _env(n, m, cfg, E, seed=77)
    env.reset(cfg)
    ids = np.unique(np.concatenate([np.arange(0, E, 263), [E - 1, 7399, 7400]]))[:64]
    idx = torch.as_tensor(ids, device="cuda:0")
    P = oracle_params_from_config(cfg, n, m)
    opmi = oracle_pmi_from_module(pmi)
    st = {k: np.ascontiguousarray(v[idx].cpu().numpy()) for k, v in env.get_state().items()}
    worst = 0.0
    for t in range(T):
        a = env.random_actions(3, t)[idx].cpu().numpy().copy()
        obs, rew4, cov = env.step_device(cfg, pmi)
        ref = oracle.step_batch(P, 2, float(cfg["cooperative"]), opmi, st, a, nthreads=8)
        assert np.array_equal(cov[idx].cpu().numpy(), ref["covered"]), t
        worst = max(worst, max_scaled_err(rew4[:, idx].double().cpu().numpy(), ref["rew4"]),
                    max_scaled_err(obs[idx].double().cpu().numpy(), ref["obs"]))
    print("configs[2] sample:", worst)
    assert worst <= TOL_TC
    env.close()
