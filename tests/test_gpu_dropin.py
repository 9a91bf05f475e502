"""Drop-in proof: the reference's OWN driver code runs against the CUDA environment.

`baseline/_ref/src` is an unmodified copy of the reference (made by __graft_entry__.build(); tests skip without it).
`src/train.py:operate_epoch` (lines 142-196) -- the rollout loop of `python main.py` -- is imported as is and given

    env   = marl_uavs_targets_tracking_b200.Environment   (n_envs = 1: the reference's shapes)
    agent = the reference's ActorCritic                    (src/models/actor_critic.py:114-148)
    pmi   = the reference's PMINetwork for MAAC-R          (src/models/PMINet.py:20-72)

for all three methods with the shipped YAML files.  The recorded transitions are then replayed step by step in the
reference's own `Environment` (same initial state, same actions): next states, the four reward lists and the covered
count must agree within the contract (1e-5; integer count exact), so the numbers the learner saw are the reference's.
"""
import random

import numpy as np
import pytest
import torch

from gpu_util import max_scaled_err

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _reference():
    import baseline
    if not baseline.available():
        pytest.skip("baseline/_ref/src absent (build() copies it where /root/reference exists)")
    return baseline, baseline.import_reference()


@pytest.mark.parametrize("method", ["MAAC", "MAAC-G", "MAAC-R"])
def test_reference_operate_epoch_drives_the_cuda_environment(method):
    baseline, ref = _reference()
    from marl_uavs_targets_tracking_b200 import Environment
    cfg = baseline.load_yaml_config(method)
    cfg["devices"] = [torch.device("cuda:0")]
    e, ac = cfg["environment"], cfg["actor_critic"]
    n, m, T = e["n_uav"], e["m_targets"], 25
    random.seed(3); np.random.seed(3); torch.manual_seed(3)

    env = Environment(n_uav=n, m_targets=m, x_max=e["x_max"], y_max=e["y_max"], na=e["na"])   # src/main.py:55-59
    agent = ref.ActorCritic(state_dim=12, hidden_dim=ac["hidden_dim"], action_dim=e["na"], actor_lr=float(ac["actor_lr"]),
                            critic_lr=float(ac["critic_lr"]), gamma=float(ac["gamma"]), device=cfg["devices"][0])
    pmi = None
    if method == "MAAC-R":
        pmi = ref.PMINetwork(hidden_dim=cfg["pmi"]["hidden_dim"], b2_size=cfg["pmi"]["b2_size"])
        with torch.no_grad():  # BatchNorm statistics away from the identity, so folding is exercised
            for bn in (pmi.bn_comm, pmi.bn_obs, pmi.bn_boundary_state, pmi.bn1):
                bn.running_mean.normal_(0, 0.3)
                bn.running_var.uniform_(0.5, 1.5)

    env.reset(cfg)                                                                              # src/train.py:232
    st0 = {k: v.cpu().numpy().copy()[0] for k, v in env.get_state().items()}
    out = ref.train.operate_epoch(cfg, env, agent, pmi, T)                                      # src/train.py:238
    tr, ret, tt_ret, bp_ret, dup_ret, avg_cov, max_cov = out

    # shapes and types the reference's learner consumes (src/train.py:250-262, actor_critic.py:150-179)
    assert len(tr["states"]) == len(tr["actions"]) == len(tr["next_states"]) == len(tr["rewards"]) == T * n
    assert all(isinstance(s, np.ndarray) and s.shape == (12,) and s.dtype == np.float64 for s in tr["states"])
    assert all(isinstance(a, int) for a in tr["actions"])
    assert cfg["step"] == T
    a_loss, c_loss, td = agent.update(tr)                                                       # src/train.py:255
    assert torch.isfinite(a_loss) and torch.isfinite(c_loss) and td.shape[0] == T * n
    assert len(env.covered_target_num) == T and len(env.position["all_uav_xs"]) == T            # traces for main.py

    # replay in the reference's own environment
    renv = ref.environment.Environment(n_uav=n, m_targets=m, x_max=e["x_max"], y_max=e["y_max"], na=e["na"])
    renv.reset(cfg)
    for i, u in enumerate(renv.uav_list):
        u.x, u.y, u.h, u.a = float(st0["ux"][i]), float(st0["uy"][i]), float(st0["uh"][i]), int(st0["ua"][i])
    for j, t in enumerate(renv.target_list):
        t.x, t.y, t.h = float(st0["tx"][j]), float(st0["ty"][j]), float(st0["th"][j])
    acts = np.asarray(tr["actions"]).reshape(T, n)
    sums = np.zeros(4)
    covs = []
    worst = 0.0
    first = np.stack([u.get_local_state() for u in renv.uav_list])
    assert max_scaled_err(np.stack(tr["states"][:n]), first) <= TOL
    for t in range(T):
        ns, rew, cov = renv.step(cfg, pmi, [int(a) for a in acts[t]])
        got_ns = np.stack(tr["next_states"][t * n:(t + 1) * n])
        worst = max(worst, max_scaled_err(got_ns, np.stack(ns)))
        worst = max(worst, max_scaled_err(np.asarray(tr["rewards"][t * n:(t + 1) * n]), np.asarray(rew["rewards"])))
        if t + 1 < T:  # the states the policy saw at the next step are these next states
            assert np.array_equal(np.stack(tr["states"][(t + 1) * n:(t + 2) * n]), got_ns)
        assert env.covered_target_num[t] == cov, (t, "covered")
        covs.append(cov)
        for q, k in enumerate(("rewards", "target_tracking_reward", "boundary_punishment", "duplicate_tracking_punishment")):
            sums[q] += sum(rew[k])
    print(method, "worst %.1e" % worst)
    assert worst <= TOL
    ref_returns = sums / (T * n)
    assert np.allclose([ret, tt_ret, bp_ret, dup_ret], ref_returns, rtol=0, atol=TOL)
    assert avg_cov == pytest.approx(np.mean(covs)) and max_cov == np.max(covs)
    env.close()
