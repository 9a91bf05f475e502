"""Learner-side rows of SURVEY.md section 8f on the GPU, through the C ABI: device prioritized replay
(csrc/replay.cuh) against the reference-generated fixture and the numpy oracle, PMI training on the device."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle.learner_oracle import ReplayOracle
from philox_ref import philox4x32_10, u53

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _buffer(capacity, alpha=0.6, **kw):
    from marl_uavs_targets_tracking_b200 import PrioritizedReplayBuffer
    return PrioritizedReplayBuffer(capacity, alpha=alpha, device=DEV, **kw)


def _t(a, dtype=None):
    return torch.as_tensor(np.asarray(a), dtype=dtype).to(DEV)


def test_device_replay_replays_the_reference_fixture():
    g = load_golden("learner_per")
    buf = _buffer(int(g["capacity"]), float(g["alpha"]))
    beta = float(g["beta"])
    for k in range(int(g["n_ops"])):
        t = "op%d" % k
        buf.add({"states": _t(g[t + ".add.states"]), "actions": _t(g[t + ".add.actions"]),
                 "rewards": _t(g[t + ".add.rewards"], torch.float32), "next_states": _t(g[t + ".add.next_states"])})
        ex = buf.export()
        assert np.array_equal(ex["priorities"].numpy(), g[t + ".after_add.priorities"])            # bit-exact
        assert (buf.pos, buf.size()) == (int(g[t + ".after_add.pos"]), int(g[t + ".after_add.size"]))
        assert np.array_equal(ex["states"].numpy(), g[t + ".after_add.states"])
        assert np.array_equal(ex["next_states"].numpy(), g[t + ".after_add.next_states"])
        assert np.array_equal(ex["actions"].numpy(), g[t + ".after_add.actions"])
        assert np.array_equal(ex["rewards"].numpy(), g[t + ".after_add.rewards"].astype(np.float32))
        sample, idx, w = buf.sample(int(g[t + ".sample.batch"]), beta, uniforms=_t(g[t + ".sample.uniforms"]))
        assert np.array_equal(idx.cpu().numpy(), g[t + ".sample.indices"])                         # index work: exact
        prob = buf.export()["probabilities"].numpy()
        assert np.allclose(prob, g[t + ".sample.prob"], rtol=3e-7, atol=0)                         # powf vs numpy pow
        assert np.allclose(w.cpu().numpy(), g[t + ".sample.weights"], rtol=2e-6, atol=0)
        assert np.array_equal(sample["states"].cpu().numpy(), g[t + ".sample.states"])
        assert np.array_equal(sample["actions"].cpu().numpy(), g[t + ".sample.actions"])
        assert np.array_equal(sample["rewards"].cpu().numpy(), g[t + ".sample.rewards"].astype(np.float32))
        buf.update_priorities(idx, _t(g[t + ".update.priorities"]))
        assert np.array_equal(buf.export()["priorities"].numpy(), g[t + ".after_update.priorities"])


def test_device_replay_matches_the_oracle_at_size():
    """100 000-slot ring, batches that wrap, repeated indices in update_priorities (last write wins), two-level scan."""
    rng = np.random.RandomState(0)
    C, D = 100_000, 12
    buf, o = _buffer(C), ReplayOracle(C)
    for it, k in enumerate((30_000, 50_000, 45_000, 5)):
        s, s2 = rng.randn(k, D).astype(np.float32), rng.randn(k, D).astype(np.float32)
        a, r = rng.randint(0, 12, k).astype(np.int32), rng.randn(k).astype(np.float32)
        buf.add({"states": _t(s), "actions": _t(a), "rewards": _t(r), "next_states": _t(s2)})
        # the oracle inserts one by one with an O(C) max each: feed it the batched equivalent
        maxp = o.priorities.max() if o.n else 1.0
        slots = (o.pos + np.arange(k)) % C
        o.states[slots], o.next_states[slots], o.actions[slots], o.rewards[slots] = s, s2, a, r
        o.priorities[slots] = maxp
        o.n, o.pos = min(o.n + k, C), (o.pos + k) % C
        assert (buf.pos, buf.size()) == (o.pos, o.n)
        u = rng.random_sample(4096)
        sample, idx, w = buf.sample(4096, 0.4, uniforms=_t(u))
        oi, ow, oprob = o.sample(4096, u, 0.4)
        prob = buf.export()["probabilities"].numpy()
        assert np.allclose(prob, oprob, rtol=5e-7, atol=0)
        gi = idx.cpu().numpy()
        # index work is exact: numpy's rule applied to the probabilities the device computed gives the same indices
        cdf = np.cumsum(prob.astype(np.float64))
        cdf /= cdf[-1]
        assert np.array_equal(gi, np.searchsorted(cdf, u, side="right"))
        # against numpy's own float32 pow the probabilities differ in the last bit (3e-7), which shifts the CDF by
        # ~1e-8: a draw lands in the neighbouring slot when its uniform is that close to a boundary (spacing 1e-5)
        assert (gi != oi).mean() <= 5e-3 and np.abs(gi - oi).max() <= 1
        same = gi == oi
        assert np.allclose(w.cpu().numpy()[same], ow[same], rtol=5e-6, atol=0)
        assert np.array_equal(sample["states"].cpu().numpy(), o.states[gi])
        assert np.array_equal(sample["next_states"].cpu().numpy(), o.next_states[gi])
        assert np.array_equal(sample["actions"].cpu().numpy(), o.actions[gi])
        newp = (np.abs(rng.randn(4096)) + 1e-3).astype(np.float32)
        dup = rng.randint(0, 4096, 4096)          # force many repeated indices
        gi2 = gi[dup]
        buf.update_priorities(_t(gi2), _t(newp))
        o.update_priorities(gi2, newp)
        assert np.array_equal(buf.export()["priorities"].numpy(), o.priorities), it


def test_device_replay_philox_draws_and_statistics():
    C = 4096
    buf = _buffer(C, seed=77)
    rng = np.random.RandomState(1)
    buf.add({"states": _t(rng.randn(C, 12).astype(np.float32)), "actions": _t(np.arange(C, dtype=np.int32)),
             "rewards": _t(np.zeros(C, np.float32)), "next_states": _t(np.zeros((C, 12), np.float32))})
    pri = (rng.rand(C).astype(np.float32) + 0.05)
    buf.update_priorities(torch.arange(C), _t(pri))
    o = ReplayOracle(C)
    o.priorities[:] = pri
    o.n, o.pos = C, 0
    # Philox path: uniforms are philox_u53 of counter (k, call number) -- reproduce on the host
    _, idx, _ = buf.sample(C, 0.4)
    x = philox4x32_10(np.arange(C), 0, 0, 0, 77)   # counter = (sample index lo, hi, call number lo, hi)
    u = u53(x[0], x[1])
    oi, _, _ = o.sample(C, u, 0.4)
    gi = idx.cpu().numpy()
    assert (gi != oi).mean() <= 5e-3 and np.abs(gi - oi).max() <= 1
    _, idx2, _ = buf.sample(C, 0.4)           # next call, next counter: different draws
    assert (idx2.cpu().numpy() != gi).mean() > 0.9
    # frequencies follow priority^alpha
    counts = np.zeros(C)
    for _ in range(40):
        _, i, _ = buf.sample(C, 0.4)
        counts += np.bincount(i.cpu().numpy(), minlength=C)
    p = o.probabilities().astype(np.float64)
    expect = p * counts.sum()
    chi2 = ((counts - expect) ** 2 / expect).sum()
    assert abs(chi2 - (C - 1)) < 6 * np.sqrt(2 * (C - 1))


def test_device_replay_edge_cases():
    from marl_uavs_targets_tracking_b200 import UavSimError
    buf = _buffer(8)
    sample, idx, w = buf.sample(4)
    assert sample == dict(states=[], actions=[], rewards=[], next_states=[]) and idx is None and w is None
    one = {"states": torch.ones(1, 12, device=DEV), "actions": torch.tensor([3], device=DEV),
           "rewards": torch.tensor([0.5], device=DEV), "next_states": torch.zeros(1, 12, device=DEV)}
    buf.add(one)
    sample, idx, w = buf.sample(4)                      # min(batch, size) = 1 draw
    assert idx.tolist() == [0] and w.tolist() == [1.0] and sample["actions"].tolist() == [3]
    assert buf.export()["priorities"][0] == 1.0         # empty buffer -> priority 1.0 (train.py:89)
    with pytest.raises(IndexError):
        buf.update_priorities([9], [1.0])
    with pytest.raises(UavSimError):
        from marl_uavs_targets_tracking_b200 import PrioritizedReplayBuffer
        PrioritizedReplayBuffer(8, device="cpu")


def test_train_pmi_on_the_device_and_reward_refresh():
    """train_pmi on cuda reaches the reference's weights to fp32 tolerance, and the environment's PMI reward follows the
    retrained weights without an explicit re-upload (version counters)."""
    from test_oracle_learner import _net_from_golden
    g = load_golden("learner_pmi_train")
    net = _net_from_golden(g, DEV)
    torch.manual_seed(int(g["seed"]))
    loss = net.train_pmi({"pmi": {"batch_size": int(g["batch_size"])}}, torch.tensor(g["data"]), int(g["n_uav"]))
    assert abs(loss - float(g["avg_loss"])) <= 1e-4 * abs(float(g["avg_loss"]))
    # Weight-by-weight comparison across devices is ill-posed here: the biases in front of a BatchNorm have zero true
    # gradient, so Adam turns their rounding noise into +-lr steps.  Every parameter stays inside Adam's step bound
    # and the trained FUNCTION agrees with the reference's.
    from marl_uavs_targets_tracking_b200 import PMINetwork as Mirror
    steps = int(g["b2_size"]) // int(g["batch_size"])
    ref_net = Mirror(hidden_dim=int(g["hidden"]))
    ref_net.load_state_dict({k[6:]: torch.tensor(np.array(g[k])) for k in g.files if k.startswith("final.")})
    for k, v in net.state_dict().items():
        if g["final." + k].dtype.kind == "f" and "running" not in k:
            assert np.abs(v.cpu().numpy() - g["final." + k]).max() <= 2.2 * steps * 1e-3, k
    x = torch.tensor(g["data"])
    with torch.no_grad():
        out_ref = ref_net.eval()(x)
        out_dev = net.eval()(x.to(DEV)).cpu()
    assert (out_ref - out_dev).abs().max() <= 5e-3 * max(1.0, float(out_ref.abs().max()))

    from marl_uavs_targets_tracking_b200 import BatchedEnvironment, PMINetwork, default_config
    cfg = default_config("MAAC-R", 10, 10)
    e = cfg["environment"]
    env = BatchedEnvironment(10, 10, e["x_max"], e["y_max"], e["na"], n_envs=64, device=DEV, seed=3)
    torch.manual_seed(0)
    pmi = PMINetwork(hidden_dim=128, b2_size=512).to(DEV)
    env.reset(cfg)
    states = []
    for t in range(8):
        env.random_actions(9, t)
        obs, rew4, _ = env.step_device(cfg, pmi.eval())
        states.append(obs.reshape(-1, 12).clone())
    pmi.train_pmi({"pmi": {"batch_size": 128}}, torch.cat(states), 10)
    env.random_actions(9, 8)
    _, rew_after, _ = env.step_device(cfg, pmi.eval())      # same handle: must notice the new weights
    # the trajectory does not depend on the PMI weights: a fresh environment stepped with the trained network only
    env2 = BatchedEnvironment(10, 10, e["x_max"], e["y_max"], e["na"], n_envs=64, device=DEV, seed=3)
    env2.reset(cfg)
    for t in range(9):
        env2.random_actions(9, t)
        _, rew_new, _ = env2.step_device(cfg, pmi.eval())
    assert torch.equal(rew_after[0], rew_new[0])
    # ... and the untrained network gives a different reward on that step
    torch.manual_seed(0)
    old = PMINetwork(hidden_dim=128, b2_size=512).to(DEV)
    env3 = BatchedEnvironment(10, 10, e["x_max"], e["y_max"], e["na"], n_envs=64, device=DEV, seed=3)
    env3.reset(cfg)
    for t in range(9):
        env3.random_actions(9, t)
        _, rew_old, _ = env3.step_device(cfg, old.eval())
    assert (rew_old[0] - rew_new[0]).abs().max() > 1e-6


@pytest.mark.parametrize("case", range(6))
def test_device_replay_random_operation_sequences(case):
    """Random capacities, alphas, batch sizes (larger than the buffer, larger than the capacity), repeated indices:
    every add / sample / update_priorities call against the numpy oracle of the reference class."""
    rng = np.random.RandomState(50 + case)
    C = int(rng.choice([1, 7, 64, 1000, 4097, 20000]))
    alpha, beta = float(rng.uniform(0.1, 1.0)), float(rng.uniform(0.1, 1.0))
    buf, o = _buffer(C, alpha), ReplayOracle(C, alpha)
    for op in range(14):
        k = int(rng.choice([1, 3, C // 2 + 1, C, C + 5, 2 * C + 3]))
        s, s2 = rng.randn(k, 12).astype(np.float32), rng.randn(k, 12).astype(np.float32)
        a, r = rng.randint(0, 12, k).astype(np.int32), rng.randn(k).astype(np.float32)
        buf.add({"states": _t(s), "actions": _t(a), "rewards": _t(r), "next_states": _t(s2)})
        o.add(s, a, r, s2)  # the reference's one-by-one loop
        assert (buf.pos, buf.size()) == (o.pos, o.n)
        ex = buf.export()
        assert np.array_equal(ex["priorities"].numpy(), o.priorities)
        assert np.array_equal(ex["states"].numpy(), o.states[:o.n]) and np.array_equal(ex["actions"].numpy(), o.actions[:o.n])
        b = int(rng.choice([1, 5, o.n, o.n + 9, 3 * C]))
        u = rng.random_sample(min(b, o.n))
        sample, idx, w = buf.sample(b, beta, uniforms=_t(u))
        oi, ow, oprob = o.sample(b, u, beta)
        gi = idx.cpu().numpy()
        assert gi.shape == oi.shape
        prob = buf.export()["probabilities"].numpy()
        cdf = np.cumsum(prob.astype(np.float64))
        cdf /= cdf[-1]
        assert np.array_equal(gi, np.searchsorted(cdf, u, side="right"))          # exact given the device's probabilities
        assert np.allclose(prob, oprob, rtol=1e-6, atol=0)
        assert (gi != oi).mean() <= 0.02 and np.abs(gi - oi).max() <= 1
        same = gi == oi
        assert np.allclose(w.cpu().numpy()[same], ow[same], rtol=1e-5, atol=0)
        assert np.array_equal(sample["next_states"].cpu().numpy(), o.next_states[gi])
        newp = (np.abs(rng.randn(len(gi))) + 1e-3).astype(np.float32)
        order = rng.permutation(len(gi))                                           # repeated indices in random order
        buf.update_priorities(_t(gi[order]), _t(newp[order]))
        o.update_priorities(gi[order], newp[order])
        assert np.array_equal(buf.export()["priorities"].numpy(), o.priorities), (case, op)
