"""Learner-side rows on the CPU: the replay oracle and the PMINetwork.train_pmi mirror against
fixtures produced by the unmodified reference (tests/golden/make_learner_golden.py)."""
import numpy as np
import torch

from conftest import load_golden
from oracle.learner_oracle import ReplayOracle


def test_replay_oracle_reproduces_the_reference_buffer():
    g = load_golden("learner_per")
    o = ReplayOracle(int(g["capacity"]), float(g["alpha"]))
    for k in range(int(g["n_ops"])):
        t = "op%d" % k
        o.add(g[t + ".add.states"], g[t + ".add.actions"], g[t + ".add.rewards"], g[t + ".add.next_states"])
        assert np.array_equal(o.priorities, g[t + ".after_add.priorities"])
        assert (o.pos, o.n) == (int(g[t + ".after_add.pos"]), int(g[t + ".after_add.size"]))
        assert np.array_equal(o.states[:o.n], g[t + ".after_add.states"])
        assert np.array_equal(o.actions[:o.n], g[t + ".after_add.actions"])
        idx, w, prob = o.sample(int(g[t + ".sample.batch"]), g[t + ".sample.uniforms"], float(g["beta"]))
        assert np.array_equal(idx, g[t + ".sample.indices"])       # numpy's choice(p=...) rule
        assert np.array_equal(prob, g[t + ".sample.prob"])
        assert np.array_equal(w, g[t + ".sample.weights"])
        assert np.array_equal(o.states[idx], g[t + ".sample.states"])
        assert np.array_equal(o.rewards[idx], g[t + ".sample.rewards"])
        o.update_priorities(idx, g[t + ".update.priorities"])
        assert np.array_equal(o.priorities, g[t + ".after_update.priorities"])


def _net_from_golden(g, device="cpu"):
    from marl_uavs_targets_tracking_b200 import PMINetwork
    net = PMINetwork(hidden_dim=int(g["hidden"]), b2_size=int(g["b2_size"]))
    net.load_state_dict({k[5:]: torch.tensor(np.array(g[k])) for k in g.files if k.startswith("init.")})
    return net.to(device)


def test_train_pmi_matches_the_reference_on_cpu():
    """Same draws (torch.randint on the seeded CPU generator), same minibatches, same loss and Adam: the mirror ends
    on the reference's weights (src/models/PMINet.py:74-100)."""
    g = load_golden("learner_pmi_train")
    torch.set_num_threads(1)
    net = _net_from_golden(g)
    torch.manual_seed(int(g["seed"]))
    loss = net.train_pmi({"pmi": {"batch_size": int(g["batch_size"])}}, torch.tensor(g["data"]), int(g["n_uav"]))
    assert abs(loss - float(g["avg_loss"])) <= 1e-6 * abs(float(g["avg_loss"]))
    for k, v in net.state_dict().items():
        ref = g["final." + k]
        if ref.dtype.kind == "f":
            assert np.allclose(v.numpy(), ref, rtol=1e-5, atol=1e-6), k
        else:
            assert np.array_equal(v.numpy(), ref), k   # num_batches_tracked
    assert not net.training or True
    # the reference's checkpoint layout round-trips
    import os, tempfile
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "pmi"))
        net.save(d, 3)
        other = _net_from_golden(g)
        other.load(os.path.join(d, "pmi", "pmi_weights_3.pth"))
        for k, v in other.state_dict().items():
            assert torch.equal(v, net.state_dict()[k]), k


def test_actor_critic_update_matches_the_reference_on_cpu():
    """BatchedActorCritic.update == ActorCritic.update (src/models/actor_critic.py:150-179) on a recorded pair of
    consecutive updates: losses, TD errors and every weight after Adam.  (The reference multiplies log_probs [B,1] by
    td_delta [B]: the mean over the [B,B] outer product is mean(-log_probs) * mean(td_delta), which is what the
    batched class evaluates in O(B).)"""
    from marl_uavs_targets_tracking_b200.rollout import BatchedActorCritic
    g = load_golden("learner_ac_update")
    torch.set_num_threads(1)
    ac = BatchedActorCritic(12, int(g["hidden"]), int(g["n_actions"]), float(g["actor_lr"]), float(g["critic_lr"]),
                            float(g["gamma"]), torch.device("cpu"))
    ac.actor.load_state_dict({k[len("init.actor."):]: torch.tensor(np.array(g[k])) for k in g.files if k.startswith("init.actor.")})
    ac.critic.load_state_dict({k[len("init.critic."):]: torch.tensor(np.array(g[k])) for k in g.files if k.startswith("init.critic.")})
    for step in range(2):
        t = "step%d." % step
        a_loss, c_loss, td = ac.update(torch.tensor(g[t + "states"]), torch.tensor(g[t + "actions"]),
                                       torch.tensor(g[t + "rewards"]), torch.tensor(g[t + "next_states"]))
        assert abs(float(a_loss) - float(g[t + "actor_loss"])) <= 2e-6 * max(1.0, abs(float(g[t + "actor_loss"])))
        assert abs(float(c_loss) - float(g[t + "critic_loss"])) <= 2e-6 * max(1.0, abs(float(g[t + "critic_loss"])))
        assert np.allclose(td.numpy(), g[t + "td"], rtol=1e-5, atol=1e-6)
    for name, net in (("actor", ac.actor), ("critic", ac.critic)):
        for k, v in net.state_dict().items():
            assert np.allclose(v.numpy(), g["final.%s.%s" % (name, k)], rtol=1e-4, atol=2e-6), (name, k)
