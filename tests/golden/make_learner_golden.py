#!/usr/bin/env python
"""Golden vectors for the learner-side rows of SURVEY.md section 8f, by EXECUTING the unmodified reference:

* `learner_pmi_train.npz`  -- `PMINetwork.train_pmi` (src/models/PMINet.py:74-100): initial weights, training data,
  the index draws, the returned average loss and the weights / BatchNorm statistics after the call.
* `learner_ac_update.npz` -- `ActorCritic.update` (src/models/actor_critic.py:150-179): two consecutive updates.
* `learner_per.npz`        -- `PrioritizedReplayBuffer` (src/train.py:73-139): a recorded sequence of add / sample /
  update_priorities calls with the priorities after every call, and for every sample call the probabilities, the
  indices numpy drew, the uniforms that produce them under numpy's own inverse-CDF rule and the importance weights.

Runs only in the build container (needs /root/reference).  Fixtures hold numbers only.

usage:  python tests/golden/make_learner_golden.py [--ref /root/reference/src]
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference(ref):
    sys.path.insert(0, ref)
    # src/train.py imports matplotlib / imageio through utils.draw_util: stub the module
    stub = types.ModuleType("utils.draw_util")
    stub.draw_animation = lambda *a, **k: None
    stub.plot_reward_curve = lambda *a, **k: None
    import utils  # noqa: F401
    sys.modules["utils.draw_util"] = stub
    for name in ("tensorboard", "torch.utils.tensorboard"):
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            m.SummaryWriter = object
            sys.modules[name] = m
    from models.PMINet import PMINetwork
    from models.actor_critic import ActorCritic
    import train as ref_train
    return PMINetwork, ref_train.PrioritizedReplayBuffer, ActorCritic


def actor_critic_case(ActorCritic):
    """ActorCritic.update (src/models/actor_critic.py:150-179): weights before, a batch, losses, TD errors, weights after
    two consecutive updates (Adam state carries over)."""
    torch.manual_seed(21)
    torch.set_num_threads(1)
    B, H, A = 96, 24, 12
    ac = ActorCritic(12, H, A, 1e-3, 2e-3, 0.95, torch.device("cpu"))
    out = {"hidden": H, "n_actions": A, "actor_lr": 1e-3, "critic_lr": 2e-3, "gamma": 0.95}
    for k, v in ac.actor.state_dict().items():
        out["init.actor." + k] = v.detach().clone().numpy()
    for k, v in ac.critic.state_dict().items():
        out["init.critic." + k] = v.detach().clone().numpy()
    rng = np.random.RandomState(4)
    for step in range(2):
        batch = {"states": [rng.randn(12).astype(np.float32) for _ in range(B)],
                 "actions": [int(a) for a in rng.randint(0, A, B)],
                 "rewards": [float(r) for r in rng.uniform(-1, 1, B)],
                 "next_states": [rng.randn(12).astype(np.float32) for _ in range(B)]}
        a_loss, c_loss, td = ac.update(batch)
        out["step%d.states" % step] = np.array(batch["states"])
        out["step%d.actions" % step] = np.array(batch["actions"], np.int64)
        out["step%d.rewards" % step] = np.array(batch["rewards"], np.float32)
        out["step%d.next_states" % step] = np.array(batch["next_states"])
        out["step%d.actor_loss" % step] = np.float64(a_loss.item())
        out["step%d.critic_loss" % step] = np.float64(c_loss.item())
        out["step%d.td" % step] = td.detach().numpy()
    for k, v in ac.actor.state_dict().items():
        out["final.actor." + k] = v.detach().numpy()
    for k, v in ac.critic.state_dict().items():
        out["final.critic." + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "learner_ac_update.npz"), **out)
    print("learner_ac_update: losses", out["step1.actor_loss"], out["step1.critic_loss"])


def pmi_train_case(PMINetwork):
    torch.manual_seed(11)
    torch.set_num_threads(1)
    H, b2, bs, n_uav, T = 32, 320, 64, 6, 40
    net = PMINetwork(hidden_dim=H, b2_size=b2)
    init = {k: v.detach().clone().numpy() for k, v in net.state_dict().items()}
    data = torch.randn(T * n_uav, 12) * 0.7
    torch.manual_seed(5)  # the draws inside train_pmi start here
    loss = net.train_pmi({"pmi": {"batch_size": bs}}, data.clone(), n_uav)
    out = {"hidden": H, "b2_size": b2, "batch_size": bs, "n_uav": n_uav, "seed": 5, "data": data.numpy(),
           "avg_loss": np.float64(loss)}
    for k, v in init.items():
        out["init." + k] = v
    for k, v in net.state_dict().items():
        out["final." + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "learner_pmi_train.npz"), **out)
    print("learner_pmi_train: avg_loss", loss)


def per_case(PER):
    rng = np.random.RandomState(3)
    cap = 50
    buf = PER(cap, alpha=0.6)
    log = {"capacity": cap, "alpha": 0.6, "beta": 0.4}
    ops = []
    step = 0

    def transitions(k):
        return {"states": [rng.randn(12).astype(np.float32) for _ in range(k)],
                "actions": [int(rng.randint(0, 12)) for _ in range(k)],
                "rewards": [float(rng.randn()) for _ in range(k)],
                "next_states": [rng.randn(12).astype(np.float32) for _ in range(k)]}

    def record_storage(tag):
        log[tag + ".priorities"] = buf.priorities.copy()
        log[tag + ".pos"] = np.int64(buf.pos)
        log[tag + ".size"] = np.int64(buf.size())
        log[tag + ".states"] = np.array([e[0] for e in buf.buffer], np.float32)
        log[tag + ".actions"] = np.array([e[1] for e in buf.buffer], np.int32)
        log[tag + ".rewards"] = np.array([e[2] for e in buf.buffer], np.float64)
        log[tag + ".next_states"] = np.array([e[3] for e in buf.buffer], np.float32)

    for k_add, k_sample in ((18, 8), (20, 16), (25, 32), (7, 64), (60, 20)):
        tag = "op%d" % step
        tr = transitions(k_add)
        log[tag + ".add.states"] = np.array(tr["states"])
        log[tag + ".add.actions"] = np.array(tr["actions"], np.int32)
        log[tag + ".add.rewards"] = np.array(tr["rewards"], np.float64)
        log[tag + ".add.next_states"] = np.array(tr["next_states"])
        buf.add(tr)
        record_storage(tag + ".after_add")
        np.random.seed(100 + step)
        sample, idx, w = buf.sample(k_sample, beta=0.4)
        # numpy's Generator-free choice(p=...) is: cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, uniforms, 'right')
        np.random.seed(100 + step)
        size = buf.size()
        pri = buf.priorities if size == cap else buf.priorities[:buf.pos]
        prob = pri ** buf.alpha
        prob /= prob.sum()
        uniforms = np.random.random_sample(len(idx))
        cdf = np.cumsum(prob.astype(np.float64))
        cdf /= cdf[-1]
        assert np.array_equal(np.searchsorted(cdf, uniforms, side="right"), idx)
        log[tag + ".sample.batch"] = np.int64(k_sample)
        log[tag + ".sample.prob"] = prob
        log[tag + ".sample.uniforms"] = uniforms
        log[tag + ".sample.indices"] = np.asarray(idx, np.int64)
        log[tag + ".sample.weights"] = np.asarray(w, np.float32)
        log[tag + ".sample.states"] = np.asarray(sample["states"], np.float32)
        log[tag + ".sample.actions"] = np.asarray(sample["actions"], np.int32)
        log[tag + ".sample.rewards"] = np.asarray(sample["rewards"], np.float64)
        new_p = np.abs(rng.randn(len(idx))).astype(np.float32) + 0.01
        buf.update_priorities(idx, new_p)
        log[tag + ".update.priorities"] = new_p
        log[tag + ".after_update.priorities"] = buf.priorities.copy()
        ops.append(tag)
        step += 1
    log["n_ops"] = np.int64(step)
    np.savez_compressed(os.path.join(HERE, "learner_per.npz"), **log)
    print("learner_per:", step, "ops, final size", buf.size())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference/src")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    PMINetwork, PER, ActorCritic = import_reference(args.ref)
    if args.only in ("", "pmi"):
        pmi_train_case(PMINetwork)
    if args.only in ("", "per"):
        per_case(PER)
    if args.only in ("", "ac"):
        actor_critic_case(ActorCritic)
