import sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from oracle.learner_oracle import ReplayOracle
from marl_uavs_targets_tracking_b200 import PrioritizedReplayBuffer
DEV="cuda:0"
_t=lambda a: torch.as_tensor(np.asarray(a)).to(DEV)
rng = np.random.RandomState(0)
C, D = 100_000, 12
buf, o = PrioritizedReplayBuffer(C, device=DEV), ReplayOracle(C)
for it, k in enumerate((30_000, 50_000, 45_000, 5)):
    s, s2 = rng.randn(k, D).astype(np.float32), rng.randn(k, D).astype(np.float32)
    a, r = rng.randint(0, 12, k).astype(np.int32), rng.randn(k).astype(np.float32)
    buf.add({"states": _t(s), "actions": _t(a), "rewards": _t(r), "next_states": _t(s2)})
    maxp = o.priorities.max() if o.n else 1.0
    slots = (o.pos + np.arange(k)) % C
    o.states[slots], o.next_states[slots], o.actions[slots], o.rewards[slots] = s, s2, a, r
    o.priorities[slots] = maxp
    o.n, o.pos = min(o.n + k, C), (o.pos + k) % C
    u = rng.random_sample(4096)
    sample, idx, w = buf.sample(4096, 0.4, uniforms=_t(u))
    oi, ow, oprob = o.sample(4096, u, 0.4)
    ex = buf.export()
    prob = ex["probabilities"].numpy()
    gi = idx.cpu().numpy()
    cdf = np.cumsum(prob.astype(np.float64)); cdf /= cdf[-1]
    si = np.searchsorted(cdf, u, side="right")
    print(it, "n", o.n, "mismatch vs numpy-prob oracle", (gi != oi).sum(), "vs own-prob oracle", (gi != si).sum(),
          "prob maxrel", np.abs(prob/oprob-1).max(), "pri equal", np.array_equal(ex["priorities"].numpy(), o.priorities))
    bad = np.nonzero(gi != oi)[0][:5]
    ocdf = np.cumsum(oprob.astype(np.float64)); ocdf/=ocdf[-1]
    for b in bad: print("   k", b, "u", u[b], "gpu", gi[b], "np", oi[b], "cdf gap own", cdf[gi[b]]-u[b], "np cdf at np idx", ocdf[oi[b]]-u[b], "cdf diff", cdf[oi[b]]-ocdf[oi[b]])
    newp = (np.abs(rng.randn(4096)) + 1e-3).astype(np.float32)
    dup = rng.randint(0, 4096, 4096)
    gi2 = gi[dup]
    buf.update_priorities(_t(gi2), _t(newp))
    o.update_priorities(gi2, newp)
