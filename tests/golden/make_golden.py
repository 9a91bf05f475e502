#!/usr/bin/env python
"""Generate golden vectors by EXECUTING the unmodified reference environment.

Runs only in the build container (needs /root/reference, which does not exist on
the GPU box).  It imports the reference's `environment.Environment` and
`models.PMINet.PMINetwork`, drives them with recorded action streams and dumps
per-step state / observation / reward / coverage plus the integer masks into
`tests/golden/*.npz`.  Nothing from the reference is copied: the fixtures hold
numbers only.

Masks are recorded with read-only wrappers around `UAV.observe_target` /
`UAV.observe_uav` (src/agent/uav.py:101-147) that re-evaluate the reference's own
private `__distance` at the moment of the call, so the Gauss-Seidel (mixed
old/new) communication mask is the one the reference actually used.

usage:  python tests/golden/make_golden.py [--ref /root/reference/src]
"""
import argparse
import json
import os
import random
import sys
from math import pi, cos, sin

import numpy as np

sys.dont_write_bytecode = True  # the reference tree is read-only


def base_config():
    # values of src/configs/*.yaml (identical environment/uav/target blocks in all four files)
    return {
        "cooperative": 0.3,
        "environment": {"n_uav": 10, "m_targets": 10, "x_max": 2000, "y_max": 2000, "na": 12},
        "uav": {"dt": 1, "v_max": 20, "h_max": 6, "dc": 500, "dp": 200,
                "alpha": 0.6, "beta": 0.2, "gamma": 0.2},
        "target": {"v_max": 5, "h_max": 6},
        "pmi": {"hidden_dim": 128, "b2_size": 3000, "batch_size": 128},
    }


def install_recorders(UAV, rec):
    orig_t, orig_u = UAV.observe_target, UAV.observe_uav

    def observe_target(self, targets_list, relative=True):
        orig_t(self, targets_list, relative)
        mask = [self._UAV__distance(t) <= self.dp for t in targets_list]
        assert sum(mask) == len(self.target_observation)
        rec["obs_mask"].append(mask)

    def observe_uav(self, uav_list, relative=True):
        orig_u(self, uav_list, relative)
        mask = [(self._UAV__distance(u) <= self.dc and u is not self) for u in uav_list]
        assert sum(mask) == len(self.uav_communication)
        rec["comm_mask"].append(mask)

    UAV.observe_target, UAV.observe_uav = observe_target, observe_uav
    return orig_t, orig_u


def make_pmi(PMINetwork, torch, hidden, seed):
    torch.manual_seed(seed)
    pmi = PMINetwork(hidden_dim=hidden, b2_size=3000)
    # randomise BN running statistics so eval-mode BN is not the identity
    g = torch.Generator().manual_seed(seed + 1)
    for bn in (pmi.bn_comm, pmi.bn_obs, pmi.bn_boundary_state, pmi.bn1):
        bn.running_mean.copy_(torch.randn(hidden, generator=g) * 0.3)
        bn.running_var.copy_(torch.rand(hidden, generator=g) + 0.5)
        bn.weight.data.copy_(1.0 + 0.2 * torch.randn(hidden, generator=g))
        bn.bias.data.copy_(0.1 * torch.randn(hidden, generator=g))
    pmi.eval()
    return pmi


def run_case(name, out_dir, mods, cfg, mode, T, seed, actions=None, init_override=None, pmi_hidden=128):
    Environment, UAV, PMINetwork, torch = mods
    env_c = cfg["environment"]
    n, m = env_c["n_uav"], env_c["m_targets"]
    if mode == "self":
        cfg["cooperative"] = 0  # src/main.py:75-76
    pmi = make_pmi(PMINetwork, torch, pmi_hidden, seed) if mode == "pmi" else None

    random.seed(seed)
    np.random.seed(seed)
    env = Environment(n_uav=n, m_targets=m, x_max=env_c["x_max"], y_max=env_c["y_max"], na=env_c["na"])
    env.reset(cfg)
    if init_override is not None:
        init_override(env)

    init = {
        "ux0": np.array([u.x for u in env.uav_list], dtype=np.float64),
        "uy0": np.array([u.y for u in env.uav_list], dtype=np.float64),
        "uh0": np.array([u.h for u in env.uav_list], dtype=np.float64),
        "ua0": np.array([u.a for u in env.uav_list], dtype=np.int32),
        "tx0": np.array([t.x for t in env.target_list], dtype=np.float64),
        "ty0": np.array([t.y for t in env.target_list], dtype=np.float64),
        "th0": np.array([t.h for t in env.target_list], dtype=np.float64),
        "obs0": np.array(env.get_states(), dtype=np.float64).reshape(n, 12),
    }
    if actions is None:
        arng = np.random.RandomState(seed + 1000)
        actions = arng.randint(0, env_c["na"], size=(T, n)).astype(np.int32)

    rec = {"obs_mask": [], "comm_mask": []}
    orig = install_recorders(UAV, rec)
    out = {k: [] for k in ("ux", "uy", "uh", "tx", "ty", "th", "obs", "rewards", "tt", "bp", "dup", "covered",
                           "nbr_mask", "dup_mask", "cover_mask", "raw")}
    try:
        for t in range(T):
            states, rew, cov = env.step(cfg, pmi, [int(a) for a in actions[t]])
            us, ts = env.uav_list, env.target_list
            out["ux"].append([u.x for u in us]); out["uy"].append([u.y for u in us]); out["uh"].append([u.h for u in us])
            out["tx"].append([q.x for q in ts]); out["ty"].append([q.y for q in ts]); out["th"].append([q.h for q in ts])
            out["obs"].append(np.array(states, dtype=np.float64).reshape(n, 12))
            out["rewards"].append([float(v) for v in rew["rewards"]])
            out["tt"].append([float(v) for v in rew["target_tracking_reward"]])
            out["bp"].append([float(v) for v in rew["boundary_punishment"]])
            out["dup"].append([float(v) for v in rew["duplicate_tracking_punishment"]])
            out["raw"].append([float(u.raw_reward) for u in us])
            out["covered"].append(int(cov))
            out["nbr_mask"].append([[(u is not o) and u._UAV__distance(o) <= u.dp for o in us] for u in us])
            out["dup_mask"].append([[(u is not o) and u._UAV__distance(o) <= 2 * u.dp for o in us] for u in us])
            out["cover_mask"].append([[UAV.distance(u.x, u.y, q.x, q.y) < u.dp for q in ts] for u in us])
    finally:
        UAV.observe_target, UAV.observe_uav = orig

    save = dict(init)
    save["actions"] = actions
    for k in ("ux", "uy", "uh", "tx", "ty", "th", "obs", "rewards", "tt", "bp", "dup", "raw"):
        save[k] = np.array(out[k], dtype=np.float64)
    save["covered"] = np.array(out["covered"], dtype=np.int32)
    save["obs_mask"] = np.array(rec["obs_mask"], dtype=bool).reshape(T, n, m)
    save["comm_mask"] = np.array(rec["comm_mask"], dtype=bool).reshape(T, n, n)
    for k in ("nbr_mask", "dup_mask"):
        save[k] = np.array(out[k], dtype=bool).reshape(T, n, n)
    save["cover_mask"] = np.array(out["cover_mask"], dtype=bool).reshape(T, n, m)
    # scenario constants (already converted the way src/environment.py:97-107 does)
    save["params_f"] = np.array([
        env_c["x_max"], env_c["y_max"], cfg["uav"]["dt"], cfg["uav"]["v_max"], pi / float(cfg["uav"]["h_max"]),
        cfg["uav"]["dc"], cfg["uav"]["dp"], cfg["target"]["v_max"], pi / float(cfg["target"]["h_max"]),
        cfg["uav"]["alpha"], cfg["uav"]["beta"], cfg["uav"]["gamma"], cfg["cooperative"]], dtype=np.float64)
    save["params_i"] = np.array([n, m, env_c["na"], T, seed, {"self": 0, "mean": 1, "pmi": 2}[mode]], dtype=np.int64)
    save["config_json"] = np.array(json.dumps(cfg))  # the exact dict handed to reset()/step()
    if pmi is not None:
        for k, v in pmi.state_dict().items():
            save["pmi." + k] = v.detach().cpu().numpy()
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, **save)
    print("%-28s n=%-3d m=%-3d T=%-4d mode=%-4s covered(mean)=%.2f  obs_hits=%d comm_hits=%d nbr_hits=%d  %.0f KB" % (
        name, n, m, T, mode, np.mean(save["covered"]), save["obs_mask"].sum(), save["comm_mask"].sum(),
        save["nbr_mask"].sum(), os.path.getsize(path) / 1024))


def origin_pass_override(na, v, dt, h_max):
    """Place UAV k so that it lands within ~0.5 m of the origin at step k+1 (alternating
    actions na/2-1, na/2), exercising the min(dist,1) weight quirk of src/agent/uav.py:162-186."""
    def rate(a):
        return (2 * (a + 1) - na - 1) * h_max / (na - 1)

    def fn(env):
        for k, u in enumerate(env.uav_list):
            steps = k + 1
            x, y = 0.3 - 0.07 * k, 0.2 + 0.05 * k
            h = -2.5 + 0.7 * k
            # walk backwards through the action stream used by the case (a_t = na/2-1 + (t % 2))
            for t in reversed(range(steps)):
                a = na // 2 - 1 + (t % 2)
                h = h - dt * rate(a)
                x -= dt * v * cos(h)
                y -= dt * v * sin(h)
            u.x, u.y, u.h = x, y, h
        for j, q in enumerate(env.target_list):
            q.x, q.y = 40.0 + 35.0 * j, 25.0 + 20.0 * j
    return fn


def corner_targets_override(env):
    """Targets aimed at walls/corners so every reflection branch of src/agent/target.py:52-58 fires,
    including the 'both out, y wins' case; UAVs near the edges so the boundary term changes branch."""
    xm, ym = env.x_max, env.y_max
    spots = [(3.0, 3.0, -2.4), (xm - 2.0, ym - 2.0, 0.8), (2.0, ym / 2, 3.0), (xm - 3.0, ym / 3, -0.2),
             (xm / 2, 2.0, -1.5), (xm / 3, ym - 1.0, 1.6), (1.0, ym - 4.0, 2.3), (xm - 4.0, 1.0, -0.7)]
    for q, (x, y, h) in zip(env.target_list, spots):
        q.x, q.y, q.h = x, y, h
    edge = [(10.0, 10.0, -2.0), (xm - 15.0, 30.0, 0.3), (190.0, ym - 5.0, 1.2), (xm / 2, ym / 2, 0.0),
            (205.0, 205.0, 2.0), (xm - 199.0, ym - 201.0, -1.0)]
    for u, (x, y, h) in zip(env.uav_list, edge):
        u.x, u.y, u.h = x, y, h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("UAVSIM_REFERENCE_SRC", "/root/reference/src"))
    ap.add_argument("--out", default=os.path.dirname(os.path.abspath(__file__)))
    ap.add_argument("--only", default="", help="comma-separated case names (default: all)")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    global run_case
    _run = run_case

    def run_case(name, *a, **k):  # noqa: F811
        if not only or name in only:
            _run(name, *a, **k)
    sys.path.insert(0, args.ref)
    import torch
    torch.set_num_threads(1)
    from environment import Environment
    from agent.uav import UAV
    from models.PMINet import PMINetwork
    mods = (Environment, UAV, PMINetwork, torch)

    def cfg(n=10, m=10, coop=0.3, **kw):
        c = base_config()
        c["environment"]["n_uav"], c["environment"]["m_targets"], c["cooperative"] = n, m, coop
        for k, v in kw.items():
            sec, key = k.split("__")
            c[sec][key] = v
        return c

    # default scenario, the three reward modes, two seeds
    for seed in (42, 7):
        run_case("d10_self_s%d" % seed, args.out, mods, cfg(), "self", 200, seed)
        run_case("d10_mean_s%d" % seed, args.out, mods, cfg(), "mean", 200, seed)
        run_case("d10_pmi_s%d" % seed, args.out, mods, cfg(), "pmi", 200, seed)
    # scaled swarm
    run_case("s64_mean_s42", args.out, mods, cfg(64, 64), "mean", 200, 42)
    run_case("s64_self_s3", args.out, mods, cfg(64, 64), "self", 60, 3)
    run_case("s64_pmi_s42", args.out, mods, cfg(64, 64), "pmi", 30, 42)
    run_case("s64_pmi_s9_long", args.out, mods, cfg(64, 64), "pmi", 200, 9)   # a whole MAAC-R episode of the scaled swarm
    run_case("s32_pmi_s5", args.out, mods, cfg(32, 32), "pmi", 80, 5)
    # ragged / degenerate sizes
    run_case("n1_m1_mean", args.out, mods, cfg(1, 1), "mean", 50, 11)
    run_case("n1_m5_pmi", args.out, mods, cfg(1, 5), "pmi", 30, 12)
    run_case("n5_m1_pmi", args.out, mods, cfg(5, 1), "pmi", 60, 13)
    run_case("n33_m7_mean", args.out, mods, cfg(33, 7), "mean", 80, 14)
    run_case("n7_m70_self", args.out, mods, cfg(7, 70), "self", 60, 15)
    run_case("n100_m3_mean", args.out, mods, cfg(100, 3), "mean", 25, 16)
    # non-default constants (non-square map, other na/dt/ranges/weights)
    odd = cfg(7, 5, 0.5, environment__x_max=1500, environment__y_max=900, environment__na=8,
              uav__dt=0.5, uav__v_max=30, uav__h_max=4, uav__dc=300, uav__dp=120,
              uav__alpha=0.5, uav__beta=0.3, uav__gamma=0.2, target__v_max=8, target__h_max=5)
    run_case("odd_mean", args.out, mods, odd, "mean", 120, 21)
    odd2 = cfg(7, 5, 0.5, environment__x_max=1500, environment__y_max=900, environment__na=8,
               uav__dt=0.5, uav__v_max=30, uav__h_max=4, uav__dc=300, uav__dp=120,
               uav__alpha=0.5, uav__beta=0.3, uav__gamma=0.2, target__v_max=8, target__h_max=5)
    run_case("odd_pmi_h64", args.out, mods, odd2, "pmi", 120, 22, pmi_hidden=64)
    # perception range wider than the communication range: the duplicate-tracking radius 2*dp (uav.py:225) exceeds
    # dc + dt*v, and every UAV sees most targets
    wide = cfg(12, 9, 0.6, environment__x_max=1200, environment__y_max=1000, environment__na=10,
               uav__dt=1, uav__v_max=25, uav__h_max=5, uav__dc=150, uav__dp=400,
               uav__alpha=0.3, uav__beta=0.3, uav__gamma=0.4, target__v_max=6, target__h_max=6)
    run_case("wide_dp_mean", args.out, mods, wide, "mean", 100, 51)
    wide2 = cfg(12, 9, 0.6, environment__x_max=1200, environment__y_max=1000, environment__na=10,
                uav__dt=1, uav__v_max=25, uav__h_max=5, uav__dc=150, uav__dp=400,
                uav__alpha=0.3, uav__beta=0.3, uav__gamma=0.4, target__v_max=6, target__h_max=6)
    run_case("wide_dp_pmi", args.out, mods, wide2, "pmi", 60, 52)
    # weight quirk: UAVs flying through the origin
    T = 12
    na = 12
    acts = np.array([[na // 2 - 1 + (t % 2)] * 8 for t in range(T)], dtype=np.int32)
    run_case("origin_mean", args.out, mods, cfg(8, 6), "mean", T, 31, actions=acts,
             init_override=origin_pass_override(na, 20, 1, pi / 6))
    run_case("origin_pmi", args.out, mods, cfg(8, 6), "pmi", T, 32, actions=acts,
             init_override=origin_pass_override(na, 20, 1, pi / 6))
    # reflections and boundary branches
    run_case("walls_mean", args.out, mods, cfg(6, 8), "mean", 40, 41, init_override=corner_targets_override)


if __name__ == "__main__":
    main()
