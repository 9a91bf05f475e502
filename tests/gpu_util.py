"""Helpers shared by the GPU parity tests and __graft_entry__.smoke()."""
import json

import numpy as np
import torch

from oracle import OracleParams, params_from_golden, pmi_from_golden  # noqa: F401  (test infrastructure)
from oracle.oracle import pmi_from_state

STATE = ("ux", "uy", "uh", "ua", "tx", "ty", "th")
MASKS = ("obs_mask", "comm_mask", "nbr_mask", "dup_mask", "cover_mask")
MODE_NAME = {0: "self", 1: "mean", 2: "pmi"}


def golden_config(g):
    return json.loads(str(g["config_json"]))


def golden_pmi_module(g):
    """Rebuild the PMI network of a fixture as the package's PMINetwork mirror."""
    from marl_uavs_targets_tracking_b200 import PMINetwork
    sd = {k[4:]: torch.tensor(np.array(g[k])) for k in g.files if k.startswith("pmi.")}
    if not sd:
        return None
    net = PMINetwork(hidden_dim=sd["fc1.weight"].shape[0])
    net.load_state_dict(sd)
    net.eval()
    return net


def oracle_params_from_config(cfg, n, m):
    from math import pi
    P = OracleParams()
    e, u, t = cfg["environment"], cfg["uav"], cfg["target"]
    P.n_uav, P.m_targets, P.na = n, m, int(e["na"])
    P.x_max, P.y_max, P.dt, P.uav_v_max = float(e["x_max"]), float(e["y_max"]), float(u["dt"]), float(u["v_max"])
    P.uav_h_max, P.dc, P.dp = pi / float(u["h_max"]), float(u["dc"]), float(u["dp"])
    P.tgt_v_max, P.tgt_h_max = float(t["v_max"]), pi / float(t["h_max"])
    P.alpha, P.beta, P.gamma = float(u["alpha"]), float(u["beta"]), float(u["gamma"])
    return P


def oracle_pmi_from_module(net):
    return pmi_from_state({k: v.detach().cpu().numpy() for k, v in net.state_dict().items()
                           if not k.endswith("num_batches_tracked")})


def max_scaled_err(got, ref):
    """|got-ref| / max(|ref|, 1): the tolerance rule of SURVEY.md section 7 ('Tolerance definition')."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    if got.size == 0:
        return 0.0
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)))
