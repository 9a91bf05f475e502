"""Scenario constants of the reference's YAML files (src/configs/MAAC*.yaml share one environment /
uav / target block) as a plain dict with the same keys, so no YAML file is needed on a GPU box."""
import copy

_DEFAULT = {
    "seed": 42,
    "cooperative": 0.3,
    "environment": {"n_uav": 10, "m_targets": 10, "x_max": 2000, "y_max": 2000, "na": 12},
    "uav": {"dt": 1, "v_max": 20, "h_max": 6, "dc": 500, "dp": 200, "alpha": 0.6, "beta": 0.2, "gamma": 0.2},
    "target": {"v_max": 5, "h_max": 6},
    "pmi": {"hidden_dim": 128, "b2_size": 3000, "batch_size": 128},
}


def default_config(method="MAAC-R", n_uav=10, m_targets=10, **overrides):
    """method: 'MAAC' (cooperative forced to 0, src/main.py:75-76), 'MAAC-G' or 'MAAC-R' (0.3)."""
    c = copy.deepcopy(_DEFAULT)
    c["exp_name"] = method
    c["environment"]["n_uav"], c["environment"]["m_targets"] = n_uav, m_targets
    if method == "MAAC":
        c["cooperative"] = 0
    for k, v in overrides.items():
        if "__" in k:
            sec, key = k.split("__", 1)
            c[sec][key] = v
        else:
            c[k] = v
    return c
