"""CPU oracle for the environment hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this package.  The product package
(`marl_uavs_targets_tracking_b200`) never does: it has no CPU path at all.

Parity pin: `uavsim_oracle.c` is checked against outputs of the reference itself
(`tests/golden/*.npz`, produced by `tests/golden/make_golden.py`) in
`tests/test_oracle_golden.py`.
"""
from .oracle import (  # noqa: F401
    OracleParams, OraclePmi, Oracle, build_oracle, params_from_golden, pmi_from_golden, MODE_SELF, MODE_MEAN,
    MODE_PMI,
)
