"""Pin the CPU oracle (oracle/uavsim_oracle.c) against outputs of the reference itself.

tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports and runs the
unmodified reference (environment.py / agent/uav.py / agent/target.py / models/PMINet.py) in the
build container.  Integer outputs must match bit for bit; positions, headings and observations are
required to be bit-equal too (same libm, same evaluation order); rewards within 1e-12 (MAAC / MAAC-G)
or 1.7e-7 x cooperative (MAAC-R: the fp32 MLP's summation order differs from torch's GEMV, a float32-epsilon effect on
the softmax weights that enters the reward scaled by `cooperative`; 5e-8 at the shipped 0.3).
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import params_from_golden, pmi_from_golden

STATE = ("ux", "uy", "uh", "tx", "ty", "th")
MASKS = ("obs_mask", "comm_mask", "nbr_mask", "dup_mask", "cover_mask")


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference(oracle, name):
    g = load_golden(name)
    P, mode, coop = params_from_golden(g)
    pmi = pmi_from_golden(g)
    st = {k: np.array(g[k + "0"]) for k in STATE + ("ua",)}
    assert np.array_equal(oracle.initial_obs(P, st["ux"], st["uy"], st["ua"]), g["obs0"])
    T = g["actions"].shape[0]
    rtol = max(5e-8, 1.7e-7 * coop) if mode == 2 else 1e-12
    for t in range(T):
        out = oracle.step(P, mode, coop, pmi, st, g["actions"][t])
        for k in STATE:
            assert np.array_equal(st[k], g[k][t]), (name, t, k)
        assert np.array_equal(st["ua"], g["actions"][t])
        assert np.array_equal(out["obs"], g["obs"][t]), (name, t, "obs")
        for k in MASKS:
            assert np.array_equal(out[k].astype(bool), g[k][t]), (name, t, k)
        assert out["covered"] == int(g["covered"][t])
        assert np.array_equal(out["tracker_cnt"], g["cover_mask"][t].sum(0))
        for k in ("tt", "bp", "dup", "raw"):
            np.testing.assert_allclose(out[k], g[k][t], rtol=0, atol=1e-15, err_msg="%s t=%d %s" % (name, t, k))
        np.testing.assert_allclose(out["rewards"], g["rewards"][t], rtol=0, atol=rtol, err_msg="%s t=%d" % (name, t))


def test_golden_covers_the_edge_cases():
    """The fixtures really exercise the branches they were built for."""
    g = load_golden("origin_mean")
    # weight quirk (src/agent/uav.py:162-186): some UAV within 1 of a relative-coordinate point
    rx = (g["tx"][:, None, :] - g["ux"][:, :, None]) / 200.0
    ry = (g["ty"][:, None, :] - g["uy"][:, :, None]) / 200.0
    w = np.sqrt((rx - g["ux"][:, :, None]) ** 2 + (ry - g["uy"][:, :, None]) ** 2)
    assert ((w < 1) & g["obs_mask"]).sum() >= 3
    g = load_golden("walls_mean")
    th = np.concatenate([g["th0"][None], g["th"]])
    assert (np.abs(np.diff(th, axis=0)) > 0).sum() >= 6          # reflections happened
    assert (g["bp"] == -1.0).any() and (g["bp"] == 0.0).any()     # outside the map / far from the walls
    g = load_golden("s64_mean_s42")
    assert (g["ux"] < 0).any() or (g["ux"] > 2000).any() or (g["uy"] < 0).any() or (g["uy"] > 2000).any()
    # MAAC-G precedence quirk (src/agent/uav.py:308-309): no neighbour -> reward exactly 0
    g = load_golden("d10_mean_s42")
    lonely = ~g["nbr_mask"].any(axis=2)
    assert lonely.any() and np.all(g["rewards"][lonely] == 0.0)


def test_oracle_batch_equals_single(oracle):
    g = load_golden("d10_mean_s42")
    P, mode, coop = params_from_golden(g)
    E = 5
    st1 = {k: np.array(g[k + "0"]) for k in STATE + ("ua",)}
    stb = {k: np.ascontiguousarray(np.broadcast_to(st1[k], (E,) + st1[k].shape)).copy() for k in st1}
    for t in range(20):
        a = g["actions"][t]
        o1 = oracle.step(P, mode, coop, None, st1, a, masks=False)
        ob = oracle.step_batch(P, mode, coop, None, stb, np.broadcast_to(a, (E, a.size)).copy(), nthreads=2)
        for e in range(E):
            assert np.array_equal(ob["obs"][e], o1["obs"])
            assert np.array_equal(ob["rew4"][0, e], o1["rewards"])
            assert ob["covered"][e] == o1["covered"]
            assert np.array_equal(ob["tracker_cnt"][e], o1["tracker_cnt"])
