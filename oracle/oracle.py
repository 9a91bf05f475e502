"""ctypes binding of oracle/liboracle.so (test infrastructure; see uavsim_oracle.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MODE_SELF, MODE_MEAN, MODE_PMI = 0, 1, 2


class OracleParams(C.Structure):
    _fields_ = [("n_uav", C.c_int32), ("m_targets", C.c_int32), ("na", C.c_int32), ("_pad", C.c_int32)] + [
        (k, C.c_double) for k in ("x_max", "y_max", "dt", "uav_v_max", "uav_h_max", "dc", "dp", "tgt_v_max",
                                  "tgt_h_max", "alpha", "beta", "gamma")]


_FP = C.POINTER(C.c_float)


class OraclePmi(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("_pad", C.c_int32),
                ("w_in", _FP * 3), ("b_in", _FP * 3), ("bn_in", (_FP * 4) * 3),
                ("w1", _FP), ("b1", _FP), ("bn1", _FP * 4), ("w2", _FP), ("b2", _FP)]


def build_oracle(force=False):
    """Compile liboracle.so with gcc (oracle/Makefile).  Building the checker is not using it."""
    src = os.path.join(_HERE, "uavsim_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def params_from_golden(g):
    pf, pi_ = g["params_f"], g["params_i"]
    P = OracleParams()
    P.n_uav, P.m_targets, P.na = int(pi_[0]), int(pi_[1]), int(pi_[2])
    (P.x_max, P.y_max, P.dt, P.uav_v_max, P.uav_h_max, P.dc, P.dp, P.tgt_v_max, P.tgt_h_max,
     P.alpha, P.beta, P.gamma) = [float(v) for v in pf[:12]]
    return P, int(pi_[5]), float(pf[12])


_BN = ("weight", "bias", "running_mean", "running_var")


def pmi_from_state(sd):
    """sd: mapping name -> float32 ndarray with the reference PMINetwork.state_dict() names."""
    keep = []

    def f(name):
        a = np.ascontiguousarray(np.asarray(sd[name], dtype=np.float32))
        keep.append(a)
        return a.ctypes.data_as(_FP)

    p = OraclePmi()
    p.hidden = int(np.asarray(sd["fc1.weight"]).shape[0])
    for b, (fc, bn) in enumerate((("fc_comm", "bn_comm"), ("fc_obs", "bn_obs"),
                                  ("fc_boundary_state", "bn_boundary_state"))):
        p.w_in[b] = f(fc + ".weight")
        p.b_in[b] = f(fc + ".bias")
        for k, nm in enumerate(_BN):
            p.bn_in[b][k] = f(bn + "." + nm)
    p.w1, p.b1 = f("fc1.weight"), f("fc1.bias")
    for k, nm in enumerate(_BN):
        p.bn1[k] = f("bn1." + nm)
    p.w2, p.b2 = f("fc2.weight"), f("fc2.bias")
    p._keep = keep
    return p


def pmi_from_golden(g):
    sd = {k[4:]: g[k] for k in g.files if k.startswith("pmi.")}
    return pmi_from_state(sd) if sd else None


class Oracle:
    """Thin object wrapper: one environment (`step`) or E environments (`step_batch`)."""

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        self.lib.oracle_step.restype = None
        self.lib.oracle_step_batch.restype = None
        self.lib.oracle_initial_obs.restype = None

    def initial_obs(self, P, ux, uy, ua):
        obs = np.empty((P.n_uav, 12), np.float64)
        self.lib.oracle_initial_obs(C.byref(P), _ptr(ux, C.c_double), _ptr(uy, C.c_double), _ptr(ua, C.c_int32),
                                    _ptr(obs, C.c_double))
        return obs

    def step(self, P, mode, coop, pmi, st, actions, masks=True):
        """st: dict of contiguous arrays ux,uy,uh (f64) ua (i32) tx,ty,th (f64), updated in place."""
        n, m = P.n_uav, P.m_targets
        out = {"obs": np.empty((n, 12)), "rewards": np.empty(n), "tt": np.empty(n), "bp": np.empty(n),
               "dup": np.empty(n), "raw": np.empty(n), "covered": np.zeros(1, np.int32),
               "tracker_cnt": np.zeros(m, np.int32)}
        mk = {}
        if masks:
            mk = {"obs_mask": np.zeros((n, m), np.uint8), "comm_mask": np.zeros((n, n), np.uint8),
                  "nbr_mask": np.zeros((n, n), np.uint8), "dup_mask": np.zeros((n, n), np.uint8),
                  "cover_mask": np.zeros((n, m), np.uint8)}
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        d, i32, u8 = C.c_double, C.c_int32, C.c_uint8
        self.lib.oracle_step(
            C.byref(P), C.c_int(mode), C.c_double(coop), C.byref(pmi) if pmi is not None else None,
            _ptr(st["ux"], d), _ptr(st["uy"], d), _ptr(st["uh"], d), _ptr(st["ua"], i32),
            _ptr(st["tx"], d), _ptr(st["ty"], d), _ptr(st["th"], d), _ptr(actions, i32),
            _ptr(out["obs"], d), _ptr(out["rewards"], d), _ptr(out["tt"], d), _ptr(out["bp"], d),
            _ptr(out["dup"], d), _ptr(out["raw"], d), _ptr(out["covered"], i32), _ptr(out["tracker_cnt"], i32),
            _ptr(mk.get("obs_mask"), u8), _ptr(mk.get("comm_mask"), u8), _ptr(mk.get("nbr_mask"), u8),
            _ptr(mk.get("dup_mask"), u8), _ptr(mk.get("cover_mask"), u8))
        out.update(mk)
        out["covered"] = int(out["covered"][0])
        return out

    def step_batch(self, P, mode, coop, pmi, st, actions, nthreads=1, want_obs=True, want_tracker=True):
        """st arrays are [E,n] / [E,m], updated in place; returns obs [E,n,12], rew4 [4,E,n], covered [E]."""
        E = st["ux"].shape[0]
        n, m = P.n_uav, P.m_targets
        obs = np.empty((E, n, 12)) if want_obs else None
        rew4 = np.empty((4, E, n))
        cov = np.zeros(E, np.int32)
        trk = np.zeros((E, m), np.int32) if want_tracker else None
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        d, i32 = C.c_double, C.c_int32
        self.lib.oracle_step_batch(
            C.byref(P), C.c_int(mode), C.c_double(coop), C.byref(pmi) if pmi is not None else None,
            C.c_int64(E), _ptr(st["ux"], d), _ptr(st["uy"], d), _ptr(st["uh"], d), _ptr(st["ua"], i32),
            _ptr(st["tx"], d), _ptr(st["ty"], d), _ptr(st["th"], d), _ptr(actions, i32),
            _ptr(obs, d), _ptr(rew4, d), _ptr(cov, i32), _ptr(trk, i32), C.c_int(nthreads))
        return {"obs": obs, "rew4": rew4, "covered": cov, "tracker_cnt": trk}
