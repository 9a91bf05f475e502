"""numpy Philox4x32-10, written independently of csrc/philox.cuh, for checking the reset and
random-policy kernels bit for bit."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
STREAM_UAV_RESET, STREAM_TGT_POS, STREAM_TGT_HEAD, STREAM_ACTION = 0, 1, 2, 3


def philox4x32_10(c0, c1, c2, c3, seed):
    c = [np.asarray(v, dtype=np.uint64) & MASK for v in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c


def u53(a, b):
    return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def below(r, n):
    return ((r * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def reset_reference(seed, E, n, m, na, x_max, y_max, env_id_offset=0):
    """What uavsim_reset must produce (src/environment.py:45-107 with Philox draws)."""
    pi = 3.141592653589793
    g = (np.arange(E, dtype=np.uint64) + np.uint64(env_id_offset))[:, None]
    i = np.arange(n, dtype=np.uint64)[None, :]
    r = philox4x32_10(i, STREAM_UAV_RESET, g, 0, seed)
    ux = np.broadcast_to((np.arange(n, dtype=np.float64) + 1.0) * float(x_max) / float(n + 1), (E, n)).copy()
    uy = np.full((E, n), float(y_max) / 2)
    uh = -pi + (pi - (-pi)) * u53(r[0], r[1])
    ua = below(r[2], na).astype(np.int32)
    t = np.arange(m, dtype=np.uint64)[None, :]
    r1 = philox4x32_10(t, STREAM_TGT_POS, g, 0, seed)
    r2 = philox4x32_10(t, STREAM_TGT_HEAD, g, 0, seed)
    tx = 0 + (float(x_max) - 0) * u53(r1[0], r1[1])
    ty = 0 + (float(y_max) - 0) * u53(r1[2], r1[3])
    th = -pi + (pi - (-pi)) * u53(r2[0], r2[1])
    return {"ux": ux, "uy": uy, "uh": uh, "ua": ua, "tx": tx, "ty": ty, "th": th}


def actions_reference(seed, step, E, n, na, env_id_offset=0):
    g = (np.arange(E, dtype=np.uint64) + np.uint64(env_id_offset))[:, None]
    i = np.arange(n, dtype=np.uint64)[None, :]
    r = philox4x32_10(i, STREAM_ACTION, g, step, seed)
    return below(r[0], na).astype(np.int32)
