"""TEST INFRASTRUCTURE (oracle) -- the counter-based noise specification shared with the CUDA engine.

The reference never seeds TensorFlow (`set_seed`, cadm/utils/utils.py:194-210, is dead code) and TF's
internal Philox streams cannot be reproduced outside TF, so "identical seeds" only has a meaning if the
noise is specified.  This file IS that specification (NumPy); cadm_b200/csrc/rng.cuh implements the same
thing on the device.  Integer outputs (Philox words, 23-bit uniforms) are bit-exact between the two; the
float transforms (erfinv, log, sin/cos) agree to a few ulp.

Generator: Philox4x32-10 (Salmon et al., SC'11), key = (seed & 0xffffffff, seed >> 32).

Counter layout (c0, c1, c2, c3):
  stream Z   (truncated-normal draws of tf.random.truncated_normal, core/utils.py:135)
      c0 = j        value block: the h*A values of one candidate, 4 per block, k = t*A + a, j = k // 4
      c1 = ni       GLOBAL candidate index (so results do not depend on the sharding)
      c2 = mi       environment index (GLOBAL: env_offset + local index when a block of environments is planned alone)
      c3 = (it << 8) | 1
      value = word k % 4 ;  u = ((w >> 9) + 0.5) * 2^-23 ;  z = sqrt(2) * erfinv(erf(sqrt(2)) * (2u - 1))
      (inverse-CDF of N(0,1) truncated to [-2, 2]; TF resamples instead -- same distribution)
  stream EPS (normal draws of tf.random.normal, core/utils.py:90)
      c0 = j        block of 4 normals of one row's D outputs, d = 4j + lane
      c1 = row id   (mi * n + ni) * p + pi   (GLOBAL candidate and environment indices)
      c2 = t        horizon step
      c3 = (it << 8) | 2
      Box-Muller: (w0, w1) -> r = sqrt(-2 ln u0), n0 = r cos(2 pi u1), n1 = r sin(2 pi u1); (w2, w3) -> n2, n3
  stream U   (uniform(-1, 1) actions of random shooting, core/utils.py:198)
      same counters as Z with c3 = (it << 8) | 3 ; value = 2u - 1
  stream UD  (uniform integer actions, core/utils.py:195) c0 = t // 4, c3 = 4 ; value = w % A
"""
import numpy as np
from scipy.special import erfinv

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)

STREAM_Z, STREAM_EPS, STREAM_U, STREAM_UD = 1, 2, 3, 4
ERF_SQRT2 = 0.9544997361036416      # erf(sqrt(2)) = P(|N(0,1)| <= 2)


def philox4x32_10(c0, c1, c2, c3, seed):
    """Vectorised Philox4x32-10.  c* broadcastable integer arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
                                           for c in (c0, c1, c2, c3)])
    c0, c1, c2, c3 = c0.copy(), c1.copy(), c2.copy(), c3.copy()
    k0 = np.uint64(int(seed) & 0xFFFFFFFF)
    k1 = np.uint64((int(seed) >> 32) & 0xFFFFFFFF)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(PHILOX_W0)) & mask
        k1 = (k1 + np.uint64(PHILOX_W1)) & mask
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u01(w):
    """23-bit uniform strictly inside (0, 1); exactly representable in float32."""
    return ((w >> np.uint32(9)).astype(np.float64) + 0.5) * (2.0 ** -23)


def trunc_normal_from_u(u):
    v = 2.0 * u - 1.0
    z = np.sqrt(2.0) * erfinv(ERF_SQRT2 * v)
    return np.clip(z, -2.0, 2.0)


def _blocks(nvals):
    return (nvals + 3) // 4


def gen_z(seed, iters, m, n, h, A, n_offset=0, dtype=np.float32, m_offset=0):
    """Truncated-normal draws [iters, m, n, h, A] for candidates n_offset .. n_offset+n-1 of environments
    m_offset .. m_offset+m-1 (the engine's "env_offset" option: a block of an environment-sharded decision)."""
    nb = _blocks(h * A)
    it = np.arange(iters)[:, None, None, None]
    mi = (np.arange(m) + m_offset)[None, :, None, None]
    ni = (np.arange(n) + n_offset)[None, None, :, None]
    j = np.arange(nb)[None, None, None, :]
    w = philox4x32_10(j, ni, mi, (it << 8) | STREAM_Z, seed)
    words = np.stack(w, axis=-1).reshape(iters, m, n, nb * 4)[..., : h * A]
    return trunc_normal_from_u(u01(words)).reshape(iters, m, n, h, A).astype(dtype)


def gen_uniform_actions(seed, m, n, h, A, it=0, n_offset=0, dtype=np.float32, m_offset=0):
    nb = _blocks(h * A)
    mi = (np.arange(m) + m_offset)[:, None, None]
    ni = (np.arange(n) + n_offset)[None, :, None]
    j = np.arange(nb)[None, None, :]
    w = philox4x32_10(j, ni, mi, (it << 8) | STREAM_U, seed)
    words = np.stack(w, axis=-1).reshape(m, n, nb * 4)[..., : h * A]
    return (2.0 * u01(words) - 1.0).reshape(m, n, h, A).astype(dtype)


def gen_discrete_actions(seed, m, n, h, A, it=0, n_offset=0, m_offset=0):
    nb = _blocks(h)
    mi = (np.arange(m) + m_offset)[:, None, None]
    ni = (np.arange(n) + n_offset)[None, :, None]
    j = np.arange(nb)[None, None, :]
    w = philox4x32_10(j, ni, mi, (it << 8) | STREAM_UD, seed)
    words = np.stack(w, axis=-1).reshape(m, n, nb * 4)[..., :h]
    return (words % np.uint32(A)).astype(np.int32)


def normals_for_rows(seed, it, t, row_ids, D):
    """N(0,1) draws [len(row_ids), D] for stream EPS."""
    nb = _blocks(D)
    rid = np.asarray(row_ids)[:, None]
    j = np.arange(nb)[None, :]
    w0, w1, w2, w3 = philox4x32_10(j, rid, t, (it << 8) | STREAM_EPS, seed)
    u0, u1, u2, u3 = u01(w0), u01(w1), u01(w2), u01(w3)
    ra, rb = np.sqrt(-2.0 * np.log(u0)), np.sqrt(-2.0 * np.log(u2))
    out = np.stack([ra * np.cos(2 * np.pi * u1), ra * np.sin(2 * np.pi * u1),
                    rb * np.cos(2 * np.pi * u3), rb * np.sin(2 * np.pi * u3)], axis=-1)
    return out.reshape(len(rid), nb * 4)[:, :D]


def gen_eps(seed, iters, h, m, n, p, E, D, dtype=np.float32, m_offset=0):
    """Normal draws in the planner's row layout [iters, h, E, R, D], R = (p/E) m n, such that the
    particle (mi, ni, pi) -- member e = pi // (p/E), row r = (pi % (p/E)) m n + mi n + ni -- gets the
    stream-EPS values of row id (mi n + ni) p + pi."""
    q = p // E
    R = q * m * n
    e = np.arange(E)[:, None]
    r = np.arange(R)[None, :]
    jq, rem = r // (m * n), r % (m * n)
    mi, ni = rem // n, rem % n
    pi = e * q + jq
    rid = (((mi + m_offset) * n + ni) * p + pi).reshape(-1)
    out = np.empty((iters, h, E, R, D), dtype=dtype)
    for it in range(iters):
        for t in range(h):
            out[it, t] = normals_for_rows(seed, it, t, rid, D).reshape(E, R, D)
    return out
