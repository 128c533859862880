"""The action stream of the rollout kernels, restated on the host (numpy): SURVEY.md 8(d) config 2 draws
u[k, b] ~ U(-uMax, uMax) "from torch.Generator(seed=0) on host ... or in-kernel Philox with the same counter scheme,
validated against the host stream".  The kernels (csrc/pg_kernels.cuh `philox_action`, C ABI pg_rollout_random /
pg_random_actions) use

    Philox2x32-10 (Salmon et al., SC'11: one 64-bit block per action, half the integer work of the 4x32 variant),
    key = Philox2x32-10(counter = (seed mod 2^32, seed div 2^32), key = 0x5047454E).r0      (once per stream)
    (r0, r1) = Philox2x32-10(counter = (b, 4 k + j), key),          b < 2^32, k < 2^30, j < 4
    U = ((r1 << 32 | r0) >> 11) * 2^-53,      u[k, j, b] = (2 U - 1) * u_max[j]            (IEEE double throughout)

and this module computes the same numbers bit for bit, so a user (or a test) can reproduce on the host exactly the controls
a `rollout(..., actions=(seed, u_max))` applied on the device.  Host-side utility: no GPU involved, nothing here is on the
hot path."""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_M2 = np.uint64(0xD256D193)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_STREAM_KEY = 0x5047454E


def philox2x32_10(c0, c1, key):
    """Vectorised Philox2x32-10: uint32 arrays (broadcastable) and a scalar key in, the two output words out."""
    c0, c1 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1))
    key = int(key) & 0xFFFFFFFF
    for _ in range(10):
        p = _M2 * c0
        c0, c1 = (p >> np.uint64(32)) ^ np.uint64(key) ^ c1, p & _MASK
        key = (key + _W0) & 0xFFFFFFFF
    return c0, c1


def stream_key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return int(philox2x32_10(np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32), _STREAM_KEY)[0])


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10: uint32 arrays (broadcastable) in, the four output words out."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & _MASK, p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def random_actions(B, T, seed, u_max, k_offset=0, b_offset=0, dtype=np.float64):
    """U [T, m, B]: the controls `rollout(..., actions=(seed, u_max, k_offset))` applies to environments
    b_offset .. b_offset + B - 1; dtype float64 (53-bit draws, double arithmetic) or float32 (the fp32 kernels: the top 24 bits
    of the same draws, float arithmetic)."""
    u_max = np.atleast_1d(np.asarray(u_max, dtype=np.float64))
    m = len(u_max)
    b = np.arange(b_offset, b_offset + B, dtype=np.uint64)[None, None, :]
    k = (np.arange(T, dtype=np.uint64) + np.uint64(k_offset))[:, None, None]
    j = np.arange(m, dtype=np.uint64)[None, :, None]
    if b_offset + B > 2 ** 32 or k_offset + T > 2 ** 30 or m > 4:
        raise ValueError("action stream: b < 2^32, k < 2^30, j < 4")
    r0, r1 = philox2x32_10(b + 0 * k + 0 * j, np.uint64(4) * k + j + 0 * b, stream_key(seed))
    if np.dtype(dtype) == np.float32:
        U = (r1 >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)
        return (np.float32(2.0) * U - np.float32(1.0)) * u_max.astype(np.float32)[None, :, None]
    bits = ((r1 << np.uint64(32)) | r0) >> np.uint64(11)
    U = bits.astype(np.float64) * 2.0 ** -53
    return (2.0 * U - 1.0) * u_max[None, :, None]
