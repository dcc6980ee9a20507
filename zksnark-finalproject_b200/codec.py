"""Limb packing between Python integers and the C-ABI layouts of include/b200zk.h.

Fr / Fq elements cross the boundary as little-endian uint64 limbs in Montgomery
form (arkworks' in-memory representation, ark-ff 0.4); MSM scalars cross as
canonical "bigint" limbs.  Points are (x, y) tuples of ints (G1) or of (c0, c1)
pairs (G2); None is the identity.
"""
import numpy as np

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
Q_MOD = int("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f624"
            "1eabfffeb153ffffb9feffffffffaaab", 16)
FR_R = (1 << 256) % R_MOD
FQ_R = (1 << 384) % Q_MOD
FR_RINV = pow(FR_R, -1, R_MOD)
FQ_RINV = pow(FQ_R, -1, Q_MOD)


def _pack(values, nbytes):
    buf = b"".join(int(v).to_bytes(nbytes, "little") for v in values)
    return np.frombuffer(buf, dtype=np.uint64).reshape(len(values), nbytes // 8).copy()


def _unpack(arr, nbytes):
    raw = np.ascontiguousarray(arr, dtype=np.uint64).tobytes()
    return [int.from_bytes(raw[i:i + nbytes], "little") for i in range(0, len(raw), nbytes)]


def fr_to_mont_limbs(values):
    """list[int] -> (n, 4) uint64, Montgomery form."""
    return _pack([(v % R_MOD) * FR_R % R_MOD for v in values], 32) if len(values) else np.zeros((0, 4), np.uint64)


def fr_from_mont_limbs(arr):
    return [v * FR_RINV % R_MOD for v in _unpack(arr, 32)]


def fr_to_bigint_limbs(values):
    """list[int] -> (n, 4) uint64, canonical (into_bigint())."""
    return _pack([v % R_MOD for v in values], 32) if len(values) else np.zeros((0, 4), np.uint64)


def fr_from_bigint_limbs(arr):
    return _unpack(arr, 32)


def fq_to_mont(v):
    return (v % Q_MOD) * FQ_R % Q_MOD


def fq_from_mont(v):
    return v * FQ_RINV % Q_MOD


def _inf_bitmap(points):
    n = len(points)
    bits = np.zeros((n + 7) // 8, dtype=np.uint8)
    for i, p in enumerate(points):
        if p is None:
            bits[i >> 3] |= 1 << (i & 7)
    return bits


def g1_to_limbs(points):
    """list of (x, y) | None -> ((n, 12) uint64 Montgomery, identity bitmap)."""
    flat = []
    for p in points:
        if p is None:
            flat += [0, 0]
        else:
            flat += [fq_to_mont(p[0]), fq_to_mont(p[1])]
    arr = _pack(flat, 48).reshape(len(points), 12) if points else np.zeros((0, 12), np.uint64)
    return arr, _inf_bitmap(points)


def g2_to_limbs(points):
    flat = []
    for p in points:
        if p is None:
            flat += [0, 0, 0, 0]
        else:
            (x0, x1), (y0, y1) = p
            flat += [fq_to_mont(x0), fq_to_mont(x1), fq_to_mont(y0), fq_to_mont(y1)]
    arr = _pack(flat, 48).reshape(len(points), 24) if points else np.zeros((0, 24), np.uint64)
    return arr, _inf_bitmap(points)


def g1_from_limbs(arr, inf_bitmap=None):
    vals = [fq_from_mont(v) for v in _unpack(arr, 48)]
    out = []
    for i in range(len(vals) // 2):
        if inf_bitmap is not None and (inf_bitmap[i >> 3] >> (i & 7)) & 1:
            out.append(None)
        else:
            out.append((vals[2 * i], vals[2 * i + 1]))
    return out


def g2_from_limbs(arr, inf_bitmap=None):
    vals = [fq_from_mont(v) for v in _unpack(arr, 48)]
    out = []
    for i in range(len(vals) // 4):
        if inf_bitmap is not None and (inf_bitmap[i >> 3] >> (i & 7)) & 1:
            out.append(None)
        else:
            out.append(((vals[4 * i], vals[4 * i + 1]), (vals[4 * i + 2], vals[4 * i + 3])))
    return out


def g1_projective_from_limbs(out18):
    """18 limbs (X, Y, Z Jacobian, Montgomery) -> (X, Y, Z) ints."""
    v = [fq_from_mont(x) for x in _unpack(out18, 48)]
    return (v[0], v[1], v[2])


def g2_projective_from_limbs(out36):
    v = [fq_from_mont(x) for x in _unpack(out36, 48)]
    return ((v[0], v[1]), (v[2], v[3]), (v[4], v[5]))
