"""ark-groth16 PreparedVerifyingKey wire format and verify_with_processed_vk -- TEST INFRASTRUCTURE ONLY
(row f4 of SURVEY.md 8(f); reference call sites src/arkworks/matrix_proof_of_work/io.rs:62-77,
src/arkworks/backend/matrix_proof.rs:134-136,199-206).

PARITY STATUS: "parity unpinned" (ark-ec 0.4.2 / ark-groth16 0.4 / ark-serialize 0.4 are not on disk).
This restatement is deliberately built differently from the product's host_pairing.hpp so that the
byte comparison between the two means something:
  * alpha_g1_beta_g2 comes from this oracle's OWN pairing (affine Miller loop with slopes over the flat
    Fq[w]/(w^12 - 2 w^6 + 2) representation, final exponentiation by plain powering), then mapped to
    arkworks' value: conjugate (x < 0) and cube (arkworks' hard part computes 3 (q^4 - q^2 + 1) / r);
  * only the G2Prepared line coefficients follow the same published formulas (homogeneous projective
    doubling / addition steps of ark-ec models/bls12/g2.rs, M-type twist order), because their exact
    scaling is part of the wire format.
"""
from oracle import bls12_381 as O

F = O.Fq2Ops
Q = O.Q_MOD


def _fq_le(v):
    return (v % Q).to_bytes(48, "little")


def _fq2_le(a):
    return _fq_le(a[0]) + _fq_le(a[1])


def g2_prepared_coeffs(q):
    """ell_coeffs of G2Prepared::from(q): 63 doubling + 5 addition steps over the bits of |x| below the top one."""
    if q is None:
        return []
    two_inv = pow(2, -1, Q)
    b = (4, 4)
    rx, ry, rz = q[0], q[1], (1, 0)
    out = []
    for bit in bin(O.BLS_X)[3:]:
        a = F.muli(F.mul(rx, ry), two_inv)
        bb, c = F.sqr(ry), F.sqr(rz)
        e = F.mul(b, F.muli(c, 3))
        f = F.muli(e, 3)
        g = F.muli(F.add(bb, f), two_inv)
        h = F.sub(F.sqr(F.add(ry, rz)), F.add(bb, c))
        i = F.sub(e, bb)
        j = F.sqr(rx)
        e2 = F.sqr(e)
        rx = F.mul(a, F.sub(bb, f))
        ry = F.sub(F.sqr(g), F.muli(e2, 3))
        rz = F.mul(bb, h)
        out.append((i, F.muli(j, 3), F.neg(h)))
        if bit == "1":
            theta = F.sub(ry, F.mul(q[1], rz))
            lam = F.sub(rx, F.mul(q[0], rz))
            c, d = F.sqr(theta), F.sqr(lam)
            e, f, g = F.mul(lam, d), F.mul(rz, c), F.mul(rx, d)
            h = F.sub(F.add(e, f), F.muli(g, 2))
            rx = F.mul(lam, h)
            ry = F.sub(F.mul(theta, F.sub(g, h)), F.mul(e, ry))
            rz = F.mul(rz, e)
            j = F.sub(F.mul(theta, q[0]), F.mul(lam, q[1]))
            out.append((j, F.neg(theta), lam))
    return out


def _prepared_bytes(q):
    coeffs = g2_prepared_coeffs(q)
    out = len(coeffs).to_bytes(8, "little")
    for c in coeffs:
        out += _fq2_le(c[0]) + _fq2_le(c[1]) + _fq2_le(c[2])
    return out + (b"\x01" if q is None else b"\x00")


def ark_pairing(P, Q2):
    """E::pairing(P, Q).0 as arkworks computes it, from this oracle's independent pairing:
    conj(f_|x|)^(3 (q^12 - 1) / r) in the flat representation."""
    f = O.miller_loop(P, Q2)
    conj = [(-c) % Q if k & 1 else c for k, c in enumerate(f)]
    return O._f12_pow(conj, 3 * ((Q ** 12 - 1) // O.R_MOD))


def fq12_flat_to_tower(t):
    """flat coefficients of w^0..w^11 -> the six Fq2 coefficients in arkworks order
    c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2 (positions w^0, w^2, w^4, w^1, w^3, w^5)."""
    out = []
    for k in (0, 2, 4, 1, 3, 5):
        b = t[k + 6]
        out.append(((t[k] + b) % Q, b % Q))
    return out


def prepare_verifying_key_bytes(alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1):
    """prepare_verifying_key(vk) then PreparedVerifyingKey::serialize_compressed."""
    out = O.g1_compress(alpha_g1) + O.g2_compress(beta_g2) + O.g2_compress(gamma_g2) + O.g2_compress(delta_g2)
    out += len(gamma_abc_g1).to_bytes(8, "little")
    for p in gamma_abc_g1:
        out += O.g1_compress(p)
    for c in fq12_flat_to_tower(ark_pairing(alpha_g1, beta_g2)):
        out += _fq2_le(c)
    out += _prepared_bytes(O.G2.neg(gamma_g2))
    out += _prepared_bytes(O.G2.neg(delta_g2))
    return out


def parse_pvk(blob):
    """The fields verification needs back out of the bytes (points decompressed, target element in tower order)."""
    at = 0

    def take(n):
        nonlocal at
        s = blob[at:at + n]
        assert len(s) == n, "truncated pvk"
        at += n
        return s
    alpha = O.g1_decompress(take(48))
    beta, gamma, delta = (O.g2_decompress(take(96)) for _ in range(3))
    cnt = int.from_bytes(take(8), "little")
    abc = [O.g1_decompress(take(48)) for _ in range(cnt)]
    target = [(int.from_bytes(take(48), "little"), int.from_bytes(take(48), "little")) for _ in range(6)]
    rest = blob[at:]
    return alpha, beta, gamma, delta, abc, target, rest


def verify_with_processed_vk(blob, public_inputs, proof_bytes):
    """Groth16::verify_with_processed_vk on wire bytes, with this oracle's own pairing:
    e(A, B) e(acc, -gamma) e(C, -delta) == alpha_g1_beta_g2  (all in arkworks' cubed normalisation)."""
    alpha, beta, gamma, delta, abc, target, _ = parse_pvk(blob)
    if len(public_inputs) + 1 != len(abc):
        raise ValueError("MalformedVerifyingKey")
    A, B, C = O.proof_deserialize_compressed(proof_bytes)
    acc = O.G1.to_jac(abc[0])
    for x, p in zip(public_inputs, abc[1:]):
        acc = O.G1.jadd(acc, O.G1.jmul(O.G1.to_jac(p), x % O.R_MOD))
    acc = O.G1.to_affine(acc)
    f = O._F12_ONE
    for P, Q2 in ((A, B), (acc, O.G2.neg(gamma)), (C, O.G2.neg(delta))):
        f = O._f12_mul(f, O.miller_loop(P, Q2))
    conj = [(-c) % Q if k & 1 else c for k, c in enumerate(f)]
    got = O._f12_pow(conj, 3 * ((Q ** 12 - 1) // O.R_MOD))
    return fq12_flat_to_tower(got) == target
