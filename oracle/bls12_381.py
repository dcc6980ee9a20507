"""BLS12-381 arithmetic in plain Python big integers -- TEST INFRASTRUCTURE ONLY.

This file is part of the parity oracle.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import it; the product path
(zksnark-finalproject_b200/) never does.

PARITY STATUS: "parity unpinned".  The arithmetic on the reference's hot path
lives in un-vendored crates (ark-ff 0.4, ark-ec ^0.4.2, ark-bls12-381 ^0.4.0,
ark-serialize ^0.4.0; /root/reference/Cargo.toml:11-15) and the reference's own
tests pin no bytes.  What pins this oracle instead:
standard BLS12-381 known answers (generator encodings, subgroup order, curve
membership), algebraic identities, and pairing verification of every proof --
see tests/test_oracle_*.py.

Restated behaviour:
  * Fr, Fq moduli and Montgomery radices (R = 2^256, 2^384) of ark-bls12-381.
  * G1: y^2 = x^3 + 4 over Fq;  G2: y^2 = x^3 + 4(1+u) over Fq2 = Fq[u]/(u^2+1).
  * Point encoding = zcash/IETF big-endian compressed form used by
    ark-bls12-381 0.4 `serialize_compressed`, which the reference calls at
    /root/reference/src/arkworks/matrix_proof_of_work/io.rs:48.
"""

# ----------------------------------------------------------------------------
# Field constants
# ----------------------------------------------------------------------------
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001   # Fr
Q_MOD = int(
    "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f624"
    "1eabfffeb153ffffb9feffffffffaaab", 16)                                   # Fq
FR_BITS = 255
FR_MONT_R = (1 << 256) % R_MOD
FQ_MONT_R = (1 << 384) % Q_MOD
FR_GENERATOR = 7                      # ark Fr::GENERATOR (multiplicative generator)
FR_TWO_ADICITY = 32
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R_MOD - 1) >> FR_TWO_ADICITY, R_MOD)   # order 2^32
BLS_X = 0xD201000000010000            # |x|, the curve parameter is -x

G1_B = 4
G2_B = (4, 4)                         # 4(1+u)

G1_GEN = (
    int("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
        "6c55e83ff97a1aeffb3af00adb22c6bb", 16),
    int("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3ed"
        "d03cc744a2888ae40caa232946c5e7e1", 16),
)
G2_GEN = (
    (int("024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d177"
         "0bac0326a805bbefd48056c8c121bdb8", 16),
     int("13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049"
         "334cf11213945d57e5ac7d055d042b7e", 16)),
    (int("0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c"
         "923ac9cc3baca289e193548608b82801", 16),
     int("0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab"
         "3f370d275cec1da1aaa9075ff05f79be", 16)),
)


# ----------------------------------------------------------------------------
# Field "operation tables": one tiny class per field so the group law below is
# written once.  Elements: Fq -> int, Fq2 -> (c0, c1).
# ----------------------------------------------------------------------------
class FqOps:
    zero = 0
    one = 1

    @staticmethod
    def add(a, b): return (a + b) % Q_MOD
    @staticmethod
    def sub(a, b): return (a - b) % Q_MOD
    @staticmethod
    def neg(a): return (-a) % Q_MOD
    @staticmethod
    def mul(a, b): return (a * b) % Q_MOD
    @staticmethod
    def sqr(a): return (a * a) % Q_MOD
    @staticmethod
    def inv(a): return pow(a, -1, Q_MOD)
    @staticmethod
    def is_zero(a): return a % Q_MOD == 0
    @staticmethod
    def muli(a, k): return (a * k) % Q_MOD
    @staticmethod
    def eq(a, b): return (a - b) % Q_MOD == 0


class Fq2Ops:
    zero = (0, 0)
    one = (1, 0)

    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % Q_MOD, (a[1] + b[1]) % Q_MOD)
    @staticmethod
    def sub(a, b): return ((a[0] - b[0]) % Q_MOD, (a[1] - b[1]) % Q_MOD)
    @staticmethod
    def neg(a): return ((-a[0]) % Q_MOD, (-a[1]) % Q_MOD)
    @staticmethod
    def mul(a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % Q_MOD, (a[0] * b[1] + a[1] * b[0]) % Q_MOD)
    @staticmethod
    def sqr(a):
        return ((a[0] + a[1]) * (a[0] - a[1]) % Q_MOD, 2 * a[0] * a[1] % Q_MOD)
    @staticmethod
    def inv(a):
        d = pow(a[0] * a[0] + a[1] * a[1], -1, Q_MOD)
        return (a[0] * d % Q_MOD, (-a[1]) * d % Q_MOD)
    @staticmethod
    def is_zero(a): return a[0] % Q_MOD == 0 and a[1] % Q_MOD == 0
    @staticmethod
    def muli(a, k): return (a[0] * k % Q_MOD, a[1] * k % Q_MOD)
    @staticmethod
    def eq(a, b): return (a[0] - b[0]) % Q_MOD == 0 and (a[1] - b[1]) % Q_MOD == 0


# ----------------------------------------------------------------------------
# Short-Weierstrass group law, a = 0.  Affine points are (x, y) or None for the
# identity; Jacobian points are (X, Y, Z) with Z == 0 for the identity.
# ----------------------------------------------------------------------------
class Curve:
    def __init__(self, F, b, gen, name):
        self.F, self.b, self.gen, self.name = F, b, gen, name

    # -- affine helpers --------------------------------------------------
    def is_on_curve(self, P):
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.eq(F.sqr(y), F.add(F.mul(F.sqr(x), x), self.b))

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    # -- Jacobian --------------------------------------------------------
    def to_jac(self, P):
        F = self.F
        return (F.one, F.one, F.zero) if P is None else (P[0], P[1], F.one)

    def to_affine(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def jdouble(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return J
        A = F.sqr(X)
        B = F.sqr(Y)
        C = F.sqr(B)
        D = F.muli(F.sub(F.sub(F.sqr(F.add(X, B)), A), C), 2)
        E = F.muli(A, 3)
        Fv = F.sqr(E)
        X3 = F.sub(Fv, F.muli(D, 2))
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.muli(C, 8))
        Z3 = F.muli(F.mul(Y, Z), 2)
        return (X3, Y3, Z3)

    def jadd(self, P, Q):
        F = self.F
        X1, Y1, Z1 = P
        X2, Y2, Z2 = Q
        if F.is_zero(Z1):
            return Q
        if F.is_zero(Z2):
            return P
        Z1Z1 = F.sqr(Z1)
        Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
        S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if F.eq(U1, U2):
            if F.eq(S1, S2):
                return self.jdouble(P)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1)
        Rr = F.sub(S2, S1)
        HH = F.sqr(H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.muli(V, 2))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def jadd_affine(self, P, Q):
        return P if Q is None else self.jadd(P, (Q[0], Q[1], self.F.one))

    def jmul(self, J, k):
        F = self.F
        acc = (F.one, F.one, F.zero)
        if k < 0:
            J = (J[0], F.neg(J[1]), J[2])
            k = -k
        for bit in bin(k)[2:] if k else "":
            acc = self.jdouble(acc)
            if bit == "1":
                acc = self.jadd(acc, J)
        return acc

    # -- affine API ------------------------------------------------------
    def add(self, P, Q):
        return self.to_affine(self.jadd(self.to_jac(P), self.to_jac(Q)))

    def mul(self, P, k):
        return self.to_affine(self.jmul(self.to_jac(P), k))

    def batch_to_affine(self, Js):
        """Montgomery batch inversion of the Z coordinates."""
        F = self.F
        pref, acc = [], F.one
        for J in Js:
            pref.append(acc)
            if not F.is_zero(J[2]):
                acc = F.mul(acc, J[2])
        inv = F.inv(acc)
        out = [None] * len(Js)
        for i in range(len(Js) - 1, -1, -1):
            X, Y, Z = Js[i]
            if F.is_zero(Z):
                continue
            zi = F.mul(inv, pref[i])
            inv = F.mul(inv, Z)
            zi2 = F.sqr(zi)
            out[i] = (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))
        return out


G1 = Curve(FqOps, G1_B, G1_GEN, "G1")
G2 = Curve(Fq2Ops, G2_B, G2_GEN, "G2")


class FixedBase:
    """Windowed fixed-base multiplication table (oracle-side setup helper)."""

    def __init__(self, curve, base, bits=255, w=8):
        self.curve, self.w = curve, w
        self.nwin = (bits + w - 1) // w
        self.table = []
        cur = curve.to_jac(base)
        for _ in range(self.nwin):
            row = [None]
            acc = (curve.F.one, curve.F.one, curve.F.zero)
            for _ in range((1 << w) - 1):
                acc = curve.jadd(acc, cur)
                row.append(acc)
            self.table.append(curve.batch_to_affine(row[1:]))
            for _ in range(w):
                cur = curve.jdouble(cur)

    def mul_jac(self, k):
        c = self.curve
        acc = (c.F.one, c.F.one, c.F.zero)
        mask = (1 << self.w) - 1
        for i in range(self.nwin):
            d = (k >> (i * self.w)) & mask
            if d:
                acc = c.jadd_affine(acc, self.table[i][d - 1])
        return acc

    def mul_many(self, ks):
        return self.curve.batch_to_affine([self.mul_jac(k % R_MOD) for k in ks])


# ----------------------------------------------------------------------------
# Serialization: ark-bls12-381 0.4 (zcash format).  SURVEY.md A.5.
# ----------------------------------------------------------------------------
def _fq_is_larger(y):
    return y > (Q_MOD - 1) // 2


def _fq2_is_larger(y):
    # "lexicographically largest": compare y with -y, c1 first then c0.
    ny = Fq2Ops.neg(y)
    return (y[1], y[0]) > (ny[1], ny[0])


def g1_compress(P):
    if P is None:
        return bytes([0xC0]) + bytes(47)
    b = bytearray(P[0].to_bytes(48, "big"))
    b[0] |= 0x80
    if _fq_is_larger(P[1]):
        b[0] |= 0x20
    return bytes(b)


def g2_compress(P):
    if P is None:
        return bytes([0xC0]) + bytes(95)
    (x0, x1), y = P
    b = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    b[0] |= 0x80
    if _fq2_is_larger(y):
        b[0] |= 0x20
    return bytes(b)


def fq_sqrt(a):
    # q = 3 mod 4
    r = pow(a, (Q_MOD + 1) // 4, Q_MOD)
    return r if r * r % Q_MOD == a % Q_MOD else None


def fq2_sqrt(a):
    """Square root in Fq2 (q = 3 mod 4) via the norm method; None if non-residue."""
    F = Fq2Ops
    if F.is_zero(a):
        return (0, 0)
    a0, a1 = a
    if a1 == 0:
        s = fq_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = fq_sqrt((-a0) % Q_MOD)
        return (0, s)
    norm = (a0 * a0 + a1 * a1) % Q_MOD
    alpha = fq_sqrt(norm)
    if alpha is None:
        return None
    inv2 = pow(2, -1, Q_MOD)
    delta = (a0 + alpha) * inv2 % Q_MOD
    x0 = fq_sqrt(delta)
    if x0 is None:
        delta = (a0 - alpha) * inv2 % Q_MOD
        x0 = fq_sqrt(delta)
        if x0 is None:
            return None
    x1 = a1 * pow(2 * x0, -1, Q_MOD) % Q_MOD
    r = (x0, x1)
    return r if F.eq(F.sqr(r), a) else None


def g1_decompress(b):
    assert len(b) == 48 and b[0] & 0x80
    if b[0] & 0x40:
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    y = fq_sqrt((x * x * x + 4) % Q_MOD)
    assert y is not None, "x not on curve"
    if _fq_is_larger(y) != bool(b[0] & 0x20):
        y = Q_MOD - y
    return (x, y)


def g2_decompress(b):
    assert len(b) == 96 and b[0] & 0x80
    if b[0] & 0x40:
        return None
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    x0 = int.from_bytes(b[48:], "big")
    x = (x0, x1)
    F = Fq2Ops
    y = fq2_sqrt(F.add(F.mul(F.sqr(x), x), G2_B))
    assert y is not None, "x not on twist"
    if _fq2_is_larger(y) != bool(b[0] & 0x20):
        y = F.neg(y)
    return (x, y)


def proof_serialize_compressed(A, B, C):
    """Proof<Bls12_381>::serialize_compressed = A(48) || B(96) || C(48)."""
    return g1_compress(A) + g2_compress(B) + g1_compress(C)


def proof_deserialize_compressed(b):
    assert len(b) == 192
    return g1_decompress(b[:48]), g2_decompress(b[48:144]), g1_decompress(b[144:])


# ----------------------------------------------------------------------------
# Limb packing helpers (the C-ABI layouts: little-endian u64 limbs, Montgomery).
# ----------------------------------------------------------------------------
def fr_to_mont(x): return x * FR_MONT_R % R_MOD
def fr_from_mont(x): return x * pow(FR_MONT_R, -1, R_MOD) % R_MOD
def fq_to_mont(x): return x * FQ_MONT_R % Q_MOD
def fq_from_mont(x): return x * pow(FQ_MONT_R, -1, Q_MOD) % Q_MOD


def int_to_limbs(x, n):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def limbs_to_int(l):
    v = 0
    for i, w in enumerate(l):
        v |= int(w) << (64 * i)
    return v


# ----------------------------------------------------------------------------
# Pairing (only for Groth16 verification in tests).  Fq12 is represented as
# Fq[w]/(w^12 - 2 w^6 + 2) with u = w^6 - 1, so an Fq2 element a + b u sitting
# at w^k is (a - b) w^k + b w^(k+6).
# ----------------------------------------------------------------------------
def _f12_mul(a, b):
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
    for i in range(22, 11, -1):          # w^12 = 2 w^6 - 2
        c = t[i]
        if c:
            t[i - 6] += 2 * c
            t[i - 12] -= 2 * c
    return [x % Q_MOD for x in t[:12]]


_F12_ONE = [1] + [0] * 11


def _f12_pow(a, e):
    r = _F12_ONE
    for bit in bin(e)[2:]:
        r = _f12_mul(r, r)
        if bit == "1":
            r = _f12_mul(r, a)
    return r


def _line(lam, xT, yT, P):
    """w^3-scaled line through the untwisted T with twist-slope lam, at P in G1:
       (lam*xT - yT) + (-lam*xP) w^2 + yP w^3."""
    F = Fq2Ops
    xP, yP = P
    c0 = F.sub(F.mul(lam, xT), yT)
    c2 = F.neg(F.muli(lam, xP))
    out = [0] * 12
    out[0] = (c0[0] - c0[1]) % Q_MOD
    out[6] = c0[1]
    out[2] = (c2[0] - c2[1]) % Q_MOD
    out[8] = c2[1]
    out[3] = yP % Q_MOD
    return out


def miller_loop(P, Q):
    """f_{|x|,Q}(P) up to factors killed by the final exponentiation."""
    if P is None or Q is None:
        return _F12_ONE
    F = Fq2Ops
    T = Q
    f = _F12_ONE
    for bit in bin(BLS_X)[3:]:
        lam = F.mul(F.muli(F.sqr(T[0]), 3), F.inv(F.muli(T[1], 2)))
        f = _f12_mul(_f12_mul(f, f), _line(lam, T[0], T[1], P))
        x3 = F.sub(F.sqr(lam), F.muli(T[0], 2))
        T = (x3, F.sub(F.mul(lam, F.sub(T[0], x3)), T[1]))
        if bit == "1":
            lam = F.mul(F.sub(Q[1], T[1]), F.inv(F.sub(Q[0], T[0])))
            f = _f12_mul(f, _line(lam, T[0], T[1], P))
            x3 = F.sub(F.sub(F.sqr(lam), T[0]), Q[0])
            T = (x3, F.sub(F.mul(lam, F.sub(T[0], x3)), T[1]))
    return f


def final_exponentiation(f):
    return _f12_pow(f, (Q_MOD ** 12 - 1) // R_MOD)


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 ?"""
    f = _F12_ONE
    for P, Q in pairs:
        f = _f12_mul(f, miller_loop(P, Q))
    return final_exponentiation(f) == _F12_ONE
