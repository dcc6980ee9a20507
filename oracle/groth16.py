"""Groth16 prove path restated in plain Python -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this file.  PARITY STATUS: "parity unpinned" (see oracle/bls12_381.py header):
the algorithm lives in un-vendored crates -- ark-groth16 ^0.4.0
(Cargo.toml:18), ark-poly ^0.4.2 (Cargo.toml:43), ark-ec ^0.4.2 (Cargo.toml:14)
-- so every function restates the published algorithm
and is anchored on the reference call sites:

  Groth16::prove          src/arkworks/backend/fibbonaci_handler.rs:110,
                          matrix_proof.rs:139-140, prime_snark.rs:119
  Groth16::setup          fibbonaci_handler.rs:107, matrix_proof.rs:129
  verify_with_processed_vk fibbonaci_handler.rs:129, matrix_proof.rs:200
  serialize_compressed    src/arkworks/matrix_proof_of_work/io.rs:48

All field values here are canonical integers (NOT Montgomery form).
"""
import random

from .bls12_381 import (R_MOD, FR_ROOT_OF_UNITY, FR_TWO_ADICITY, FR_GENERATOR,
                        G1, G2, G1_GEN, G2_GEN, FixedBase,
                        proof_serialize_compressed, pairing_product_is_one)


# ----------------------------------------------------------------------------
# ark-poly Radix2EvaluationDomain<Fr>  
# ----------------------------------------------------------------------------
class Radix2EvaluationDomain:
    def __init__(self, num_coeffs, offset=1):
        size = 1
        log = 0
        while size < num_coeffs:
            size <<= 1
            log += 1
        if log > FR_TWO_ADICITY:
            raise ValueError("PolynomialDegreeTooLarge")
        self.size, self.log_size = size, log
        self.group_gen = pow(FR_ROOT_OF_UNITY, 1 << (FR_TWO_ADICITY - log), R_MOD)
        self.group_gen_inv = pow(self.group_gen, -1, R_MOD)
        self.size_inv = pow(size, -1, R_MOD)
        self.offset = offset % R_MOD
        self.offset_inv = pow(self.offset, -1, R_MOD)

    def get_coset(self, offset):
        d = Radix2EvaluationDomain(self.size)
        d.offset = offset % R_MOD
        d.offset_inv = pow(d.offset, -1, R_MOD)
        return d

    @staticmethod
    def _ntt_core(v, w):
        """In-place iterative radix-2 NTT, natural order in and out."""
        n = len(v)
        j = 0
        for i in range(1, n):                      # bit reversal
            bit = n >> 1
            while j & bit:
                j ^= bit
                bit >>= 1
            j |= bit
            if i < j:
                v[i], v[j] = v[j], v[i]
        m = 2
        while m <= n:
            wm = pow(w, n // m, R_MOD)
            half = m >> 1
            tw = [1] * half
            for k in range(1, half):
                tw[k] = tw[k - 1] * wm % R_MOD
            for s in range(0, n, m):
                for k in range(half):
                    t = v[s + k + half] * tw[k] % R_MOD
                    u = v[s + k]
                    v[s + k] = (u + t) % R_MOD
                    v[s + k + half] = (u - t) % R_MOD
            m <<= 1
        return v

    def fft(self, v):
        """v_hat[k] = sum_j v[j] (offset * w^k)^j ; zero-pads to the domain size."""
        v = [x % R_MOD for x in v] + [0] * (self.size - len(v))
        if self.offset != 1:
            g = 1
            for i in range(self.size):
                v[i] = v[i] * g % R_MOD
                g = g * self.offset % R_MOD
        return self._ntt_core(v, self.group_gen)

    def ifft(self, v):
        v = [x % R_MOD for x in v] + [0] * (self.size - len(v))
        self._ntt_core(v, self.group_gen_inv)
        g = self.size_inv
        for i in range(self.size):
            v[i] = v[i] * g % R_MOD
            if self.offset != 1:
                g = g * self.offset_inv % R_MOD
        return v

    def evaluate_vanishing_polynomial(self, tau):
        return (pow(tau, self.size, R_MOD) - pow(self.offset, self.size, R_MOD)) % R_MOD

    def evaluate_all_lagrange_coefficients(self, tau):
        """L_i(tau), i < size, for the base domain (offset 1)."""
        n = self.size
        z = (pow(tau, n, R_MOD) - 1) % R_MOD
        if z == 0:
            out = [0] * n
            w = 1
            for i in range(n):
                if w == tau % R_MOD:
                    out[i] = 1
                w = w * self.group_gen % R_MOD
            return out
        zn = z * self.size_inv % R_MOD
        out = []
        w = 1
        for i in range(n):
            out.append(zn * w % R_MOD * pow((tau - w) % R_MOD, -1, R_MOD) % R_MOD)
            w = w * self.group_gen % R_MOD
        return out


# ----------------------------------------------------------------------------
# R1CS container:  rows are lists of (coeff, column) like ark-relations'
# ConstraintMatrices; z = instance (z[0] = 1) || witness.
# ----------------------------------------------------------------------------
class R1CS:
    def __init__(self, num_instance, num_witness, a, b, c):
        self.num_instance = num_instance          # includes the constant 1
        self.num_witness = num_witness
        self.a, self.b, self.c = a, b, c
        self.num_constraints = len(a)

    @property
    def num_variables(self):
        return self.num_instance + self.num_witness

    def is_satisfied(self, z):
        for ra, rb, rc in zip(self.a, self.b, self.c):
            if (evaluate_constraint(ra, z) * evaluate_constraint(rb, z)
                    - evaluate_constraint(rc, z)) % R_MOD:
                return False
        return True


def evaluate_constraint(row, z):
    acc = 0
    for coeff, col in row:
        acc += coeff * z[col]
    return acc % R_MOD


def constraint_evaluations(r1cs, z):
    """The a, b, c vectors LibsnarkReduction builds before its FFTs (A.3):
    a[i] = <A_i, z>, b[i] = <B_i, z>, c[i] = <C_i, z> for i < num_constraints,
    a[num_constraints + j] = z[j] for j < num_instance, zero elsewhere."""
    dom = Radix2EvaluationDomain(r1cs.num_constraints + r1cs.num_instance)
    n = dom.size
    a = [0] * n
    b = [0] * n
    c = [0] * n
    for i in range(r1cs.num_constraints):
        a[i] = evaluate_constraint(r1cs.a[i], z)
        b[i] = evaluate_constraint(r1cs.b[i], z)
        c[i] = evaluate_constraint(r1cs.c[i], z)
    for j in range(r1cs.num_instance):
        a[r1cs.num_constraints + j] = z[j] % R_MOD
    return a, b, c


def witness_map_from_evals(a, b, c):
    """LibsnarkReduction::witness_map_from_matrices after the SpMV (A.3):
    3 iFFT, 3 coset FFT (g = 7), (ab - c)/Z_H on the coset, 1 coset iFFT."""
    n = len(a)
    dom = Radix2EvaluationDomain(n)
    coset = dom.get_coset(FR_GENERATOR)
    a = coset.fft(dom.ifft(a))
    b = coset.fft(dom.ifft(b))
    c = coset.fft(dom.ifft(c))
    zinv = pow(dom.evaluate_vanishing_polynomial(FR_GENERATOR), -1, R_MOD)
    ab = [(x * y - w) % R_MOD * zinv % R_MOD for x, y, w in zip(a, b, c)]
    return coset.ifft(ab)


# ----------------------------------------------------------------------------
# ark-ec VariableBaseMSM::msm_bigint restated (A.4): signed-digit windows,
# c = 3 if N < 32 else ceil(log2 N)*69/100 + 2, running-sum bucket reduction,
# Horner combine high -> low.
# ----------------------------------------------------------------------------
def _ln_without_floats(n):
    # ark_std::log2(n) = ceil(log2 n); then * 69 / 100 in integers
    lg = (n - 1).bit_length() if n > 1 else 0
    return lg * 69 // 100


def ark_window_size(n):
    return 3 if n < 32 else _ln_without_floats(n) + 2


def make_digits(k, w, num_bits=255):
    """ark-ec make_digits: signed radix-2^w digits, last window takes the carry."""
    radix = 1 << w
    window_mask = radix - 1
    digits_count = (num_bits + w - 1) // w
    digits = []
    carry = 0
    for i in range(digits_count):
        coef = carry + ((k >> (i * w)) & window_mask)
        carry = (coef + radix // 2) >> w
        d = coef - (carry << w)
        if i == digits_count - 1:
            d += carry << w
        digits.append(d)
    return digits


def msm_bigint(curve, bases, scalars):
    """Returns a Jacobian point.  bases: affine or None; scalars: canonical ints."""
    F = curve.F
    size = min(len(bases), len(scalars))
    ident = (F.one, F.one, F.zero)
    if size == 0:
        return ident
    c = ark_window_size(size)
    digits = [make_digits(s % R_MOD, c) for s in scalars[:size]]
    nwin = len(digits[0])
    window_sums = []
    for w in range(nwin):
        buckets = [ident] * (1 << c)     # ark-ec 0.4.2 allocates 2^c: the top digit keeps its carry
        for i in range(size):
            d = digits[i][w]
            if d > 0:
                buckets[d - 1] = curve.jadd_affine(buckets[d - 1], bases[i])
            elif d < 0:
                buckets[-d - 1] = curve.jadd_affine(buckets[-d - 1], curve.neg(bases[i]))
        running = ident
        res = ident
        for bk in reversed(buckets):
            running = curve.jadd(running, bk)
            res = curve.jadd(res, running)
        window_sums.append(res)
    total = window_sums[-1]
    for ws in reversed(window_sums[:-1]):
        for _ in range(c):
            total = curve.jdouble(total)
        total = curve.jadd(total, ws)
    return total


def msm_naive(curve, bases, scalars):
    F = curve.F
    acc = (F.one, F.one, F.zero)
    for P, k in zip(bases, scalars):
        if P is not None:
            acc = curve.jadd(acc, curve.jmul(curve.to_jac(P), k % R_MOD))
    return acc


# ----------------------------------------------------------------------------
# Groth16 setup / prove / verify  (A.4)
# ----------------------------------------------------------------------------
class ProvingKey:
    pass


def setup(r1cs, seed=0xB2000004, toxic=None, g1_gen=None, g2_gen=None):
    """generate_parameters_with_qap semantics; by default with the standard
    generators and the oracle's own toxic waste (pk layouts match ark-groth16's
    ProvingKey so the keys are interchangeable).  g1_gen / g2_gen: the random
    generators Groth16::setup draws (oracle/ark_rng.py setup_draws)."""
    rnd = random.Random(seed)
    if toxic is None:
        toxic = [rnd.randrange(1, R_MOD) for _ in range(5)]
    alpha, beta, gamma, delta, tau = toxic
    nc, l, m = r1cs.num_constraints, r1cs.num_instance, r1cs.num_variables
    dom = Radix2EvaluationDomain(nc + l)
    n = dom.size
    lag = dom.evaluate_all_lagrange_coefficients(tau)
    At = [0] * m
    Bt = [0] * m
    Ct = [0] * m
    for j in range(l):
        At[j] = lag[nc + j]
    for i in range(nc):
        u = lag[i]
        for coeff, col in r1cs.a[i]:
            At[col] = (At[col] + u * coeff) % R_MOD
        for coeff, col in r1cs.b[i]:
            Bt[col] = (Bt[col] + u * coeff) % R_MOD
        for coeff, col in r1cs.c[i]:
            Ct[col] = (Ct[col] + u * coeff) % R_MOD
    zt = dom.evaluate_vanishing_polynomial(tau)
    ginv = pow(gamma, -1, R_MOD)
    dinv = pow(delta, -1, R_MOD)
    fb1 = FixedBase(G1, G1_GEN if g1_gen is None else g1_gen)
    fb2 = FixedBase(G2, G2_GEN if g2_gen is None else g2_gen)
    pk = ProvingKey()
    pk.domain_size, pk.num_instance, pk.num_variables = n, l, m
    pk.a_query = fb1.mul_many(At)
    pk.b_g1_query = fb1.mul_many(Bt)
    pk.b_g2_query = fb2.mul_many(Bt)
    hs, t = [], zt * dinv % R_MOD
    for _ in range(n - 1):
        hs.append(t)
        t = t * tau % R_MOD
    pk.h_query = fb1.mul_many(hs)
    abc = [(beta * At[i] + alpha * Bt[i] + Ct[i]) % R_MOD for i in range(m)]
    pk.l_query = fb1.mul_many([x * dinv % R_MOD for x in abc[l:]])
    pk.alpha_g1, pk.beta_g1, pk.delta_g1 = fb1.mul_many([alpha, beta, delta])
    pk.beta_g2, pk.gamma_g2, pk.delta_g2 = fb2.mul_many([beta, gamma, delta])
    pk.gamma_abc_g1 = fb1.mul_many([x * ginv % R_MOD for x in abc[:l]])
    return pk


def create_proof_with_assignment(pk, r, s, h, z):
    """ark-groth16 prover.rs create_proof_with_assignment (A.4).  z is the full
    assignment (z[0] = 1); returns affine (A, B, C)."""
    l = pk.num_instance
    j1 = G1.to_jac
    h_acc = msm_bigint(G1, pk.h_query, h)                 # min(len) semantics
    l_acc = msm_bigint(G1, pk.l_query, z[l:])
    d1 = j1(pk.delta_g1)

    def calc(curve, query, vk_param, delta, scalar):
        acc = msm_bigint(curve, query[1:], z[1:])
        res = curve.jmul(curve.to_jac(delta), scalar)
        res = curve.jadd_affine(res, query[0])
        res = curve.jadd(res, acc)
        return curve.jadd_affine(res, vk_param)

    g_a = calc(G1, pk.a_query, pk.alpha_g1, pk.delta_g1, r)
    if r % R_MOD == 0:
        g1_b = (1, 1, 0)
    else:
        g1_b = calc(G1, pk.b_g1_query, pk.beta_g1, pk.delta_g1, s)
    g2_b = calc(G2, pk.b_g2_query, pk.beta_g2, pk.delta_g2, s)
    g_c = G1.jmul(g_a, s)
    g_c = G1.jadd(g_c, G1.jmul(g1_b, r))
    g_c = G1.jadd(g_c, G1.jmul(d1, -(r * s % R_MOD)))
    g_c = G1.jadd(g_c, l_acc)
    g_c = G1.jadd(g_c, h_acc)
    return G1.to_affine(g_a), G2.to_affine(g2_b), G1.to_affine(g_c)


def prove(pk, r1cs, z, r, s):
    """create_proof_with_reduction: witness map then the MSM/combine stage.
    Returns (A, B, C) affine and the 192 serialized bytes."""
    a, b, c = constraint_evaluations(r1cs, z)
    h = witness_map_from_evals(a, b, c)
    A, B, C = create_proof_with_assignment(pk, r, s, h, z)
    return (A, B, C), proof_serialize_compressed(A, B, C)


def verify(pk, public_inputs, proof):
    """e(A,B) = e(alpha,beta) e(sum x_i gamma_abc_i, gamma) e(C, delta);
    public_inputs excludes the constant 1 (as Groth16::verify takes them)."""
    A, B, C = proof
    acc = G1.to_jac(pk.gamma_abc_g1[0])
    for x, P in zip(public_inputs, pk.gamma_abc_g1[1:]):
        acc = G1.jadd(acc, G1.jmul(G1.to_jac(P), x % R_MOD))
    L = G1.to_affine(acc)
    return pairing_product_is_one([
        (A, B), (G1.neg(pk.alpha_g1), pk.beta_g2),
        (G1.neg(L), pk.gamma_g2), (G1.neg(C), pk.delta_g2)])
