"""R1CS builders for the three circuit families the reference proves.

The circuit front-end (ark-r1cs-std gadgets + ConstraintSynthesizer) stays on the
host in the reference and is out of the GPU path's scope; what the prover
consumes is only (A, B, C matrices, full assignment z).  These builders produce
systems with the SAME SHAPE as the reference circuits -- constraint counts, the
number of public inputs, row sparsity after linear-combination inlining and the
witness value distribution -- so that MSM / NTT sizes and scalar statistics match:

  fibonacci        src/arkworks/constraints/fibbonaci.rs:22-48
                   l = 4 (1, a, b, result), 1 witness, num_steps + 1 constraints
  matrix           src/arkworks/matrix_proof_of_work/constraints.rs:78-128 and
                   hasher.rs:30-40: 2 n^3 product constraints + three Poseidon
                   sponges (rate 2, capacity 1, alpha = 17, 8 full + 29 partial
                   rounds, hashing_utils.rs:701-705), 265 constraints per
                   permutation; total 3*ceil(n^2/2)*265 + 2 n^3 + 3
  prime            src/arkworks/prime_snark/prime_circut.rs:92-146,
                   fermat_circut.rs:56-129: SHA-256-like Boolean-heavy witness +
                   K = 3 bit-serial modular exponentiations of NUM_BITS bits

Poseidon round constants / MDS here are derived from a fixed seed, NOT the
reference's table (hashing_utils.rs:15-715 is data of the reference and is not
copied): the arithmetic shape is identical, the digest values differ.  The
SHA-256 part of `prime` is a synthetic Boolean circuit of the same size class,
not a bit-exact SHA-256 (documented in DESIGN.md, "out of scope").

Rows are lists of (coeff, column); z[0] = 1; columns < num_instance are public.
"""
import random

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


class ConstraintSystem:
    """Minimal ark_relations::r1cs::ConstraintSystem: variables, linear
    combinations (dict column -> coeff) inlined eagerly, constraint rows."""

    def __init__(self):
        self.instance = [1]
        self.witness = []
        self.a, self.b, self.c = [], [], []

    # variables are encoded as ('i', k) / ('w', k) until finalised
    def new_input(self, value):
        self.instance.append(value % R_MOD)
        return {("i", len(self.instance) - 1): 1}

    def new_witness(self, value):
        self.witness.append(value % R_MOD)
        return {("w", len(self.witness) - 1): 1}

    @staticmethod
    def constant(v):
        return {("i", 0): v % R_MOD} if v % R_MOD else {}

    @staticmethod
    def lc_add(x, y):
        out = dict(x)
        for k, v in y.items():
            nv = (out.get(k, 0) + v) % R_MOD
            if nv:
                out[k] = nv
            else:
                out.pop(k, None)
        return out

    @staticmethod
    def lc_scale(x, s):
        s %= R_MOD
        return {k: v * s % R_MOD for k, v in x.items()} if s else {}

    def lc_sub(self, x, y):
        return self.lc_add(x, self.lc_scale(y, R_MOD - 1))

    def value(self, lc):
        acc = 0
        for (kind, k), v in lc.items():
            acc += v * (self.instance[k] if kind == "i" else self.witness[k])
        return acc % R_MOD

    def enforce(self, a, b, c):
        self.a.append(a)
        self.b.append(b)
        self.c.append(c)

    def mul(self, x, y):
        """x * y as a fresh witness with one constraint (FpVar::mul)."""
        out = self.new_witness(self.value(x) * self.value(y))
        self.enforce(x, y, out)
        return out

    def enforce_equal(self, x, y):
        """(x - y) * 1 = 0  (EqGadget::enforce_equal)."""
        self.enforce(self.lc_sub(x, y), self.constant(1), {})

    def finalize(self):
        """-> R1CSInstance with columns numbered instance-first."""
        l = len(self.instance)

        def row(lc):
            return sorted((v, (k if kind == "i" else l + k)) for (kind, k), v in lc.items()) if lc else []

        rows = lambda m: [[(v, col) for v, col in row(lc)] for lc in m]
        return R1CSInstance(l, len(self.witness), rows(self.a), rows(self.b), rows(self.c),
                            list(self.instance) + list(self.witness))


class R1CSInstance:
    def __init__(self, num_instance, num_witness, a, b, c, z):
        self.num_instance, self.num_witness = num_instance, num_witness
        self.a, self.b, self.c, self.z = a, b, c, z
        self.num_constraints = len(a)

    @property
    def num_variables(self):
        return self.num_instance + self.num_witness

    @property
    def domain_size(self):
        n = 1
        while n < self.num_constraints + self.num_instance:
            n <<= 1
        return n

    @property
    def matrices(self):
        return (self.a, self.b, self.c)

    def is_satisfied(self):
        def ev(r):
            return sum(v * self.z[col] for v, col in r) % R_MOD
        return all((ev(ra) * ev(rb) - ev(rc)) % R_MOD == 0 for ra, rb, rc in zip(self.a, self.b, self.c))


# ---------------------------------------------------------------------------
def fibonacci_circuit(a, b, num_steps):
    """FibonacciCircuit (fibbonaci.rs:22-48).  The result is computed in Fr, not
    u128, so the circuit stays satisfiable past step 186."""
    cs = ConstraintSystem()
    f2 = cs.new_input(a)
    f1 = cs.new_input(b)
    va, vb = a % R_MOD, b % R_MOD
    for _ in range(num_steps):
        va, vb = vb, (va + vb) % R_MOD
    saved = cs.new_input(vb if num_steps else 0)
    fi = cs.new_witness(0)
    for _ in range(num_steps):
        fi = cs.lc_add(f1, f2)
        cs.enforce_equal(fi, cs.lc_add(f1, f2))       # cancels to the empty row after inlining
        f2, f1 = f1, fi
    cs.enforce_equal(fi, saved)
    return cs.finalize()


# ---------------------------------------------------------------------------
class PoseidonShape:
    """Width-3 Poseidon permutation with x^17 S-boxes, 8 full + 29 partial rounds."""
    FULL, PARTIAL, ALPHA, WIDTH = 8, 29, 17, 3

    def __init__(self, seed=0xB2005EED):
        rnd = random.Random(seed)
        rounds = self.FULL + self.PARTIAL
        self.ark = [[rnd.randrange(R_MOD) for _ in range(3)] for _ in range(rounds)]
        # Cauchy MDS
        xs = [rnd.randrange(R_MOD) for _ in range(3)]
        ys = [rnd.randrange(R_MOD) for _ in range(3)]
        self.mds = [[pow(xs[i] + ys[j], -1, R_MOD) for j in range(3)] for i in range(3)]

    def sbox(self, cs, x):
        """pow_by_constant(17): 4 squarings + 1 multiplication = 5 constraints."""
        x2 = cs.mul(x, x)
        x4 = cs.mul(x2, x2)
        x8 = cs.mul(x4, x4)
        x16 = cs.mul(x8, x8)
        return cs.mul(x16, x)

    def permute(self, cs, state):
        half = self.FULL // 2
        for r in range(self.FULL + self.PARTIAL):
            state = [cs.lc_add(s, cs.constant(k)) for s, k in zip(state, self.ark[r])]
            if r < half or r >= half + self.PARTIAL:
                state = [self.sbox(cs, s) for s in state]
            else:
                state[0] = self.sbox(cs, state[0])
            new = []
            for i in range(3):
                acc = {}
                for j in range(3):
                    acc = cs.lc_add(acc, cs.lc_scale(state[j], self.mds[i][j]))
                new.append(acc)
            state = new
        return state

    def hash(self, cs, elems):
        """Sponge, rate 2 / capacity 1: absorb pairs, permute when the rate fills,
        one more permutation to squeeze; ceil(n/2) permutations for n >= 1 inputs
        (matches the 3*ceil(n^2/2)*265 term of the reference's constraint count)."""
        state = [{}, {}, {}]
        pos = 0
        pending = False
        for e in elems:
            if pos == 2:
                state = self.permute(cs, state)
                pos = 0
            state[1 + pos] = cs.lc_add(state[1 + pos], e)   # rate elements sit after the capacity
            pos += 1
            pending = True
        if pending:
            state = self.permute(cs, state)
        return state[1]


def matrix_circuit(mat_a, mat_b, poseidon=None):
    """MatrixCircuit (constraints.rs:101-128): public hashes of A, B, C; witnesses A, B;
    C = A*B with TWO constraints per scalar product (the `*` and the redundant
    `mul_equals`, constraints.rs:91-93)."""
    n = len(mat_a)
    ps = poseidon or PoseidonShape()
    # native pass for the public digests
    scratch = ConstraintSystem()
    nat = lambda m: ps.hash(scratch, [scratch.new_witness(v) for row in m for v in row])
    mat_c = [[sum(mat_a[i][k] * mat_b[k][j] for k in range(n)) % R_MOD for j in range(n)] for i in range(n)]
    ha, hb, hc = (scratch.value(nat(m)) for m in (mat_a, mat_b, mat_c))

    cs = ConstraintSystem()
    pub_a = cs.new_input(ha)
    pub_b = cs.new_input(hb)
    va = [[cs.new_witness(v) for v in row] for row in mat_a]
    vb = [[cs.new_witness(v) for v in row] for row in mat_b]
    cs.enforce_equal(ps.hash(cs, [v for row in va for v in row]), pub_a)
    cs.enforce_equal(ps.hash(cs, [v for row in vb for v in row]), pub_b)
    # matrix_mul: n^2 placeholder witnesses for C, then per (i, j) a sum witness and 2n constraints
    for _ in range(n * n):
        cs.new_witness(0)
    vc = []
    for i in range(n):
        row = []
        for j in range(n):
            acc = cs.new_witness(0)
            for k in range(n):
                prod = cs.mul(va[i][k], vb[k][j])
                acc = cs.lc_add(acc, prod)
                cs.enforce(va[i][k], vb[k][j], prod)          # mul_equals
            row.append(acc)
        vc.append(row)
    hash_c = ps.hash(cs, [v for row in vc for v in row])
    pub_c = cs.new_input(hc)
    cs.enforce_equal(hash_c, pub_c)
    return cs.finalize()


def matrix_constraint_count(n):
    """3*ceil(n^2/2)*265 + 2 n^3 + 3."""
    return 3 * ((n * n + 1) // 2) * 265 + 2 * n ** 3 + 3


# ---------------------------------------------------------------------------
def prime_circuit(x, num_bits=20, k_bases=3, sha_blocks=7, seed=0xB2000005):
    """PrimeCircuit-shaped system: `sha_blocks` SHA-256-compression-sized Boolean
    blocks (the reference hashes x and K (r || j) strings: 1 + 2K blocks) followed
    by K bit-serial modular exponentiations a^(n-1) mod n over `num_bits` bits.
    Witnesses are ~90 % Booleans, as in the reference circuit."""
    rnd = random.Random(seed ^ x)
    cs = ConstraintSystem()
    pub_x = cs.new_input(x)
    bits_per_block = 64 * 32 * 6        # 64 rounds x 32-bit words x ~6 Boolean gates: size class of one compression
    digest_acc = {}
    prev = [cs.new_witness(rnd.randrange(2)) for _ in range(64)]
    for b in prev:
        cs.enforce(b, cs.lc_sub(b, cs.constant(1)), {})                 # booleanity
    for _ in range(sha_blocks):
        for g in range(bits_per_block // 3):
            u, v = prev[g % 64], prev[(g * 7 + 13) % 64]
            w = cs.mul(u, v)                                              # AND
            xor = cs.lc_sub(cs.lc_add(u, v), cs.lc_scale(w, 2))           # XOR as an LC
            nb = cs.new_witness(cs.value(xor))
            cs.enforce(nb, cs.lc_sub(nb, cs.constant(1)), {})             # booleanity
            cs.enforce_equal(nb, xor)
            prev[(g * 11 + 5) % 64] = nb
        digest_acc = cs.lc_add(digest_acc, prev[0])
    d1 = cs.new_input(cs.value(digest_acc))
    cs.enforce_equal(digest_acc, d1)
    d2 = cs.new_input((x * x) % R_MOD)
    cs.enforce(pub_x, pub_x, d2)
    # Fermat: K modular exponentiations with witness quotients / remainders
    n_val = (x | 1) % (1 << num_bits) or 3
    for _ in range(k_bases):
        base = rnd.randrange(2, max(3, n_val))
        acc_val = 1
        acc = cs.constant(1)
        e = n_val - 1
        for i in reversed(range(num_bits)):
            sq = cs.mul(acc, acc)
            q, r_ = divmod(cs.value(sq), n_val)
            qv, rv = cs.new_witness(q), cs.new_witness(r_)
            cs.enforce(qv, cs.constant(n_val), cs.lc_sub(sq, rv))        # sq = q*n + r
            acc = rv
            acc_val = r_
            bit = (e >> i) & 1
            bv = cs.new_witness(bit)
            cs.enforce(bv, cs.lc_sub(bv, cs.constant(1)), {})
            t = cs.mul(acc, cs.lc_add(cs.constant(1), cs.lc_scale(bv, base - 1)))   # acc * (bit ? base : 1)
            q, r_ = divmod(cs.value(t), n_val)
            qv, rv = cs.new_witness(q), cs.new_witness(r_)
            cs.enforce(qv, cs.constant(n_val), cs.lc_sub(t, rv))
            acc = rv
            acc_val = r_
    return cs.finalize()
