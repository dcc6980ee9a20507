"""arkworks / rand RNG draws of the prove path -- TEST INFRASTRUCTURE ONLY (row a8 of SURVEY.md 8(a)).

PARITY STATUS: "parity unpinned" -- restated from the published crates (rand 0.8.5 / rand_core 0.6 /
rand_chacha 0.3, ark-ff 0.4, ark-ec 0.4.2, ark-groth16 0.4: /root/reference/Cargo.toml:10-22), none of
which is on disk.  The ChaCha block function is pinned by the two ChaCha20 known answers of RFC 7539
(tests/test_oracle_rng.py); the 12-round variant, the PCG32 seed expansion and the arkworks sampling
rules follow SURVEY.md A.6.  Purpose: reproduce, from the seed alone, every random value of the one
fixed-seed route of the reference (src/arkworks/backend/fibbonaci_handler.rs:99-110:
`StdRng::seed_from_u64(42)`, then Groth16::setup and Groth16::prove on the same stream), so that a
real arkworks dump can be matched end to end the day one is available.

  StdRng               rand::rngs::StdRng = ChaCha12Rng; seed_from_u64 = PCG32 expansion (rand_core)
  fr_rand / fq_rand    ark_ff Fp::rand: N x next_u64 -> limbs (LE), shave the top bits, reject >= p;
                       the accepted limbs ARE the Montgomery representation (value = limbs * R^-1)
  g1_rand / g2_rand    ark_ec short_weierstrass Projective::rand: x <- BaseField::rand, greatest <- bool,
                       y from x (smaller / larger root), then multiplication by the cofactor
  setup_draws          ark_groth16 generate_random_parameters_with_reduction: alpha, beta, gamma, delta,
                       g1, g2, then tau = domain.sample_element_outside_domain
  prove_draws          create_random_proof_with_reduction: r then s
"""
from oracle import bls12_381 as O

MASK32 = 0xFFFFFFFF
MASK64 = (1 << 64) - 1


def _rotl(x, n):
    return ((x << n) | (x >> (32 - n))) & MASK32


def chacha_block(state, rounds):
    """state: 16 u32 words; returns the 16 output words (RFC 7539 2.3 with `rounds` rounds)."""
    x = list(state)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & MASK32; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & MASK32; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & MASK32; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & MASK32; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & MASK32 for a, b in zip(x, state)]


CHACHA_CONST = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574]      # "expand 32-byte k"


class StdRng:
    """rand 0.8 StdRng (ChaCha12Rng, 64-bit block counter in words 12-13, stream id 0 in words 14-15)."""
    ROUNDS = 12

    def __init__(self, seed32):
        assert len(seed32) == 32
        self.key = [int.from_bytes(seed32[4 * i:4 * i + 4], "little") for i in range(8)]
        self.counter = 0
        self.buf = []

    @classmethod
    def from_seed(cls, seed32):
        return cls(bytes(seed32))

    @classmethod
    def seed_from_u64(cls, state):
        """rand_core::SeedableRng::seed_from_u64: PCG32 (XSH RR) output words fill the 32-byte seed."""
        MUL, INC = 6364136223846793005, 11634580027462260723
        seed = b""
        for _ in range(8):
            state = (state * MUL + INC) & MASK64
            xorshifted = (((state >> 18) ^ state) >> 27) & MASK32
            rot = state >> 59
            x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & MASK32
            seed += x.to_bytes(4, "little")
        return cls(seed)

    def _refill(self):
        st = CHACHA_CONST + self.key + [self.counter & MASK32, (self.counter >> 32) & MASK32, 0, 0]
        self.buf = chacha_block(st, self.ROUNDS)
        self.counter += 1

    def next_u32(self):
        if not self.buf:
            self._refill()
        return self.buf.pop(0)

    def next_u64(self):
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)

    def gen_bool(self):
        """rand Standard for bool: the sign bit of next_u32."""
        return (self.next_u32() >> 31) == 1


def _fp_rand(rng, nlimbs, modulus, bits):
    shave = 64 * nlimbs - bits
    while True:
        limbs = [rng.next_u64() for _ in range(nlimbs)]
        limbs[-1] &= MASK64 >> shave
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < modulus:
            return v                                    # Montgomery representation


def fr_rand(rng):
    """Fr::rand -> canonical integer value (the drawn limbs are the Montgomery form)."""
    return O.fr_from_mont(_fp_rand(rng, 4, O.R_MOD, 255))


def fq_rand(rng):
    return O.fq_from_mont(_fp_rand(rng, 6, O.Q_MOD, 381))


def fq2_rand(rng):
    c0 = fq_rand(rng)
    return (c0, fq_rand(rng))


G1_COFACTOR = 0x396C8C005555E1568C00AAAB0000AAAB
G2_COFACTOR = int("5d543a95414e7f1091d50792876a202cd91de4547085abaa68a205b2e5a7ddfa628f1cb4d9e82ef2"
                  "1537e293a6691ae1616ec6e786f0c70cf1c38e31c7238e5", 16)


def _fq2_lt(a, b):
    """ark_ff QuadExtField Ord: c1 first, then c0."""
    return (a[1], a[0]) < (b[1], b[0])


def g1_rand(rng):
    while True:
        x = fq_rand(rng)
        greatest = rng.gen_bool()
        y = O.fq_sqrt((x * x * x + 4) % O.Q_MOD)
        if y is None:
            continue
        ny = (-y) % O.Q_MOD
        small, large = (y, ny) if y < ny else (ny, y)
        return O.G1.mul((x, large if greatest else small), G1_COFACTOR)


def g2_rand(rng):
    F = O.Fq2Ops
    while True:
        x = fq2_rand(rng)
        greatest = rng.gen_bool()
        y = O.fq2_sqrt(F.add(F.mul(F.sqr(x), x), (4, 4)))
        if y is None:
            continue
        ny = F.neg(y)
        small, large = (y, ny) if _fq2_lt(y, ny) else (ny, y)
        return O.G2.mul((x, large if greatest else small), G2_COFACTOR)


def setup_draws(rng, domain_size):
    """ark-groth16 0.4 generator draw order: alpha, beta, gamma, delta, g1, g2, tau."""
    alpha, beta, gamma, delta = (fr_rand(rng) for _ in range(4))
    g1 = g1_rand(rng)
    g2 = g2_rand(rng)
    while True:
        tau = fr_rand(rng)
        if (pow(tau, domain_size, O.R_MOD) - 1) % O.R_MOD:
            break
    return {"alpha": alpha, "beta": beta, "gamma": gamma, "delta": delta, "g1": g1, "g2": g2, "tau": tau}


def prove_draws(rng):
    r = fr_rand(rng)
    return r, fr_rand(rng)
