"""GPU parity tests: every C-ABI entry point against the oracle, bit-exact.

All of these call through libb200zk.so (ctypes -> extern "C" -> CUDA); nothing is
computed on the CPU except by the checker (oracle/).  Small cases use the Python
big-int oracle, larger ones the C++ restatement, full-size ones size-independent
properties (round trips, linearity, known-multiplier bases, pairing verification).
"""
import random

import numpy as np
import pytest

from oracle import bls12_381 as O
from oracle import groth16 as OG
from helpers import oracle_r1cs, pk_limbs, scalar_mix

pytestmark = pytest.mark.gpu
R = O.R_MOD


@pytest.fixture(scope="module")
def codec(b2z):
    return b2z.codec


def rand_fr_limbs(n, seed):
    """n uniform canonical Fr elements as (n, 4) uint64 (used as Montgomery or bigint limbs)."""
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64)
    a = (a[:, 0::2] | (a[:, 1::2] << np.uint64(32))).astype(np.uint64)
    a[:, 3] &= np.uint64((1 << 62) - 1)          # < 2^254 < r
    return np.ascontiguousarray(a)


# ------------------------------------------------------------------------------- NTT
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13])
def test_ntt_matches_python_oracle(b2z, ctx, codec, log_n):
    rnd = random.Random(log_n)
    n = 1 << log_n
    v = [rnd.choice([0, 1, R - 1]) if rnd.random() < 0.05 else rnd.randrange(R) for _ in range(n)]
    L = codec.fr_to_mont_limbs(v)
    d = OG.Radix2EvaluationDomain(n)
    dom = b2z.Radix2EvaluationDomain(ctx, n)
    assert codec.fr_from_mont_limbs(dom.fft(L)) == d.fft(list(v))
    assert codec.fr_from_mont_limbs(dom.ifft(L)) == d.ifft(list(v))
    dc, domc = d.get_coset(7), dom.get_coset(7)
    assert codec.fr_from_mont_limbs(domc.fft(L)) == dc.fft(list(v))
    assert codec.fr_from_mont_limbs(domc.ifft(L)) == dc.ifft(list(v))


def test_ntt_zero_padding_and_other_coset(b2z, ctx, codec):
    """fft of a short vector zero-pads (ark-poly); coset offset other than 7."""
    v = [5, 6, 7]
    dom = b2z.Radix2EvaluationDomain(ctx, 8)
    assert codec.fr_from_mont_limbs(dom.fft(codec.fr_to_mont_limbs(v))) == OG.Radix2EvaluationDomain(8).fft(list(v))
    g = 0x1234567890ABCDEF
    want = OG.Radix2EvaluationDomain(8).get_coset(g).fft(list(v))
    assert codec.fr_from_mont_limbs(dom.get_coset(g).fft(codec.fr_to_mont_limbs(v))) == want


@pytest.mark.parametrize("log_n", [14, 16, 17, 18, 20])
def test_ntt_matches_cpu_oracle(b2z, ctx, cpu_oracle, log_n):
    L = rand_fr_limbs(1 << log_n, log_n)
    g = b2z.codec.fr_to_mont_limbs([7])
    dom = b2z.Radix2EvaluationDomain(ctx, 1 << log_n)
    assert np.array_equal(dom.fft(L), cpu_oracle.ntt(L))
    assert np.array_equal(dom.ifft(L), cpu_oracle.ntt(L, True))
    assert np.array_equal(dom.get_coset(7).fft(L), cpu_oracle.ntt(L, False, g))
    assert np.array_equal(dom.get_coset(7).ifft(L), cpu_oracle.ntt(L, True, g))


@pytest.mark.parametrize("log_n", [22, 24])
def test_ntt_full_size_properties(b2z, ctx, log_n):
    """BASELINE sweep sizes: inverse(forward(v)) == v and linearity, bit-exact."""
    n = 1 << log_n
    L = rand_fr_limbs(n, 7)
    dom = b2z.Radix2EvaluationDomain(ctx, n)
    f = dom.fft(L)
    assert np.array_equal(dom.ifft(f), L)
    c = dom.get_coset(7)
    assert np.array_equal(c.ifft(c.fft(L)), L)
    # delta at index 1 transforms to the powers of w: check w^(n/2) = -1 and w^(n/4)^2 = -1 entries
    delta = np.zeros((n, 4), dtype=np.uint64)
    delta[1] = b2z.codec.fr_to_mont_limbs([1])[0]
    pw = dom.fft(delta)
    vals = b2z.codec.fr_from_mont_limbs(pw[[0, 1, n // 4, n // 2]])
    w = pow(O.FR_ROOT_OF_UNITY, 1 << (32 - log_n), R)
    assert vals == [1, w, pow(w, n // 4, R), R - 1]


# ------------------------------------------------------------------------------- witness map
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 8, 10, 11, 12])
def test_witness_map_matches_python_oracle(b2z, ctx, codec, log_n):
    rnd = random.Random(50 + log_n)
    n = 1 << log_n
    a, b, c = ([rnd.randrange(R) for _ in range(n)] for _ in range(3))
    got = b2z.LibsnarkReduction.witness_map_from_evaluations(ctx, *(codec.fr_to_mont_limbs(x) for x in (a, b, c)))
    assert codec.fr_from_mont_limbs(got) == OG.witness_map_from_evals(a, b, c)


def test_witness_map_golden(b2z, ctx, codec, golden):
    w = golden["witness_map"]
    a, b, c = ([int(x, 16) for x in w[k]] for k in "abc")
    got = b2z.LibsnarkReduction.witness_map_from_evaluations(ctx, *(codec.fr_to_mont_limbs(x) for x in (a, b, c)))
    assert ["%064x" % x for x in codec.fr_from_mont_limbs(got)] == w["h"]
    g = golden["ntt"]
    v = codec.fr_to_mont_limbs([int(x, 16) for x in g["input"]])
    dom = b2z.Radix2EvaluationDomain(ctx, 16)
    hx = lambda arr: ["%064x" % x for x in codec.fr_from_mont_limbs(arr)]
    assert hx(dom.fft(v)) == g["fft"] and hx(dom.ifft(v)) == g["ifft"]
    assert hx(dom.get_coset(7).fft(v)) == g["coset_fft"] and hx(dom.get_coset(7).ifft(v)) == g["coset_ifft"]


@pytest.mark.parametrize("log_n", [15, 17, 19])
def test_witness_map_matches_cpu_oracle(b2z, ctx, cpu_oracle, log_n):
    n = 1 << log_n
    a, b, c = rand_fr_limbs(n, 1), rand_fr_limbs(n, 2), rand_fr_limbs(n, 3)
    got = b2z.LibsnarkReduction.witness_map_from_evaluations(ctx, a, b, c)
    assert np.array_equal(got, cpu_oracle.witness_map(a, b, c))


def test_witness_map_from_matrices_quotient(b2z, ctx, codec, circuits):
    """h from the real row evaluations of a satisfied system is the exact quotient:
    a(x) b(x) - c(x) == h(x) (x^n - 1) at a random point (host big-int check)."""
    inst = circuits.matrix_circuit([[1, 2, 3], [4, 5, 6], [7, 8, 9]], [[1] * 3] * 3)
    h = codec.fr_from_mont_limbs(b2z.LibsnarkReduction.witness_map_from_matrices(
        ctx, inst.matrices, inst.num_instance, inst.num_constraints, inst.z))
    a, b, c = OG.constraint_evaluations(oracle_r1cs(inst), inst.z)
    n = len(a)
    d = OG.Radix2EvaluationDomain(n)
    x = 0xABCDEF12345
    ev = lambda cf: sum(v * pow(x, i, R) for i, v in enumerate(cf)) % R
    assert (ev(d.ifft(list(a))) * ev(d.ifft(list(b))) - ev(d.ifft(list(c)))) % R == ev(h) * (pow(x, n, R) - 1) % R
    assert h[n - 1] == 0


# ------------------------------------------------------------------------------- fixed base / MSM
def test_fixed_base_matches_oracle(b2z, ctx, codec):
    rnd = random.Random(4)
    ks = [0, 1, 2, R - 1, 1 << 254] + [rnd.randrange(R) for _ in range(24)]
    out, inf = b2z.FixedBase.msm_g1(ctx, codec.fr_to_bigint_limbs(ks))
    assert codec.g1_from_limbs(out, inf) == [O.G1.mul(O.G1_GEN, k % R) for k in ks]
    out2, inf2 = b2z.FixedBase.msm_g2(ctx, codec.fr_to_bigint_limbs(ks[:10]))
    assert codec.g2_from_limbs(out2, inf2) == [O.G2.mul(O.G2_GEN, k % R) for k in ks[:10]]


def _msm_known_multipliers(b2z, ctx, codec, group, n, kind, seed, with_identity=False):
    """Bases k_i G made on the GPU, so the expected MSM result is (sum k_i s_i) G."""
    rnd = random.Random(seed)
    curve = O.G1 if group == 1 else O.G2
    ks = [rnd.randrange(1, R) for _ in range(n)]
    if with_identity and n > 4:
        ks[1] = 0                    # identity base (a_query holds these for unused variables)
        ks[3] = ks[2]                # repeated base  -> P + P inside a bucket
        ks[4] = R - ks[2]            # negated base   -> P + (-P)
    fb = b2z.FixedBase.msm_g1 if group == 1 else b2z.FixedBase.msm_g2
    bases, inf = fb(ctx, codec.fr_to_bigint_limbs(ks))
    sc = scalar_mix(rnd, n, kind)
    if with_identity and n > 4:
        sc[3] = sc[2]
        sc[4] = sc[2]
    msm = b2z.VariableBaseMSM.msm_bigint_g1 if group == 1 else b2z.VariableBaseMSM.msm_bigint_g2
    out = msm(ctx, bases, codec.fr_to_bigint_limbs(sc), inf)
    proj = codec.g1_projective_from_limbs(out) if group == 1 else codec.g2_projective_from_limbs(out)
    want = curve.mul(curve.gen, sum(k * s for k, s in zip(ks, sc)) % R)
    return curve.to_affine(proj), want, (bases, inf, sc)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 100, 1000, 4097])
@pytest.mark.parametrize("kind", ["uniform", "witness"])
def test_msm_g1_small(b2z, ctx, codec, n, kind):
    got, want, _ = _msm_known_multipliers(b2z, ctx, codec, 1, n, kind, n, with_identity=True)
    assert got == want


@pytest.mark.parametrize("kind", ["zero", "equal", "max"])
def test_msm_g1_adversarial(b2z, ctx, codec, kind):
    got, want, _ = _msm_known_multipliers(b2z, ctx, codec, 1, 3000, kind, 17)
    assert got == want


def test_msm_empty_and_mismatch(b2z, ctx, codec):
    out = b2z.VariableBaseMSM.msm_bigint_g1(ctx, np.zeros((0, 12), np.uint64), np.zeros((0, 4), np.uint64))
    assert O.G1.to_affine(codec.g1_projective_from_limbs(out)) is None          # identity, Z = 0
    out = b2z.VariableBaseMSM.msm_bigint_g2(ctx, np.zeros((0, 24), np.uint64), np.zeros((0, 4), np.uint64))
    assert O.G2.to_affine(codec.g2_projective_from_limbs(out)) is None
    bases, inf = b2z.FixedBase.msm_g1(ctx, codec.fr_to_bigint_limbs([3, 5, 7]))
    # msm_bigint truncates to the shorter input; msm refuses a mismatch
    out = b2z.VariableBaseMSM.msm_bigint_g1(ctx, bases, codec.fr_to_bigint_limbs([2, 2]), inf)
    assert O.G1.to_affine(codec.g1_projective_from_limbs(out)) == O.G1.mul(O.G1_GEN, 16)
    with pytest.raises(ValueError):
        b2z.VariableBaseMSM.msm_g1(ctx, bases, codec.fr_to_bigint_limbs([2, 2]), inf)


def test_msm_golden(b2z, ctx, codec, golden):
    m = golden["msm"]
    ks = [int(x, 16) for x in m["base_multipliers"]]
    sc = codec.fr_to_bigint_limbs([int(x, 16) for x in m["scalars"]])
    b1, i1 = b2z.FixedBase.msm_g1(ctx, codec.fr_to_bigint_limbs(ks))
    got = O.G1.to_affine(codec.g1_projective_from_limbs(b2z.VariableBaseMSM.msm_bigint_g1(ctx, b1, sc, i1)))
    assert O.g1_compress(got).hex() == m["g1_result_compressed"]
    b2, i2 = b2z.FixedBase.msm_g2(ctx, codec.fr_to_bigint_limbs(ks[:4]))
    got2 = O.G2.to_affine(codec.g2_projective_from_limbs(b2z.VariableBaseMSM.msm_bigint_g2(ctx, b2, sc[:4], i2)))
    assert O.g2_compress(got2).hex() == m["g2_result_compressed"]


@pytest.mark.parametrize("n,kind", [(1 << 14, "uniform"), (1 << 16, "witness"), (70001, "uniform")])
def test_msm_g1_vs_cpu_oracle(b2z, ctx, codec, cpu_oracle, n, kind):
    got, want, (bases, inf, sc) = _msm_known_multipliers(b2z, ctx, codec, 1, n, kind, n)
    assert got == want
    out, is_inf = cpu_oracle.msm_g1(bases, codec.fr_to_bigint_limbs(sc), inf)
    assert not is_inf and codec.g1_from_limbs(out.reshape(1, -1))[0] == got


@pytest.mark.parametrize("n,kind", [(1, "uniform"), (5, "witness"), (64, "uniform"), (2000, "witness"),
                                    (1 << 13, "uniform")])
def test_msm_g2(b2z, ctx, codec, n, kind):
    got, want, _ = _msm_known_multipliers(b2z, ctx, codec, 2, n, kind, n, with_identity=True)
    assert got == want


def test_msm_g2_vs_cpu_oracle(b2z, ctx, codec, cpu_oracle):
    got, want, (bases, inf, sc) = _msm_known_multipliers(b2z, ctx, codec, 2, 1 << 12, "witness", 5)
    out, is_inf = cpu_oracle.msm_g2(bases, codec.fr_to_bigint_limbs(sc), inf)
    assert got == want and codec.g2_from_limbs(out.reshape(1, -1))[0] == got


def test_msm_full_size_linearity(b2z, ctx, codec):
    """BASELINE sweep size 2^20 on known-multiplier bases: exact result and
    msm(s) + msm(t) == msm(s + t)."""
    n = 1 << 20
    ks = rand_fr_limbs(n, 11)
    bases, inf = b2z.FixedBase.msm_g1(ctx, ks)
    s, t = rand_fr_limbs(n, 12), rand_fr_limbs(n, 13)
    to_int = codec.fr_from_bigint_limbs
    ki, si, ti = to_int(ks), to_int(s), to_int(t)
    run = lambda sc: O.G1.to_affine(codec.g1_projective_from_limbs(b2z.VariableBaseMSM.msm_bigint_g1(ctx, bases, sc, inf)))
    ps, pt = run(s), run(t)
    assert ps == O.G1.mul(O.G1_GEN, sum(k * x for k, x in zip(ki, si)) % R)
    st = codec.fr_to_bigint_limbs([(x + y) % R for x, y in zip(si, ti)])
    assert run(st) == O.G1.add(ps, pt)


# ------------------------------------------------------------------------------- Groth16 prove
def _gpu_pk(b2z, codec, opk):
    return b2z.ProvingKey(*pk_limbs(codec, opk))


def _prove_gpu(b2z, ctx, codec, pk, inst, r, s):
    a, b, c = b2z.LibsnarkReduction.constraint_evaluations(inst.matrices, inst.num_instance, inst.num_constraints,
                                                           inst.z)
    return b2z.Groth16.create_proof_with_reduction(ctx, pk, a, b, c, codec.fr_to_mont_limbs(inst.z), r, s), (a, b, c)


def test_prove_golden_fibonacci(b2z, ctx, codec, circuits, golden):
    case = golden["proof_fibonacci_0_1_10"]
    inst = circuits.fibonacci_circuit(0, 1, 10)
    opk = OG.setup(oracle_r1cs(inst), seed=case["setup_seed"])
    pk = _gpu_pk(b2z, codec, opk)
    got, _ = _prove_gpu(b2z, ctx, codec, pk, inst, int(case["r"], 16), int(case["s"], 16))
    assert got.hex() == case["proof"]
    assert OG.verify(opk, inst.z[1:inst.num_instance], O.proof_deserialize_compressed(got))
    # proving twice on the same uploaded key gives the same bytes
    again, _ = _prove_gpu(b2z, ctx, codec, pk, inst, int(case["r"], 16), int(case["s"], 16))
    assert again == got
    pk.free()


def test_prove_golden_matrix_2x2(b2z, ctx, codec, circuits, golden):
    case = golden["proof_matrix_2x2"]
    inst = circuits.matrix_circuit([[1, 2], [3, 4]], [[4, 3], [2, 1]])
    opk = OG.setup(oracle_r1cs(inst), seed=case["setup_seed"])
    pk = _gpu_pk(b2z, codec, opk)
    got, _ = _prove_gpu(b2z, ctx, codec, pk, inst, int(case["r"], 16), int(case["s"], 16))
    assert got.hex() == case["proof"]
    pk.free()


def test_prove_r_zero_and_s_zero(b2z, ctx, codec, circuits):
    inst = circuits.fibonacci_circuit(1, 1, 4)
    r1 = oracle_r1cs(inst)
    opk = OG.setup(r1)
    pk = _gpu_pk(b2z, codec, opk)
    for r, s in ((0, 12345), (777, 0), (0, 0)):
        got, _ = _prove_gpu(b2z, ctx, codec, pk, inst, r, s)
        assert got == OG.prove(opk, r1, inst.z, r, s)[1]
    pk.free()


def _vk_as_oracle(codec, vk):
    class V:
        pass
    v = V()
    v.alpha_g1 = codec.g1_from_limbs(vk.alpha_g1.reshape(1, -1))[0]
    v.beta_g2, v.gamma_g2, v.delta_g2 = (codec.g2_from_limbs(x.reshape(1, -1))[0]
                                         for x in (vk.beta_g2, vk.gamma_g2, vk.delta_g2))
    v.gamma_abc_g1 = codec.g1_from_limbs(*vk.gamma_abc_g1)
    return v


def _setup_prove_verify(b2z, ctx, codec, cpu_oracle, inst, seed):
    """GPU key generation -> GPU prove -> bytes == C++ oracle prove on the same key,
    and the proof passes the pairing check (the reference's acceptance test)."""
    rnd = random.Random(seed)
    toxic = [rnd.randrange(1, R) for _ in range(5)]
    pk, vk = b2z.Groth16.generate_parameters_with_qap(ctx, inst.matrices, inst.num_constraints, inst.num_instance,
                                                       inst.num_variables, *toxic)
    r, s = rnd.randrange(R), rnd.randrange(R)
    got, (a, b, c) = _prove_gpu(b2z, ctx, codec, pk, inst, r, s)
    cpk = cpu_oracle.CpuProvingKey(pk.num_variables, pk.num_instance, pk.domain_size, pk.a_query, pk.b_g1_query,
                                   pk.b_g2_query, pk.h_query, pk.l_query, pk.alpha_g1, pk.beta_g1, pk.delta_g1,
                                   pk.beta_g2, pk.delta_g2)
    rs = codec.fr_to_mont_limbs([r, s])
    want = cpk.prove(a, b, c, codec.fr_to_mont_limbs(inst.z), rs[0], rs[1])
    assert got == want
    assert OG.verify(_vk_as_oracle(codec, vk), inst.z[1:inst.num_instance], O.proof_deserialize_compressed(got))
    pk.free()


def test_prove_fibonacci_1000(b2z, ctx, codec, cpu_oracle, circuits):
    """BASELINE config 1 (n = 1000 steps, domain 2^10, 5 variables)."""
    _setup_prove_verify(b2z, ctx, codec, cpu_oracle, circuits.fibonacci_circuit(0, 1, 1000), 1)


def test_prove_matrix_4x4(b2z, ctx, codec, cpu_oracle, circuits):
    inst = circuits.matrix_circuit([[(3 * i + j) % 7 for j in range(4)] for i in range(4)], [[1] * 4] * 4)
    _setup_prove_verify(b2z, ctx, codec, cpu_oracle, inst, 2)


def test_prove_prime_shape(b2z, ctx, codec, cpu_oracle, circuits):
    """BASELINE config 3 shape at reduced size: Boolean-heavy witness (~90 % of scalars in {0, 1})."""
    _setup_prove_verify(b2z, ctx, codec, cpu_oracle, circuits.prime_circuit(5, num_bits=20, sha_blocks=1), 3)


def test_prove_matrix_16x16_full_size(b2z, ctx, codec, cpu_oracle, circuits):
    """BASELINE config 2: 16x16 matrices, 109 955 constraints, domain 2^17 -- full size."""
    inst = circuits.matrix_circuit([[1] * 16 for _ in range(16)], [[1] * 16 for _ in range(16)])   # bench/matrix.py:10-11
    assert inst.num_constraints == 109955 and inst.domain_size == 1 << 17
    _setup_prove_verify(b2z, ctx, codec, cpu_oracle, inst, 4)


def test_prove_matrix_20x20_reference_fixture(b2z, ctx, codec, cpu_oracle):
    """The reference's own large test case (matrix_proof_of_work/constraints.rs:280-285): 20x20 matrices with
    A[i][j] = i, B[i][j] = j -- 175 003 constraints, domain 2^18; rows evaluated on the GPU, proof bytes equal
    the C++ oracle's on the same key and pass the pairing check."""
    import importlib
    fast = importlib.import_module("zksnark-finalproject_b200.circuits_fast")
    n = 20
    cm, z_int = fast.matrix_circuit_fast([[i] * n for i in range(n)], [[j for j in range(n)] for _ in range(n)])
    assert cm.num_constraints == 3 * ((n * n + 1) // 2) * 265 + 2 * n ** 3 + 3 and cm.domain_size == 1 << 18
    rnd = random.Random(2020)
    pk, vk = b2z.Groth16.generate_parameters_with_qap(ctx, cm, cm.num_constraints, cm.num_instance_variables,
                                                       cm.num_variables, *[rnd.randrange(1, R) for _ in range(5)])
    z = codec.fr_to_mont_limbs(z_int)
    r, s = rnd.randrange(R), rnd.randrange(R)
    got = b2z.Groth16.create_proof_with_matrices(ctx, pk, cm, z, r, s)
    a, b, c = b2z.LibsnarkReduction.constraint_evaluations_device(ctx, cm, z)
    cpk = cpu_oracle.CpuProvingKey(pk.num_variables, pk.num_instance, pk.domain_size, pk.a_query, pk.b_g1_query,
                                   pk.b_g2_query, pk.h_query, pk.l_query, pk.alpha_g1, pk.beta_g1, pk.delta_g1,
                                   pk.beta_g2, pk.delta_g2)
    rs = codec.fr_to_mont_limbs([r, s])
    assert got == cpk.prove(a, b, c, z, rs[0], rs[1])
    assert OG.verify(_vk_as_oracle(codec, vk), z_int[1:cm.num_instance_variables], O.proof_deserialize_compressed(got))
    cm.free()
    pk.free()


# ------------------------------------------------------------------------------- constraint matrices on the device
def test_r1cs_rows_witness_map_and_prove_on_device(b2z, ctx, codec, circuits):
    """Row f1: the evaluate_constraint loop, witness_map_from_matrices and the whole prove with
    only z crossing PCIe -- each equal to the host-row path / the oracle."""
    import importlib
    fast = importlib.import_module("zksnark-finalproject_b200.circuits_fast")
    A = [[(3 * i + j + 1) for j in range(4)] for i in range(4)]
    B = [[(i * j + 2) for j in range(4)] for i in range(4)]
    cm, z_int = fast.matrix_circuit_fast(A, B)
    inst = circuits.matrix_circuit(A, B)
    assert z_int == inst.z
    z = codec.fr_to_mont_limbs(z_int)
    a_h, b_h, c_h = b2z.LibsnarkReduction.constraint_evaluations(inst.matrices, inst.num_instance,
                                                                 inst.num_constraints, inst.z)
    a_d, b_d, c_d = b2z.LibsnarkReduction.constraint_evaluations_device(ctx, cm, z)
    assert np.array_equal(a_d, a_h) and np.array_equal(b_d, b_h) and np.array_equal(c_d, c_h)
    h = b2z.LibsnarkReduction.witness_map_from_matrices(ctx, cm, cm.num_instance_variables, cm.num_constraints, z)
    want = OG.witness_map_from_evals(*(codec.fr_from_mont_limbs(x) for x in (a_h, b_h, c_h)))
    assert codec.fr_from_mont_limbs(h) == want
    rnd = random.Random(77)
    toxic = [rnd.randrange(1, R) for _ in range(5)]
    pk_rows, _ = b2z.Groth16.generate_parameters_with_qap(ctx, inst.matrices, inst.num_constraints, inst.num_instance,
                                                           inst.num_variables, *toxic)
    pk_dev, _ = b2z.Groth16.generate_parameters_with_qap(ctx, cm, cm.num_constraints, cm.num_instance_variables,
                                                          cm.num_variables, *toxic)
    for f in b2z.ProvingKey.FIELDS:        # key generation through the GPU SpMV gives the same key
        assert np.array_equal(getattr(pk_rows, f)[0], getattr(pk_dev, f)[0]), f
    r, s = rnd.randrange(R), rnd.randrange(R)
    p1 = b2z.Groth16.create_proof_with_reduction(ctx, pk_dev, a_h, b_h, c_h, z, r, s)
    p2 = b2z.Groth16.create_proof_with_matrices(ctx, pk_dev, cm, z, r, s)
    assert p1 == p2
    cm.free()
    pk_rows.free()
    pk_dev.free()


@pytest.mark.parametrize("world", [2, 3])
def test_prove_point_sharded_equals_whole_key(b2z, ctx, codec, circuits, world):
    """the shards of one key (here all on one GPU, one after the other) produce
    partial sums whose host combination is byte-identical to the single-GPU proof."""
    inst = circuits.matrix_circuit([[1, 2, 3], [4, 5, 6], [7, 8, 9]], [[2, 0, 1], [1, 1, 1], [3, 2, 1]])
    rnd = random.Random(world)
    toxic = [rnd.randrange(1, R) for _ in range(5)]
    pk, _ = b2z.Groth16.generate_parameters_with_qap(ctx, inst.matrices, inst.num_constraints, inst.num_instance,
                                                      inst.num_variables, *toxic)
    r, s = rnd.randrange(R), rnd.randrange(R)
    want, (a, b, c) = _prove_gpu(b2z, ctx, codec, pk, inst, r, s)
    z = codec.fr_to_mont_limbs(inst.z)
    parts = []
    for k in range(world):
        shard = b2z.ProvingKey(pk.num_variables, pk.num_instance, pk.domain_size, pk.a_query, pk.b_g1_query,
                               pk.b_g2_query, pk.h_query, pk.l_query, pk.alpha_g1, pk.beta_g1, pk.delta_g1,
                               pk.beta_g2, pk.delta_g2).upload(ctx, rank=k, world=world)
        parts.append(b2z.Groth16.create_proof_partial(ctx, shard, a, b, c, z, r, s))
        shard.free()
    assert b2z.Groth16.combine(parts) == want
    pk.free()


def test_prove_uneven_shards(b2z, ctx, codec, circuits):
    """b2z_pk_upload_slice: shards of different sizes tile the key and combine to the whole-key proof."""
    inst = circuits.matrix_circuit([[1, 2, 3], [4, 5, 6], [7, 8, 9]], [[2, 0, 1], [1, 1, 1], [3, 2, 1]])
    rnd = random.Random(12)
    pk, _ = b2z.Groth16.generate_parameters_with_qap(ctx, inst.matrices, inst.num_constraints, inst.num_instance,
                                                      inst.num_variables, *[rnd.randrange(1, R) for _ in range(5)])
    want, (a, b, c) = _prove_gpu(b2z, ctx, codec, pk, inst, 5, 6)
    z = codec.fr_to_mont_limbs(inst.z)
    weights = [85, 100, 7, 100]
    parts = []
    for k in range(len(weights)):
        shard = b2z.ProvingKey(pk.num_variables, pk.num_instance, pk.domain_size, pk.a_query, pk.b_g1_query,
                               pk.b_g2_query, pk.h_query, pk.l_query, pk.alpha_g1, pk.beta_g1, pk.delta_g1,
                               pk.beta_g2, pk.delta_g2).upload(ctx, rank=k, world=len(weights), weights=weights)
        parts.append(b2z.Groth16.create_proof_partial(ctx, shard, a, b, c, z, 5, 6))
        shard.free()
    assert b2z.Groth16.combine(parts) == want
    pk.free()


@pytest.mark.parametrize("world", [1, 3])
def test_prove_sharded_with_distributed_witness_map(b2z, ctx, codec, world):
    """b2z_groth16_shard_begin / b2z_r1cs_coset_evals / b2z_groth16_shard_finish: every shard gets the three
    coset evaluation vectors from "somewhere" (here: computed once on this GPU into torch buffers, as rank
    j would before broadcasting them) and the combined proof equals the whole-key proof."""
    import importlib
    import torch
    fast = importlib.import_module("zksnark-finalproject_b200.circuits_fast")
    A = [[(5 * i + j + 1) for j in range(3)] for i in range(3)]
    B = [[(i + 2 * j + 3) for j in range(3)] for i in range(3)]
    cm, z_int = fast.matrix_circuit_fast(A, B)
    rnd = random.Random(31 + world)
    pk, _ = b2z.Groth16.generate_parameters_with_qap(ctx, cm, cm.num_constraints, cm.num_instance_variables,
                                                      cm.num_variables, *[rnd.randrange(1, R) for _ in range(5)])
    z = codec.fr_to_mont_limbs(z_int)
    r, s = rnd.randrange(R), rnd.randrange(R)
    want = b2z.Groth16.create_proof_with_matrices(ctx, pk, cm, z, r, s)
    parts = []
    for k in range(world):
        shard = b2z.ProvingKey(pk.num_variables, pk.num_instance, pk.domain_size, pk.a_query, pk.b_g1_query,
                               pk.b_g2_query, pk.h_query, pk.l_query, pk.alpha_g1, pk.beta_g1, pk.delta_g1,
                               pk.beta_g2, pk.delta_g2).upload(ctx, rank=k, world=world)
        bufs = [torch.empty((pk.domain_size, 4), dtype=torch.int64, device="cuda") for _ in range(3)]
        if k % 2 == 0:       # either call may come first; the first one carries z
            keep = b2z.Groth16.shard_begin(ctx, shard, cm, z, r, s)
            for j in range(3):
                b2z.Groth16.coset_evals(ctx, cm, j, bufs[j].data_ptr())
        else:
            for j in range(3):
                b2z.Groth16.coset_evals(ctx, cm, j, bufs[j].data_ptr(), z if j == 0 else None)
            keep = b2z.Groth16.shard_begin(ctx, shard, cm, None, r, s)
        torch.cuda.synchronize()
        parts.append(b2z.Groth16.shard_finish(ctx, shard, *(t.data_ptr() for t in bufs)))
        del keep
        shard.free()
    assert b2z.Groth16.combine(parts) == want
    cm.free()
    pk.free()


def test_prove_degenerate_shapes(b2z, ctx, codec, cpu_oracle, circuits):
    """Edge shapes: a system with a single witness and nearly empty rows (Fibonacci, 0 steps: one
    constraint, domain 8), more shards than variables, and an all-zero witness vector."""
    inst = circuits.fibonacci_circuit(3, 4, 0)
    assert inst.num_constraints == 1 and inst.num_witness == 1
    _setup_prove_verify(b2z, ctx, codec, cpu_oracle, inst, 21)
    # shards outnumber the 5 variables: some ranks hold no query points at all
    rnd = random.Random(5)
    toxic = [rnd.randrange(1, R) for _ in range(5)]
    inst = circuits.fibonacci_circuit(0, 0, 3)          # all-zero assignment except the constant 1
    assert inst.is_satisfied() and sum(inst.z) == 1
    pk, _ = b2z.Groth16.generate_parameters_with_qap(ctx, inst.matrices, inst.num_constraints, inst.num_instance,
                                                      inst.num_variables, *toxic)
    want, (a, b, c) = _prove_gpu(b2z, ctx, codec, pk, inst, 11, 22)
    z = codec.fr_to_mont_limbs(inst.z)
    parts = []
    for k in range(7):
        shard = b2z.ProvingKey(pk.num_variables, pk.num_instance, pk.domain_size, pk.a_query, pk.b_g1_query,
                               pk.b_g2_query, pk.h_query, pk.l_query, pk.alpha_g1, pk.beta_g1, pk.delta_g1,
                               pk.beta_g2, pk.delta_g2).upload(ctx, rank=k, world=7)
        parts.append(b2z.Groth16.create_proof_partial(ctx, shard, a, b, c, z, 11, 22))
        shard.free()
    assert b2z.Groth16.combine(parts) == want
    pk.free()


def test_prove_from_page_locked_assignment(b2z, ctx, codec):
    """b2z_host_register: the same proof bytes whether z is pageable or page-locked caller memory."""
    import importlib
    fast = importlib.import_module("zksnark-finalproject_b200.circuits_fast")
    cm, z_int = fast.matrix_circuit_fast([[1, 2], [3, 4]], [[4, 3], [2, 1]])
    rnd = random.Random(99)
    pk, _ = b2z.Groth16.generate_parameters_with_qap(ctx, cm, cm.num_constraints, cm.num_instance_variables,
                                                      cm.num_variables, *[rnd.randrange(1, R) for _ in range(5)])
    z = codec.fr_to_mont_limbs(z_int)
    want = b2z.Groth16.create_proof_with_matrices(ctx, pk, cm, z, 21, 34)
    zp = ctx.pin(z.copy())
    got = b2z.Groth16.create_proof_with_matrices(ctx, pk, cm, zp, 21, 34)
    ctx.unpin(zp)
    assert got == want
    with pytest.raises(b2z._ffi.B2zError):
        ctx.unpin(zp)                       # not registered any more
    cm.free()
    pk.free()


def test_prove_from_concurrent_host_threads(b2z, ctx, codec, circuits):
    """The reference serves proofs from several actix worker threads (src/main.rs:37).  Three host
    threads prove at the same time -- two with their own context and key copy, one sharing the session
    context with the main thread (calls on one context serialise) -- and every proof equals the
    single-threaded bytes."""
    import threading
    inst = circuits.matrix_circuit([[1, 2, 3], [4, 5, 6], [7, 8, 9]], [[9, 8, 7], [6, 5, 4], [3, 2, 1]])
    opk = OG.setup(oracle_r1cs(inst), seed=77)
    limbs = pk_limbs(codec, opk)
    a, b, c = b2z.LibsnarkReduction.constraint_evaluations(inst.matrices, inst.num_instance, inst.num_constraints, inst.z)
    z = codec.fr_to_mont_limbs(inst.z)
    pairs = [(3 + 5 * i, 11 + 7 * i) for i in range(4)]
    shared_pk = b2z.ProvingKey(*limbs).upload(ctx)
    want = [b2z.Groth16.create_proof_with_reduction(ctx, shared_pk, a, b, c, z, r, s) for r, s in pairs]
    assert len(set(want)) == len(pairs)
    results, errors = {}, []

    def worker(name, own):
        try:
            wctx = b2z.Context(0) if own else ctx
            wpk = b2z.ProvingKey(*limbs).upload(wctx) if own else shared_pk
            out = []
            for _ in range(3):
                out.append([b2z.Groth16.create_proof_with_reduction(wctx, wpk, a, b, c, z, r, s) for r, s in pairs])
            results[name] = out
            if own:
                wpk.free()
                wctx.close()
        except Exception as e:          # surfaced below: a thread must not die silently
            errors.append((name, repr(e)))

    threads = [threading.Thread(target=worker, args=(n, own)) for n, own in (("w0", True), ("w1", True), ("shared", False))]
    for t in threads:
        t.start()
    main_out = [b2z.Groth16.create_proof_with_reduction(ctx, shared_pk, a, b, c, z, r, s) for r, s in pairs]
    for t in threads:
        t.join()
    assert not errors, errors
    assert main_out == want
    for name in ("w0", "w1", "shared"):
        assert all(rep == want for rep in results[name]), name
    shared_pk.free()


# ------------------------------------------------------------------------------- error behaviour
def test_error_codes(b2z, ctx):
    import ctypes
    L = b2z._ffi.lib()
    assert L.b2z_ntt_fr(ctx.handle, None, 4, 0, None) == b2z._ffi.B2Z_EINVAL
    buf = np.zeros((2, 4), dtype=np.uint64)
    assert L.b2z_ntt_fr(ctx.handle, buf.ctypes.data, 33, 0, None) == b2z._ffi.B2Z_ESIZE
    assert b"2^32" in L.b2z_last_error(ctx.handle)
    assert L.b2z_witness_map(ctx.handle, buf.ctypes.data, None, buf.ctypes.data, 1, buf.ctypes.data) == b2z._ffi.B2Z_EINVAL
    h = ctypes.c_void_p()
    assert L.b2z_ctx_create(99, ctypes.byref(h)) == b2z._ffi.B2Z_EINVAL
    # the context is still usable after an error
    dom = b2z.Radix2EvaluationDomain(ctx, 2)
    assert b2z.codec.fr_from_mont_limbs(dom.fft(b2z.codec.fr_to_mont_limbs([1, 2]))) == [3, R - 1]


def test_error_codes_of_the_shard_entry_points(b2z, ctx, codec):
    """Misuse of the sharded entry points is reported as a status, never a crash, and leaves the context usable."""
    import ctypes
    import importlib
    import torch
    fast = importlib.import_module("zksnark-finalproject_b200.circuits_fast")
    L, F = b2z._ffi.lib(), b2z._ffi
    cm, z_int = fast.matrix_circuit_fast([[1, 2], [3, 4]], [[4, 3], [2, 1]])
    rnd = random.Random(8)
    pk, _ = b2z.Groth16.generate_parameters_with_qap(ctx, cm, cm.num_constraints, cm.num_instance_variables,
                                                      cm.num_variables, *[rnd.randrange(1, R) for _ in range(5)])
    z = codec.fr_to_mont_limbs(z_int)
    want = b2z.Groth16.create_proof_with_matrices(ctx, pk, cm, z, 3, 4)
    fresh = b2z.ConstraintMatrices(cm.num_instance_variables, cm.num_witness_variables, cm.num_constraints,
                                   cm.a, cm.b, cm.c).upload(ctx)
    rs = codec.fr_to_mont_limbs([3, 4])
    buf = torch.empty((pk.domain_size, 4), dtype=torch.int64, device="cuda")
    ptr = ctypes.c_void_p(buf.data_ptr())
    # no assignment was ever uploaded for these matrices
    assert L.b2z_groth16_shard_begin(ctx.handle, pk._handle, fresh._handle, None, rs[0:1].ctypes.data,
                                     rs[1:2].ctypes.data) == F.B2Z_EINVAL
    assert b"no assignment" in L.b2z_last_error(ctx.handle)
    assert L.b2z_r1cs_coset_evals(ctx.handle, fresh._handle, 1, None, ptr) == F.B2Z_EINVAL
    assert L.b2z_r1cs_coset_evals(ctx.handle, fresh._handle, 3, z.ctypes.data, ptr) == F.B2Z_EINVAL
    assert L.b2z_groth16_shard_finish(ctx.handle, pk._handle, ptr, None, ptr, None) == F.B2Z_EINVAL
    # slices: from <= to <= den
    with pytest.raises(b2z._ffi.B2zError):
        b2z.ProvingKey(pk.num_variables, pk.num_instance, pk.domain_size, pk.a_query, pk.b_g1_query, pk.b_g2_query,
                       pk.h_query, pk.l_query, pk.alpha_g1, pk.beta_g1, pk.delta_g1, pk.beta_g2,
                       pk.delta_g2).upload(ctx, rank=2, world=2)
    h = ctypes.c_void_p()
    assert L.b2z_host_register(ctx.handle, None, 16) == F.B2Z_EINVAL
    # still proving correctly afterwards
    assert b2z.Groth16.create_proof_with_matrices(ctx, pk, cm, z, 3, 4) == want
    fresh.free()
    cm.free()
    pk.free()
