"""CPU tests that pin the Python big-int oracle: standard BLS12-381 known
answers, algebraic identities, pairing verification (the reference's own
acceptance criterion: constraints/fibbonaci.rs:231, matrix_proof_of_work/
constraints.rs:266-271), and the committed golden vectors."""
import random

import pytest

from oracle import bls12_381 as O
from oracle import groth16 as OG
from helpers import oracle_r1cs

# Standard compressed generators (zcash / IETF BLS12-381 encoding), SURVEY.md A.5
G1_GEN_COMPRESSED = ("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
                     "6c55e83ff97a1aeffb3af00adb22c6bb")
G2_GEN_COMPRESSED = ("93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049"
                     "334cf11213945d57e5ac7d055d042b7e024aa2b2f08f0a91260805272dc51051"
                     "c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")


def test_field_constants():
    assert O.R_MOD.bit_length() == 255 and O.Q_MOD.bit_length() == 381
    assert (O.R_MOD - 1) % (1 << 32) == 0 and (O.R_MOD - 1) % (1 << 33) != 0          # two-adicity 32
    w = O.FR_ROOT_OF_UNITY
    assert pow(w, 1 << 32, O.R_MOD) == 1 and pow(w, 1 << 31, O.R_MOD) != 1
    assert w == 0x16A2A19EDFE81F20D09B681922C813B4B63683508C2280B93829971F439F0D2B
    assert O.FR_MONT_R == 0x1824B159ACC5056F998C4FEFECBC4FF55884B7FA0003480200000001FFFFFFFE
    assert (-pow(O.R_MOD, -1, 1 << 64)) % (1 << 64) == 0xFFFFFFFEFFFFFFFF
    assert (-pow(O.Q_MOD, -1, 1 << 64)) % (1 << 64) == 0x89F3FFFCFFFCFFFD


def test_generators_and_encoding_kat(golden):
    assert O.G1.is_on_curve(O.G1_GEN) and O.G2.is_on_curve(O.G2_GEN)
    assert O.G1.mul(O.G1_GEN, O.R_MOD) is None and O.G2.mul(O.G2_GEN, O.R_MOD) is None
    assert O.g1_compress(O.G1_GEN).hex() == G1_GEN_COMPRESSED == golden["kat"]["g1_generator_compressed"]
    assert O.g2_compress(O.G2_GEN).hex() == G2_GEN_COMPRESSED == golden["kat"]["g2_generator_compressed"]
    assert O.g1_compress(None) == bytes([0xC0]) + bytes(47)
    assert O.g2_compress(None) == bytes([0xC0]) + bytes(95)


def test_serialization_roundtrip():
    rnd = random.Random(3)
    for _ in range(6):
        k = rnd.randrange(1, O.R_MOD)
        p, q = O.G1.mul(O.G1_GEN, k), O.G2.mul(O.G2_GEN, k)
        for pt in (p, O.G1.neg(p)):
            assert O.g1_decompress(O.g1_compress(pt)) == pt
        for pt in (q, O.G2.neg(q)):
            assert O.g2_decompress(O.g2_compress(pt)) == pt
    # exactly one of P, -P carries the sort flag
    p = O.G1.mul(O.G1_GEN, 12345)
    assert (O.g1_compress(p)[0] ^ O.g1_compress(O.G1.neg(p))[0]) & 0x20


def test_pairing_bilinear():
    p5, q7 = O.G1.mul(O.G1_GEN, 5), O.G2.mul(O.G2_GEN, 7)
    assert O.pairing_product_is_one([(p5, q7), (O.G1.neg(O.G1.mul(O.G1_GEN, 35)), O.G2_GEN)])
    assert not O.pairing_product_is_one([(p5, q7), (O.G1.neg(O.G1.mul(O.G1_GEN, 34)), O.G2_GEN)])


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 8])
def test_ntt_inverse_and_definition(log_n):
    rnd = random.Random(log_n)
    n = 1 << log_n
    v = [rnd.randrange(O.R_MOD) for _ in range(n)]
    d = OG.Radix2EvaluationDomain(n)
    dc = d.get_coset(7)
    assert d.ifft(d.fft(list(v))) == v and dc.ifft(dc.fft(list(v))) == v
    if n <= 32:   # check against the definition  v_hat[k] = sum_j v[j] (g w^k)^j
        for dom in (d, dc):
            want = [sum(v[j] * pow(dom.offset * pow(dom.group_gen, k, O.R_MOD), j, O.R_MOD) for j in range(n)) % O.R_MOD
                    for k in range(n)]
            assert dom.fft(list(v)) == want


def test_domain_rules():
    assert OG.Radix2EvaluationDomain(1001 + 4).size == 1024
    assert OG.Radix2EvaluationDomain(1024).size == 1024 and OG.Radix2EvaluationDomain(1025).size == 2048
    with pytest.raises(ValueError):
        OG.Radix2EvaluationDomain((1 << 32) + 1)
    d = OG.Radix2EvaluationDomain(8).get_coset(7)
    assert d.evaluate_vanishing_polynomial(5) == (pow(5, 8, O.R_MOD) - pow(7, 8, O.R_MOD)) % O.R_MOD


def test_witness_map_quotient_identity(circuits):
    """a(x) b(x) - c(x) = h(x) (x^n - 1) at a random point for a satisfied system."""
    inst = circuits.matrix_circuit([[1, 2], [3, 4]], [[4, 3], [2, 1]])
    assert inst.is_satisfied()
    a, b, c = OG.constraint_evaluations(oracle_r1cs(inst), inst.z)
    h = OG.witness_map_from_evals(a, b, c)
    n = len(a)
    assert h[n - 1] == 0
    d = OG.Radix2EvaluationDomain(n)
    x = 0x1234567 
    ev = lambda coeffs: sum(cf * pow(x, i, O.R_MOD) for i, cf in enumerate(coeffs)) % O.R_MOD
    A, B, C = ev(d.ifft(list(a))), ev(d.ifft(list(b))), ev(d.ifft(list(c)))
    assert (A * B - C) % O.R_MOD == ev(h) * (pow(x, n, O.R_MOD) - 1) % O.R_MOD


def test_msm_window_rule_and_digits():
    assert [OG.ark_window_size(n) for n in (1, 31, 32, 1 << 16, 1 << 20, 1 << 22, 1 << 24, 1 << 26)] == \
        [3, 3, 5, 13, 15, 17, 18, 19]
    rnd = random.Random(9)
    for w in (3, 5, 13, 17):
        for k in [0, 1, O.R_MOD - 1] + [rnd.randrange(O.R_MOD) for _ in range(50)]:
            d = OG.make_digits(k, w)
            assert sum(x << (w * i) for i, x in enumerate(d)) == k
            assert all(-(1 << (w - 1)) <= x < (1 << (w - 1)) for x in d[:-1])


@pytest.mark.parametrize("n", [0, 1, 5, 33])
def test_msm_matches_naive(n):
    rnd = random.Random(n)
    pts = [O.G1.mul(O.G1_GEN, rnd.randrange(1, 1 << 40)) for _ in range(n)]
    if n > 3:
        pts[2] = None                       # identity entries occur in a_query / b_query
        pts[3] = pts[1]                     # repeated base
    sc = [rnd.choice([0, 1, O.R_MOD - 1, rnd.randrange(O.R_MOD)]) for _ in range(n)]
    assert O.G1.to_affine(OG.msm_bigint(O.G1, pts, sc)) == O.G1.to_affine(OG.msm_naive(O.G1, pts, sc))
    if 0 < n <= 5:
        pts2 = [O.G2.mul(O.G2_GEN, rnd.randrange(1, 1 << 30)) for _ in range(n)]
        assert O.G2.to_affine(OG.msm_bigint(O.G2, pts2, sc)) == O.G2.to_affine(OG.msm_naive(O.G2, pts2, sc))


def test_golden_ntt_witness_map_msm(golden):
    g = golden["ntt"]
    v = [int(x, 16) for x in g["input"]]
    d = OG.Radix2EvaluationDomain(len(v))
    dc = d.get_coset(7)
    hx = lambda xs: ["%064x" % x for x in xs]
    assert hx(d.fft(list(v))) == g["fft"] and hx(d.ifft(list(v))) == g["ifft"]
    assert hx(dc.fft(list(v))) == g["coset_fft"] and hx(dc.ifft(list(v))) == g["coset_ifft"]
    w = golden["witness_map"]
    a, b, c = ([int(x, 16) for x in w[k]] for k in "abc")
    assert hx(OG.witness_map_from_evals(a, b, c)) == w["h"]
    m = golden["msm"]
    ks = [int(x, 16) for x in m["base_multipliers"]]
    sc = [int(x, 16) for x in m["scalars"]]
    # independent of the bucket method: the result is (sum k_i s_i) G
    want = O.G1.mul(O.G1_GEN, sum(k * s for k, s in zip(ks, sc)) % O.R_MOD)
    assert O.g1_compress(want).hex() == m["g1_result_compressed"]
    want2 = O.G2.mul(O.G2_GEN, sum(k * s for k, s in zip(ks[:4], sc[:4])) % O.R_MOD)
    assert O.g2_compress(want2).hex() == m["g2_result_compressed"]


def _golden_proof_case(inst, case):
    r1 = oracle_r1cs(inst)
    pk = OG.setup(r1, seed=case["setup_seed"])
    proof, raw = OG.prove(pk, r1, inst.z, int(case["r"], 16), int(case["s"], 16))
    assert raw.hex() == case["proof"]
    assert O.proof_deserialize_compressed(raw) == proof
    assert OG.verify(pk, inst.z[1:inst.num_instance], proof)
    bad = list(inst.z[1:inst.num_instance])
    bad[-1] = (bad[-1] + 1) % O.R_MOD
    assert not OG.verify(pk, bad, proof)


def test_golden_proof_fibonacci(golden, circuits):
    inst = circuits.fibonacci_circuit(0, 1, 10)
    assert inst.z[3] == 89                              # fibbonaci_handler.rs:34-65 checks fibonacci(10) == 89
    _golden_proof_case(inst, golden["proof_fibonacci_0_1_10"])


@pytest.mark.slow
def test_golden_proof_matrix_2x2(golden, circuits):
    _golden_proof_case(circuits.matrix_circuit([[1, 2], [3, 4]], [[4, 3], [2, 1]]), golden["proof_matrix_2x2"])


def test_prove_r_zero_branch(circuits):
    """ark-groth16 skips B1 when r == 0; the proof must still verify."""
    inst = circuits.fibonacci_circuit(1, 1, 4)
    r1 = oracle_r1cs(inst)
    pk = OG.setup(r1)
    proof, _ = OG.prove(pk, r1, inst.z, 0, 12345)
    assert OG.verify(pk, inst.z[1:inst.num_instance], proof)
