"""Row f4: the host-side verifier and the PreparedVerifyingKey wire format behind the C ABI
(b2z_groth16_prepare_verifying_key / b2z_groth16_verify_with_processed_vk; reference: io.rs:62-77,
matrix_proof.rs:134-136,199-206), byte-checked against the independent restatement in oracle/pvk.py.
No GPU is involved (verification is host work in the reference too)."""
import random

import numpy as np
import pytest

from oracle import bls12_381 as O
from oracle import groth16 as OG
from oracle import pvk as OP
from helpers import oracle_r1cs


@pytest.fixture(scope="module")
def case(b2z, circuits):
    inst = circuits.matrix_circuit([[1, 2], [3, 4]], [[4, 3], [2, 1]])
    r1 = oracle_r1cs(inst)
    opk = OG.setup(r1, seed=31)
    codec = b2z.codec
    one1 = lambda p: codec.g1_to_limbs([p])[0].reshape(-1)
    one2 = lambda p: codec.g2_to_limbs([p])[0].reshape(-1)
    vk = b2z.VerifyingKey(one1(opk.alpha_g1), one2(opk.beta_g2), one2(opk.gamma_g2), one2(opk.delta_g2),
                          codec.g1_to_limbs(opk.gamma_abc_g1))
    (A, B, C), proof = OG.prove(opk, r1, inst.z, 12345, 67890)
    return inst, opk, vk, proof


def test_pvk_bytes_equal_the_oracle(b2z, case):
    inst, opk, vk, _ = case
    got = vk.prepare()
    want = OP.prepare_verifying_key_bytes(opk.alpha_g1, opk.beta_g2, opk.gamma_g2, opk.delta_g2, opk.gamma_abc_g1)
    l = len(opk.gamma_abc_g1)
    assert len(got) == 48 + 3 * 96 + 8 + 48 * l + 576 + 2 * (8 + 68 * 288 + 1)
    head = 48 + 3 * 96 + 8 + 48 * l
    assert got[:head] == want[:head]                                  # vk part (zcash compressed points)
    assert got[head:head + 576] == want[head:head + 576]              # alpha_g1_beta_g2 from two different pairings
    assert got == want                                                # line coefficients and flags


def test_verify_with_processed_vk(b2z, case):
    inst, opk, vk, proof = case
    pvk = vk.prepare()
    pub = inst.z[1:inst.num_instance]
    V = b2z.Groth16.verify_with_processed_vk
    assert V(pvk, pub, proof) is True
    assert OP.verify_with_processed_vk(pvk, pub, proof) is True        # the oracle reads the product's bytes
    bad_pub = list(pub)
    bad_pub[0] = (bad_pub[0] + 1) % O.R_MOD
    assert V(pvk, bad_pub, proof) is False
    assert OP.verify_with_processed_vk(pvk, bad_pub, proof) is False
    # another valid proof of the same statement (different r, s) verifies; swapping A and C does not
    r1 = oracle_r1cs(inst)
    _, proof2 = OG.prove(opk, r1, inst.z, 1, 2)
    assert proof2 != proof and V(pvk, pub, proof2) is True
    assert V(pvk, pub, proof[144:192] + proof[48:144] + proof[0:48]) is False
    # wrong number of inputs / corrupted encodings -> SynthesisError (MalformedVerifyingKey), not a crash
    with pytest.raises(b2z.SynthesisError):
        V(pvk, pub[:-1], proof)
    with pytest.raises(b2z.SynthesisError):
        V(pvk[:-1], pub, proof)
    broken = bytearray(proof)
    broken[0] &= 0x7f                                                  # compression flag cleared
    with pytest.raises(b2z.SynthesisError):
        V(pvk, pub, bytes(broken))
    not_on_curve = bytearray(proof)
    not_on_curve[47] ^= 1
    try:
        assert V(pvk, pub, bytes(not_on_curve)) is False               # another x: off-curve, out of subgroup or wrong
    except b2z.SynthesisError:
        pass


def test_verifier_accepts_the_golden_fibonacci_proof(b2z, circuits, golden):
    """The committed golden proof (tests/golden/vectors.json) through prepare + verify."""
    g = golden["proof_fibonacci_0_1_10"]
    inst = circuits.fibonacci_circuit(0, 1, 10)
    opk = OG.setup(oracle_r1cs(inst), seed=g["setup_seed"])
    codec = b2z.codec
    one1 = lambda p: codec.g1_to_limbs([p])[0].reshape(-1)
    one2 = lambda p: codec.g2_to_limbs([p])[0].reshape(-1)
    vk = b2z.VerifyingKey(one1(opk.alpha_g1), one2(opk.beta_g2), one2(opk.gamma_g2), one2(opk.delta_g2),
                          codec.g1_to_limbs(opk.gamma_abc_g1))
    assert b2z.Groth16.verify_with_processed_vk(vk.prepare(), inst.z[1:inst.num_instance], bytes.fromhex(g["proof"]))


def test_identity_and_edge_points_in_the_wire_format(b2z):
    """gamma_abc_g1 entries that are the identity serialize as c0 00.. and survive the round trip."""
    codec = b2z.codec
    one1 = lambda p: codec.g1_to_limbs([p])[0].reshape(-1)
    one2 = lambda p: codec.g2_to_limbs([p])[0].reshape(-1)
    rnd = random.Random(5)
    pts = [O.G1.mul(O.G1_GEN, rnd.randrange(1, O.R_MOD)), None, O.G1.mul(O.G1_GEN, 7)]
    g2s = [O.G2.mul(O.G2_GEN, rnd.randrange(1, O.R_MOD)) for _ in range(3)]
    vk = b2z.VerifyingKey(one1(pts[0]), one2(g2s[0]), one2(g2s[1]), one2(g2s[2]), codec.g1_to_limbs(pts))
    got = vk.prepare()
    assert got == OP.prepare_verifying_key_bytes(pts[0], g2s[0], g2s[1], g2s[2], pts)
    alpha, beta, gamma, delta, abc, _, _ = OP.parse_pvk(got)
    assert abc == pts and (alpha, beta, gamma, delta) == (pts[0], g2s[0], g2s[1], g2s[2])
