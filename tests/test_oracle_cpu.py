"""Pins the C++ CPU restatement (oracle/cpu/ark_cpu.cpp, also the timed CPU
baseline) against the Python big-int oracle."""
import random

import pytest

from oracle import bls12_381 as O
from oracle import groth16 as OG
from helpers import oracle_r1cs, pk_limbs, scalar_mix


@pytest.fixture(scope="module")
def codec(b2z):
    return b2z.codec


@pytest.mark.parametrize("log_n", [0, 1, 4, 9])
def test_ntt_and_witness_map(cpu_oracle, codec, log_n):
    rnd = random.Random(100 + log_n)
    n = 1 << log_n
    v = [rnd.randrange(O.R_MOD) for _ in range(n)]
    d = OG.Radix2EvaluationDomain(n)
    dc = d.get_coset(7)
    L, g = codec.fr_to_mont_limbs(v), codec.fr_to_mont_limbs([7])
    dec = codec.fr_from_mont_limbs
    assert dec(cpu_oracle.ntt(L)) == d.fft(list(v))
    assert dec(cpu_oracle.ntt(L, True)) == d.ifft(list(v))
    assert dec(cpu_oracle.ntt(L, False, g)) == dc.fft(list(v))
    assert dec(cpu_oracle.ntt(L, True, g)) == dc.ifft(list(v))
    a, b, c = ([rnd.randrange(O.R_MOD) for _ in range(n)] for _ in range(3))
    got = dec(cpu_oracle.witness_map(*(codec.fr_to_mont_limbs(x) for x in (a, b, c))))
    assert got == OG.witness_map_from_evals(a, b, c)


@pytest.mark.parametrize("n,kind", [(1, "uniform"), (7, "witness"), (40, "uniform"), (200, "witness"), (64, "max"),
                                    (64, "zero")])
def test_msm_g1(cpu_oracle, codec, n, kind):
    rnd = random.Random(n)
    pts = [O.G1.mul(O.G1_GEN, rnd.randrange(1, O.R_MOD)) for _ in range(n)]
    if n > 4:
        pts[1] = None
        pts[3] = pts[2]
    sc = scalar_mix(rnd, n, kind)
    L, inf = codec.g1_to_limbs(pts)
    out, is_inf = cpu_oracle.msm_g1(L, codec.fr_to_bigint_limbs(sc), inf)
    got = None if is_inf else codec.g1_from_limbs(out.reshape(1, -1))[0]
    assert got == O.G1.to_affine(OG.msm_bigint(O.G1, pts, sc))


@pytest.mark.parametrize("n", [1, 6, 35])
def test_msm_g2(cpu_oracle, codec, n):
    rnd = random.Random(n)
    pts = [O.G2.mul(O.G2_GEN, rnd.randrange(1, 1 << 64)) for _ in range(n)]
    sc = scalar_mix(rnd, n, "witness")
    L, inf = codec.g2_to_limbs(pts)
    out, is_inf = cpu_oracle.msm_g2(L, codec.fr_to_bigint_limbs(sc), inf)
    got = None if is_inf else codec.g2_from_limbs(out.reshape(1, -1))[0]
    assert got == O.G2.to_affine(OG.msm_bigint(O.G2, pts, sc))


def test_prove_bytes_match_python_oracle(cpu_oracle, b2z, codec, circuits, golden):
    inst = circuits.fibonacci_circuit(0, 1, 10)
    case = golden["proof_fibonacci_0_1_10"]
    opk = OG.setup(oracle_r1cs(inst), seed=case["setup_seed"])
    cpk = cpu_oracle.CpuProvingKey(*pk_limbs(codec, opk))
    a, b, c = b2z.LibsnarkReduction.constraint_evaluations(inst.matrices, inst.num_instance, inst.num_constraints,
                                                           inst.z)
    rs = codec.fr_to_mont_limbs([int(case["r"], 16), int(case["s"], 16)])
    got = cpk.prove(a, b, c, codec.fr_to_mont_limbs(inst.z), rs[0], rs[1])
    assert got.hex() == case["proof"]
    # threads do not change the bytes
    cpu_oracle.set_threads(1)
    assert cpk.prove(a, b, c, codec.fr_to_mont_limbs(inst.z), rs[0], rs[1]) == got
    cpu_oracle.set_threads(4)


def _key_equal(codec, key, opk):
    """CpuKey / product ProvingKey arrays == Python-oracle key, element by element."""
    g1 = lambda pair: codec.g1_from_limbs(pair[0], pair[1])
    g2 = lambda pair: codec.g2_from_limbs(pair[0], pair[1])
    assert g1(key.a_query) == opk.a_query
    assert g1(key.b_g1_query) == opk.b_g1_query
    assert g2(key.b_g2_query) == opk.b_g2_query
    assert g1(key.h_query) == opk.h_query
    assert g1(key.l_query) == opk.l_query
    one1 = lambda x: codec.g1_from_limbs(x.reshape(1, -1))[0]
    one2 = lambda x: codec.g2_from_limbs(x.reshape(1, -1))[0]
    assert (one1(key.alpha_g1), one1(key.beta_g1), one1(key.delta_g1)) == (opk.alpha_g1, opk.beta_g1, opk.delta_g1)
    assert (one2(key.beta_g2), one2(key.delta_g2)) == (opk.beta_g2, opk.delta_g2)


@pytest.mark.parametrize("which", ["fibonacci", "matrix2"])
def test_cpu_setup_matches_python_oracle(cpu_oracle, b2z, codec, circuits, which):
    """ark_cpu_groth16_setup (the reference arm's key generation) == OG.setup on the same toxic waste."""
    inst = circuits.fibonacci_circuit(0, 1, 10) if which == "fibonacci" else \
        circuits.matrix_circuit([[1, 2], [3, 4]], [[4, 3], [2, 1]])
    r1 = oracle_r1cs(inst)
    toxic = [random.Random(7).randrange(1, O.R_MOD) for _ in range(5)]
    opk = OG.setup(r1, toxic=toxic)
    cm = b2z.ConstraintMatrices.from_rows(inst.num_instance, inst.num_witness, inst.a, inst.b, inst.c)
    key = cpu_oracle.groth16_setup(cm.a, cm.b, cm.c, cm.num_constraints, cm.num_instance_variables, cm.num_variables,
                                   toxic)
    _key_equal(codec, key, opk)
    assert codec.g2_from_limbs(key.gamma_g2.reshape(1, -1))[0] == opk.gamma_g2
    assert codec.g1_from_limbs(*key.gamma_abc_g1) == opk.gamma_abc_g1
    # the row evaluation of the same library == the oracle's, and the proof from this key verifies
    z = codec.fr_to_mont_limbs(inst.z)
    a, b, c = cpu_oracle.constraint_evals(cm.a, cm.b, cm.c, cm.num_constraints, cm.num_instance_variables, z)
    ea, eb, ec = OG.constraint_evaluations(r1, inst.z)
    assert [codec.fr_from_mont_limbs(x) for x in (a, b, c)] == [ea, eb, ec]
    rs = codec.fr_to_mont_limbs([5, 9])
    proof = cpu_oracle.proving_key_of(key).prove(a, b, c, z, rs[0], rs[1])
    assert proof == OG.prove(opk, r1, inst.z, 5, 9)[1]
