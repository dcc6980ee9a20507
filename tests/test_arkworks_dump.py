"""Parity against dumps of the REAL reference (tools/dump_vectors.rs run inside
ArielElb/zkSnark-FinalProject), whenever such files exist under tests/golden/.

Every `tests/golden/ark_dump_*.json` holds inputs and outputs of arkworks at the seams this library
replaces (format: tools/dump_vectors.rs).  The CPU oracle is checked in the `not gpu` suite, the CUDA
path in the `gpu` suite.  Files with "producer": "oracle" (tests/golden/make_dump_like.py) are written
by this repository's own oracle: they keep the loader and both checkers exercised, they pin nothing.
A file with "producer": "arkworks" turns "parity unpinned"  into pinned parity.
"""
import glob
import json
import os

import numpy as np
import pytest

from oracle import bls12_381 as O
from oracle import groth16 as OG

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMPS = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ark_dump_*.json")))
IDS = [os.path.basename(p)[len("ark_dump_"):-len(".json")] for p in DUMPS]
ints = lambda xs: [int(x) for x in xs]


def _load(path):
    with open(path) as f:
        d = json.load(f)
    assert d["producer"] in ("arkworks", "oracle")
    return d


def _oracle_pk(g):
    p = g["pk"]
    g1s = lambda xs: [O.g1_decompress(bytes.fromhex(x)) for x in xs]
    g2s = lambda xs: [O.g2_decompress(bytes.fromhex(x)) for x in xs]
    pk = OG.ProvingKey()
    pk.num_instance = g["num_instance"]
    pk.num_variables = g["num_instance"] + g["num_witness"]
    pk.domain_size = OG.Radix2EvaluationDomain(g["num_constraints"] + g["num_instance"]).size
    pk.alpha_g1, pk.beta_g1, pk.delta_g1 = g1s([p["alpha_g1"], p["beta_g1"], p["delta_g1"]])
    pk.beta_g2, pk.gamma_g2, pk.delta_g2 = g2s([p["beta_g2"], p["gamma_g2"], p["delta_g2"]])
    pk.gamma_abc_g1 = g1s(p["gamma_abc_g1"])
    pk.a_query, pk.b_g1_query, pk.h_query, pk.l_query = (g1s(p[k]) for k in ("a_query", "b_g1_query", "h_query", "l_query"))
    pk.b_g2_query = g2s(p["b_g2_query"])
    assert len(pk.a_query) == pk.num_variables and len(pk.l_query) == g["num_witness"]
    assert len(pk.h_query) == pk.domain_size - 1
    return pk


def _rows(g):
    return tuple([[(int(c), int(j)) for c, j in row] for row in g[k]] for k in "abc")


def test_a_dump_exists():
    assert DUMPS, "tests/golden/make_dump_like.py writes at least the oracle-made dump"


@pytest.mark.parametrize("path", DUMPS, ids=IDS)
def test_oracle_matches_dump(path):
    d = _load(path)
    t = d["ntt"]
    v = ints(t["input"])
    dom = OG.Radix2EvaluationDomain(1 << t["log_n"])
    cos = dom.get_coset(7)
    assert dom.fft(list(v)) == ints(t["fft"]) and dom.ifft(list(v)) == ints(t["ifft"])
    assert cos.fft(list(v)) == ints(t["coset_fft"]) and cos.ifft(list(v)) == ints(t["coset_ifft"])
    m = d["msm"]
    sc = ints(m["scalars"])
    g1 = [O.g1_decompress(bytes.fromhex(x)) for x in m["bases_g1"]]
    g2 = [O.g2_decompress(bytes.fromhex(x)) for x in m["bases_g2"]]
    assert O.g1_compress(O.G1.to_affine(OG.msm_bigint(O.G1, g1, sc))).hex() == m["g1_result"]
    assert O.g2_compress(O.G2.to_affine(OG.msm_bigint(O.G2, g2, sc))).hex() == m["g2_result"]
    g = d["groth16"]
    pk = _oracle_pk(g)
    ra, rb, rc = _rows(g)
    r1 = OG.R1CS(g["num_instance"], g["num_witness"], ra, rb, rc)
    z = ints(g["z"])
    assert r1.is_satisfied(z) if hasattr(r1, "is_satisfied") else True
    a, b, c = OG.constraint_evaluations(r1, z)
    assert OG.witness_map_from_evals(a, b, c) == ints(g["h"])
    proof, raw = OG.prove(pk, r1, z, int(g["r"]), int(g["s"]))
    assert raw.hex() == g["proof"]
    assert OG.verify(pk, z[1:g["num_instance"]], proof)


@pytest.mark.gpu
@pytest.mark.parametrize("path", DUMPS, ids=IDS)
def test_gpu_matches_dump(path, b2z, ctx):
    codec = b2z.codec
    d = _load(path)
    t = d["ntt"]
    v = codec.fr_to_mont_limbs(ints(t["input"]))
    dom = b2z.Radix2EvaluationDomain(ctx, 1 << t["log_n"])
    back = lambda arr: codec.fr_from_mont_limbs(arr)
    assert back(dom.fft(v)) == ints(t["fft"]) and back(dom.ifft(v)) == ints(t["ifft"])
    assert back(dom.get_coset(7).fft(v)) == ints(t["coset_fft"]) and back(dom.get_coset(7).ifft(v)) == ints(t["coset_ifft"])
    m = d["msm"]
    sc = codec.fr_to_bigint_limbs(ints(m["scalars"]))
    b1, i1 = codec.g1_to_limbs([O.g1_decompress(bytes.fromhex(x)) for x in m["bases_g1"]])
    b2, i2 = codec.g2_to_limbs([O.g2_decompress(bytes.fromhex(x)) for x in m["bases_g2"]])
    got1 = O.G1.to_affine(codec.g1_projective_from_limbs(b2z.VariableBaseMSM.msm_bigint_g1(ctx, b1, sc, i1)))
    got2 = O.G2.to_affine(codec.g2_projective_from_limbs(b2z.VariableBaseMSM.msm_bigint_g2(ctx, b2, sc, i2)))
    assert O.g1_compress(got1).hex() == m["g1_result"] and O.g2_compress(got2).hex() == m["g2_result"]
    g = d["groth16"]
    opk = _oracle_pk(g)
    from helpers import pk_limbs
    pk = b2z.ProvingKey(*pk_limbs(codec, opk))
    rows = _rows(g)
    z_int = ints(g["z"])
    a, b, c = b2z.LibsnarkReduction.constraint_evaluations(rows, g["num_instance"], g["num_constraints"], z_int)
    h = b2z.LibsnarkReduction.witness_map_from_evaluations(ctx, a, b, c)
    assert back(h) == ints(g["h"])
    proof = b2z.Groth16.create_proof_with_reduction(ctx, pk, a, b, c, codec.fr_to_mont_limbs(z_int), int(g["r"]), int(g["s"]))
    assert proof.hex() == g["proof"]
    pk.free()
