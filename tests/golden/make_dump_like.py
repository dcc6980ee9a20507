#!/usr/bin/env python3
"""Writes tests/golden/ark_dump_oracle_*.json: the SAME format tools/dump_vectors.rs produces from real
arkworks, but computed by this repository's Python oracle ("producer": "oracle").  These files keep
tests/test_arkworks_dump.py (loader + checks) exercised until a maintainer of the reference drops a
real dump ("producer": "arkworks") next to them -- which then pins parity against the reference itself.

Run:  python tests/golden/make_dump_like.py
"""
import importlib
import json
import os
import random
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from oracle import bls12_381 as O            # noqa: E402
from oracle import groth16 as OG             # noqa: E402
circuits = importlib.import_module("zksnark-finalproject_b200.circuits")

dec = lambda v: str(v % O.R_MOD)
g1h = lambda p: O.g1_compress(p).hex()
g2h = lambda p: O.g2_compress(p).hex()


def ntt_section(rnd):
    n = 16
    v = [rnd.randrange(O.R_MOD) for _ in range(n)]
    d = OG.Radix2EvaluationDomain(n)
    c = d.get_coset(7)
    return {"log_n": 4, "input": [dec(x) for x in v], "fft": [dec(x) for x in d.fft(list(v))],
            "ifft": [dec(x) for x in d.ifft(list(v))], "coset_fft": [dec(x) for x in c.fft(list(v))],
            "coset_ifft": [dec(x) for x in c.ifft(list(v))]}


def msm_section(rnd):
    n = 40
    sc = [rnd.randrange(O.R_MOD) for _ in range(n)]
    sc[0:4] = [0, 1, O.R_MOD - 1, 65535]
    g1 = [O.G1.mul(O.G1_GEN, rnd.randrange(1, O.R_MOD)) for _ in range(n)]
    g2 = [O.G2.mul(O.G2_GEN, rnd.randrange(1, O.R_MOD)) for _ in range(n)]
    return {"scalars": [dec(s) for s in sc], "bases_g1": [g1h(p) for p in g1],
            "g1_result": g1h(O.G1.to_affine(OG.msm_bigint(O.G1, g1, sc))),
            "bases_g2": [g2h(p) for p in g2], "g2_result": g2h(O.G2.to_affine(OG.msm_bigint(O.G2, g2, sc)))}


def groth16_section(inst, rnd):
    r1 = OG.R1CS(inst.num_instance, inst.num_witness, inst.a, inst.b, inst.c)
    pk = OG.setup(r1, toxic=[rnd.randrange(1, O.R_MOD) for _ in range(5)])
    r, s = rnd.randrange(O.R_MOD), rnd.randrange(O.R_MOD)
    _, raw = OG.prove(pk, r1, inst.z, r, s)
    a, b, c = OG.constraint_evaluations(r1, inst.z)
    h = OG.witness_map_from_evals(a, b, c)
    rows = lambda m: [[[dec(v), j] for v, j in row] for row in m]
    return {"num_instance": inst.num_instance, "num_witness": inst.num_witness, "num_constraints": inst.num_constraints,
            "a": rows(inst.matrices[0]), "b": rows(inst.matrices[1]), "c": rows(inst.matrices[2]),
            "z": [dec(x) for x in inst.z], "r": dec(r), "s": dec(s), "h": [dec(x) for x in h],
            "pk": {"alpha_g1": g1h(pk.alpha_g1), "beta_g1": g1h(pk.beta_g1), "delta_g1": g1h(pk.delta_g1),
                   "beta_g2": g2h(pk.beta_g2), "gamma_g2": g2h(pk.gamma_g2), "delta_g2": g2h(pk.delta_g2),
                   "gamma_abc_g1": [g1h(p) for p in pk.gamma_abc_g1], "a_query": [g1h(p) for p in pk.a_query],
                   "b_g1_query": [g1h(p) for p in pk.b_g1_query], "b_g2_query": [g2h(p) for p in pk.b_g2_query],
                   "h_query": [g1h(p) for p in pk.h_query], "l_query": [g1h(p) for p in pk.l_query]},
            "proof": raw.hex()}


def write(case, inst, seed):
    rnd = random.Random(seed)
    out = {"producer": "oracle", "case": case, "ntt": ntt_section(rnd), "msm": msm_section(rnd),
           "groth16": groth16_section(inst, rnd)}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ark_dump_oracle_%s.json" % case)
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path)


if __name__ == "__main__":
    write("fibonacci_0_1_10", circuits.fibonacci_circuit(0, 1, 10), 0xB2000004)
