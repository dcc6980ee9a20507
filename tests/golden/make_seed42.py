#!/usr/bin/env python3
"""Writes tests/golden/fibonacci_seed42.json: every random draw and the proof of the reference's fixed-seed route
(src/arkworks/backend/fibbonaci_handler.rs:99-110, StdRng::seed_from_u64(42), a = 0, b = 1, num_of_rounds = 10),
derived from the seed alone by oracle/ark_rng.py + oracle/groth16.py.  "producer": "oracle" -- it pins the CUDA path
and the C++ oracle to the Python restatement; a real arkworks run of the same route would replace it byte for byte."""
import importlib
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ark_rng as A, bls12_381 as O, groth16 as OG      # noqa: E402
from helpers import oracle_r1cs                                      # noqa: E402

circuits = importlib.import_module("zksnark-finalproject_b200.circuits")
inst = circuits.fibonacci_circuit(0, 1, 10)
r1 = oracle_r1cs(inst)
rng = A.StdRng.seed_from_u64(42)
n = 1
while n < r1.num_constraints + r1.num_instance:
    n <<= 1
d = A.setup_draws(rng, n)
r, s = A.prove_draws(rng)
opk = OG.setup(r1, toxic=[d["alpha"], d["beta"], d["gamma"], d["delta"], d["tau"]], g1_gen=d["g1"], g2_gen=d["g2"])
proof, blob = OG.prove(opk, r1, inst.z, r, s)
assert OG.verify(opk, inst.z[1:inst.num_instance], proof)
out = {"producer": "oracle", "route": "fibbonaci_handler.rs:99-110", "seed": 42, "a": 0, "b": 1, "num_of_rounds": 10,
       "draws": {k: hex(v) for k, v in d.items() if isinstance(v, int)},
       "g1_generator": O.g1_compress(d["g1"]).hex(), "g2_generator": O.g2_compress(d["g2"]).hex(),
       "r": hex(r), "s": hex(s), "proof": blob.hex()}
with open(os.path.join(os.path.dirname(__file__), "fibonacci_seed42.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out["proof"])
