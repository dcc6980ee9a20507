#!/usr/bin/env python3
"""Writes tests/golden/witness_vectors.json from oracle/witness.py (hashlib + Python integers; "producer": "oracle"):
the prime route's search for the reference's seeds (x = 5: prime_circut.rs:361), modpow tables, and Poseidon digests
under the repository's seed-derived parameters.  The native code (csrc/witness.cu) is NOT involved in making them."""
import importlib
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import witness as OW      # noqa: E402

circuits = importlib.import_module("zksnark-finalproject_b200.circuits")
ps = circuits.PoseidonShape()
cfg = OW.PoseidonConfig(ps.FULL, ps.PARTIAL, ps.ALPHA, ps.ark, ps.mds, 2, 1)
out = {"producer": "oracle", "prime_search": [], "modpow": [], "poseidon": []}
for x in (5, 0, 1, 42, 2 ** 64 - 1):
    j, c = OW.prime_search(x, 200)
    out["prime_search"].append({"x": x, "i": 200, "j": j, "digest": c["digest"].hex(), "is_prime": c["is_prime"],
                                "quotient": hex(c["quotient"]), "remainder": c["remainder"], "a": hex(c["a"])})
for base, mod, exp, nb in ((2, 506183, 506182, 20), (7, 1048573, 1048572, 20), (3, 2 ** 61 - 1, 12345678901234567, 61)):
    w = OW.mod_pow_generate_witnesses(base, mod, exp, nb)
    out["modpow"].append({"base": base, "modulus": mod, "exponent": exp, "num_bits": nb, "result": w["result"],
                          "bits": w["bits"], "mod_vals": [[hex(v) for v in row] for row in w["mod_vals"]],
                          "mod_pow_vals": [[hex(v) for v in row] for row in w["mod_pow_vals"]]})
for elems in ([], [1], [1, 2], [1, 2, 3], list(range(1, 17)), [OW.R_MOD - 1] * 5):
    out["poseidon"].append({"params": "circuits.PoseidonShape(seed 0xB2005EED)", "elems": [hex(e) for e in elems],
                            "digest": hex(OW.poseidon_hash(cfg, elems))})
with open(os.path.join(os.path.dirname(__file__), "witness_vectors.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out["prime_search"]), "searches,", len(out["modpow"]), "modpow tables,", len(out["poseidon"]), "digests")
