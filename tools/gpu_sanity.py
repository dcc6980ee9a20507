#!/usr/bin/env python3
"""First-contact GPU sanity run (development aid; the real checks live in tests/).
Compares every C-ABI entry point with the Python oracle on small inputs."""
import importlib
import os
import random
import sys
import time
import traceback

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
b = importlib.import_module("zksnark-finalproject_b200")
circuits = importlib.import_module("zksnark-finalproject_b200.circuits")
from oracle import bls12_381 as O          # noqa: E402
from oracle import groth16 as OG           # noqa: E402

codec = b.codec
rnd = random.Random(1234)
ctx = b.Context(0)
results = []


def check(name, fn):
    t = time.time()
    try:
        ok = fn()
    except Exception:
        traceback.print_exc()
        ok = False
    results.append((name, ok))
    print("%-40s %s  (%.2fs)" % (name, "OK" if ok else "FAIL", time.time() - t), flush=True)


def t_ntt(logn, coset):
    def run():
        n = 1 << logn
        v = [rnd.randrange(O.R_MOD) for _ in range(n)]
        dom = OG.Radix2EvaluationDomain(n)
        if coset:
            dom = dom.get_coset(7)
        want_f = dom.fft(list(v))
        want_i = dom.ifft(list(v))
        d = b.Radix2EvaluationDomain(ctx, n)
        if coset:
            d = d.get_coset(7)
        got_f = codec.fr_from_mont_limbs(d.fft(codec.fr_to_mont_limbs(v)))
        got_i = codec.fr_from_mont_limbs(d.ifft(codec.fr_to_mont_limbs(v)))
        return got_f == want_f and got_i == want_i
    return run


for logn in (0, 1, 2, 3, 4, 5, 8, 10, 11, 12, 13):
    check("ntt logn=%d" % logn, t_ntt(logn, False))
for logn in (1, 4, 10, 12):
    check("ntt coset logn=%d" % logn, t_ntt(logn, True))


def t_wm(logn):
    def run():
        n = 1 << logn
        a = [rnd.randrange(O.R_MOD) for _ in range(n)]
        bb = [rnd.randrange(O.R_MOD) for _ in range(n)]
        # c = a*b on the domain makes h a genuine quotient, but the map itself is defined for any c
        c = [x * y % O.R_MOD for x, y in zip(a, bb)]
        want = OG.witness_map_from_evals(a, bb, c)
        got = codec.fr_from_mont_limbs(b.LibsnarkReduction.witness_map_from_evaluations(
            ctx, codec.fr_to_mont_limbs(a), codec.fr_to_mont_limbs(bb), codec.fr_to_mont_limbs(c)))
        return got == want
    return run


for logn in (0, 1, 3, 7, 10, 12, 13):
    check("witness_map logn=%d" % logn, t_wm(logn))


def t_fixed():
    ks = [0, 1, 2, O.R_MOD - 1] + [rnd.randrange(O.R_MOD) for _ in range(20)]
    out, inf = b.FixedBase.msm_g1(ctx, codec.fr_to_bigint_limbs(ks))
    got = codec.g1_from_limbs(out, inf)
    want = [O.G1.mul(O.G1_GEN, k) for k in ks]
    out2, inf2 = b.FixedBase.msm_g2(ctx, codec.fr_to_bigint_limbs(ks[:8]))
    got2 = codec.g2_from_limbs(out2, inf2)
    want2 = [O.G2.mul(O.G2_GEN, k) for k in ks[:8]]
    return got == want and got2 == want2


check("fixed base g1/g2", t_fixed)


def t_msm(group, n, kind):
    def run():
        curve = O.G1 if group == 1 else O.G2
        ks = [rnd.randrange(1, O.R_MOD) for _ in range(n)]
        if group == 1:
            out, inf = b.FixedBase.msm_g1(ctx, codec.fr_to_bigint_limbs(ks))
        else:
            out, inf = b.FixedBase.msm_g2(ctx, codec.fr_to_bigint_limbs(ks))
        if kind == "uniform":
            sc = [rnd.randrange(O.R_MOD) for _ in range(n)]
        elif kind == "small":
            sc = [rnd.choice([0, 1, 1, 2, rnd.randrange(1 << 16), O.R_MOD - 1]) for _ in range(n)]
        else:
            sc = [O.R_MOD - 1] * n
        want_k = sum(k * s for k, s in zip(ks, sc)) % O.R_MOD
        want = curve.mul(curve.gen, want_k)
        fn = b.VariableBaseMSM.msm_bigint_g1 if group == 1 else b.VariableBaseMSM.msm_bigint_g2
        res = fn(ctx, out, codec.fr_to_bigint_limbs(sc), inf)
        if group == 1:
            X, Y, Z = codec.g1_projective_from_limbs(res)
            got = O.G1.to_affine((X, Y, Z))
        else:
            X, Y, Z = codec.g2_projective_from_limbs(res)
            got = O.G2.to_affine((X, Y, Z))
        return got == want
    return run


for n in (1, 2, 31, 100, 1000, 5000):
    check("msm g1 n=%d uniform" % n, t_msm(1, n, "uniform"))
check("msm g1 n=3000 small", t_msm(1, 3000, "small"))
check("msm g1 n=500 allmax", t_msm(1, 500, "max"))
check("msm g1 n=70000 uniform", t_msm(1, 70000, "uniform"))
for n in (1, 50, 2000):
    check("msm g2 n=%d uniform" % n, t_msm(2, n, "uniform"))
check("msm g2 n=1500 small", t_msm(2, 1500, "small"))


def pk_from_oracle(opk):
    q1 = codec.g1_to_limbs
    q2 = codec.g2_to_limbs
    one1 = lambda p: q1([p])[0].reshape(-1)
    one2 = lambda p: q2([p])[0].reshape(-1)
    return b.ProvingKey(opk.num_variables, opk.num_instance, opk.domain_size, q1(opk.a_query), q1(opk.b_g1_query),
                        q2(opk.b_g2_query), q1(opk.h_query), q1(opk.l_query), one1(opk.alpha_g1), one1(opk.beta_g1),
                        one1(opk.delta_g1), one2(opk.beta_g2), one2(opk.delta_g2))


def t_prove(inst, name):
    def run():
        r1 = OG.R1CS(inst.num_instance, inst.num_witness, inst.a, inst.b, inst.c)
        opk = OG.setup(r1)
        r, s = rnd.randrange(O.R_MOD), rnd.randrange(O.R_MOD)
        (A, B, C), want = OG.prove(opk, r1, inst.z, r, s)
        assert OG.verify(opk, inst.z[1:inst.num_instance], (A, B, C)), "oracle proof does not verify"
        pk = pk_from_oracle(opk)
        a, bb, c = b.LibsnarkReduction.constraint_evaluations(inst.matrices, inst.num_instance, inst.num_constraints,
                                                              inst.z)
        got = b.Groth16.create_proof_with_reduction(ctx, pk, a, bb, c, codec.fr_to_mont_limbs(inst.z), r, s)
        if got != want:
            print("   want", want.hex())
            print("   got ", got.hex())
        pk.free()
        return got == want
    return run


check("prove fibonacci(10)", t_prove(circuits.fibonacci_circuit(0, 1, 10), "fib10"))
check("prove matrix 2x2", t_prove(circuits.matrix_circuit([[1, 2], [3, 4]], [[4, 3], [2, 1]]), "m2"))
bad = [n for n, ok in results if not ok]
print("FAILED: %s" % bad if bad else "ALL OK")
sys.exit(1 if bad else 0)
