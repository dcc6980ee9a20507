#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import importlib, os, random, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
b = importlib.import_module("zksnark-finalproject_b200")
fast = importlib.import_module("zksnark-finalproject_b200.circuits_fast")
codec = b.codec
R = codec.R_MOD
rnd = random.Random(3)
ctx = b.Context(0)
# NTT / coset / witness map at a multi-pass size and a tiny size
for n in (4, 1 << 13):
    v = codec.fr_to_mont_limbs([rnd.randrange(R) for _ in range(n)])
    d = b.Radix2EvaluationDomain(ctx, n)
    assert np.array_equal(d.ifft(d.fft(v)), v)
    c = d.get_coset(7)
    assert np.array_equal(c.ifft(c.fft(v)), v)
# generic MSMs incl. skewed scalars, identity bases
ks = [rnd.randrange(R) for _ in range(3000)]
ks[5] = 0
bases, inf = b.FixedBase.msm_g1(ctx, codec.fr_to_bigint_limbs(ks))
sc = [rnd.choice([0, 1, 1, 1, 2, R - 1, rnd.randrange(R)]) for _ in ks]
b.VariableBaseMSM.msm_bigint_g1(ctx, bases, codec.fr_to_bigint_limbs(sc), inf)
bases2, inf2 = b.FixedBase.msm_g2(ctx, codec.fr_to_bigint_limbs(ks[:700]))
b.VariableBaseMSM.msm_bigint_g2(ctx, bases2, codec.fr_to_bigint_limbs(sc[:700]), inf2)
# key generation, whole-key proof, sharded proof, matrices on device
cm, z_int = fast.matrix_circuit_fast([[1, 2, 3], [4, 5, 6], [7, 8, 9]], [[1, 1, 1], [1, 1, 1], [1, 1, 1]])
toxic = [rnd.randrange(1, R) for _ in range(5)]
pk, vk = b.Groth16.generate_parameters_with_qap(ctx, cm, cm.num_constraints, cm.num_instance_variables,
                                                cm.num_variables, *toxic)
z = codec.fr_to_mont_limbs(z_int)
p1 = b.Groth16.create_proof_with_matrices(ctx, pk, cm, z, 12345, 67890)
a, bb, c = b.LibsnarkReduction.constraint_evaluations_device(ctx, cm, z)
parts = []
for k in range(3):
    sh = b.ProvingKey(pk.num_variables, pk.num_instance, pk.domain_size, pk.a_query, pk.b_g1_query, pk.b_g2_query,
                      pk.h_query, pk.l_query, pk.alpha_g1, pk.beta_g1, pk.delta_g1, pk.beta_g2, pk.delta_g2)
    sh.upload(ctx, rank=k, world=3)
    parts.append(b.Groth16.create_proof_partial(ctx, sh, a, bb, c, z, 12345, 67890))
    sh.free()
assert b.Groth16.combine(parts) == p1
print("sanitize_small OK")
