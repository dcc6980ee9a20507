#!/usr/bin/env python3
"""Development aid: per-kernel-span timeline of ONE proof of a bench workload."""
import ctypes, importlib, os, random, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
b = importlib.import_module("zksnark-finalproject_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
ctx = b.Context(0)
inst = bench.build_instance(name)
pk, vk = b.Groth16.generate_parameters_with_qap(ctx, inst.cm, inst.num_constraints, inst.num_instance,
                                                inst.num_variables, *bench.toxic_waste())
z = b.codec.fr_to_mont_limbs(inst.z)
for i in range(4):
    if i == 3:
        ctx._lib.b2z_profile_enable(ctx.handle, 1)
    b.Groth16.create_proof_with_matrices(ctx, pk, inst.cm, z, 123456789, 987654321)
N = 4096
ph = (ctypes.c_int * N)(); t0 = (ctypes.c_double * N)(); t1 = (ctypes.c_double * N)()
n = ctx._lib.b2z_profile_spans(ctx.handle, N, ph, t0, t1)
names = ["ntt_pass", "wm_pointwise", "msm_sort", "msm_accum_g1", "msm_accum_g2", "msm_reduce", "finalize"]
base = min(t0[i] for i in range(n))
for i in sorted(range(n), key=lambda i: t0[i]):
    print("%-14s %8.3f -> %8.3f  (%.3f ms)" % (names[ph[i]], t0[i] - base, t1[i] - base, t1[i] - t0[i]))
