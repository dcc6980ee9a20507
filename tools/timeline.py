#!/usr/bin/env python3
"""Development aid: per-kernel-span timeline of ONE proof of a bench workload."""
import ctypes, importlib, os, random, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
b = importlib.import_module("zksnark-finalproject_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1      # > 1: time shard 0 of a point-sharded key on this one GPU
ctx = b.Context(0)
inst = bench.build_instance(name)
pk, vk = b.Groth16.generate_parameters_with_qap(ctx, inst.cm, inst.num_constraints, inst.num_instance,
                                                inst.num_variables, *bench.toxic_waste())
z = ctx.pin(b.codec.fr_to_mont_limbs(inst.z))
if world > 1:
    shard = b.ProvingKey(pk.num_variables, pk.num_instance, pk.domain_size, pk.a_query, pk.b_g1_query, pk.b_g2_query,
                         pk.h_query, pk.l_query, pk.alpha_g1, pk.beta_g1, pk.delta_g1, pk.beta_g2, pk.delta_g2)
    pk.free()
    pk = shard.upload(ctx, rank=0, world=world)
import time
for i in range(4):
    if i == 3:
        ctx._lib.b2z_profile_enable(ctx.handle, 1)
    t0 = time.perf_counter()
    if world > 1:
        b.Groth16.create_proof_partial_with_matrices(ctx, pk, inst.cm, z, 123456789, 987654321)
    else:
        b.Groth16.create_proof_with_matrices(ctx, pk, inst.cm, z, 123456789, 987654321)
    print("call %d: %.3f ms wall" % (i, (time.perf_counter() - t0) * 1e3))
N = 4096
ph = (ctypes.c_int * N)(); t0 = (ctypes.c_double * N)(); t1 = (ctypes.c_double * N)()
n = ctx._lib.b2z_profile_spans(ctx.handle, N, ph, t0, t1)
names = ["ntt_pass", "wm_pointwise", "msm_sort", "msm_accum_g1", "msm_accum_g2", "msm_reduce", "r1cs_eval",
         "msm_accum_affine"]
base = min(t0[i] for i in range(n))
for i in sorted(range(n), key=lambda i: t0[i]):
    print("%-14s %8.3f -> %8.3f  (%.3f ms)" % (names[ph[i]], t0[i] - base, t1[i] - base, t1[i] - t0[i]))
