#!/usr/bin/env python3
"""Development aid: aggregate proofs/s of W host threads proving concurrently on ONE GPU, each with
its own context and key copy -- the way the reference's actix workers (src/main.rs:37) would use the
library.  Usage: workers.py [workload] [max_workers] [proofs_per_worker]"""
import importlib, json, os, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
b = importlib.import_module("zksnark-finalproject_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
max_w = int(sys.argv[2]) if len(sys.argv) > 2 else 3
per = int(sys.argv[3]) if len(sys.argv) > 3 else 20
inst = bench.build_instance(name)
ctx0 = b.Context(0)
pk0, vk = b.Groth16.generate_parameters_with_qap(ctx0, inst.cm, inst.num_constraints, inst.num_instance,
                                                 inst.num_variables, *bench.toxic_waste())
z = ctx0.pin(b.codec.fr_to_mont_limbs(inst.z))
want = b.Groth16.create_proof_with_matrices(ctx0, pk0, inst.cm, z, 5, 7)
out = {}
for W in range(1, max_w + 1):
    ctxs = [ctx0] + [b.Context(0) for _ in range(W - 1)]
    pks = [pk0] + [b.ProvingKey(pk0.num_variables, pk0.num_instance, pk0.domain_size, pk0.a_query, pk0.b_g1_query,
                                pk0.b_g2_query, pk0.h_query, pk0.l_query, pk0.alpha_g1, pk0.beta_g1, pk0.delta_g1,
                                pk0.beta_g2, pk0.delta_g2).upload(c) for c in ctxs[1:]]
    cms = [inst.cm] + [b.ConstraintMatrices(inst.cm.num_instance_variables, inst.cm.num_witness_variables,
                                            inst.cm.num_constraints, inst.cm.a, inst.cm.b, inst.cm.c) for _ in ctxs[1:]]
    bad = []

    def work(i, n):
        for _ in range(n):
            if b.Groth16.create_proof_with_matrices(ctxs[i], pks[i], cms[i], z, 5, 7) != want:
                bad.append(i)

    for i in range(W):
        work(i, 2)                                     # warm-up (uploads the matrices of each context)
    ts = [threading.Thread(target=work, args=(i, per)) for i in range(W)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    out[W] = {"proofs_per_s": W * per / dt, "ms_per_proof_per_worker": dt / per * 1e3, "mismatches": len(bad)}
    print(W, out[W], flush=True)
    for p, c, m in zip(pks[1:], ctxs[1:], cms[1:]):
        m.free(); p.free(); c.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"workload": name, "workers": out}, open("gpurun_out/workers_%s.json" % name, "w"), indent=1)
