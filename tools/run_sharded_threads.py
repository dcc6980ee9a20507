#!/usr/bin/env python3
"""Point-sharded proving of ONE proof across the GPUs of a box from ONE process: a host thread
per device, each with its own context, key shard and matrix copy; the main thread sums the
1344-byte partials (b2z_groth16_combine, host only).  No NCCL, no launcher -- the shape a Rust
server with one worker thread per GPU would use.

    python tools/run_sharded_threads.py --gpus 8 --size 64 --steps 5
"""
import argparse, importlib, json, os, random, sys, threading, time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
b = importlib.import_module("zksnark-finalproject_b200")
fast = importlib.import_module("zksnark-finalproject_b200.circuits_fast")

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=8)
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--owner-weight", type=int, default=100,
                help="points given to each of GPUs 0..2 per 100 given to each other GPU (with --distributed-wm)")
ap.add_argument("--distributed-wm", action="store_true",
                help="GPUs 0..2 transform a, b, c once and push the coset evaluations to their peers (copy engines)")
args = ap.parse_args()
codec = b.codec
R = codec.R_MOD
n, G = args.size, args.gpus
cm, z_int = fast.matrix_circuit_fast([[1] * n for _ in range(n)], [[1] * n for _ in range(n)])
ctxs = [b.Context(k) for k in range(G)]
rnd = random.Random(0xB2000004)
toxic = [rnd.randrange(1, R) for _ in range(5)]
t0 = time.time()
pk, vk = b.Groth16.generate_parameters_with_qap(ctxs[0], cm, cm.num_constraints, cm.num_instance_variables,
                                                cm.num_variables, *toxic)
t_keygen = time.time() - t0
pk.free()
cm.free()
shards, cms = [None] * G, [None] * G
weights = None
if args.distributed_wm and args.owner_weight != 100:
    weights = [args.owner_weight if k < min(3, G) else 100 for k in range(G)]


def setup(k):
    shards[k] = b.ProvingKey(pk.num_variables, pk.num_instance, pk.domain_size, pk.a_query, pk.b_g1_query,
                             pk.b_g2_query, pk.h_query, pk.l_query, pk.alpha_g1, pk.beta_g1, pk.delta_g1,
                             pk.beta_g2, pk.delta_g2).upload(ctxs[k], rank=k, world=G, weights=weights)
    cms[k] = b.ConstraintMatrices(cm.num_instance_variables, cm.num_witness_variables, cm.num_constraints,
                                  cm.a, cm.b, cm.c).upload(ctxs[k])


ts = [threading.Thread(target=setup, args=(k,)) for k in range(G)]
[t.start() for t in ts]
[t.join() for t in ts]
z = ctxs[0].pin(codec.fr_to_mont_limbs(z_int))
r, s = rnd.randrange(R), rnd.randrange(R)
parts = [None] * G
go = [threading.Semaphore(0) for _ in range(G)]
if args.distributed_wm:
    import torch
    # bufs[k][j]: coset evaluations of matrix j on device k
    bufs = [[torch.empty((pk.domain_size, 4), dtype=torch.int64, device="cuda:%d" % k) for _ in range(3)] for k in range(G)]
    ready = [threading.Event() for _ in range(3)]
    owners = [j % G for j in range(3)]
done = threading.Semaphore(0)
stop = False


def worker(k):
    while True:
        go[k].acquire()
        if stop:
            return
        if args.distributed_wm:
            uploaded = False
            for j in range(3):
                if owners[j] == k:          # transform first: the accumulations would starve it
                    b.Groth16.coset_evals(ctxs[k], cms[k], j, bufs[k][j].data_ptr(), None if uploaded else z)
                    uploaded = True
                    with torch.cuda.device(k):
                        for peer in range(G):
                            if peer != k:   # peer-to-peer DMA: needs no SM on either side
                                bufs[peer][j].copy_(bufs[k][j], non_blocking=True)
            keep = b.Groth16.shard_begin(ctxs[k], shards[k], cms[k], None if uploaded else z, r, s)
            if uploaded:
                torch.cuda.synchronize(k)
                for j in range(3):
                    if owners[j] == k:
                        ready[j].set()
            for j in range(3):
                ready[j].wait()
            parts[k] = b.Groth16.shard_finish(ctxs[k], shards[k], *(t.data_ptr() for t in bufs[k]))
        else:
            parts[k] = b.Groth16.create_proof_partial_with_matrices(ctxs[k], shards[k], cms[k], z, r, s)
        done.release()


workers = [threading.Thread(target=worker, args=(k,), daemon=True) for k in range(G)]
[t.start() for t in workers]


def step():
    if args.distributed_wm:
        for e in ready:
            e.clear()
    for k in range(G):
        go[k].release()
    for _ in range(G):
        done.acquire()
    return b.Groth16.combine(parts)


proof = step()
step()
w0 = time.perf_counter()
for _ in range(args.steps):
    p2 = step()
wall = (time.perf_counter() - w0) / args.steps
assert p2 == proof
stop = True
[g.release() for g in go]
from oracle import bls12_381 as O, groth16 as OG


class V:
    pass


v = V()
v.alpha_g1 = codec.g1_from_limbs(vk.alpha_g1.reshape(1, -1))[0]
v.beta_g2, v.gamma_g2, v.delta_g2 = (codec.g2_from_limbs(x.reshape(1, -1))[0] for x in (vk.beta_g2, vk.gamma_g2, vk.delta_g2))
v.gamma_abc_g1 = codec.g1_from_limbs(*vk.gamma_abc_g1)
ok = OG.verify(v, z_int[1:cm.num_instance_variables], O.proof_deserialize_compressed(proof))
line = {"mode": "point-sharded single proof, one process, one host thread per GPU"
                + (", distributed witness map (peer copies)" if args.distributed_wm else ""), "n_gpus": G,
        "workload": "matrix %dx%d" % (n, n), "num_constraints": cm.num_constraints, "domain": cm.domain_size,
        "ms_per_proof_wall": wall * 1e3, "shard_weights": weights, "proof_verifies": bool(ok), "keygen_s": t_keygen,
        "collective": "none (host sum of %d B per GPU)" % b._ffi.PARTIAL_BYTES,
        "inputs": "z in page-locked host memory (H2D to every GPU inside the timed region); rows evaluated on the GPU"}
print(json.dumps(line))
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/sharded_threads%s_%d_%d.json" % ("_dwm" if args.distributed_wm else "", n, G), "w").write(json.dumps(line) + "\n")
