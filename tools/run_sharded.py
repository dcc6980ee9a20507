#!/usr/bin/env python3
"""Point-sharded proving of ONE proof across the GPUs of a box (BASELINE configs[4]):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29517 tools/run_sharded.py --size 64 --steps 5

Every rank builds the same circuit and key (key generation on its own GPU), uploads ITS shard of
the key, and each step all ranks compute their partial sums, all_gather 1344 bytes over NCCL and
combine on the host.  Rank 0 checks the proof with the pairing verifier and prints one JSON line.
Timing: CUDA events + barrier, max over ranks.
"""
import argparse, importlib, json, os, random, sys, time
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
b = importlib.import_module("zksnark-finalproject_b200")
fast = importlib.import_module("zksnark-finalproject_b200.circuits_fast")

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--distributed-wm", action="store_true",
                help="transform a, b, c once in the box (ranks 0..2) and broadcast the coset evaluations")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
real_stdout = os.dup(1)
os.dup2(2, 1)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
codec = b.codec
R = codec.R_MOD
n = args.size
cm, z_int = fast.matrix_circuit_fast([[1] * n for _ in range(n)], [[1] * n for _ in range(n)])
ctx = b.Context(local)
rnd = random.Random(0xB2000004)
toxic = [rnd.randrange(1, R) for _ in range(5)]
t0 = time.time()
pk, vk = b.Groth16.generate_parameters_with_qap(ctx, cm, cm.num_constraints, cm.num_instance_variables,
                                                cm.num_variables, *toxic)
t_keygen = time.time() - t0
pk.upload(ctx, rank=rank, world=world)
cm.upload(ctx)
z = ctx.pin(codec.fr_to_mont_limbs(z_int))      # page-locked: the 61 MB upload per proof runs at PCIe speed
a, bb, c = b.LibsnarkReduction.constraint_evaluations_device(ctx, cm, z)
r, s = rnd.randrange(R), rnd.randrange(R)


bufs = [torch.empty((pk.domain_size, 4), dtype=torch.int64, device="cuda") for _ in range(3)] if args.distributed_wm else None


def step():
    if args.distributed_wm:
        return b.Groth16.create_proof_sharded_distributed(ctx, pk, cm, z, r, s, buffers=bufs)
    return b.Groth16.create_proof_sharded(ctx, pk, None, None, None, z, r, s, cm=cm)


proof = step()
step()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); torch.cuda.synchronize()
w0 = time.perf_counter()
e0.record()
for _ in range(args.steps):
    p2 = step()
e1.record()
dist.barrier(); torch.cuda.synchronize()
wall = (time.perf_counter() - w0) / args.steps
t = torch.tensor([e0.elapsed_time(e1) / args.steps, wall * 1e3], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert p2 == proof
if rank == 0:
    from oracle import bls12_381 as O, groth16 as OG
    class V: pass
    v = V()
    v.alpha_g1 = codec.g1_from_limbs(vk.alpha_g1.reshape(1, -1))[0]
    v.beta_g2, v.gamma_g2, v.delta_g2 = (codec.g2_from_limbs(x.reshape(1, -1))[0] for x in (vk.beta_g2, vk.gamma_g2, vk.delta_g2))
    v.gamma_abc_g1 = codec.g1_from_limbs(*vk.gamma_abc_g1)
    ok = OG.verify(v, z_int[1:cm.num_instance_variables], O.proof_deserialize_compressed(proof))
    line = {"mode": "point-sharded single proof" + (", distributed witness map" if args.distributed_wm else ""), "n_gpus": world, "workload": "matrix %dx%d" % (n, n),
            "num_constraints": cm.num_constraints, "domain": cm.domain_size, "ms_per_proof_device": float(t[0]),
            "ms_per_proof_wall": float(t[1]), "proof_verifies": bool(ok), "keygen_s": t_keygen,
            "collective": ("3 broadcasts of %d MB (coset evaluations) + " % (cm.domain_size * 32 >> 20) if args.distributed_wm else "")
                          + "all_gather of %d B per rank (NCCL)" % b._ffi.PARTIAL_BYTES,
            "inputs": "z in host memory on every rank (H2D inside the timed region); rows evaluated on the GPU"}
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
dist.barrier()
dist.destroy_process_group()
