"""Point-sharded proving, CPU side (world_size 2, gloo): the sharding rule and the host
combine (b2z_groth16_combine needs no GPU).  Each rank computes its partial sums with the
Python oracle, the partials travel through torch.distributed all_gather, and the combined
192 bytes must equal the golden proof."""
import json
import os
import subprocess
import sys

import numpy as np

from oracle import bls12_381 as O
from oracle import groth16 as OG
from helpers import oracle_r1cs

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def xyzz_limbs(codec, curve, P):
    """affine oracle point -> XYZZ limbs as the device leaves them (ZZ = ZZZ = 1, or all zero)."""
    width = 12 if curve is O.G1 else 24
    if P is None:
        return np.zeros(2 * width, dtype=np.uint64).tobytes()
    pts = (codec.g1_to_limbs if curve is O.G1 else codec.g2_to_limbs)([P])[0].reshape(-1)
    one = np.zeros(width // 2, dtype=np.uint64)
    one[:6] = codec._pack([codec.fq_to_mont(1)], 48).reshape(-1)
    return np.concatenate([pts, one, one]).tobytes()


def oracle_partial(codec, opk, inst, h, r, s, rank, world):
    """The partial sums rank `rank` owes, by the rule in include/b200zk.h (b2z_pk_upload_shard)."""
    R = O.R_MOD
    m, l, n = opk.num_variables, opk.num_instance, opk.domain_size
    # variables: block-cyclic, blocks of 64 (rank k takes blocks k, k + world, ...); h positions: contiguous chunks
    BLOCK = 64
    mine = [i for i in range(m) if world == 1 or (i // BLOCK) % world == rank]
    wit = [i for i in mine if i >= l]
    h_lo, h_hi = n * rank // world, n * (rank + 1) // world
    z = inst.z
    j1, j2 = O.G1, O.G2
    pick = lambda arr, ids, off=0: [arr[i - off] for i in ids]
    A = OG.msm_naive(j1, pick(opk.a_query, mine), pick(z, mine))
    B = OG.msm_naive(j2, pick(opk.b_g2_query, mine), pick(z, mine))
    B1 = OG.msm_naive(j1, pick(opk.b_g1_query, mine), pick(z, mine))
    Lp = OG.msm_naive(j1, pick(opk.l_query, wit, l), pick(z, wit))
    if rank == 0:
        A = j1.jadd(A, OG.msm_naive(j1, [opk.alpha_g1, opk.delta_g1], [1, r]))
        B = j2.jadd(B, OG.msm_naive(j2, [opk.beta_g2, opk.delta_g2], [1, s]))
        B1 = j1.jadd(B1, OG.msm_naive(j1, [opk.beta_g1, opk.delta_g1], [1, s]))
        Lp = j1.jadd(Lp, j1.jmul(j1.to_jac(opk.delta_g1), (-(r * s)) % R))
    sA, rB1 = j1.jmul(A, s), j1.jmul(B1, r)
    # h positions are in bit-reversed order on the device
    lg = n.bit_length() - 1
    br = lambda p: int(format(p, "0%db" % lg)[::-1], 2) if lg else 0
    idx = [br(p) for p in range(h_lo, h_hi)]
    Ch = OG.msm_naive(j1, [opk.h_query[i] if i < n - 1 else None for i in idx], [h[i] for i in idx])
    return b"".join([xyzz_limbs(codec, O.G1, j1.to_affine(P)) for P in (A, sA, rB1, Lp, Ch)] +
                    [xyzz_limbs(codec, O.G2, j2.to_affine(B))])


def test_combine_matches_golden_single_process(b2z, circuits, golden):
    case = golden["proof_fibonacci_0_1_10"]
    inst = circuits.fibonacci_circuit(0, 1, 10)
    r1 = oracle_r1cs(inst)
    opk = OG.setup(r1, seed=case["setup_seed"])
    r, s = int(case["r"], 16), int(case["s"], 16)
    h = OG.witness_map_from_evals(*OG.constraint_evaluations(r1, inst.z))
    for world in (1, 2, 3, 5):
        parts = [oracle_partial(b2z.codec, opk, inst, h, r, s, k, world) for k in range(world)]
        assert all(len(p) == b2z._ffi.PARTIAL_BYTES for p in parts)
        assert b2z.Groth16.combine(parts).hex() == case["proof"], world


WORKER = r'''
import importlib, json, os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch, torch.distributed as dist
from oracle import groth16 as OG
from helpers import oracle_r1cs
from test_sharded_host import oracle_partial
b2z = importlib.import_module("zksnark-finalproject_b200")
circuits = importlib.import_module("zksnark-finalproject_b200.circuits")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
case = json.load(open(os.path.join(sys.argv[1], "tests", "golden", "vectors.json")))["proof_fibonacci_0_1_10"]
inst = circuits.fibonacci_circuit(0, 1, 10)
r1 = oracle_r1cs(inst)
opk = OG.setup(r1, seed=case["setup_seed"])
h = OG.witness_map_from_evals(*OG.constraint_evaluations(r1, inst.z))
mine = oracle_partial(b2z.codec, opk, inst, h, int(case["r"], 16), int(case["s"], 16), rank, world)
t = torch.frombuffer(bytearray(mine), dtype=torch.uint8)
parts = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(parts, t)
proof = b2z.Groth16.combine([bytes(x.numpy().tobytes()) for x in parts])
assert proof.hex() == case["proof"], "rank %d: combined proof differs" % rank
# the tile-sharded prover's way (b2z_dist_prove ends like this): partial sums meet in POSIX shared memory, ranks
# synchronise by the host barrier on its flag words, every rank combines -- no collective at all; twice in a row
import ctypes, numpy as np
from multiprocessing import shared_memory
L = b2z._ffi.lib()
nbytes = int(L.b2z_dist_shared_bytes(world))
name = [None]
if rank == 0:
    shm = shared_memory.SharedMemory(create=True, size=nbytes)
    shm.buf[:nbytes] = bytes(nbytes)
    name[0] = shm.name
dist.broadcast_object_list(name, src=0)
if rank != 0:
    shm = shared_memory.SharedMemory(name=name[0])
buf = np.ndarray((nbytes,), dtype=np.uint8, buffer=shm.buf)
epoch = ctypes.c_uint32(0)
part = np.frombuffer(mine, dtype=np.uint8).copy()
for _ in range(2):
    out = np.zeros(192, dtype=np.uint8)
    st = L.b2z_dist_combine_shared(buf.ctypes.data_as(ctypes.c_void_p), rank, world, ctypes.byref(epoch),
                                   part.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    assert st == 0 and out.tobytes().hex() == case["proof"], "rank %d: shared-memory combine differs" % rank
assert epoch.value == 4
dist.barrier()
del buf
shm.close()
if rank == 0:
    shm.unlink()
dist.destroy_process_group()
print("rank %d ok" % rank)
'''


def test_sharded_proof_over_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout
