"""Shared helpers for the parity tests (oracle <-> C-ABI layouts)."""
import random

from oracle import bls12_381 as O
from oracle import groth16 as OG


def oracle_r1cs(inst):
    return OG.R1CS(inst.num_instance, inst.num_witness, inst.a, inst.b, inst.c)


def pk_limbs(codec, opk):
    """Oracle ProvingKey -> the limb arrays b2z_pk_desc / CpuProvingKey take."""
    q1, q2 = codec.g1_to_limbs, codec.g2_to_limbs
    one1 = lambda p: q1([p])[0].reshape(-1)
    one2 = lambda p: q2([p])[0].reshape(-1)
    return (opk.num_variables, opk.num_instance, opk.domain_size, q1(opk.a_query), q1(opk.b_g1_query),
            q2(opk.b_g2_query), q1(opk.h_query), q1(opk.l_query), one1(opk.alpha_g1), one1(opk.beta_g1),
            one1(opk.delta_g1), one2(opk.beta_g2), one2(opk.delta_g2))


def scalar_mix(rnd, n, kind):
    """Scalar distributions of SURVEY.md 8(d): uniform, witness-like, adversarial."""
    R = O.R_MOD
    if kind == "uniform":
        return [rnd.randrange(R) for _ in range(n)]
    if kind == "witness":
        out = []
        for _ in range(n):
            u = rnd.random()
            out.append(rnd.randrange(2) if u < 0.3 else rnd.randrange(1 << 16) if u < 0.5 else rnd.randrange(R))
        return out
    if kind == "zero":
        return [0] * n
    if kind == "equal":
        return [rnd.randrange(R)] * n
    if kind == "max":
        return [R - 1] * n
    raise ValueError(kind)


def rng(seed):
    return random.Random(seed)
