"""GPU parity of the batched-affine bucket accumulation (csrc/accum_affine.cuh, msm_accum_affine_kernel).

The kernel is chosen by segment length (>= 128 point references per thread), which small inputs never reach; the two
environment knobs of csrc/msm_impl.inc force it here: B2Z_AFFINE_MIN_SEG = 0 (always) and B2Z_ACCUM_MAX_SEGS (few,
long segments, so the pairwise tree runs several rounds and runs are cut by segment boundaries).  Every result is
compared with the known-multiplier answer of the Python oracle and with the XYZZ kernel on the same inputs
(replaces VariableBaseMSM::msm_bigint, ark-ec ^0.4.2 -- /root/reference/Cargo.toml:14).
"""
import os

import numpy as np
import pytest

from oracle import bls12_381 as O
from test_gpu_parity import _msm_known_multipliers, _prove_gpu, codec  # noqa: F401  (fixture re-export)

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


AFFINE = dict(B2Z_AFFINE_MIN_SEG=0, B2Z_AFFINE_G2=1)
XYZZ = dict(B2Z_AFFINE_MIN_SEG=4000000000)


@pytest.mark.parametrize("segs", [1, 7, 64, 1000])
@pytest.mark.parametrize("n,kind", [(33, "uniform"), (1000, "witness"), (4097, "uniform"), (4097, "witness")])
def test_affine_g1_small_with_special_points(b2z, ctx, codec, n, kind, segs):
    with _Env(B2Z_ACCUM_MAX_SEGS=segs, **AFFINE):
        got, want, _ = _msm_known_multipliers(b2z, ctx, codec, 1, n, kind, n + segs, with_identity=True)
    assert got == want


@pytest.mark.parametrize("kind", ["zero", "equal", "max"])
def test_affine_g1_adversarial(b2z, ctx, codec, kind):
    with _Env(B2Z_ACCUM_MAX_SEGS=37, **AFFINE):
        got, want, _ = _msm_known_multipliers(b2z, ctx, codec, 1, 3000, kind, 17)
    assert got == want


def test_affine_g1_repeated_point(b2z, ctx, codec):
    """Every base the same point and every scalar equal: each tree level is made of doublings only."""
    n = 2048
    bases1, inf1 = b2z.FixedBase.msm_g1(ctx, codec.fr_to_bigint_limbs([12345]))
    bases = np.repeat(bases1, n, axis=0)
    for s in (1, 0x1234567, O.R_MOD - 2):
        sc = codec.fr_to_bigint_limbs([s] * n)
        with _Env(B2Z_ACCUM_MAX_SEGS=5, **AFFINE):
            out = b2z.VariableBaseMSM.msm_bigint_g1(ctx, bases, sc)
        got = O.G1.to_affine(codec.g1_projective_from_limbs(out))
        assert got == O.G1.mul(O.G1_GEN, 12345 * s * n % O.R_MOD)


@pytest.mark.parametrize("n,kind,segs", [(1 << 16, "witness", 0), (1 << 16, "uniform", 2000), (70001, "uniform", 301)])
def test_affine_g1_equals_xyzz_and_oracle(b2z, ctx, codec, n, kind, segs):
    with _Env(B2Z_ACCUM_MAX_SEGS=segs, **AFFINE):
        got, want, (bases, inf, sc) = _msm_known_multipliers(b2z, ctx, codec, 1, n, kind, n)
    assert got == want
    with _Env(**XYZZ):
        out = b2z.VariableBaseMSM.msm_bigint_g1(ctx, bases, codec.fr_to_bigint_limbs(sc), inf)
    assert O.G1.to_affine(codec.g1_projective_from_limbs(out)) == got


@pytest.mark.parametrize("n,kind,segs", [(5, "witness", 1), (2000, "witness", 9), (1 << 13, "uniform", 100),
                                         (1 << 13, "witness", 0)])
def test_affine_g2(b2z, ctx, codec, n, kind, segs):
    with _Env(B2Z_ACCUM_MAX_SEGS=segs, **AFFINE):
        got, want, _ = _msm_known_multipliers(b2z, ctx, codec, 2, n, kind, n, with_identity=True)
    assert got == want


def test_affine_proof_bytes_equal_xyzz(b2z, ctx, codec, circuits):
    """One proof of a 6x6 matrix circuit with every accumulation forced through either kernel: same 192 bytes."""
    inst = circuits.matrix_circuit([[(3 * i + j) % 7 for j in range(6)] for i in range(6)], [[1] * 6] * 6)
    toxic = [3, 5, 7, 11, 13]
    proofs = []
    for env in (dict(B2Z_ACCUM_MAX_SEGS=50, **AFFINE), XYZZ):
        with _Env(**env):
            pk, vk = b2z.Groth16.generate_parameters_with_qap(ctx, inst.matrices, inst.num_constraints,
                                                              inst.num_instance, inst.num_variables, *toxic)
            got, _ = _prove_gpu(b2z, ctx, codec, pk, inst, 1234567, 7654321)
            proofs.append(bytes(got))
            pk.free()
    assert proofs[0] == proofs[1]
