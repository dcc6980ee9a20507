"""CPU-only checks of the product's host side: the C-ABI library loads and
exports what include/b200zk.h declares, fails loudly without a GPU, and the
limb algorithms of the device code (run through the host carry-flag emulation)
agree with the oracle."""
import ctypes
import os
import random
import re

import numpy as np
import pytest

from oracle import bls12_381 as O

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_header_symbols_are_exported(b2z):
    hdr = open(os.path.join(ROOT, "include", "b200zk.h")).read()
    declared = set(re.findall(r"B2Z_API\s+[\w\s\*]+?\b(b2z_\w+)\s*\(", hdr))
    assert len(declared) >= 17
    L = ctypes.CDLL(b2z._ffi.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), "header declares %s but the library does not export it" % name
    assert declared == set(b2z._ffi.SIGNATURES), "ctypes table out of sync with the header"


def test_no_cpu_fallback_without_gpu(b2z):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(b2z._ffi.B2zError) as e:
        b2z.Context(0)
    assert e.value.status == b2z._ffi.B2Z_ECUDA


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "zksnark-finalproject_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".inc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "ark_cpu" not in txt, f


def _arr32(x, n):
    return np.array([(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)], dtype=np.uint32)


def _val(a):
    return sum(int(v) << (32 * i) for i, v in enumerate(a))


@pytest.mark.parametrize("field,p,n", [(0, O.R_MOD, 8), (1, O.Q_MOD, 12)])
def test_device_field_arithmetic_on_host(b2z, field, p, n):
    L = b2z._ffi.lib()
    R = 1 << (32 * n)
    rnd = random.Random(field)
    edge = [0, 1, p - 1, p, p + 1, 2 * p - 1]
    for it in range(3000):
        a = rnd.choice(edge) if rnd.random() < 0.15 else rnd.randrange(2 * p)
        b = rnd.choice(edge) if rnd.random() < 0.15 else rnd.randrange(2 * p)
        A, B = _arr32(a, n), _arr32(b, n)
        out = np.zeros(n, dtype=np.uint32)
        for op, want in ((0, a * b * pow(R, -1, p)), (1, a + b), (2, a - b)):
            assert L.b2z_host_field_op(field, op, A.ctypes.data, B.ctypes.data, out.ctypes.data) == 0
            r = _val(out)
            assert r < 2 * p and (r - want) % p == 0, (field, op, hex(a), hex(b))
        L.b2z_host_field_op(field, 3, A.ctypes.data, None, out.ctypes.data)
        assert _val(out) == a % p
        # raw product with the operand precondition of mont.cuh (Fr: first operand canonical)
        a1 = a % p if field == 0 else a
        L.b2z_host_field_op(field, 5, _arr32(a1, n).ctypes.data, B.ctypes.data, out.ctypes.data)
        r = _val(out)
        assert r < 2 * p and (r * R - a1 * b) % p == 0
    a = rnd.randrange(1, p)
    out = np.zeros(n, dtype=np.uint32)
    L.b2z_host_field_op(field, 4, _arr32(a * R % p, n).ctypes.data, None, out.ctypes.data)
    assert _val(out) == pow(a, -1, p) * R % p


def test_division_step_inversion_on_host(b2z):
    """csrc/inv_gcd.cuh against pow(a, -1, q): random, lazy (>= q), tiny, huge and power-of-two inputs; 0 -> 0."""
    L = b2z._ffi.lib()
    q, n = O.Q_MOD, 12
    R = 1 << 384
    rnd = random.Random(7)
    cases = [1, 2, 3, q - 1, q - 2, (q - 1) // 2, (q + 1) // 2, 1 << 380, (1 << 380) + 1, R % q, pow(R, -1, q)]
    cases += [1 << k for k in range(1, 380, 13)] + [rnd.randrange(1, q) for _ in range(1500)]
    out = np.zeros(n, dtype=np.uint32)
    worst = 0
    for a in cases:
        for lazy in (0, q):                      # the canonical value and its lazy twin a + q
            if a + lazy >= 2 * q:
                continue
            batches = L.b2z_host_fq_inv_gcd(_arr32(a + lazy, n).ctypes.data, out.ctypes.data)
            assert 0 < batches < 40, (hex(a), batches)
            worst = max(worst, batches)
            # input is x = a (taken as a Montgomery value a' R); output must be a'^-1 R = R^2 / x
            assert _val(out) == pow(a, -1, q) * R * R % q, hex(a)
    assert worst <= 37                            # (49 * 382 + 57) / 17 = 1104 division steps at most
    assert L.b2z_host_fq_inv_gcd(_arr32(0, n).ctypes.data, out.ctypes.data) == 0 and _val(out) == 0
    assert L.b2z_host_fq_inv_gcd(_arr32(q, n).ctypes.data, out.ctypes.data) == 0 and _val(out) == 0


@pytest.mark.parametrize("group", [1, 2])
def test_device_point_arithmetic_on_host(b2z, group):
    L = b2z._ffi.lib()
    codec = b2z.codec
    curve = O.G1 if group == 1 else O.G2
    rnd = random.Random(group)
    pts = [curve.mul(curve.gen, rnd.randrange(1, 1 << 20)) for _ in range(10)]
    pts += [pts[0], pts[0], curve.neg(pts[1]), pts[1], pts[4]]          # doubling, cancellation, repeats
    neg = [rnd.randrange(2) for _ in pts]
    limbs = (codec.g1_to_limbs if group == 1 else codec.g2_to_limbs)(pts)[0]
    out = np.zeros(12 * group, dtype=np.uint64)
    negs = np.array(neg, dtype=np.uint8)
    rc = L.b2z_host_point_sum(group, limbs.ctypes.data, negs.ctypes.data, len(pts), out.ctypes.data)
    want = None
    for p, ng in zip(pts, neg):
        want = curve.add(want, curve.neg(p) if ng else p)
    assert rc == 0
    got = (codec.g1_from_limbs if group == 1 else codec.g2_from_limbs)(out.reshape(1, -1))[0]
    assert got == want
    two = (codec.g1_to_limbs if group == 1 else codec.g2_to_limbs)([pts[3], pts[3]])[0]
    rc = L.b2z_host_point_sum(group, two.ctypes.data, np.array([0, 1], dtype=np.uint8).ctypes.data, 2, out.ctypes.data)
    assert rc == 1                                                          # P + (-P) is the identity


@pytest.mark.parametrize("group", [1, 2])
def test_host_epilogue_planes_horner(b2z, group):
    """host_fq.hpp (the prover's host epilogue): the serial last step of every MSM -- planes[0] +
    2^chunk_log * sum 2^(k-1) planes[k] over XYZZ points -- and the zcash serialization, vs the oracle."""
    L = b2z._ffi.lib()
    curve = O.G1 if group == 1 else O.G2
    rnd = random.Random(40 + group)
    q = O.Q_MOD
    for nplanes, chunk_log in ((0, 3), (1, 3), (2, 3), (13, 3), (7, 2)):
        pts = [curve.mul(curve.gen, rnd.randrange(1, 1 << 30)) for _ in range(nplanes)]
        if nplanes > 4:
            pts[2] = None                      # an empty plane
            pts[4] = pts[3]                    # forces a doubling inside an addition
        words = []
        for p in pts:
            if p is None:
                coords = [0] * (4 * group)
            elif group == 1:
                t = rnd.randrange(1, q)        # a non-trivial XYZZ representative: (x t^2, y t^3, t^2, t^3)
                coords = [p[0] * t * t % q, p[1] * pow(t, 3, q) % q, t * t % q, pow(t, 3, q)]
            else:
                t = (rnd.randrange(1, q), rnd.randrange(q))
                F = O.Fq2Ops
                t2 = F.sqr(t); t3 = F.mul(t2, t)
                coords = [v for pair in (F.mul(p[0], t2), F.mul(p[1], t3), t2, t3) for v in pair]
            for v in coords:
                words.append(_arr32(O.fq_to_mont(v), 12))
        planes = np.concatenate(words) if words else np.zeros(1, dtype=np.uint32)
        out = np.zeros(48 * group, dtype=np.uint8)
        assert L.b2z_host_planes_horner(group, planes.ctypes.data, nplanes, chunk_log, out.ctypes.data) == 0
        want = pts[0] if nplanes else None
        if nplanes > 1:
            h = None
            for k in range(nplanes - 1, 0, -1):
                h = curve.add(curve.add(h, h), pts[k])
            for _ in range(chunk_log):
                h = curve.add(h, h)
            want = curve.add(want, h)
        assert out.tobytes() == (O.g1_compress if group == 1 else O.g2_compress)(want), (nplanes, chunk_log)


@pytest.mark.parametrize("c", [4, 5, 8, 11, 13, 15, 16, 17, 20])
def test_msm_digit_recoding(b2z, c):
    L = b2z._ffi.lib()
    rnd = random.Random(c)
    cases = [0, 1, O.R_MOD - 1, (1 << 254), (1 << 254) + (1 << 253)] + [rnd.randrange(O.R_MOD) for _ in range(300)]
    cases += [rnd.randrange(1 << 16) for _ in range(50)]
    for k in cases:
        d = np.zeros(64, dtype=np.int32)
        w = L.b2z_host_msm_digits(_arr32(k % O.R_MOD, 8).ctypes.data, c, d.ctypes.data)
        assert sum(int(d[i]) << (c * i) for i in range(w)) == k % O.R_MOD
        assert all(-(1 << (c - 1)) < int(d[i]) <= (1 << (c - 1)) for i in range(w))     # fits 2^(c-1) buckets
        assert int(d[w - 1]) >= 0


def test_codec_roundtrip(b2z):
    codec = b2z.codec
    rnd = random.Random(1)
    v = [0, 1, O.R_MOD - 1] + [rnd.randrange(O.R_MOD) for _ in range(20)]
    assert codec.fr_from_mont_limbs(codec.fr_to_mont_limbs(v)) == v
    assert codec.fr_from_bigint_limbs(codec.fr_to_bigint_limbs(v)) == v
    # Montgomery one = R mod r
    assert codec.fr_to_mont_limbs([1])[0].tolist() == [0x00000001FFFFFFFE, 0x5884B7FA00034802, 0x998C4FEFECBC4FF5,
                                                        0x1824B159ACC5056F]
    pts = [O.G1_GEN, None, O.G1.mul(O.G1_GEN, 77)]
    L, inf = codec.g1_to_limbs(pts)
    assert codec.g1_from_limbs(L, inf) == pts and L.shape == (3, 12)
    pts2 = [O.G2_GEN, None]
    L2, inf2 = codec.g2_to_limbs(pts2)
    assert codec.g2_from_limbs(L2, inf2) == pts2 and L2.shape == (2, 24)


def test_circuit_shapes(circuits):
    fib = circuits.fibonacci_circuit(0, 1, 1000)                # BASELINE config 1
    assert (fib.num_constraints, fib.num_instance, fib.num_witness, fib.domain_size) == (1001, 4, 1, 1024)
    assert fib.is_satisfied() and fib.a[0] == [] and fib.b[0] == [(1, 0)]
    for n in (2, 3, 4):
        m = circuits.matrix_circuit([[i * n + j for j in range(n)] for i in range(n)], [[1] * n for _ in range(n)])
        assert m.num_constraints == circuits.matrix_constraint_count(n) and m.num_instance == 4
        assert m.is_satisfied()
    assert circuits.matrix_constraint_count(16) == 109955 and circuits.matrix_constraint_count(64) == 2152451
    p = circuits.prime_circuit(5, num_bits=8, sha_blocks=1)
    assert p.is_satisfied() and p.num_instance == 4
    bools = sum(1 for v in p.z if v in (0, 1))
    assert bools > 0.6 * len(p.z)


@pytest.mark.parametrize("n", [2, 3])
def test_vectorised_matrix_circuit_equals_symbolic_builder(b2z, circuits, n):
    import importlib
    fast = importlib.import_module("zksnark-finalproject_b200.circuits_fast")
    A = [[(3 * i + j + 1) for j in range(n)] for i in range(n)]
    B = [[(i * j + 2) for j in range(n)] for i in range(n)]
    slow = circuits.matrix_circuit(A, B)
    cm, z = fast.matrix_circuit_fast(A, B)
    assert z == slow.z and cm.num_constraints == slow.num_constraints and cm.num_variables == slow.num_variables
    for got, want in zip(cm.rows(), (slow.a, slow.b, slow.c)):
        assert all(sorted(x) == sorted(y) for x, y in zip(got, want))
    assert cm.domain_size == slow.domain_size


def test_domain_too_large_maps_to_polynomial_degree_too_large(b2z):
    with pytest.raises(b2z.PolynomialDegreeTooLarge):
        b2z.Radix2EvaluationDomain(None, (1 << 32) + 1)
    assert issubclass(b2z.PolynomialDegreeTooLarge, b2z.SynthesisError)


def test_uploaded_handles_are_bound_to_their_context_and_shard(b2z):
    """ProvingKey.upload / ConstraintMatrices.upload never hand back a handle that was made for another
    context or another shard (host logic only: the handle is faked, nothing touches the GPU)."""
    import numpy as np
    z12, z24 = np.zeros((1, 12), np.uint64), np.zeros((1, 24), np.uint64)
    pk = b2z.ProvingKey(1, 1, 1, (z12, None), (z12, None), (z24, None), (z12[:0], None), (z12[:0], None),
                        z12[0], z12[0], z12[0], z24[0], z24[0])
    ctx_a, ctx_b = object(), object()
    pk._ctx, pk._handle, pk._spec = ctx_a, object(), (0, 2, None)
    assert pk.upload(ctx_a, rank=0, world=2) is pk
    with pytest.raises(ValueError):
        pk.upload(ctx_b, rank=0, world=2)
    with pytest.raises(ValueError):
        pk.upload(ctx_a, rank=1, world=2)
    with pytest.raises(ValueError):
        pk.upload(ctx_a)
    with pytest.raises(ValueError):
        b2z.Groth16.create_proof_partial(ctx_b, pk, None, None, None, None, 1, 1)
    rp = np.zeros(1, np.uint64)
    cm = b2z.ConstraintMatrices(1, 0, 0, (rp, np.zeros(0, np.uint32), np.zeros((0, 4), np.uint64)),
                                (rp, np.zeros(0, np.uint32), np.zeros((0, 4), np.uint64)),
                                (rp, np.zeros(0, np.uint32), np.zeros((0, 4), np.uint64)))
    cm._ctx, cm._handle = ctx_a, object()
    assert cm.upload(ctx_a) is cm
    with pytest.raises(ValueError):
        cm.upload(ctx_b)
    pk._handle = cm._handle = None          # nothing real to free


def _accum_case(rnd, npoints, run_lengths, special):
    """sorted references + bucket offsets with the given run lengths; `special` plants adversarial runs."""
    sorted_refs, offsets = [], [0]
    for b, R in enumerate(run_lengths):
        kind = special.get(b)
        if kind == "same":                       # one point over and over: a doubling at every tree level
            refs = [5] * R
        elif kind == "cancel":                   # P, -P, P, -P ...: identity markers travel up the tree
            refs = [(7 | (0x80000000 if i & 1 else 0)) for i in range(R)]
        elif kind == "mixed":                    # P, P, -P, -P, Q, ...
            refs = ([9, 9, 9 | 0x80000000, 9 | 0x80000000] * (R // 4 + 1))[:R]
        elif kind == "dup":                      # two equal points stored at different indices
            refs = [0, 1] * (R // 2) + [0] * (R & 1)
        else:
            refs = [rnd.randrange(npoints) | (0x80000000 if rnd.random() < 0.4 else 0) for _ in range(R)]
        sorted_refs += refs
        offsets.append(len(sorted_refs))
    return np.array(sorted_refs, dtype=np.uint32), np.array(offsets, dtype=np.uint32)


@pytest.mark.parametrize("group", [1, 2])
def test_batched_affine_accumulation_on_host(b2z, group):
    """csrc/accum_affine.cuh (tree rounds with a shared inversion + XYZZ finish) == the XYZZ running sum, bucket by
    bucket, for ragged runs, runs cut by segment boundaries, doublings and cancellations."""
    L = b2z._ffi.lib()
    codec = b2z.codec
    curve = O.G1 if group == 1 else O.G2
    rnd = random.Random(40 + group)
    npoints = 48 if group == 1 else 24
    pts = [curve.mul(curve.gen, rnd.randrange(1, 1 << 24)) for _ in range(npoints)]
    pts[1] = pts[0]
    limbs = (codec.g1_to_limbs if group == 1 else codec.g2_to_limbs)(pts)[0]
    big = 700 if group == 1 else 260
    runs = [0, 1, 2, 3, 0, 5, 64, 33, big, 1, 1, 1, 130, 0, 0, 97, 40, 2, 31, 64, 50, 7]
    special = {6: "same", 7: "cancel", 12: "mixed", 15: "dup", 18: "same", 19: "cancel"}
    sorted_refs, offsets = _accum_case(rnd, npoints, runs, special)
    nb = len(runs)
    stats = np.zeros(7, dtype=np.uint32)
    total = int(offsets[-1])
    for nseg, cap in ((1, 0), (2, 0), (3, 0), (5, 0), (13, 0), (40, 0), (total // 8, 0)):
        bad = L.b2z_host_accum_affine(group, limbs.ctypes.data, sorted_refs.ctypes.data, offsets.ctypes.data, nb, nseg,
                                      cap, stats.ctypes.data)
        assert bad == 0, (group, nseg, cap, bad)
        _, rounds, batched, doublings, free, abandoned, finish = (int(v) for v in stats)
        if nseg <= 5 and cap == 0:
            # the tree really ran: most additions were batched, doublings and division-free pairs occurred
            assert rounds >= 3 and batched > total // 2 and doublings > 0 and free > 0, tuple(stats)
            assert batched + finish <= total
        if nseg == total // 8:
            assert rounds == 0 and finish > 0          # 8-reference segments go straight to the XYZZ finish
