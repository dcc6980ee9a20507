"""Key-generation scalars on the host (csrc/setup_host.cu; include/b200zk.h row f3) and the host mirror's native
key-generation path built on them.  Two kinds of checks, both without a GPU:

* every helper against the Python-integer formulas of ark_groth16's generate_parameters_with_qap /
  LibsnarkReduction::instance_map_with_evaluation (reached from Groth16::setup,
  /root/reference/src/arkworks/backend/matrix_proof.rs:128-131) -- bit-exact;
* Groth16.generate_parameters_with_qap run twice over a RECORDING stand-in for the GPU entry points (b2z_spmv_fr
  computed with integers, b2z_fixed_base_mul_g1/g2 recording their scalar arrays): the native path hands the GPU
  exactly the arrays the integer path does, in the same order.
"""
import ctypes
import importlib
import os
import random

import numpy as np
import pytest

from oracle import bls12_381 as O

R = O.R_MOD


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _mont(codec, v):
    return np.ascontiguousarray(codec.fr_to_mont_limbs([v]))


@pytest.mark.parametrize("log_n,count", [(0, 1), (1, 2), (3, 8), (5, 19), (10, 1024), (12, 4096 - 7), (14, 12345), (16, 1 << 16)])
def test_lagrange_coefficients(b2z, log_n, count):
    codec, L = b2z.codec, b2z._ffi.lib()
    rnd = random.Random(log_n)
    tau = rnd.randrange(2, R)
    n = 1 << log_n
    w = pow(O.FR_ROOT_OF_UNITY, 1 << (32 - log_n), R)
    zt = (pow(tau, n, R) - 1) % R
    zn = zt * pow(n, -1, R) % R
    want, cur = [], 1
    for _ in range(count):
        want.append(zn * cur % R * pow((tau - cur) % R, -1, R) % R)
        cur = cur * w % R
    for threads in (1, 3, 0):
        out = np.full((count, 4), 0x5A5A5A5A5A5A5A5A, dtype=np.uint64)
        assert L.b2z_fr_lagrange_at(log_n, _ptr(_mont(codec, tau)), count, threads, _ptr(out)) == 0
        assert codec.fr_from_mont_limbs(out) == want, threads
    if count == n and n > 1:
        assert sum(want) % R == 1                      # the Lagrange basis sums to one


def test_lagrange_rejects_a_domain_point_and_bad_sizes(b2z):
    codec, L, ffi = b2z.codec, b2z._ffi.lib(), b2z._ffi
    out = np.zeros((8, 4), dtype=np.uint64)
    w8 = pow(O.FR_ROOT_OF_UNITY, 1 << 29, R)
    assert L.b2z_fr_lagrange_at(3, _ptr(_mont(codec, pow(w8, 5, R))), 8, 0, _ptr(out)) == ffi.B2Z_EINVAL
    assert L.b2z_fr_lagrange_at(3, _ptr(_mont(codec, 5)), 9, 0, _ptr(out)) == ffi.B2Z_EINVAL        # count > n
    assert L.b2z_fr_lagrange_at(33, _ptr(_mont(codec, 5)), 8, 0, _ptr(out)) == ffi.B2Z_EINVAL
    bad = np.full((1, 4), 2 ** 64 - 1, dtype=np.uint64)
    assert L.b2z_fr_lagrange_at(3, _ptr(bad), 8, 0, _ptr(out)) == ffi.B2Z_EINVAL


@pytest.mark.parametrize("count", [0, 1, 2, 1000, 4096, 50001])
def test_geometric_lincomb_and_into_bigint(b2z, count):
    codec, L = b2z.codec, b2z._ffi.lib()
    rnd = random.Random(count)
    base, scale = rnd.randrange(R), rnd.randrange(R)
    for threads in (1, 4, 0):
        out = np.zeros((count, 4), dtype=np.uint64)
        assert L.b2z_fr_geometric(_ptr(_mont(codec, base)), _ptr(_mont(codec, scale)), count, threads, _ptr(out)) == 0
        want, cur = [], scale
        for _ in range(count):
            want.append(cur)
            cur = cur * base % R
        assert codec.fr_from_mont_limbs(out) == want
    xs, ys, zs = ([rnd.randrange(R) for _ in range(count)] for _ in range(3))
    if count > 3:
        xs[0], ys[1], zs[2], xs[3] = 0, R - 1, 1, R - 1
    X, Y, Z = (np.ascontiguousarray(codec.fr_to_mont_limbs(v)) if count else np.zeros((0, 4), np.uint64) for v in (xs, ys, zs))
    a, b, c = rnd.randrange(R), rnd.randrange(R), 1
    out = np.zeros((count, 4), dtype=np.uint64)
    assert L.b2z_fr_lincomb3(count, _ptr(_mont(codec, a)), _ptr(X), _ptr(_mont(codec, b)), _ptr(Y), _ptr(_mont(codec, c)),
                             _ptr(Z), 0, _ptr(out)) == 0
    assert codec.fr_from_mont_limbs(out) == [(a * x + b * y + z) % R for x, y, z in zip(xs, ys, zs)]
    # one term, in place
    inplace = X.copy()
    assert L.b2z_fr_lincomb3(count, _ptr(_mont(codec, a)), _ptr(inplace), None, None, None, None, 3, _ptr(inplace)) == 0
    assert codec.fr_from_mont_limbs(inplace) == [a * x % R for x in xs]
    # two terms with unit coefficients (the instance adjustment of the A query)
    assert L.b2z_fr_lincomb3(count, _ptr(_mont(codec, 1)), _ptr(X), _ptr(_mont(codec, 1)), _ptr(Y), None, None, 0, _ptr(out)) == 0
    assert codec.fr_from_mont_limbs(out) == [(x + y) % R for x, y in zip(xs, ys)]
    big = np.zeros((count, 4), dtype=np.uint64)
    assert L.b2z_fr_into_bigint(count, _ptr(X), 0, _ptr(big)) == 0
    assert np.array_equal(big, codec.fr_to_bigint_limbs(xs) if count else big)
    # lazily reduced vector elements (value + r < 2^256) are accepted and reduced
    if count:
        lazy = codec._pack([x * codec.FR_R % R + R for x in xs], 32)
        assert L.b2z_fr_into_bigint(count, _ptr(lazy), 0, _ptr(big)) == 0
        assert np.array_equal(big, codec.fr_to_bigint_limbs(xs))


# ---- the whole key-generation path over a recording stand-in for the GPU entry points
class _FakeLib:
    """b2z_spmv_fr with Python integers; b2z_fixed_base_mul_g1/g2 record their scalars and return a digest of them in
    place of points (so that the keys of two runs can be compared as well); the host helpers are the real library."""

    def __init__(self, real, codec):
        self.real, self.codec, self.calls = real, codec, []

    def __getattr__(self, name):
        return getattr(self.real, name)

    @staticmethod
    def _arr(p, shape, dtype=np.uint64):
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dtype=dtype)
        ct = {np.uint64: ctypes.c_uint64, np.uint32: ctypes.c_uint32, np.uint8: ctypes.c_uint8}[dtype]
        return np.ctypeslib.as_array((ct * n).from_address(p.value)).reshape(shape)

    def b2z_spmv_fr(self, handle, nrows, ncols, rp, cols, cf, x, y):
        rp = self._arr(rp, (nrows + 1,))
        nnz = int(rp[-1])
        cols = self._arr(cols, (nnz,), np.uint32)
        cf = self.codec.fr_from_mont_limbs(self._arr(cf, (nnz, 4)))
        xv = self.codec.fr_from_mont_limbs(self._arr(x, (ncols, 4)))
        self.calls.append(("spmv", self._arr(x, (ncols, 4)).copy(), rp.copy(), cols.copy()))
        out = []
        for r in range(nrows):
            acc = 0
            for k in range(int(rp[r]), int(rp[r + 1])):
                acc += cf[k] * xv[int(cols[k])]
            out.append(acc % R)
        self._arr(y, (nrows, 4))[:] = self.codec.fr_to_mont_limbs(out)
        return 0

    def _fixed(self, tag, scalars, n, out_points, out_inf, limbs):
        s = self._arr(scalars, (n, 4)).copy()
        self.calls.append((tag, s))
        pts = self._arr(out_points, (n, limbs))
        pts[:] = 0
        pts[:, :4] = s                                                # "point" = its scalar: keys become comparable
        self._arr(out_inf, ((n + 7) // 8,), np.uint8)[:] = 0
        return 0

    def b2z_fixed_base_mul_g1(self, handle, scalars, n, out_points, out_inf):
        return self._fixed("g1", scalars, n, out_points, out_inf, 12)

    def b2z_fixed_base_mul_g2(self, handle, scalars, n, out_points, out_inf):
        return self._fixed("g2", scalars, n, out_points, out_inf, 24)


class _FakeCtx:
    handle = None

    def __init__(self, lib):
        self._lib = lib

    def check(self, st):
        assert st == 0, st


def _keys_equal(b2z, pk1, pk2, vk1, vk2):
    for f in b2z.ProvingKey.FIELDS:
        assert np.array_equal(getattr(pk1, f)[0], getattr(pk2, f)[0]), f
    for f in ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2"):
        assert np.array_equal(getattr(pk1, f), getattr(pk2, f)), f
    assert np.array_equal(vk1.gamma_abc_g1[0], vk2.gamma_abc_g1[0])


@pytest.mark.parametrize("case", ["matrix2", "matrix5", "fibonacci", "prime"])
def test_native_key_generation_feeds_the_gpu_the_same_arrays(b2z, case):
    circuits = importlib.import_module("zksnark-finalproject_b200.circuits")
    fast = importlib.import_module("zksnark-finalproject_b200.circuits_fast")
    if case.startswith("matrix"):
        n = int(case[6:])
        rnd = random.Random(n)
        cm, _ = fast.matrix_circuit_fast([[rnd.randrange(1 << 64) for _ in range(n)] for _ in range(n)],
                                         [[rnd.randrange(1 << 64) for _ in range(n)] for _ in range(n)])
    else:
        inst = circuits.fibonacci_circuit(0, 1, 10) if case == "fibonacci" else \
            circuits.prime_circuit(5, num_bits=6, k_bases=1, sha_blocks=1)
        cm = b2z.ConstraintMatrices.from_rows(inst.num_instance, inst.num_witness, inst.a, inst.b, inst.c)
    toxic = [random.Random(7).randrange(1, R) for _ in range(5)]
    runs = []
    for python_path in (True, False):
        lib = _FakeLib(b2z._ffi.lib(), b2z.codec)
        ctx = _FakeCtx(lib)
        if python_path:
            os.environ["B2Z_SETUP_PYTHON"] = "1"
        else:
            os.environ.pop("B2Z_SETUP_PYTHON", None)
        try:
            pk, vk = b2z.Groth16.generate_parameters_with_qap(ctx, cm, cm.num_constraints, cm.num_instance_variables,
                                                              cm.num_variables, *toxic)
        finally:
            os.environ.pop("B2Z_SETUP_PYTHON", None)
        runs.append((lib.calls, pk, vk))
    (c1, pk1, vk1), (c2, pk2, vk2) = runs
    assert [c[0] for c in c1] == [c[0] for c in c2] and len(c1) == 3 + 2 + 6
    for a, b in zip(c1, c2):
        for x, y in zip(a[1:], b[1:]):
            assert x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y), a[0]
    _keys_equal(b2z, pk1, pk2, vk1, vk2)


def test_native_scalar_pipeline_at_2p20(b2z):
    """Thread partitioning at a BASELINE-class size: Lagrange vector and h-query scalars of a 2^20 domain against a
    strided sample of the integer formulas plus global identities."""
    codec, L = b2z.codec, b2z._ffi.lib()
    log_n, n = 20, 1 << 20
    tau, scale = 0x1234567890ABCDEF1234567890ABCDEF % R, 0xFEDCBA9876543210 % R
    lag = np.zeros((n, 4), dtype=np.uint64)
    assert L.b2z_fr_lagrange_at(log_n, _ptr(_mont(codec, tau)), n, 0, _ptr(lag)) == 0
    w = pow(O.FR_ROOT_OF_UNITY, 1 << (32 - log_n), R)
    zn = (pow(tau, n, R) - 1) * pow(n, -1, R) % R
    idx = list(range(0, n, 4099)) + [n - 1, n // 2, n // 16 - 1, n // 16, n // 16 + 1]
    got = codec.fr_from_mont_limbs(lag[idx])
    for i, g in zip(idx, got):
        wi = pow(w, i, R)
        assert g == zn * wi % R * pow((tau - wi) % R, -1, R) % R, i
    one_thread = np.zeros((n, 4), dtype=np.uint64)
    assert L.b2z_fr_lagrange_at(log_n, _ptr(_mont(codec, tau)), n, 1, _ptr(one_thread)) == 0
    assert np.array_equal(lag, one_thread)
    hs = np.zeros((n - 1, 4), dtype=np.uint64)
    assert L.b2z_fr_geometric(_ptr(_mont(codec, tau)), _ptr(_mont(codec, scale)), n - 1, 0, _ptr(hs)) == 0
    got = codec.fr_from_mont_limbs(hs[[i for i in idx if i < n - 1]])
    for i, g in zip([i for i in idx if i < n - 1], got):
        assert g == scale * pow(tau, i, R) % R, i
