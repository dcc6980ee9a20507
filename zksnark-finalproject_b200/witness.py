"""Host-side witness generation over the C ABI (include/b200zk.h): native, multithreaded,
no GPU.  Mirrors the reference's native helpers next to its gadgets --

  poseidon_hash            hasher()  (src/arkworks/matrix_proof_of_work/hasher.rs:17-27)
  matrix_circuit_witness   the assignment MatrixCircuit::generate_constraints computes
                           (matrix_proof_of_work/constraints.rs:78-128), in the variable order of circuits.matrix_circuit
  fibonacci_witness        FibonacciCircuit (constraints/fibbonaci.rs:22-48)
  modpow_witnesses         mod_pow_generate_witnesses() (prime_snark/utils/modulo.rs:31-89)
  prime_search             the prove_prime loop over check_if_next_is_prime (backend/prime_snark.rs:60-70,
                           prime_snark/prime_circut.rs:165-195, fermat_circut.rs:131-141), spread over host threads

The Poseidon parameters are arguments (PoseidonParams); nothing here carries the reference's constant table.
"""
import ctypes

import numpy as np

from . import _ffi, codec


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _check(status, what):
    if status != _ffi.B2Z_OK:
        raise _ffi.B2zError(status, what)


class PoseidonParams:
    """ark_crypto_primitives::sponge::poseidon::PoseidonConfig<Fr>: full_rounds, partial_rounds, alpha, ark, mds, rate,
    capacity (values as Python integers mod r)."""

    def __init__(self, full_rounds, partial_rounds, alpha, ark, mds, rate, capacity):
        self.full_rounds, self.partial_rounds, self.alpha = int(full_rounds), int(partial_rounds), int(alpha)
        self.rate, self.capacity = int(rate), int(capacity)
        self.width = self.rate + self.capacity
        self.ark = [list(r) for r in ark]
        self.mds = [list(r) for r in mds]
        if len(self.ark) != self.full_rounds + self.partial_rounds or any(len(r) != self.width for r in self.ark):
            raise ValueError("ark must be (full_rounds + partial_rounds) x width")
        if len(self.mds) != self.width or any(len(r) != self.width for r in self.mds):
            raise ValueError("mds must be width x width")
        self._ark = np.ascontiguousarray(codec.fr_to_mont_limbs([v for r in self.ark for v in r]))
        self._mds = np.ascontiguousarray(codec.fr_to_mont_limbs([v for r in self.mds for v in r]))

    @classmethod
    def from_shape(cls, ps):
        """From circuits.PoseidonShape (seed-derived constants of the same shape as the reference's parameters)."""
        return cls(ps.FULL, ps.PARTIAL, ps.ALPHA, ps.ark, ps.mds, ps.WIDTH - 1, 1)

    def desc(self):
        d = _ffi.PoseidonDesc()
        d.full_rounds, d.partial_rounds, d.alpha = self.full_rounds, self.partial_rounds, self.alpha
        d.width, d.rate, d.capacity = self.width, self.rate, self.capacity
        d.ark, d.mds = _ptr(self._ark), _ptr(self._mds)
        return d


def _default_params():
    from .circuits import PoseidonShape
    return PoseidonParams.from_shape(PoseidonShape())


def _fr_matrix(m):
    """n x n matrix of integers, or an (n*n, 4) uint64 array of Montgomery limbs -> (n, limbs)."""
    if isinstance(m, np.ndarray) and m.dtype == np.uint64 and m.ndim == 2 and m.shape[1] == 4:
        n = int(round(m.shape[0] ** 0.5))
        if n * n != m.shape[0]:
            raise ValueError("limb array is not a square matrix")
        return n, np.ascontiguousarray(m)
    n = len(m)
    if any(len(row) != n for row in m):
        raise ValueError("matrix is not square")
    return n, np.ascontiguousarray(codec.fr_to_mont_limbs([int(v) for row in m for v in row]))


def poseidon_hash(elems, params=None):
    """Digest (integer mod r) of a list of field elements: absorb everything, squeeze one element."""
    params = params or _default_params()
    x = np.ascontiguousarray(codec.fr_to_mont_limbs([int(v) for v in elems])) if len(elems) else np.zeros((0, 4), np.uint64)
    out = np.zeros((1, 4), dtype=np.uint64)
    d = params.desc()
    _check(_ffi.lib().b2z_poseidon_hash(ctypes.byref(d), _ptr(x), len(elems), _ptr(out)), "b2z_poseidon_hash")
    return codec.fr_from_mont_limbs(out)[0]


def matrix_circuit_num_variables(n, params=None):
    params = params or _default_params()
    d = params.desc()
    return int(_ffi.lib().b2z_matrix_circuit_num_variables(ctypes.byref(d), int(n)))


def matrix_circuit_witness(mat_a, mat_b, params=None, threads=0, out=None):
    """-> (m, 4) uint64 array of Montgomery limbs: the full assignment z of circuits.matrix_circuit(mat_a, mat_b)
    (same variable order), ready for Groth16.create_proof_with_matrices.  `out`: a preallocated (page-locked) array."""
    params = params or _default_params()
    n, a = _fr_matrix(mat_a)
    nb, b = _fr_matrix(mat_b)
    if nb != n:
        raise ValueError("A and B differ in size")
    d = params.desc()
    L = _ffi.lib()
    m = int(L.b2z_matrix_circuit_num_variables(ctypes.byref(d), n))
    if m == 0:
        raise ValueError("bad Poseidon parameters or n = 0")
    if out is None:
        out = np.empty((m, 4), dtype=np.uint64)
    elif out.shape != (m, 4) or out.dtype != np.uint64 or not out.flags["C_CONTIGUOUS"]:
        raise ValueError("out must be a contiguous (%d, 4) uint64 array" % m)
    _check(L.b2z_matrix_circuit_witness(ctypes.byref(d), n, _ptr(a), _ptr(b), int(threads), _ptr(out), m),
           "b2z_matrix_circuit_witness")
    return out


def fibonacci_witness(a, b, num_steps):
    """-> (5, 4) uint64 Montgomery limbs: [1, a, b, F(num_steps) | 0]."""
    ab = np.ascontiguousarray(codec.fr_to_mont_limbs([int(a), int(b)]))
    out = np.zeros((5, 4), dtype=np.uint64)
    _check(_ffi.lib().b2z_fibonacci_witness(_ptr(ab[0:1]), _ptr(ab[1:2]), int(num_steps), _ptr(out)), "b2z_fibonacci_witness")
    return out


def modpow_witnesses(base, modulus, exponent, num_bits):
    """mod_pow_generate_witnesses: {'mod_vals': [(num, q, remainder)] * num_bits, 'mod_pow_vals': [...], 'bits': [...],
    'result': base^exponent mod modulus}; base, modulus < 2^63."""
    mv = np.zeros((num_bits, 5), dtype=np.uint64)
    pv = np.zeros((num_bits, 5), dtype=np.uint64)
    bits = np.zeros(num_bits, dtype=np.uint8)
    res = ctypes.c_uint64()
    _check(_ffi.lib().b2z_modpow_witnesses(int(base), int(modulus), int(exponent), int(num_bits), _ptr(mv), _ptr(pv),
                                           _ptr(bits), ctypes.byref(res)), "b2z_modpow_witnesses")
    rows = lambda t: [(int(r[0]) | (int(r[1]) << 64), int(r[2]) | (int(r[3]) << 64), int(r[4])) for r in t]
    return {"mod_vals": rows(mv), "mod_pow_vals": rows(pv), "bits": [int(x) for x in bits], "result": int(res.value)}


def sha256(data):
    """SHA-256 of a byte string through the library (the native side of the prime route hashes with the sha2 crate)."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8).copy() if len(data) else np.zeros(0, np.uint8)
    out = np.zeros(32, dtype=np.uint8)
    _ffi.lib().b2z_sha256(_ptr(buf) if len(data) else None, len(data), _ptr(out))
    return out.tobytes()


def prime_search(x, j_first, j_last, num_bits=20, k_bases=3, threads=0):
    """First j in [j_first, j_last] for which check_if_next_is_prime(Fr::from(x), j) succeeds.
    -> dict(found, j, digest (32 bytes), is_prime, quotient, remainder (the candidate), a); found = False: the entry
    describes j_last."""
    xl = np.ascontiguousarray(codec.fr_to_mont_limbs([int(x)]))
    out = _ffi.PrimeCheck()
    found = ctypes.c_int32()
    _check(_ffi.lib().b2z_prime_search(_ptr(xl), int(j_first), int(j_last), int(num_bits), int(k_bases), int(threads),
                                       ctypes.byref(out), ctypes.byref(found)), "b2z_prime_search")
    big = lambda limbs: sum(int(v) << (64 * i) for i, v in enumerate(limbs))
    return {"found": bool(found.value), "j": int(out.j), "digest": bytes(out.digest), "is_prime": bool(out.is_prime),
            "quotient": big(out.quotient), "remainder": int(out.remainder), "a": big(out.a)}
