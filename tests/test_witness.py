"""Native, multithreaded witness generation (csrc/witness.cu; include/b200zk.h row f5) against independent Python
restatements: oracle/witness.py (Poseidon sponge with arkworks' PoseidonSponge structure, mod_pow_generate_witnesses,
check_if_next_is_prime on hashlib + Python integers), the symbolic circuit builders (big-int arithmetic) and the
committed vectors tests/golden/witness_vectors.json
(/root/reference/src/arkworks/matrix_proof_of_work/hasher.rs:17-27, constraints.rs:78-128,
prime_snark/utils/modulo.rs:31-89, prime_snark/prime_circut.rs:149-195, constraints/fibbonaci.rs:22-48).
Host only: no GPU needed."""
import hashlib
import importlib
import json
import os
import random

import numpy as np
import pytest

from oracle import bls12_381 as O
from oracle import witness as OW

R = O.R_MOD
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def W(b2z):
    return b2z.witness


@pytest.fixture(scope="module")
def circuits():
    return importlib.import_module("zksnark-finalproject_b200.circuits")


@pytest.fixture(scope="module")
def fast():
    return importlib.import_module("zksnark-finalproject_b200.circuits_fast")


def _py_hash(p, elems):
    """oracle/witness.py on the parameters of a witness.PoseidonParams"""
    return OW.poseidon_hash(OW.PoseidonConfig(p.full_rounds, p.partial_rounds, p.alpha, p.ark, p.mds, p.rate, p.capacity),
                            elems)


def _params(W, rnd, full, partial, alpha, rate, capacity):
    w = rate + capacity
    return W.PoseidonParams(full, partial, alpha, [[rnd.randrange(R) for _ in range(w)] for _ in range(full + partial)],
                            [[rnd.randrange(R) for _ in range(w)] for _ in range(w)], rate, capacity)


@pytest.mark.parametrize("full,partial,alpha,rate,capacity", [(8, 29, 17, 2, 1), (8, 31, 5, 2, 1), (4, 3, 5, 3, 1),
                                                              (2, 0, 3, 1, 1), (8, 5, 257, 4, 2), (6, 7, 7, 5, 3)])
def test_poseidon_hash_against_python_sponge(W, full, partial, alpha, rate, capacity):
    rnd = random.Random(alpha * 1000 + rate)
    p = _params(W, rnd, full, partial, alpha, rate, capacity)
    for count in (0, 1, rate - 1, rate, rate + 1, 2 * rate, 2 * rate + 1, 37):
        elems = [rnd.randrange(R) for _ in range(count)]
        assert W.poseidon_hash(elems, p) == _py_hash(p, elems), count
    edge = [0, 1, R - 1, R - 2, 2 ** 64, 2 ** 255 % R]
    assert W.poseidon_hash(edge, p) == _py_hash(p, edge)


def test_poseidon_hash_equals_the_circuit_builders_digest(W, circuits):
    ps = circuits.PoseidonShape()
    rnd = random.Random(11)
    for count in (1, 2, 3, 8, 9):
        elems = [rnd.randrange(R) for _ in range(count)]
        cs = circuits.ConstraintSystem()
        want = cs.value(ps.hash(cs, [cs.new_witness(v) for v in elems]))
        assert W.poseidon_hash(elems) == want


@pytest.mark.parametrize("n", [1, 2, 3, 5])
def test_matrix_witness_equals_symbolic_builder_and_satisfies(W, b2z, circuits, n):
    rnd = random.Random(n)
    a = [[rnd.randrange(1 << 64) for _ in range(n)] for _ in range(n)]
    b = [[rnd.randrange(1 << 64) for _ in range(n)] for _ in range(n)]
    inst = circuits.matrix_circuit(a, b)
    want = b2z.codec.fr_to_mont_limbs(inst.z)
    assert W.matrix_circuit_num_variables(n) == inst.num_variables == len(inst.z)
    for threads in (1, 2, 3, 4, 7, 0):
        got = W.matrix_circuit_witness(a, b, threads=threads)
        assert np.array_equal(got, want), threads
    # and the native assignment satisfies the builder's constraint system
    inst.z = b2z.codec.fr_from_mont_limbs(W.matrix_circuit_witness(a, b))
    assert inst.is_satisfied()
    inst.z[len(inst.z) // 2] = (inst.z[len(inst.z) // 2] + 1) % R
    assert not inst.is_satisfied()


@pytest.mark.parametrize("n,entries", [(8, "ones"), (16, "ones"), (16, "u64"), (12, "field")])
def test_matrix_witness_equals_vectorised_builder(W, b2z, fast, n, entries):
    """The reference's bench posts all-ones matrices (bench/matrix.py:10-11); its tests use u64 entries
    (constraints.rs:326-330); full-size field elements exercise the reduction paths."""
    rnd = random.Random(n)
    gen = {"ones": lambda: 1, "u64": lambda: rnd.randrange(1 << 64), "field": lambda: rnd.randrange(R)}[entries]
    a = [[gen() for _ in range(n)] for _ in range(n)]
    b = [[gen() for _ in range(n)] for _ in range(n)]
    cm, z = fast.matrix_circuit_fast(a, b)
    want = b2z.codec.fr_to_mont_limbs(z)
    for threads in (1, 3, 6, 0):
        out = np.full(want.shape, 0xA5A5A5A5A5A5A5A5, dtype=np.uint64)       # every element must be written
        got = W.matrix_circuit_witness(a, b, threads=threads, out=out)
        assert got is out and np.array_equal(got, want), threads
    # limb-array inputs (what a caller that already holds Fr elements passes)
    la = b2z.codec.fr_to_mont_limbs([v for row in a for v in row])
    lb = b2z.codec.fr_to_mont_limbs([v for row in b for v in row])
    assert np.array_equal(W.matrix_circuit_witness(la, lb), want)


def test_matrix_witness_with_other_parameters(W):
    """Layout with another S-box / sponge geometry: x^5 (3 witnesses per S-box), rate 3."""
    rnd = random.Random(99)
    p = _params(W, rnd, 4, 5, 5, 3, 1)
    n = 3
    a = [[rnd.randrange(R) for _ in range(n)] for _ in range(n)]
    b = [[rnd.randrange(R) for _ in range(n)] for _ in range(n)]
    z = W.matrix_circuit_witness(a, b, p)
    N, T, pv = n * n, 3, (4 * 4 + 5) * 3
    assert z.shape[0] == 4 + 2 * N + 2 * T * pv + N + N * (1 + n) + T * pv == W.matrix_circuit_num_variables(n, p)
    codec = importlib.import_module("zksnark-finalproject_b200.codec")
    v = codec.fr_from_mont_limbs(z)
    flat = lambda m: [x for row in m for x in row]
    c = [sum(a[i][k] * b[k][j] for k in range(n)) % R for i in range(n) for j in range(n)]
    assert v[0] == 1 and v[4:4 + N] == flat(a) and v[4 + N:4 + 2 * N] == flat(b)
    assert v[1:4] == [_py_hash(p, flat(a)), _py_hash(p, flat(b)), _py_hash(p, c)]
    w_mm = 4 + 2 * N + 2 * T * pv + N
    assert v[w_mm - N:w_mm] == [0] * N
    for i in range(n):
        for j in range(n):
            cell = v[w_mm + (i * n + j) * (1 + n):w_mm + (i * n + j + 1) * (1 + n)]
            assert cell == [0] + [a[i][k] * b[k][j] % R for k in range(n)]
    # first S-box of the digest of A: state = (0, a00, a01, a02) + ark[0], then x^2, x^4, x^5 of state[0]
    x = p.ark[0][0] % R
    assert v[4 + 2 * N:4 + 2 * N + 3] == [x * x % R, pow(x, 4, R), pow(x, 5, R)]


def test_fibonacci_witness(W, b2z, circuits):
    for a, b, steps in ((0, 1, 10), (0, 1, 1000), (3, 4, 0), (R - 1, R - 2, 5), (7, 11, 1)):
        inst = circuits.fibonacci_circuit(a, b, steps)
        assert np.array_equal(W.fibonacci_witness(a, b, steps), b2z.codec.fr_to_mont_limbs(inst.z)), (a, b, steps)


_py_modpow_witnesses = OW.mod_pow_generate_witnesses


def test_modpow_witnesses(W):
    rnd = random.Random(3)
    cases = [(2, 1000003, 1000002, 20), (5, 7, 6, 20), (3, 2 ** 61 - 1, 2 ** 61 - 2, 64), (2, 3, 0, 8), (10, 11, 1, 1)]
    for _ in range(40):
        bits = rnd.choice([8, 20, 32, 63])
        mod = rnd.randrange(2, 1 << rnd.choice([8, 20, 40, 62]))
        cases.append((rnd.randrange(1, 1 << 62), mod, rnd.randrange(1 << bits), bits + rnd.randrange(3)))
    for base, mod, exp, nb in cases:
        got = W.modpow_witnesses(base, mod, exp, nb)
        assert got == _py_modpow_witnesses(base, mod, exp, nb), (base, mod, exp, nb)
        if got["result"] < mod:
            assert got["result"] == pow(base, exp, mod)
    # Fermat test of the reference's seed x = 5 -> candidate bases: a^(n-1) mod n == 1 for a prime n
    assert W.modpow_witnesses(2, 1048583, 1048582, 21)["result"] == 1


def test_witness_entry_points_reject_bad_arguments(W, b2z):
    ffi = b2z._ffi
    p = W._default_params()
    with pytest.raises(ffi.B2zError) as e:                      # entry >= r is not a field element
        W.matrix_circuit_witness(np.full((4, 4), 2 ** 64 - 1, dtype=np.uint64), np.zeros((4, 4), dtype=np.uint64), p)
    assert e.value.status == ffi.B2Z_EINVAL
    a = b2z.codec.fr_to_mont_limbs([1, 2, 3, 4])
    d = p.desc()
    L = ffi.lib()
    m = W.matrix_circuit_num_variables(2, p)
    small = np.zeros((m - 1, 4), dtype=np.uint64)
    assert L.b2z_matrix_circuit_witness(d, 2, a.ctypes.data, a.ctypes.data, 0, small.ctypes.data, m - 1) == ffi.B2Z_ESIZE
    assert L.b2z_matrix_circuit_witness(d, 0, a.ctypes.data, a.ctypes.data, 0, small.ctypes.data, m) == ffi.B2Z_EINVAL
    assert L.b2z_matrix_circuit_witness(d, 2, None, a.ctypes.data, 0, small.ctypes.data, m) == ffi.B2Z_EINVAL
    bad = W.PoseidonParams(8, 29, 17, p.ark, p.mds, 2, 1)
    bad.full_rounds = 7                                           # odd number of full rounds
    with pytest.raises(ffi.B2zError):
        W.poseidon_hash([1, 2], bad)
    assert W.matrix_circuit_num_variables(2, bad) == 0
    with pytest.raises(ffi.B2zError):
        W.modpow_witnesses(2, 1, 5, 8)                            # modulus < 2
    with pytest.raises(ffi.B2zError):
        W.modpow_witnesses(2, 7, 256, 8)                          # exponent does not fit num_bits
    with pytest.raises(ffi.B2zError):
        W.modpow_witnesses(2, 1 << 63, 3, 8)                      # modulus >= 2^63


_py_check_if_next_is_prime = OW.check_if_next_is_prime


def test_sha256_known_answers(W):
    assert W.sha256(b"").hex() == "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"
    assert W.sha256(b"abc").hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    assert (W.sha256(b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq").hex()
            == "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1")
    rnd = random.Random(1)
    for n in (1, 55, 56, 57, 63, 64, 65, 119, 120, 128, 1000):
        data = bytes(rnd.randrange(256) for _ in range(n))
        assert W.sha256(data) == hashlib.sha256(data).digest(), n


@pytest.mark.parametrize("x", [5, 0, 123456789, 2 ** 64 - 1])          # x = 5: the reference's own test seed
def test_prime_search_follows_the_reference_loop(W, x):
    # every j on its own (single-j ranges), against the Python restatement
    first = None
    for j in range(0, 40):
        want = _py_check_if_next_is_prime(x, j)
        got = W.prime_search(x, j, j)
        assert {k: got[k] for k in want} == want and got["j"] == j and got["found"] == want["is_prime"], j
        if want["is_prime"] and first is None:
            first = j
    # the loop of prove_prime: first success in 0..=i, whatever the thread count
    for threads in (1, 2, 3, 8, 0):
        got = W.prime_search(x, 0, 39, threads=threads)
        if first is None:
            assert not got["found"] and got["j"] == 39
        else:
            assert got["found"] and got["j"] == first and got["is_prime"]
            assert got["remainder"] == _py_check_if_next_is_prime(x, first)["remainder"]
    if first is not None and first > 0:                            # a range that ends before the first hit
        got = W.prime_search(x, 0, first - 1, threads=4)
        assert not got["found"] and got["j"] == first - 1
        assert got["digest"] == _py_check_if_next_is_prime(x, first - 1)["digest"]


def test_prime_search_other_widths_and_bad_arguments(W, b2z):
    rnd = random.Random(8)
    for _ in range(20):
        x, j, nb, k = rnd.randrange(R), rnd.randrange(1 << 40), rnd.choice([1, 8, 20, 31, 32, 33, 62]), rnd.choice([1, 3, 5])
        want = _py_check_if_next_is_prime(x, j, nb, k)
        got = W.prime_search(x, j, j, num_bits=nb, k_bases=k)
        assert {kk: got[kk] for kk in want} == want, (x, j, nb, k)
    for bad in (dict(num_bits=0), dict(num_bits=63), dict(k_bases=0)):
        with pytest.raises(b2z._ffi.B2zError):
            W.prime_search(5, 0, 3, **bad)
    with pytest.raises(b2z._ffi.B2zError):
        W.prime_search(5, 4, 3)


def test_golden_witness_vectors(W):
    """tests/golden/witness_vectors.json (made by make_witness_golden.py from oracle/witness.py alone)."""
    with open(os.path.join(ROOT, "tests", "golden", "witness_vectors.json")) as f:
        g = json.load(f)
    assert g["producer"] == "oracle"
    for c in g["prime_search"]:
        got = W.prime_search(c["x"], 0, c["i"])
        assert got["found"] and got["j"] == c["j"] and got["digest"].hex() == c["digest"]
        assert got["remainder"] == c["remainder"] and got["quotient"] == int(c["quotient"], 16) and got["a"] == int(c["a"], 16)
    for c in g["modpow"]:
        got = W.modpow_witnesses(c["base"], c["modulus"], c["exponent"], c["num_bits"])
        assert got["result"] == c["result"] and got["bits"] == c["bits"]
        assert got["mod_vals"] == [tuple(int(v, 16) for v in row) for row in c["mod_vals"]]
        assert got["mod_pow_vals"] == [tuple(int(v, 16) for v in row) for row in c["mod_pow_vals"]]
    for c in g["poseidon"]:
        assert W.poseidon_hash([int(e, 16) for e in c["elems"]]) == int(c["digest"], 16)
