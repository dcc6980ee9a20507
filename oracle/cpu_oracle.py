"""ctypes front-end of oracle/cpu/libark_cpu.so -- TEST INFRASTRUCTURE ONLY
(see the header of oracle/cpu/ark_cpu.cpp).  Array conventions are those of
include/b200zk.h: uint64 limb arrays, Fr/Fq in Montgomery form, MSM scalars
canonical."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "cpu", "libark_cpu.so")
_lib = None
vp = ctypes.c_void_p


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.ark_cpu_set_threads.argtypes = [ctypes.c_int]
        L.ark_cpu_hardware_threads.restype = ctypes.c_int
        L.ark_cpu_ntt.argtypes = [vp, ctypes.c_uint32, ctypes.c_int, vp]
        L.ark_cpu_witness_map.argtypes = [vp, vp, vp, ctypes.c_uint32, vp]
        L.ark_cpu_msm_g1.argtypes = [vp, vp, vp, ctypes.c_uint64, vp]
        L.ark_cpu_msm_g1.restype = ctypes.c_int
        L.ark_cpu_msm_g2.argtypes = [vp, vp, vp, ctypes.c_uint64, vp]
        L.ark_cpu_msm_g2.restype = ctypes.c_int
        L.ark_cpu_pk_new.restype = vp
        L.ark_cpu_pk_new.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32] + [vp] * 15
        L.ark_cpu_pk_free.argtypes = [vp]
        L.ark_cpu_prove.argtypes = [vp] * 8
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(vp) if a is not None else None


def _c(a, dt=np.uint64):
    return np.ascontiguousarray(a, dtype=dt) if a is not None else None


def hardware_threads():
    return lib().ark_cpu_hardware_threads()


def set_threads(t):
    lib().ark_cpu_set_threads(int(t))


def ntt(data, inverse=False, coset_gen=None):
    a = _c(data).copy()
    n = a.shape[0]
    g = _c(coset_gen)
    lib().ark_cpu_ntt(_p(a), n.bit_length() - 1, int(inverse), _p(g))
    return a


def witness_map(a, b, c):
    a, b, c = _c(a).copy(), _c(b).copy(), _c(c).copy()
    n = a.shape[0]
    h = np.empty_like(a)
    lib().ark_cpu_witness_map(_p(a), _p(b), _p(c), n.bit_length() - 1, _p(h))
    return h


def msm_g1(bases, scalars, inf=None):
    bases, scalars, inf = _c(bases), _c(scalars), _c(inf, np.uint8)
    n = min(bases.shape[0], scalars.shape[0])
    out = np.zeros(12, dtype=np.uint64)
    is_inf = lib().ark_cpu_msm_g1(_p(bases), _p(inf), _p(scalars), n, _p(out))
    return out, bool(is_inf)


def msm_g2(bases, scalars, inf=None):
    bases, scalars, inf = _c(bases), _c(scalars), _c(inf, np.uint8)
    n = min(bases.shape[0], scalars.shape[0])
    out = np.zeros(24, dtype=np.uint64)
    is_inf = lib().ark_cpu_msm_g2(_p(bases), _p(inf), _p(scalars), n, _p(out))
    return out, bool(is_inf)


class CpuProvingKey:
    """Holds an ark_cpu_pk built from the same limb arrays a b2z_pk_desc takes."""

    def __init__(self, num_variables, num_instance, domain_size, a, b1, b2, h, l, alpha, beta1, delta1, beta2, delta2):
        self._keep = []

        def q(pair):
            arr, inf = _c(pair[0]), _c(pair[1], np.uint8)
            self._keep += [arr, inf]
            return _p(arr), _p(inf)

        def one(x):
            x = _c(x)
            self._keep.append(x)
            return _p(x)

        args = [num_variables, num_instance, domain_size.bit_length() - 1]
        for pair in (a, b1, b2, h, l):
            args += list(q(pair))
        args += [one(alpha), one(beta1), one(delta1), one(beta2), one(delta2)]
        self.domain_size = domain_size
        self.handle = lib().ark_cpu_pk_new(*args)

    def prove(self, a, b, c, z, r_mont, s_mont):
        a, b, c, z = _c(a).copy(), _c(b).copy(), _c(c).copy(), _c(z)
        r_mont, s_mont = _c(r_mont), _c(s_mont)
        out = np.zeros(192, dtype=np.uint8)
        lib().ark_cpu_prove(self.handle, _p(a), _p(b), _p(c), _p(z), _p(r_mont), _p(s_mont), _p(out))
        return out.tobytes()

    def __del__(self):
        try:
            if self.handle:
                lib().ark_cpu_pk_free(self.handle)
                self.handle = None
        except Exception:
            pass
