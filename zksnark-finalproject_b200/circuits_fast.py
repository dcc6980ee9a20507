"""Vectorised builder for LARGE matrix-multiplication circuits (BASELINE config 5: 64x64,
2 152 451 constraints, domain 2^22).

Produces exactly the system `circuits.matrix_circuit` builds -- same variable numbering, same
constraint order, same coefficients (tests compare the two row by row at small n) -- but as CSR
arrays assembled with numpy: the 265 constraint rows of a Poseidon permutation are extracted ONCE
from the symbolic builder as a template and tiled over all permutations; only the six rows per
permutation that touch the incoming sponge state, and the witness values, are computed per
permutation.  Shape source: src/arkworks/matrix_proof_of_work/constraints.rs:78-128,
hasher.rs:30-40 (see circuits.py).
"""
import numpy as np

from . import codec
from .circuits import ConstraintSystem, PoseidonShape, R_MOD
from .groth16 import ConstraintMatrices

PERM_VARS = 265          # 53 S-boxes x 5 witnesses
PERM_ROWS = 265


class _Template:
    """Rows of one permutation in terms of: const (kind 0), placeholder state P0..P2 (kind 1),
    internal variable k (kind 2)."""

    def __init__(self, ps):
        cs = ConstraintSystem()
        ph = [cs.new_witness(0) for _ in range(3)]
        out = ps.permute(cs, ph)
        assert len(cs.a) == PERM_ROWS and len(cs.witness) == 3 + PERM_VARS

        def conv(lc):
            terms = []
            for (kind, k), v in lc.items():
                if kind == "i":
                    terms.append((0, 0, v))
                elif k < 3:
                    terms.append((1, k, v))
                else:
                    terms.append((2, k - 3, v))
            return terms

        self.rows = [[conv(m[i]) for i in range(PERM_ROWS)] for m in (cs.a, cs.b, cs.c)]
        self.out = [conv(lc) for lc in out]           # three LCs over internal variables
        # rows (matrix, index) that mention a placeholder
        self.patch = sorted({(mi, i) for mi in range(3) for i in range(PERM_ROWS)
                             if any(t[0] == 1 for t in self.rows[mi][i])})


def _native_permute(ps, state):
    """Numeric run of PoseidonShape.permute; returns (new_state, the 265 witness values in allocation order)."""
    wit = []
    half = ps.FULL // 2
    for r in range(ps.FULL + ps.PARTIAL):
        state = [(s + k) % R_MOD for s, k in zip(state, ps.ark[r])]
        full = r < half or r >= half + ps.PARTIAL
        for i in range(3 if full else 1):
            x = state[i]
            x2 = x * x % R_MOD
            x4 = x2 * x2 % R_MOD
            x8 = x4 * x4 % R_MOD
            x16 = x8 * x8 % R_MOD
            x17 = x16 * x % R_MOD
            wit += [x2, x4, x8, x16, x17]
            state[i] = x17
        state = [sum(ps.mds[i][j] * state[j] for j in range(3)) % R_MOD for i in range(3)]
    return state, wit


class _Coo:
    """COO accumulator with a dictionary of distinct coefficients."""

    def __init__(self):
        self.rows, self.cols, self.cidx = [], [], []
        self.table, self.index = [], {}

    def coef(self, v):
        v %= R_MOD
        i = self.index.get(v)
        if i is None:
            i = len(self.table)
            self.table.append(v)
            self.index[v] = i
        return i

    def add_arrays(self, rows, cols, cidx):
        self.rows.append(np.asarray(rows, dtype=np.int64))
        self.cols.append(np.asarray(cols, dtype=np.int64))
        self.cidx.append(np.asarray(cidx, dtype=np.int64))

    def csr(self, num_rows):
        if self.rows:
            rows = np.concatenate(self.rows)
            cols = np.concatenate(self.cols)
            cidx = np.concatenate(self.cidx)
        else:
            rows = cols = cidx = np.zeros(0, dtype=np.int64)
        order = np.argsort(rows, kind="stable")
        rows, cols, cidx = rows[order], cols[order], cidx[order]
        rp = np.zeros(num_rows + 1, dtype=np.uint64)
        np.add.at(rp, rows + 1, 1)
        rp = np.cumsum(rp).astype(np.uint64)
        table = codec.fr_to_mont_limbs(self.table) if self.table else np.zeros((0, 4), np.uint64)
        return rp, cols.astype(np.uint32), table[cidx] if len(cidx) else np.zeros((0, 4), np.uint64)


def matrix_circuit_fast(mat_a, mat_b, poseidon=None):
    """-> (ConstraintMatrices, z as a list of ints).  Same system as circuits.matrix_circuit."""
    n = len(mat_a)
    ps = poseidon or PoseidonShape()
    tpl = _Template(ps)
    N = n * n
    T = (N + 1) // 2                                   # permutations per hash
    mat_c = [[sum(mat_a[i][k] * mat_b[k][j] for k in range(n)) % R_MOD for j in range(n)] for i in range(n)]

    # ---- variable layout (instance first): [1, hash_a, hash_b, hash_c | witnesses]
    L = 4
    w_a = L                                            # A entries
    w_b = w_a + N                                      # B entries
    w_ha = w_b + N                                     # hash(A) internals
    w_hb = w_ha + T * PERM_VARS
    w_cph = w_hb + T * PERM_VARS                       # n^2 placeholder witnesses for C
    w_mm = w_cph + N                                   # per (i, j): sum witness + n products
    w_hc = w_mm + N * (1 + n)
    num_vars = w_hc + T * PERM_VARS
    z = [0] * num_vars
    z[0] = 1
    flat = lambda m: [v % R_MOD for row in m for v in row]
    z[w_a:w_a + N] = flat(mat_a)
    z[w_b:w_b + N] = flat(mat_b)
    prod_col = lambda i, j, k: w_mm + (i * n + j) * (1 + n) + 1 + k
    for i in range(n):
        for j in range(n):
            for k in range(n):
                z[prod_col(i, j, k)] = mat_a[i][k] * mat_b[k][j] % R_MOD

    # ---- constraint layout
    r_ha = 0
    r_eq_a = r_ha + T * PERM_ROWS
    r_hb = r_eq_a + 1
    r_eq_b = r_hb + T * PERM_ROWS
    r_mm = r_eq_b + 1
    r_hc = r_mm + 2 * n ** 3
    r_eq_c = r_hc + T * PERM_ROWS
    num_rows = r_eq_c + 1

    coo = [_Coo(), _Coo(), _Coo()]

    # generic template rows, tiled: arrays per matrix of (row_in_perm, kind, k, coef)
    generic = []
    for mi in range(3):
        rr, kk, kind, cc = [], [], [], []
        for i in range(PERM_ROWS):
            if (mi, i) in tpl.patch:
                continue
            for knd, k, v in tpl.rows[mi][i]:
                rr.append(i); kind.append(knd); kk.append(k); cc.append(coo[mi].coef(v))
        generic.append((np.array(rr, np.int64), np.array(kind, np.int64), np.array(kk, np.int64), np.array(cc, np.int64)))

    def sponge(elems, elem_vals, row_base, var_base):
        """elems: per absorbed element a list of (coef, column); returns (digest LC terms, digest value)."""
        # tiled generic rows
        t_idx = np.arange(T, dtype=np.int64)
        for mi in range(3):
            rr, kind, kk, cc = generic[mi]
            if len(rr) == 0:
                continue
            rows = (row_base + t_idx[:, None] * PERM_ROWS + rr[None, :]).reshape(-1)
            cols = np.where(kind[None, :] == 2, var_base + t_idx[:, None] * PERM_VARS + kk[None, :], 0).reshape(-1)
            coo[mi].add_arrays(rows, cols, np.broadcast_to(cc[None, :], (T, len(cc))).reshape(-1))
        # per-permutation: incoming state as LC (for the patch rows) and as value (for the witnesses)
        state_lc = [[], [], []]
        state_val = [0, 0, 0]
        e = 0
        for t in range(T):
            take = min(2, len(elems) - e)
            for p in range(take):
                state_lc[1 + p] = state_lc[1 + p] + list(elems[e + p])
                state_val[1 + p] = (state_val[1 + p] + elem_vals[e + p]) % R_MOD
            e += take
            base_v = var_base + t * PERM_VARS
            for mi, i in tpl.patch:
                acc = {}
                for knd, k, v in tpl.rows[mi][i]:
                    if knd == 0:
                        acc[0] = (acc.get(0, 0) + v) % R_MOD
                    elif knd == 2:
                        col = base_v + k
                        acc[col] = (acc.get(col, 0) + v) % R_MOD
                    else:
                        for cf, col in state_lc[k]:
                            acc[col] = (acc.get(col, 0) + v * cf) % R_MOD
                items = [(col, v) for col, v in acc.items() if v]
                if items:
                    coo[mi].add_arrays([row_base + t * PERM_ROWS + i] * len(items), [c for c, _ in items],
                                       [coo[mi].coef(v) for _, v in items])
            new_val, wit = _native_permute(ps, list(state_val))
            z[base_v:base_v + PERM_VARS] = wit
            state_lc = [[(v, base_v + k) for knd, k, v in tpl.out[i]] for i in range(3)]
            for i in range(3):
                assert all(knd == 2 for knd, _, _ in tpl.out[i])
            state_val = new_val
        return state_lc[1], state_val[1]

    def enforce_equal(row, lc_terms, pub_col):
        items = {}
        for cf, col in lc_terms:
            items[col] = (items.get(col, 0) + cf) % R_MOD
        items[pub_col] = (items.get(pub_col, 0) - 1) % R_MOD
        its = [(c, v) for c, v in items.items() if v]
        coo[0].add_arrays([row] * len(its), [c for c, _ in its], [coo[0].coef(v) for _, v in its])
        coo[1].add_arrays([row], [0], [coo[1].coef(1)])

    single = lambda base: [[(1, base + i)] for i in range(N)]
    dig_a, val_a = sponge(single(w_a), flat(mat_a), r_ha, w_ha)
    enforce_equal(r_eq_a, dig_a, 1)
    dig_b, val_b = sponge(single(w_b), flat(mat_b), r_hb, w_hb)
    enforce_equal(r_eq_b, dig_b, 2)
    # matrix multiplication: two identical rows per scalar product
    ii, jj, kk = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    ii, jj, kk = ii.reshape(-1), jj.reshape(-1), kk.reshape(-1)
    base_rows = r_mm + 2 * ((ii * n + jj) * n + kk)
    rows2 = np.concatenate([base_rows, base_rows + 1])
    one = [c.coef(1) for c in coo]
    two = lambda x: np.concatenate([x, x])
    coo[0].add_arrays(rows2, two(w_a + ii * n + kk), np.full(2 * n ** 3, one[0]))
    coo[1].add_arrays(rows2, two(w_b + kk * n + jj), np.full(2 * n ** 3, one[1]))
    coo[2].add_arrays(rows2, two(w_mm + (ii * n + jj) * (1 + n) + 1 + kk), np.full(2 * n ** 3, one[2]))
    # the reference's `sum` starts as a zero WITNESS that stays in the linear combination (constraints.rs:85-90)
    c_elems = [[(1, w_mm + (i * n + j) * (1 + n))] + [(1, prod_col(i, j, k)) for k in range(n)]
               for i in range(n) for j in range(n)]
    dig_c, val_c = sponge(c_elems, flat(mat_c), r_hc, w_hc)
    enforce_equal(r_eq_c, dig_c, 3)
    z[1], z[2], z[3] = val_a, val_b, val_c

    mats = [c.csr(num_rows) for c in coo]
    cm = ConstraintMatrices(L, num_vars - L, num_rows, mats[0], mats[1], mats[2])
    return cm, z
