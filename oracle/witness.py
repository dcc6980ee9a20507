"""TEST INFRASTRUCTURE -- plain-Python restatement of the NATIVE (non-gadget) helpers on the reference's witness side,
the checker for csrc/witness.cu (row f5).  Parity unpinned: no Rust toolchain here, so nothing below has been compared
with a run of the reference; SHA-256 comes from hashlib, the rest is integer arithmetic following the cited lines.

  poseidon_permute / poseidon_hash   PoseidonSponge as hasher() uses it -- absorb a flattened matrix, squeeze one element
                                     (src/arkworks/matrix_proof_of_work/hasher.rs:17-27; round structure as in the in-tree
                                     copy hashing/hashing_utils.rs:737-802; the sponge itself is ark-crypto-primitives
                                     ^0.4.0, Cargo.toml:20).  Parameters are arguments: the table hashing_utils.rs:15-715
                                     is not carried.
  mod_pow_generate_witnesses         src/arkworks/prime_snark/utils/modulo.rs:21-89 (get_mod_vals, the squaring chain, the
                                     bit-serial running product, the padding rows); table length is an argument (382 there)
  check_if_next_is_prime             src/arkworks/prime_snark/prime_circut.rs:149-195 with init_randomness,
                                     generate_bases_native (utils/hasher.rs:51-76) and fermat_test (fermat_circut.rs:131-141)
  prime_search                       the loop of prove_prime, src/arkworks/backend/prime_snark.rs:60-70
"""
import hashlib

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
NUM_BITS = 20      # prime_snark/utils/constants.rs:5
K = 3              # prime_snark/utils/constants.rs:4


class PoseidonConfig:
    """ark_crypto_primitives::sponge::poseidon::PoseidonConfig: full_rounds, partial_rounds, alpha, ark, mds, rate, capacity."""

    def __init__(self, full_rounds, partial_rounds, alpha, ark, mds, rate, capacity):
        self.full_rounds, self.partial_rounds, self.alpha = full_rounds, partial_rounds, alpha
        self.ark, self.mds, self.rate, self.capacity = ark, mds, rate, capacity
        self.width = rate + capacity


def poseidon_permute(p, state):
    """hashing_utils.rs:778-802: for every round ARK, S-box (whole state in the first and last full_rounds / 2 rounds,
    state[0] otherwise), MDS."""
    half = p.full_rounds // 2
    for r in range(p.full_rounds + p.partial_rounds):
        state = [(s + k) % R_MOD for s, k in zip(state, p.ark[r])]
        if r < half or r >= half + p.partial_rounds:
            state = [pow(s, p.alpha, R_MOD) for s in state]
        else:
            state[0] = pow(state[0], p.alpha, R_MOD)
        state = [sum(p.mds[i][j] * state[j] for j in range(p.width)) % R_MOD for i in range(p.width)]
    return state


def poseidon_hash(p, elems):
    """sponge.absorb(&elems); sponge.squeeze_native_field_elements(1)[0] (hasher.rs:23-25): rate slots sit after the
    capacity slots, a full rate is permuted before the next element goes in, squeezing permutes first."""
    state = [0] * p.width
    pos = 0
    for e in elems:
        if pos == p.rate:
            state = poseidon_permute(p, state)
            pos = 0
        state[p.capacity + pos] = (state[p.capacity + pos] + e) % R_MOD
        pos += 1
    return poseidon_permute(p, state)[p.capacity]


def mod_pow_generate_witnesses(base, div, exp, num_bits=382):
    vals = lambda num: (num, num // div, num % div)                      # get_mod_vals, modulo.rs:21-30
    power, mod_pow_vals = base, []
    for _ in range(num_bits):                                            # modulo.rs:49-53
        power = power * power
        mod_pow_vals.append(vals(power))
        power %= div
    cur, res, bits, v = base, 1, [0] * num_bits, []
    counter = 0
    while exp > 0:                                                       # modulo.rs:55-74
        elem = exp & 1
        bits[counter] = elem
        res *= (cur - 1) * elem + 1
        v.append(vals(res))
        if res > div:
            res %= div
        exp >>= 1
        cur = cur * cur % div
        counter += 1
    v += [(res, 0, res)] * (num_bits - counter)                          # modulo.rs:75-81
    return {"mod_vals": v, "mod_pow_vals": mod_pow_vals, "bits": bits, "result": res}


def _le32(v):
    return (v % R_MOD).to_bytes(32, "little")                            # Fr::into_bigint().to_bytes_le()


def check_if_next_is_prime(x, j, num_bits=NUM_BITS, k=K):
    xb = _le32(x + j)
    a_j = hashlib.sha256(xb).digest()                                    # prime_circut.rs:167-172
    num = int.from_bytes(a_j, "little")
    q, p = num >> num_bits, num & ((1 << num_bits) - 1)                  # get_mod_vals(a_j, 1 << NUM_BITS), :186-187
    r = hashlib.sha256(xb + a_j + j.to_bytes(8, "little")).digest()      # init_randomness, :149-158
    a = int.from_bytes(r, "little") % R_MOD                              # Fr::from_le_bytes_mod_order, :183
    is_prime = False
    if p:                                                                # (the reference divides by zero for p = 0)
        for jj in range(k):                                              # generate_bases_native + fermat_test
            base = int.from_bytes(hashlib.sha256(_le32(a) + _le32(jj)).digest(), "little") % p
            if pow(base, p - 1, p) == 1:
                is_prime = True
                break
    return {"digest": a_j, "is_prime": is_prime, "quotient": q, "remainder": p, "a": a}


def prime_search(x, i, num_bits=NUM_BITS, k=K):
    """prove_prime: for j in 0..=i, the first j whose candidate passes; None when there is none."""
    for j in range(i + 1):
        c = check_if_next_is_prime(x, j, num_bits, k)
        if c["is_prime"]:
            return j, c
    return None, c
