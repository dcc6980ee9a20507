// Witness generation on the host, multithreaded (include/b200zk.h): no GPU, no context.
//
// Once a proof takes tens of milliseconds, computing the assignment z is the next bottleneck (the Python builder of
// this repository needs seconds for the 64x64 circuit).  The reference computes its witnesses during synthesis
// (ark-r1cs-std gadgets), with two native helpers next to them whose behaviour is restated here:
//   * hasher()  -- Poseidon sponge digest of a flattened matrix: absorb everything, squeeze one element
//                  (src/arkworks/matrix_proof_of_work/hasher.rs:17-27; permutation structure as in the in-tree copy
//                  hashing/hashing_utils.rs:737-802: ARK, S-box on the whole state in full rounds and on state[0] in
//                  partial rounds, MDS; rounds full/2 | partial | full/2);
//   * mod_pow_generate_witnesses() -- per-bit (value, quotient, remainder) tables of a right-to-left
//                  square-and-multiply (src/arkworks/prime_snark/utils/modulo.rs:31-89).
// The Poseidon parameters (round constants, MDS) are ARGUMENTS: the reference's table (hashing_utils.rs:15-715) is
// data of the reference and is not carried here; a Rust caller passes its PoseidonConfig's `ark` / `mds`.
//
// b2z_matrix_circuit_witness produces the full assignment of the matrix-multiplication circuit
// (matrix_proof_of_work/constraints.rs:78-128) in the variable order of this repository's circuit builder
// (zksnark-finalproject_b200/circuits.py: matrix_circuit), as Montgomery limbs ready for b2z_groth16_prove_r1cs.
// Threads: the digests of A and B are independent sequential chains (one thread each), the n^3 scalar products and
// the n^2 sums are split over the remaining threads, the digest of C follows the sums.
#include <algorithm>
#include <cstring>
#include <new>
#include <system_error>
#include <thread>
#include <vector>

#include "../../include/b200zk.h"
#include "host_fr.hpp"

using namespace b2z::hostfr;

namespace {

// Fr arithmetic: host_fr.hpp

// ---- Poseidon
constexpr uint32_t kMaxWidth = 8;

struct Poseidon {
  uint32_t full = 0, partial = 0, width = 0, rate = 0, capacity = 0;
  uint64_t alpha = 0;
  uint32_t sbox_vars = 0;            // witnesses one S-box allocates: squarings + multiplications of the power chain
  std::vector<Fr> ark, mds;
};

// x^alpha by left-to-right square-and-multiply (FpVar::pow_by_constant); every squaring and every multiplication
// is one witness; they are appended to *wit when it is not NULL.  alpha = 17: x^2, x^4, x^8, x^16, x^17.
inline Fr sbox(const Poseidon& P, const Fr& x, Fr** wit) {
  int top = 63;
  while (!((P.alpha >> top) & 1)) top--;
  Fr acc = x;
  for (int b = top - 1; b >= 0; b--) {
    acc = fr_mul(acc, acc);
    if (wit) *(*wit)++ = acc;
    if ((P.alpha >> b) & 1) {
      acc = fr_mul(acc, x);
      if (wit) *(*wit)++ = acc;
    }
  }
  return acc;
}

uint32_t sbox_var_count(uint64_t alpha) {
  int top = 63;
  while (!((alpha >> top) & 1)) top--;
  uint32_t n = 0;
  for (int b = top - 1; b >= 0; b--) n += 1 + (uint32_t)((alpha >> b) & 1);
  return n;
}

uint32_t perm_vars(const Poseidon& P) { return (P.full * P.width + P.partial) * P.sbox_vars; }

void permute(const Poseidon& P, Fr* state, Fr* wit) {
  const uint32_t half = P.full / 2, w = P.width;
  Fr tmp[kMaxWidth];
  Fr** wp = wit ? &wit : nullptr;
  for (uint32_t r = 0; r < P.full + P.partial; r++) {
    for (uint32_t i = 0; i < w; i++) state[i] = fr_add(state[i], P.ark[(size_t)r * w + i]);
    const bool full_round = r < half || r >= half + P.partial;
    const uint32_t ns = full_round ? w : 1;
    for (uint32_t i = 0; i < ns; i++) state[i] = sbox(P, state[i], wp);
    for (uint32_t i = 0; i < w; i++) {
      Fr acc = fr_zero();
      for (uint32_t j = 0; j < w; j++) acc = fr_add(acc, fr_mul(P.mds[(size_t)i * w + j], state[j]));
      tmp[i] = acc;
    }
    for (uint32_t i = 0; i < w; i++) state[i] = tmp[i];
  }
}

// sponge: absorb `count` elements (rate slots sit after the capacity; permute when the rate is full), one more
// permutation, squeeze state[capacity].  wit (may be NULL): perm_vars() witnesses per permutation, in order.
Fr sponge_digest(const Poseidon& P, const Fr* elems, size_t count, Fr* wit) {
  Fr state[kMaxWidth];
  for (uint32_t i = 0; i < P.width; i++) state[i] = fr_zero();
  const uint32_t pv = perm_vars(P);
  uint32_t pos = 0;
  bool pending = false;
  for (size_t e = 0; e < count; e++) {
    if (pos == P.rate) {
      permute(P, state, wit);
      if (wit) wit += pv;
      pos = 0;
    }
    state[P.capacity + pos] = fr_add(state[P.capacity + pos], elems[e]);
    pos++;
    pending = true;
  }
  // squeezing always permutes first (PoseidonSponge in absorbing mode), also after an empty absorb
  (void)pending;
  permute(P, state, wit);
  return state[P.capacity];
}

size_t sponge_permutations(const Poseidon& P, size_t count) { return count ? (count + P.rate - 1) / P.rate : 1; }

bool load_poseidon(const b2z_poseidon_desc* d, Poseidon* P) {
  if (d == nullptr || d->ark == nullptr || d->mds == nullptr) return false;
  if (d->width < 2 || d->width > kMaxWidth || d->rate == 0 || d->capacity == 0 || d->rate + d->capacity != d->width) return false;
  if (d->alpha < 2 || (d->full_rounds & 1) || d->full_rounds + d->partial_rounds == 0 || d->full_rounds + d->partial_rounds > 4096)
    return false;
  P->full = d->full_rounds; P->partial = d->partial_rounds; P->width = d->width; P->rate = d->rate;
  P->capacity = d->capacity; P->alpha = d->alpha; P->sbox_vars = sbox_var_count(d->alpha);
  const size_t na = (size_t)(P->full + P->partial) * P->width, nm = (size_t)P->width * P->width;
  P->ark.resize(na);
  P->mds.resize(nm);
  for (size_t i = 0; i < na; i++) {
    if (!fr_is_canonical(d->ark + 4 * i)) return false;
    P->ark[i] = fr_load(d->ark + 4 * i);
  }
  for (size_t i = 0; i < nm; i++) {
    if (!fr_is_canonical(d->mds + 4 * i)) return false;
    P->mds[i] = fr_load(d->mds + 4 * i);
  }
  return true;
}

uint32_t pick_threads(uint32_t want) {
  if (want == 0) {
    want = std::thread::hardware_concurrency();
    if (want == 0) want = 4;
  }
  return std::min<uint32_t>(want, 256);
}

struct MatrixLayout {
  uint64_t N, T, pv, w_a, w_b, w_ha, w_hb, w_cph, w_mm, w_hc, num_vars;
};
MatrixLayout matrix_layout(const Poseidon& P, uint32_t n) {
  MatrixLayout L;
  L.N = (uint64_t)n * n;
  L.T = sponge_permutations(P, L.N);
  L.pv = perm_vars(P);
  L.w_a = 4;                                   // instance: 1, digest(A), digest(B), digest(C)
  L.w_b = L.w_a + L.N;
  L.w_ha = L.w_b + L.N;                        // S-box witnesses of the digest of A ...
  L.w_hb = L.w_ha + L.T * L.pv;                // ... of B
  L.w_cph = L.w_hb + L.T * L.pv;               // n^2 placeholder witnesses for C (zero)
  L.w_mm = L.w_cph + L.N;                      // per (i, j): the zero-initialised sum witness, then n products
  L.w_hc = L.w_mm + L.N * (1 + (uint64_t)n);
  L.num_vars = L.w_hc + L.T * L.pv;
  return L;
}

}  // namespace

extern "C" {

b2z_status b2z_poseidon_hash(const b2z_poseidon_desc* params, const uint64_t* elems, uint64_t count, uint64_t digest_out[4]) {
  if (digest_out == nullptr || (count && elems == nullptr)) return B2Z_EINVAL;
  try {
    Poseidon P;
    if (!load_poseidon(params, &P)) return B2Z_EINVAL;
    std::vector<Fr> in(count);
    for (uint64_t i = 0; i < count; i++) {
      if (!fr_is_canonical(elems + 4 * i)) return B2Z_EINVAL;
      in[i] = fr_load(elems + 4 * i);
    }
    fr_store(digest_out, sponge_digest(P, in.data(), count, nullptr));
    return B2Z_OK;
  } catch (const std::bad_alloc&) {
    return B2Z_ENOMEM;
  }
}

uint64_t b2z_matrix_circuit_num_variables(const b2z_poseidon_desc* params, uint32_t n) {
  Poseidon P;
  if (n == 0 || !load_poseidon(params, &P)) return 0;
  return matrix_layout(P, n).num_vars;
}

b2z_status b2z_matrix_circuit_witness(const b2z_poseidon_desc* params, uint32_t n, const uint64_t* a, const uint64_t* b,
                                      uint32_t threads, uint64_t* z_out, uint64_t z_capacity) {
  if (n == 0 || a == nullptr || b == nullptr || z_out == nullptr) return B2Z_EINVAL;
  try {
    Poseidon P;
    if (!load_poseidon(params, &P)) return B2Z_EINVAL;
    const MatrixLayout L = matrix_layout(P, n);
    if (z_capacity < L.num_vars) return B2Z_ESIZE;
    for (uint64_t i = 0; i < L.N; i++)
      if (!fr_is_canonical(a + 4 * i) || !fr_is_canonical(b + 4 * i)) return B2Z_EINVAL;
    Fr* z = reinterpret_cast<Fr*>(z_out);
    static_assert(sizeof(Fr) == 32, "Fr is four packed limbs");
    z[0] = fr_from_u64(1);
    std::memcpy(z + L.w_a, a, L.N * 32);
    std::memcpy(z + L.w_b, b, L.N * 32);
    std::memset(z + L.w_cph, 0, L.N * 32);
    std::vector<Fr> c(L.N);
    const uint32_t nt = pick_threads(threads);
    // phase 1: digests of A and B (one thread each) beside the products and sums (the other threads, rows i)
    auto products = [&](uint32_t lo, uint32_t hi) {
      for (uint32_t i = lo; i < hi; i++)
        for (uint32_t j = 0; j < n; j++) {
          Fr* cell = z + L.w_mm + ((uint64_t)i * n + j) * (1 + (uint64_t)n);
          cell[0] = fr_zero();                                     // the sum starts as a zero WITNESS
          Fr acc = fr_zero();
          for (uint32_t k = 0; k < n; k++) {
            const Fr p = fr_mul(z[L.w_a + (uint64_t)i * n + k], z[L.w_b + (uint64_t)k * n + j]);
            cell[1 + k] = p;
            acc = fr_add(acc, p);
          }
          c[(uint64_t)i * n + j] = acc;
        }
    };
    auto digest_a = [&]() { z[1] = sponge_digest(P, z + L.w_a, L.N, z + L.w_ha); };
    auto digest_b = [&]() { z[2] = sponge_digest(P, z + L.w_b, L.N, z + L.w_hb); };
    if (nt <= 1) {
      digest_a();
      digest_b();
      products(0, n);
    } else {
      std::vector<std::thread> pool;
      pool.emplace_back(digest_a);
      if (nt >= 3) pool.emplace_back(digest_b);
      const uint32_t workers = nt >= 4 ? std::min<uint32_t>(nt - 2, n) : 1;
      std::vector<std::thread> mm;
      for (uint32_t t = 1; t < workers; t++)
        mm.emplace_back(products, (uint32_t)((uint64_t)n * t / workers), (uint32_t)((uint64_t)n * (t + 1) / workers));
      products(0, (uint32_t)((uint64_t)n / workers));
      if (nt < 3) digest_b();
      for (auto& t : mm) t.join();
      // phase 2 can start as soon as the sums are there: the digest of C runs on this thread while A / B finish
      z[3] = sponge_digest(P, c.data(), L.N, z + L.w_hc);
      for (auto& t : pool) t.join();
      return B2Z_OK;
    }
    z[3] = sponge_digest(P, c.data(), L.N, z + L.w_hc);
    return B2Z_OK;
  } catch (const std::bad_alloc&) {
    return B2Z_ENOMEM;
  } catch (const std::system_error&) {
    return B2Z_ENOMEM;
  }
}

b2z_status b2z_fibonacci_witness(const uint64_t a[4], const uint64_t b[4], uint64_t num_steps, uint64_t z_out[20]) {
  if (a == nullptr || b == nullptr || z_out == nullptr) return B2Z_EINVAL;
  if (!fr_is_canonical(a) || !fr_is_canonical(b)) return B2Z_EINVAL;
  Fr f2 = fr_load(a), f1 = fr_load(b);
  for (uint64_t i = 0; i < num_steps; i++) {
    const Fr fi = fr_add(f1, f2);
    f2 = f1;
    f1 = fi;
  }
  Fr* z = reinterpret_cast<Fr*>(z_out);
  z[0] = fr_from_u64(1);
  z[1] = fr_load(a);
  z[2] = fr_load(b);
  z[3] = num_steps ? f1 : fr_zero();
  z[4] = fr_zero();                       // the one allocated witness (fibbonaci.rs:30) keeps its initial value
  return B2Z_OK;
}

b2z_status b2z_modpow_witnesses(uint64_t base, uint64_t modulus, uint64_t exponent, uint32_t num_bits, uint64_t* mod_vals,
                                uint64_t* mod_pow_vals, uint8_t* bits, uint64_t* result) {
  if (modulus < 2 || modulus >> 63 || base >> 63 || num_bits == 0 || num_bits > 4096 || mod_vals == nullptr ||
      mod_pow_vals == nullptr || bits == nullptr || result == nullptr)
    return B2Z_EINVAL;
  if (num_bits < 64 && (exponent >> num_bits)) return B2Z_EINVAL;
  auto put = [](uint64_t* row, u128 num, uint64_t div) {          // (num lo, num hi, q lo, q hi, remainder)
    const u128 q = num / div;
    row[0] = (uint64_t)num; row[1] = (uint64_t)(num >> 64);
    row[2] = (uint64_t)q; row[3] = (uint64_t)(q >> 64);
    row[4] = (uint64_t)(num % div);
  };
  std::memset(bits, 0, num_bits);
  // the squaring chain: power <- power^2, recorded BEFORE the reduction
  u128 power = base;
  for (uint32_t i = 0; i < num_bits; i++) {
    power = power * power;
    put(mod_pow_vals + 5 * (size_t)i, power, modulus);
    power %= modulus;
  }
  // the running result, bit by bit from the least significant one
  u128 res = 1, cur = base;
  uint64_t e = exponent;
  uint32_t counter = 0;
  while (e > 0) {
    const uint64_t bit = e & 1;
    bits[counter] = (uint8_t)bit;
    res *= bit ? cur : (u128)1;
    put(mod_vals + 5 * (size_t)counter, res, modulus);
    if (res > modulus) res %= modulus;
    e >>= 1;
    cur = (cur * cur) % modulus;
    counter++;
  }
  for (uint32_t i = counter; i < num_bits; i++) {                  // padding rows: (res, 0, res)
    uint64_t* row = mod_vals + 5 * (size_t)i;
    row[0] = (uint64_t)res; row[1] = (uint64_t)(res >> 64); row[2] = 0; row[3] = 0; row[4] = (uint64_t)res;
  }
  *result = (uint64_t)res;
  return B2Z_OK;
}

}  // extern "C"

// ---- SHA-256 (FIPS 180-4) and the prime route's native search -------------------------------------------------------
namespace {

const uint32_t kShaK[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

struct Sha256 {
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  uint8_t buf[64];
  uint64_t len = 0;
  void block(const uint8_t* p) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
      w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
      const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      const uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + kShaK[i] + w[i];
      const uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  void update(const uint8_t* p, size_t n) {
    while (n) {
      const size_t fill = (size_t)(len & 63), take = std::min<size_t>(n, 64 - fill);
      std::memcpy(buf + fill, p, take);
      len += take; p += take; n -= take;
      if (((len & 63) == 0)) block(buf);
    }
  }
  void finalize(uint8_t out[32]) {
    const uint64_t bits = len * 8;
    const uint8_t one = 0x80, zero = 0;
    update(&one, 1);
    while ((len & 63) != 56) update(&zero, 1);
    uint8_t lb[8];
    for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(lb, 8);
    for (int i = 0; i < 8; i++) {
      out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16);
      out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i];
    }
  }
};

// canonical little-endian bytes of an Fr element given in Montgomery form (into_bigint().to_bytes_le())
inline void fr_to_le_bytes(const Fr& mont, uint8_t out[32]) {
  const Fr one{{1, 0, 0, 0}};
  const Fr c = fr_mul(mont, one);                 // a R * 1 / R = a
  std::memcpy(out, c.l, 32);                      // x86-64: limbs are little-endian already
}
// 256-bit little-endian integer (4 limbs) mod a 64-bit modulus
inline uint64_t mod_u256(const uint64_t v[4], uint64_t m) {
  u128 rem = 0;
  for (int i = 3; i >= 0; i--) rem = ((rem << 64) | v[i]) % m;
  return (uint64_t)rem;
}
inline uint64_t modpow_u64(uint64_t base, uint64_t e, uint64_t m) {
  if (m == 1) return 0;
  u128 res = 1, b = base % m;
  while (e) {
    if (e & 1) res = res * b % m;
    b = b * b % m;
    e >>= 1;
  }
  return (uint64_t)res;
}

struct PrimeCheck {
  uint8_t digest[32];      // a_j = SHA-256(le32(x + j))
  bool is_prime;
  uint64_t q[4];           // a_j >> num_bits
  uint64_t remainder;      // a_j mod 2^num_bits: the candidate
  uint64_t a[4];           // Fr::from_le_bytes_mod_order(SHA-256(le32(x + j) || a_j || le64(j))), canonical
};

// check_if_next_is_prime(x, j) (prime_circut.rs:165-195) with fermat_test / generate_bases_native inlined
PrimeCheck prime_check(const Fr& x_mont, uint64_t j, uint32_t num_bits, uint32_t k_bases) {
  PrimeCheck R;
  uint8_t xb[32];
  fr_to_le_bytes(fr_add(x_mont, fr_from_u64(j)), xb);
  Sha256 s1;
  s1.update(xb, 32);
  s1.finalize(R.digest);
  uint64_t aj[4];
  std::memcpy(aj, R.digest, 32);                  // BigUint::from_bytes_le
  // a_j = q * 2^num_bits + remainder
  const uint32_t ws = num_bits / 64, bs = num_bits % 64;
  for (int i = 0; i < 4; i++) {
    uint64_t lo = (i + ws < 4) ? aj[i + ws] : 0, hi = (i + ws + 1 < 4) ? aj[i + ws + 1] : 0;
    R.q[i] = bs ? (lo >> bs) | (hi << (64 - bs)) : lo;
  }
  R.remainder = num_bits >= 64 ? aj[0] : aj[0] & ((1ull << num_bits) - 1);
  // r = SHA-256(le32(x + j) || a_j || j as 8 little-endian bytes), reduced mod the scalar field
  uint8_t jb[8], rb[32];
  for (int i = 0; i < 8; i++) jb[i] = (uint8_t)(j >> (8 * i));
  Sha256 s2;
  s2.update(xb, 32);
  s2.update(R.digest, 32);
  s2.update(jb, 8);
  s2.finalize(rb);
  std::memcpy(R.a, rb, 32);
  while (geq_mod(R.a)) sub_mod(R.a);              // 2^256 < 3 r: at most two subtractions
  // fermat_test(a, p): bases SHA-256(le32(a) || le32(jj)) mod p; "prime" when ANY base passes (fermat_circut.rs:131-141)
  const uint64_t p = R.remainder;
  R.is_prime = false;
  if (p == 0) return R;                           // the reference divides by zero here (BigUint panic)
  uint8_t ab[32];
  std::memcpy(ab, R.a, 32);
  for (uint32_t jj = 0; jj < k_bases && !R.is_prime; jj++) {
    uint8_t jjb[32] = {0}, d[32];
    for (int i = 0; i < 4; i++) jjb[i] = (uint8_t)(jj >> (8 * i));
    Sha256 s3;
    s3.update(ab, 32);
    s3.update(jjb, 32);
    s3.finalize(d);
    uint64_t v[4];
    std::memcpy(v, d, 32);
    const uint64_t base = mod_u256(v, p);
    if (modpow_u64(base, p - 1, p) == 1) R.is_prime = true;
  }
  return R;
}

}  // namespace

extern "C" {

void b2z_sha256(const uint8_t* data, uint64_t len, uint8_t digest_out[32]) {
  Sha256 s;
  if (len) s.update(data, (size_t)len);
  s.finalize(digest_out);
}

b2z_status b2z_prime_search(const uint64_t x[4], uint64_t j_first, uint64_t j_last, uint32_t num_bits, uint32_t k_bases,
                            uint32_t threads, b2z_prime_check* out, int32_t* found) {
  if (x == nullptr || out == nullptr || found == nullptr || j_last < j_first || num_bits == 0 || num_bits > 62 ||
      k_bases == 0 || k_bases > 64 || !fr_is_canonical(x))
    return B2Z_EINVAL;
  try {
    const Fr xm = fr_load(x);
    const uint64_t total = j_last - j_first + 1;
    uint32_t nt = pick_threads(threads);
    if ((uint64_t)nt > total) nt = (uint32_t)total;
    // ranks take j = j_first + t, j_first + t + nt, ...: the smallest hit of all ranks is the reference's loop exit
    std::vector<uint64_t> hit(nt, UINT64_MAX);
    std::vector<PrimeCheck> res(nt);
    std::vector<PrimeCheck> last(nt);
    auto work = [&](uint32_t t) {
      for (uint64_t off = t; off < total; off += nt) {
        const uint64_t j = j_first + off;
        bool stop = false;
        for (uint32_t u = 0; u < nt; u++)
          if (__atomic_load_n(&hit[u], __ATOMIC_RELAXED) < j) stop = true;      // somebody found a smaller j
        if (stop) return;
        PrimeCheck c = prime_check(xm, j, num_bits, k_bases);
        if (j == j_last) last[t] = c;
        if (c.is_prime) {
          res[t] = c;
          __atomic_store_n(&hit[t], j, __ATOMIC_RELAXED);
          return;
        }
      }
    };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < nt; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    uint32_t best = nt;
    for (uint32_t t = 0; t < nt; t++)
      if (hit[t] != UINT64_MAX && (best == nt || hit[t] < hit[best])) best = t;
    const PrimeCheck* c;
    uint64_t j;
    if (best != nt) {
      c = &res[best]; j = hit[best]; *found = 1;
    } else {                                       // none: report j_last, as the reference's loop leaves check_result
      c = &last[(uint32_t)((total - 1) % nt)]; j = j_last; *found = 0;
    }
    out->j = j;
    std::memcpy(out->digest, c->digest, 32);
    out->is_prime = c->is_prime ? 1 : 0;
    std::memcpy(out->quotient, c->q, 32);
    out->remainder = c->remainder;
    std::memcpy(out->a, c->a, 32);
    return B2Z_OK;
  } catch (const std::bad_alloc&) {
    return B2Z_ENOMEM;
  } catch (const std::system_error&) {
    return B2Z_ENOMEM;
  }
}

}  // extern "C"
