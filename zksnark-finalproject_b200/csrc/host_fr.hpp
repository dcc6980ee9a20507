// Scalar-field (Fr) arithmetic on the HOST: 4 x u64 Montgomery limbs (R = 2^256), canonical representatives -- arkworks'
// in-memory form of ark_bls12_381::Fr.  Used by the host-only translation units (witness.cu: witness generation;
// setup_host.cu: key-generation scalars).  Device code has its own 32-bit-limb arithmetic (mont.cuh).
#pragma once
#include <cstdint>
#include <cstring>

namespace b2z {
namespace hostfr {

typedef unsigned __int128 u128;

// ---- Fr, 4 x u64 Montgomery (R = 2^256), canonical representatives
static const uint64_t kMod[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
static const uint64_t kInv = 0xfffffffeffffffffull;   // -r^-1 mod 2^64
static const uint64_t kR2[4] = {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull};

struct Fr {
  uint64_t l[4];
};

inline bool geq_mod(const uint64_t t[4]) {
  for (int i = 3; i >= 0; i--)
    if (t[i] != kMod[i]) return t[i] > kMod[i];
  return true;
}
inline void sub_mod(uint64_t t[4]) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    const u128 d = (u128)t[i] - kMod[i] - borrow;
    t[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
}
inline Fr fr_add(const Fr& a, const Fr& b) {
  Fr o;
  uint64_t c = 0;
  for (int i = 0; i < 4; i++) {
    const u128 s = (u128)a.l[i] + b.l[i] + c;
    o.l[i] = (uint64_t)s;
    c = (uint64_t)(s >> 64);
  }
  if (c || geq_mod(o.l)) sub_mod(o.l);      // r < 2^255: a + b < 2^256, c is always 0; kept for clarity
  return o;
}
// CIOS Montgomery product
inline Fr fr_mul(const Fr& a, const Fr& b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 4; j++) {
      const u128 x = (u128)a.l[j] * b.l[i] + t[j] + c;
      t[j] = (uint64_t)x;
      c = (uint64_t)(x >> 64);
    }
    u128 x = (u128)t[4] + c;
    t[4] = (uint64_t)x;
    t[5] = (uint64_t)(x >> 64);
    const uint64_t m = t[0] * kInv;
    x = (u128)m * kMod[0] + t[0];
    c = (uint64_t)(x >> 64);
    for (int j = 1; j < 4; j++) {
      x = (u128)m * kMod[j] + t[j] + c;
      t[j - 1] = (uint64_t)x;
      c = (uint64_t)(x >> 64);
    }
    x = (u128)t[4] + c;
    t[3] = (uint64_t)x;
    t[4] = t[5] + (uint64_t)(x >> 64);
  }
  Fr o;
  std::memcpy(o.l, t, 32);
  if (t[4] || geq_mod(o.l)) sub_mod(o.l);
  return o;
}
inline Fr fr_zero() { return Fr{{0, 0, 0, 0}}; }
inline Fr fr_from_u64(uint64_t v) {
  Fr a{{v, 0, 0, 0}}, r2;
  std::memcpy(r2.l, kR2, 32);
  return fr_mul(a, r2);
}
inline bool fr_is_canonical(const uint64_t* l) { return !geq_mod(l); }
inline Fr fr_load(const uint64_t* p) {
  Fr a;
  std::memcpy(a.l, p, 32);
  return a;
}
inline void fr_store(uint64_t* p, const Fr& a) { std::memcpy(p, a.l, 32); }

inline Fr fr_one() { return fr_from_u64(1); }
inline bool fr_is_zero(const Fr& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
inline Fr fr_sub(const Fr& a, const Fr& b) {
  Fr o;
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    const u128 d = (u128)a.l[i] - b.l[i] - borrow;
    o.l[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  if (borrow) {                                  // a < b: add the modulus back
    uint64_t c = 0;
    for (int i = 0; i < 4; i++) {
      const u128 s = (u128)o.l[i] + kMod[i] + c;
      o.l[i] = (uint64_t)s;
      c = (uint64_t)(s >> 64);
    }
  }
  return o;
}
// a^e for a 256-bit exponent (little-endian limbs), left to right
inline Fr fr_pow(const Fr& a, const uint64_t e[4]) {
  Fr acc = fr_one();
  bool started = false;
  for (int i = 255; i >= 0; i--) {
    if (started) acc = fr_mul(acc, acc);
    if ((e[i >> 6] >> (i & 63)) & 1) {
      acc = started ? fr_mul(acc, a) : a;
      started = true;
    }
  }
  return acc;
}
inline Fr fr_pow_u64(const Fr& a, uint64_t e) {
  const uint64_t ee[4] = {e, 0, 0, 0};
  return fr_pow(a, ee);
}
// a^(r - 2); 0 -> 0
inline Fr fr_inv(const Fr& a) {
  const uint64_t e[4] = {0xfffffffeffffffffull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
  return fr_pow(a, e);
}
// Montgomery -> canonical integer limbs (into_bigint())
inline Fr fr_into_bigint(const Fr& a) {
  const Fr one{{1, 0, 0, 0}};
  return fr_mul(a, one);
}
// the 2^32-th root of unity 7^((r - 1) / 2^32) (ark_bls12_381::FrConfig::TWO_ADIC_ROOT_OF_UNITY), Montgomery form
inline Fr fr_two_adic_root() {
  const Fr canon{{0x3829971f439f0d2bull, 0xb63683508c2280b9ull, 0xd09b681922c813b4ull, 0x16a2a19edfe81f20ull}};
  Fr r2;
  std::memcpy(r2.l, kR2, 32);
  return fr_mul(canon, r2);
}

}  // namespace hostfr
}  // namespace b2z
