// Host-side BLS12-381 pairing and the ark-groth16 verifier over the reference's wire formats -- row f4 of
// `prepare_verifying_key` + `PreparedVerifyingKey::serialize_compressed` (what encode_pvk
// base64-encodes, /root/reference/src/arkworks/matrix_proof_of_work/io.rs:62-68, built at
// src/arkworks/backend/matrix_proof.rs:134-136) and `Groth16::verify_with_processed_vk`
// (matrix_proof.rs:199-206).  Verification is O(1) host work (three Miller loops and one final exponentiation),
// exactly where arkworks does it; nothing here touches the GPU.
//
// Restated from the published crates (ark-ec 0.4.2 models/bls12, ark-groth16 0.4 verifier.rs / data_structures.rs,
// ark-serialize 0.4; none on disk -- "parity unpinned"):
//   tower      Fq2 = Fq[u]/(u^2 + 1), Fq6 = Fq2[v]/(v^3 - (1 + u)), Fq12 = Fq6[w]/(w^2 - v); M-type twist
//   G2Prepared line coefficients in homogeneous projective coordinates: 63 doubling steps and 5 addition steps
//              over the bits of |x| = 0xd201000000010000 (MSB skipped) = 68 (Fq2, Fq2, Fq2) triples
//   miller     f <- f^2 * ell(coeff, P) per bit (+ one more ell on set bits), ell = mul_by_014(c0, c1 * Px, c2 * Py);
//              conjugated at the end because x is negative
//   final exp  easy part (q^6 - 1)(q^2 + 1), hard part 3 (q^4 - q^2 + 1) / r by the x-power chain (the factor 3 is
//              part of arkworks' result: alpha_g1_beta_g2 = e(alpha, beta)^3 in textbook terms)
//   wire       G1 / G2 compressed zcash encoding (big-endian), Fq12 = 12 x 48 bytes little-endian canonical,
//              Vec<T> = u64 LE length + items, bool = 1 byte
#pragma once
#include <vector>

#include "host_fq.hpp"

namespace b2z {
namespace host {

// ------------------------------------------------------------------ Fq helpers
inline Fq fq_from_u64(uint64_t v) {   // Montgomery form of a small integer: v * R = v * one
  Fq acc = fq_zero(), base = fq_one();
  while (v) {
    if (v & 1) acc = fq_add(acc, base);
    base = fq_dbl(base);
    v >>= 1;
  }
  return acc;
}
inline Fq fq_pow(const Fq& a, const uint64_t* e, int nlimbs) {
  Fq r = fq_one();
  for (int i = nlimbs * 64 - 1; i >= 0; i--) {
    r = fq_sqr(r);
    if ((e[i >> 6] >> (i & 63)) & 1) r = fq_mul(r, a);
  }
  return r;
}
// q = 3 mod 4: sqrt(a) = a^((q+1)/4); returns false if a is a non-residue
inline bool fq_sqrt(const Fq& a, Fq* out) {
  uint64_t e[6];
  std::memcpy(e, kQ, sizeof(e));
  e[0] += 1;   // q + 1 (no carry: low limb ends in ...aaab)
  for (int i = 0; i < 6; i++) e[i] = (e[i] >> 2) | (i + 1 < 6 ? e[i + 1] << 62 : 0);
  const Fq r = fq_pow(a, e, 6);
  *out = r;
  return fq_eq(fq_sqr(r), a);
}
// canonical little-endian bytes <-> Montgomery element (ark-serialize Fp encoding)
inline void fq_write_le(uint8_t* dst, const Fq& a) {
  uint64_t c[6];
  fq_to_canonical(a, c);
  std::memcpy(dst, c, 48);
}
static const uint64_t kR2[6] = {0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull,
                                0x67eb88a9939d83c0ull, 0x9a793e85b519952dull, 0x11988fe592cae3aaull};   // R^2 mod q
inline bool fq_from_canonical(const uint64_t c[6], Fq* out) {
  if (fq_geq_q(c)) return false;
  Fq x, r2;
  std::memcpy(x.l, c, 48);
  std::memcpy(r2.l, kR2, 48);
  *out = fq_mul(x, r2);
  return true;
}
inline bool fq_read_le(const uint8_t* src, Fq* out) {
  uint64_t c[6];
  std::memcpy(c, src, 48);
  return fq_from_canonical(c, out);
}
inline bool fq_read_be(const uint8_t* src, uint8_t mask_first, Fq* out) {
  uint64_t c[6];
  for (int i = 0; i < 6; i++) {
    uint64_t v = 0;
    for (int b = 0; b < 8; b++) {
      uint8_t byte = src[8 * i + b];
      if (i == 0 && b == 0) byte &= mask_first;
      v = (v << 8) | byte;
    }
    c[5 - i] = v;
  }
  return fq_from_canonical(c, out);
}

// ------------------------------------------------------------------ Fq2 (more of it than the prover's epilogue needs)
inline Fq2 fq2_zero() { return Fq2{fq_zero(), fq_zero()}; }
inline Fq2 fq2_one() { return Fq2{fq_one(), fq_zero()}; }
inline Fq2 fq2_neg(const Fq2& a) { return Fq2{fq_neg(a.c0), fq_neg(a.c1)}; }
inline Fq2 fq2_conj(const Fq2& a) { return Fq2{a.c0, fq_neg(a.c1)}; }
inline bool fq2_eq(const Fq2& a, const Fq2& b) { return fq_eq(a.c0, b.c0) && fq_eq(a.c1, b.c1); }
inline Fq2 fq2_mul_fq(const Fq2& a, const Fq& k) { return Fq2{fq_mul(a.c0, k), fq_mul(a.c1, k)}; }
inline Fq2 fq2_mul_xi(const Fq2& a) {   // * (1 + u)
  return Fq2{fq_sub(a.c0, a.c1), fq_add(a.c0, a.c1)};
}
inline Fq2 fq2_pow(const Fq2& a, const uint64_t* e, int nlimbs) {
  Fq2 r = fq2_one();
  for (int i = nlimbs * 64 - 1; i >= 0; i--) {
    r = fq2_sqr(r);
    if ((e[i >> 6] >> (i & 63)) & 1) r = fq2_mul(r, a);
  }
  return r;
}
// square root by the norm method (any root: the caller picks the sign by the wire flag)
inline bool fq2_sqrt(const Fq2& a, Fq2* out) {
  if (fq2_is_zero(a)) { *out = a; return true; }
  Fq s;
  if (fq_is_zero(a.c1)) {
    if (fq_sqrt(a.c0, &s)) { *out = Fq2{s, fq_zero()}; return true; }
    if (fq_sqrt(fq_neg(a.c0), &s)) { *out = Fq2{fq_zero(), s}; return true; }
    return false;
  }
  Fq alpha;
  if (!fq_sqrt(fq_add(fq_sqr(a.c0), fq_sqr(a.c1)), &alpha)) return false;
  const Fq inv2 = fq_inv(fq_from_u64(2));
  Fq x0;
  if (!fq_sqrt(fq_mul(fq_add(a.c0, alpha), inv2), &x0)) {
    if (!fq_sqrt(fq_mul(fq_sub(a.c0, alpha), inv2), &x0)) return false;
  }
  const Fq x1 = fq_mul(a.c1, fq_inv(fq_dbl(x0)));
  *out = Fq2{x0, x1};
  return fq2_eq(fq2_sqr(*out), a);
}

// ------------------------------------------------------------------ Fq6, Fq12
struct Fq6 {
  Fq2 c0, c1, c2;
};
struct Fq12 {
  Fq6 c0, c1;
};
inline Fq6 fq6_zero() { return Fq6{fq2_zero(), fq2_zero(), fq2_zero()}; }
inline Fq6 fq6_one() { return Fq6{fq2_one(), fq2_zero(), fq2_zero()}; }
inline Fq6 fq6_add(const Fq6& a, const Fq6& b) { return Fq6{fq2_add(a.c0, b.c0), fq2_add(a.c1, b.c1), fq2_add(a.c2, b.c2)}; }
inline Fq6 fq6_sub(const Fq6& a, const Fq6& b) { return Fq6{fq2_sub(a.c0, b.c0), fq2_sub(a.c1, b.c1), fq2_sub(a.c2, b.c2)}; }
inline Fq6 fq6_neg(const Fq6& a) { return Fq6{fq2_neg(a.c0), fq2_neg(a.c1), fq2_neg(a.c2)}; }
inline bool fq6_eq(const Fq6& a, const Fq6& b) { return fq2_eq(a.c0, b.c0) && fq2_eq(a.c1, b.c1) && fq2_eq(a.c2, b.c2); }
inline Fq6 fq6_mul_v(const Fq6& a) { return Fq6{fq2_mul_xi(a.c2), a.c0, a.c1}; }   // * v  (v^3 = 1 + u)
inline Fq6 fq6_mul(const Fq6& a, const Fq6& b) {
  const Fq2 t0 = fq2_mul(a.c0, b.c0), t1 = fq2_mul(a.c1, b.c1), t2 = fq2_mul(a.c2, b.c2);
  const Fq2 x0 = fq2_add(t0, fq2_mul_xi(fq2_sub(fq2_sub(fq2_mul(fq2_add(a.c1, a.c2), fq2_add(b.c1, b.c2)), t1), t2)));
  const Fq2 x1 = fq2_add(fq2_sub(fq2_sub(fq2_mul(fq2_add(a.c0, a.c1), fq2_add(b.c0, b.c1)), t0), t1), fq2_mul_xi(t2));
  const Fq2 x2 = fq2_add(fq2_sub(fq2_sub(fq2_mul(fq2_add(a.c0, a.c2), fq2_add(b.c0, b.c2)), t0), t2), t1);
  return Fq6{x0, x1, x2};
}
inline Fq6 fq6_inv(const Fq6& a) {
  const Fq2 t0 = fq2_sub(fq2_sqr(a.c0), fq2_mul_xi(fq2_mul(a.c1, a.c2)));
  const Fq2 t1 = fq2_sub(fq2_mul_xi(fq2_sqr(a.c2)), fq2_mul(a.c0, a.c1));
  const Fq2 t2 = fq2_sub(fq2_sqr(a.c1), fq2_mul(a.c0, a.c2));
  const Fq2 d = fq2_inv(fq2_add(fq2_mul(a.c0, t0), fq2_mul_xi(fq2_add(fq2_mul(a.c2, t1), fq2_mul(a.c1, t2)))));
  return Fq6{fq2_mul(t0, d), fq2_mul(t1, d), fq2_mul(t2, d)};
}
inline Fq12 fq12_one() { return Fq12{fq6_one(), fq6_zero()}; }
inline bool fq12_eq(const Fq12& a, const Fq12& b) { return fq6_eq(a.c0, b.c0) && fq6_eq(a.c1, b.c1); }
inline Fq12 fq12_mul(const Fq12& a, const Fq12& b) {
  const Fq6 t0 = fq6_mul(a.c0, b.c0), t1 = fq6_mul(a.c1, b.c1);
  const Fq6 m = fq6_mul(fq6_add(a.c0, a.c1), fq6_add(b.c0, b.c1));
  return Fq12{fq6_add(t0, fq6_mul_v(t1)), fq6_sub(fq6_sub(m, t0), t1)};
}
inline Fq12 fq12_sqr(const Fq12& a) { return fq12_mul(a, a); }
inline Fq12 fq12_conj(const Fq12& a) { return Fq12{a.c0, fq6_neg(a.c1)}; }   // = inverse on the cyclotomic subgroup
inline Fq12 fq12_inv(const Fq12& a) {
  const Fq6 d = fq6_inv(fq6_sub(fq6_mul(a.c0, a.c0), fq6_mul_v(fq6_mul(a.c1, a.c1))));
  return Fq12{fq6_mul(a.c0, d), fq6_neg(fq6_mul(a.c1, d))};
}
inline bool fq12_is_zero(const Fq12& a) {
  const Fq2* c = &a.c0.c0;
  for (int i = 0; i < 6; i++)
    if (!fq2_is_zero(c[i])) return false;
  return true;
}

// gamma = (1 + u)^((q - 1) / 6): the Frobenius twist constant; the other five are its powers
inline const Fq2& frobenius_gamma() {
  static const Fq2 g = [] {
    uint64_t e[6];
    std::memcpy(e, kQ, sizeof(e));
    e[0] -= 1;
    uint64_t rem = 0;   // e = (q - 1) / 6 by long division from the top limb
    for (int i = 5; i >= 0; i--) {
      const u128 cur = ((u128)rem << 64) | e[i];
      e[i] = (uint64_t)(cur / 6);
      rem = (uint64_t)(cur % 6);
    }
    return fq2_pow(Fq2{fq_one(), fq_one()}, e, 6);
  }();
  return g;
}
// a^q
inline Fq12 fq12_frobenius(const Fq12& a) {
  const Fq2& g1 = frobenius_gamma();
  const Fq2 g2 = fq2_sqr(g1), g3 = fq2_mul(g2, g1), g4 = fq2_sqr(g2), g5 = fq2_mul(g4, g1);
  Fq12 r;
  r.c0.c0 = fq2_conj(a.c0.c0);
  r.c0.c1 = fq2_mul(fq2_conj(a.c0.c1), g2);
  r.c0.c2 = fq2_mul(fq2_conj(a.c0.c2), g4);
  r.c1.c0 = fq2_mul(fq2_conj(a.c1.c0), g1);
  r.c1.c1 = fq2_mul(fq2_conj(a.c1.c1), g3);
  r.c1.c2 = fq2_mul(fq2_conj(a.c1.c2), g5);
  return r;
}

static const uint64_t kBlsX = 0xd201000000010000ull;   // |x|; x is negative
inline Fq12 fq12_exp_by_x(const Fq12& a) {             // a^x for a in the cyclotomic subgroup
  Fq12 r = fq12_one();
  for (int i = 63; i >= 0; i--) {
    r = fq12_sqr(r);
    if ((kBlsX >> i) & 1) r = fq12_mul(r, a);
  }
  return fq12_conj(r);
}

// ------------------------------------------------------------------ curve points in affine form
struct G1Aff {
  Fq x, y;
  bool inf;
};
struct G2Aff {
  Fq2 x, y;
  bool inf;
};
inline G1Xyzz g1_from_affine(const G1Aff& p) {
  G1Xyzz r = g1_identity();
  if (!p.inf) { r.x = p.x; r.y = p.y; r.zz = fq_one(); r.zzz = fq_one(); }
  return r;
}
inline G2Xyzz g2_from_affine(const G2Aff& p) {
  G2Xyzz r = g2_identity();
  if (!p.inf) { r.x = p.x; r.y = p.y; r.zz = fq2_one(); r.zzz = fq2_one(); }
  return r;
}
inline G2Xyzz g2_mul_scalar(const G2Xyzz& p, const uint64_t* k, int nlimbs) {
  G2Xyzz acc = g2_identity();
  for (int i = nlimbs * 64 - 1; i >= 0; i--) {
    acc = g2_dbl(acc);
    if ((k[i >> 6] >> (i & 63)) & 1) acc = g2_add(acc, p);
  }
  return acc;
}
static const uint64_t kFrModulus[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull,
                                       0x73eda753299d7d48ull};
inline bool g1_in_subgroup(const G1Aff& p) {
  if (p.inf) return true;
  return fq_is_zero(g1_mul_scalar(g1_from_affine(p), reinterpret_cast<const uint32_t*>(kFrModulus)).zz);
}
inline bool g2_in_subgroup(const G2Aff& p) {
  if (p.inf) return true;
  return fq2_is_zero(g2_mul_scalar(g2_from_affine(p), kFrModulus, 4).zz);
}

// zcash-format decompression with arkworks' validation (on curve, in the r-order subgroup)
inline bool g1_deserialize(const uint8_t* src, G1Aff* out) {
  const uint8_t flags = src[0];
  if (!(flags & 0x80)) return false;                      // uncompressed encoding is not used on this path
  if (flags & 0x40) {
    if (flags & 0x3f) return false;
    for (int i = 1; i < 48; i++) if (src[i]) return false;
    out->inf = true; out->x = out->y = fq_zero();
    return true;
  }
  Fq x, y;
  if (!fq_read_be(src, 0x1f, &x)) return false;
  if (!fq_sqrt(fq_add(fq_mul(fq_sqr(x), x), fq_from_u64(4)), &y)) return false;
  if (fq_is_larger(y) != ((flags & 0x20) != 0)) y = fq_neg(y);
  *out = G1Aff{x, y, false};
  return g1_in_subgroup(*out);
}
inline bool fq2_is_larger(const Fq2& y) { return fq_is_zero(y.c1) ? fq_is_larger(y.c0) : fq_is_larger(y.c1); }
inline bool g2_deserialize(const uint8_t* src, G2Aff* out) {
  const uint8_t flags = src[0];
  if (!(flags & 0x80)) return false;
  if (flags & 0x40) {
    if (flags & 0x3f) return false;
    for (int i = 1; i < 96; i++) if (src[i]) return false;
    out->inf = true; out->x = out->y = fq2_zero();
    return true;
  }
  Fq2 x, y;
  if (!fq_read_be(src, 0x1f, &x.c1) || !fq_read_be(src + 48, 0xff, &x.c0)) return false;
  const Fq four = fq_from_u64(4);
  if (!fq2_sqrt(fq2_add(fq2_mul(fq2_sqr(x), x), Fq2{four, four}), &y)) return false;
  if (fq2_is_larger(y) != ((flags & 0x20) != 0)) y = fq2_neg(y);
  *out = G2Aff{x, y, false};
  return g2_in_subgroup(*out);
}
inline void g1_serialize_affine(uint8_t* dst, const G1Aff& p) { g1_serialize(dst, g1_from_affine(p)); }
inline void g2_serialize_affine(uint8_t* dst, const G2Aff& p) { g2_serialize(dst, g2_from_affine(p)); }

// ------------------------------------------------------------------ G2Prepared (ark-ec models/bls12/g2.rs)
struct EllCoeff {
  Fq2 c0, c1, c2;
};
struct G2Prepared {
  std::vector<EllCoeff> ell;
  bool inf = true;
};
inline G2Prepared g2_prepare(const G2Aff& q) {
  G2Prepared out;
  if (q.inf) return out;
  out.inf = false;
  const Fq two_inv = fq_inv(fq_from_u64(2));
  const Fq four = fq_from_u64(4);
  const Fq2 coeff_b{four, four};
  Fq2 rx = q.x, ry = q.y, rz = fq2_one();
  auto dbl_step = [&]() {
    const Fq2 a = fq2_mul_fq(fq2_mul(rx, ry), two_inv);
    const Fq2 b = fq2_sqr(ry), c = fq2_sqr(rz);
    const Fq2 e = fq2_mul(coeff_b, fq2_add(fq2_dbl(c), c));
    const Fq2 f = fq2_add(fq2_dbl(e), e);
    const Fq2 g = fq2_mul_fq(fq2_add(b, f), two_inv);
    const Fq2 h = fq2_sub(fq2_sqr(fq2_add(ry, rz)), fq2_add(b, c));
    const Fq2 i = fq2_sub(e, b);
    const Fq2 j = fq2_sqr(rx);
    const Fq2 e2 = fq2_sqr(e);
    rx = fq2_mul(a, fq2_sub(b, f));
    ry = fq2_sub(fq2_sqr(g), fq2_add(fq2_dbl(e2), e2));
    rz = fq2_mul(b, h);
    out.ell.push_back(EllCoeff{i, fq2_add(fq2_dbl(j), j), fq2_neg(h)});     // M-type twist order
  };
  auto add_step = [&]() {
    const Fq2 theta = fq2_sub(ry, fq2_mul(q.y, rz));
    const Fq2 lambda = fq2_sub(rx, fq2_mul(q.x, rz));
    const Fq2 c = fq2_sqr(theta), d = fq2_sqr(lambda);
    const Fq2 e = fq2_mul(lambda, d), f = fq2_mul(rz, c), g = fq2_mul(rx, d);
    const Fq2 h = fq2_sub(fq2_add(e, f), fq2_dbl(g));
    rx = fq2_mul(lambda, h);
    ry = fq2_sub(fq2_mul(theta, fq2_sub(g, h)), fq2_mul(e, ry));
    rz = fq2_mul(rz, e);
    const Fq2 j = fq2_sub(fq2_mul(theta, q.x), fq2_mul(lambda, q.y));
    out.ell.push_back(EllCoeff{j, fq2_neg(theta), lambda});
  };
  for (int i = 62; i >= 0; i--) {   // bits of |x| below the leading one
    dbl_step();
    if ((kBlsX >> i) & 1) add_step();
  }
  return out;
}

// f * (c0 + c1 v + c4 v w): the sparse line value of an M-type twist
inline Fq12 fq12_mul_by_014(const Fq12& f, const Fq2& c0, const Fq2& c1, const Fq2& c4) {
  Fq12 s;
  s.c0 = Fq6{c0, c1, fq2_zero()};
  s.c1 = Fq6{fq2_zero(), c4, fq2_zero()};
  return fq12_mul(f, s);
}

struct PairingInput {
  G1Aff p;
  const G2Prepared* q;
};
inline Fq12 multi_miller_loop(const std::vector<PairingInput>& in) {
  std::vector<PairingInput> pairs;
  for (const auto& x : in)
    if (!x.p.inf && !x.q->inf) pairs.push_back(x);
  std::vector<size_t> at(pairs.size(), 0);
  Fq12 f = fq12_one();
  auto ell = [&](size_t k) {
    const EllCoeff& c = pairs[k].q->ell[at[k]++];
    f = fq12_mul_by_014(f, c.c0, fq2_mul_fq(c.c1, pairs[k].p.x), fq2_mul_fq(c.c2, pairs[k].p.y));
  };
  for (int i = 62; i >= 0; i--) {
    f = fq12_sqr(f);
    for (size_t k = 0; k < pairs.size(); k++) ell(k);
    if ((kBlsX >> i) & 1)
      for (size_t k = 0; k < pairs.size(); k++) ell(k);
  }
  return fq12_conj(f);   // x < 0
}

// ark-ec Bls12::final_exponentiation; returns false for f = 0
inline bool final_exponentiation(const Fq12& f, Fq12* out) {
  if (fq12_is_zero(f)) return false;
  // easy part: f^((q^6 - 1)(q^2 + 1))
  Fq12 r = fq12_mul(fq12_conj(f), fq12_inv(f));
  r = fq12_mul(fq12_frobenius(fq12_frobenius(r)), r);
  // hard part: r^(3 (q^4 - q^2 + 1) / r_order) = r^3 * (r^((x-1)^2 (x+q)))^(x^2 + q^2 - 1)
  Fq12 y0 = fq12_sqr(r);
  Fq12 y1 = fq12_exp_by_x(r);
  Fq12 y2 = fq12_conj(r);
  y1 = fq12_mul(y1, y2);
  y2 = fq12_exp_by_x(y1);
  y1 = fq12_conj(y1);
  y1 = fq12_mul(y1, y2);
  y2 = fq12_exp_by_x(y1);
  y1 = fq12_frobenius(y1);
  y1 = fq12_mul(y1, y2);
  r = fq12_mul(r, y0);
  y0 = fq12_exp_by_x(y1);
  y2 = fq12_exp_by_x(y0);
  y0 = fq12_frobenius(fq12_frobenius(y1));
  y1 = fq12_conj(y1);
  y1 = fq12_mul(y1, y2);
  y1 = fq12_mul(y1, y0);
  *out = fq12_mul(r, y1);
  return true;
}

// Fr element (Montgomery limbs, R = 2^256) -> canonical integer: one Montgomery reduction of (a, 0)
static const uint64_t kFrInv = 0xfffffffeffffffffull;   // -r^-1 mod 2^64
inline bool fr_mont_to_canonical(const uint64_t a[4], uint64_t out[4]) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] != kFrModulus[i]) { if (a[i] > kFrModulus[i]) return false; break; }
    if (i == 0) return false;   // a == r
  }
  uint64_t t[5] = {a[0], a[1], a[2], a[3], 0};
  for (int i = 0; i < 4; i++) {
    const uint64_t m = t[0] * kFrInv;
    u128 x = (u128)m * kFrModulus[0] + t[0];
    uint64_t c = (uint64_t)(x >> 64);
    for (int j = 1; j < 4; j++) {
      x = (u128)m * kFrModulus[j] + t[j] + c;
      t[j - 1] = (uint64_t)x;
      c = (uint64_t)(x >> 64);
    }
    x = (u128)t[4] + c;
    t[3] = (uint64_t)x;
    t[4] = (uint64_t)(x >> 64);
  }
  bool ge = t[4] != 0;
  if (!ge) {
    ge = true;
    for (int i = 3; i >= 0; i--)
      if (t[i] != kFrModulus[i]) { ge = t[i] > kFrModulus[i]; break; }
  }
  if (ge) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)t[i] - kFrModulus[i] - borrow;
      t[i] = (uint64_t)d;
      borrow = (uint64_t)(d >> 64) & 1;
    }
  }
  std::memcpy(out, t, 32);
  return true;
}

// ------------------------------------------------------------------ PreparedVerifyingKey and its wire format
struct PreparedVk {
  G1Aff alpha_g1;
  G2Aff beta_g2, gamma_g2, delta_g2;
  std::vector<G1Aff> gamma_abc_g1;
  Fq12 alpha_g1_beta_g2;
  G2Prepared gamma_g2_neg_pc, delta_g2_neg_pc;
};
inline G2Aff g2_neg_affine(const G2Aff& p) { return p.inf ? p : G2Aff{p.x, fq2_neg(p.y), false}; }

// ark_groth16::prepare_verifying_key
inline bool prepare_vk(PreparedVk* k) {
  const G2Prepared beta = g2_prepare(k->beta_g2);
  if (!final_exponentiation(multi_miller_loop({PairingInput{k->alpha_g1, &beta}}), &k->alpha_g1_beta_g2)) return false;
  k->gamma_g2_neg_pc = g2_prepare(g2_neg_affine(k->gamma_g2));
  k->delta_g2_neg_pc = g2_prepare(g2_neg_affine(k->delta_g2));
  return true;
}

inline void put_u64(std::vector<uint8_t>& out, uint64_t v) {
  for (int i = 0; i < 8; i++) out.push_back((uint8_t)(v >> (8 * i)));
}
inline void put_fq2(std::vector<uint8_t>& out, const Fq2& a) {
  const size_t at = out.size();
  out.resize(at + 96);
  fq_write_le(&out[at], a.c0);
  fq_write_le(&out[at + 48], a.c1);
}
inline void put_prepared(std::vector<uint8_t>& out, const G2Prepared& p) {
  put_u64(out, p.ell.size());
  for (const auto& c : p.ell) { put_fq2(out, c.c0); put_fq2(out, c.c1); put_fq2(out, c.c2); }
  out.push_back(p.inf ? 1 : 0);
}
// PreparedVerifyingKey::serialize_compressed: vk | alpha_g1_beta_g2 | gamma_g2_neg_pc | delta_g2_neg_pc
inline std::vector<uint8_t> pvk_serialize(const PreparedVk& k) {
  std::vector<uint8_t> out;
  size_t at = 0;
  auto grow = [&](size_t n) { at = out.size(); out.resize(at + n); return &out[at]; };
  g1_serialize_affine(grow(48), k.alpha_g1);
  g2_serialize_affine(grow(96), k.beta_g2);
  g2_serialize_affine(grow(96), k.gamma_g2);
  g2_serialize_affine(grow(96), k.delta_g2);
  put_u64(out, k.gamma_abc_g1.size());
  for (const auto& p : k.gamma_abc_g1) g1_serialize_affine(grow(48), p);
  const Fq2* c = &k.alpha_g1_beta_g2.c0.c0;
  for (int i = 0; i < 6; i++) put_fq2(out, c[i]);
  put_prepared(out, k.gamma_g2_neg_pc);
  put_prepared(out, k.delta_g2_neg_pc);
  return out;
}

struct Reader {
  const uint8_t* p;
  size_t left;
  bool take(size_t n, const uint8_t** out) {
    if (left < n) return false;
    *out = p; p += n; left -= n;
    return true;
  }
  bool u64(uint64_t* v) {
    const uint8_t* s;
    if (!take(8, &s)) return false;
    *v = 0;
    for (int i = 0; i < 8; i++) *v |= (uint64_t)s[i] << (8 * i);
    return true;
  }
  bool fq2(Fq2* a) {
    const uint8_t* s;
    return take(96, &s) && fq_read_le(s, &a->c0) && fq_read_le(s + 48, &a->c1);
  }
  bool prepared(G2Prepared* g) {
    uint64_t n;
    if (!u64(&n) || n > 4096) return false;
    g->ell.resize(n);
    for (auto& c : g->ell)
      if (!fq2(&c.c0) || !fq2(&c.c1) || !fq2(&c.c2)) return false;
    const uint8_t* s;
    if (!take(1, &s) || *s > 1) return false;
    g->inf = *s == 1;
    return true;
  }
};
inline bool pvk_deserialize(const uint8_t* bytes, size_t len, PreparedVk* k) {
  Reader r{bytes, len};
  const uint8_t* s;
  if (!r.take(48, &s) || !g1_deserialize(s, &k->alpha_g1)) return false;
  if (!r.take(96, &s) || !g2_deserialize(s, &k->beta_g2)) return false;
  if (!r.take(96, &s) || !g2_deserialize(s, &k->gamma_g2)) return false;
  if (!r.take(96, &s) || !g2_deserialize(s, &k->delta_g2)) return false;
  uint64_t n;
  if (!r.u64(&n) || n > (1u << 24)) return false;
  k->gamma_abc_g1.resize(n);
  for (auto& p : k->gamma_abc_g1)
    if (!r.take(48, &s) || !g1_deserialize(s, &p)) return false;
  Fq2* c = &k->alpha_g1_beta_g2.c0.c0;
  for (int i = 0; i < 6; i++)
    if (!r.fq2(&c[i])) return false;
  return r.prepared(&k->gamma_g2_neg_pc) && r.prepared(&k->delta_g2_neg_pc) && r.left == 0;
}

// Groth16::verify_with_processed_vk.  inputs: canonical 4-limb integers (into_bigint()).
// returns 0 = ok (see *valid), 1 = malformed (input count / encodings), as SynthesisError::MalformedVerifyingKey
inline int verify_with_processed_vk(const PreparedVk& k, const uint64_t* inputs, uint64_t n_inputs, const uint8_t proof[192],
                                    bool* valid) {
  *valid = false;
  if (n_inputs + 1 != k.gamma_abc_g1.size()) return 1;
  for (const G2Prepared* g : {&k.gamma_g2_neg_pc, &k.delta_g2_neg_pc})
    if (!g->inf && g->ell.size() != 68) return 1;
  G1Aff a, c;
  G2Aff b;
  if (!g1_deserialize(proof, &a) || !g2_deserialize(proof + 48, &b) || !g1_deserialize(proof + 144, &c)) return 1;
  G1Xyzz acc = g1_from_affine(k.gamma_abc_g1[0]);
  for (uint64_t i = 0; i < n_inputs; i++)
    acc = g1_add(acc, g1_mul_scalar(g1_from_affine(k.gamma_abc_g1[i + 1]), reinterpret_cast<const uint32_t*>(inputs + 4 * i)));
  G1Aff prepared_inputs;
  prepared_inputs.inf = g1_to_affine(acc, &prepared_inputs.x, &prepared_inputs.y);
  const G2Prepared pb = g2_prepare(b);
  const Fq12 ml = multi_miller_loop({PairingInput{a, &pb}, PairingInput{prepared_inputs, &k.gamma_g2_neg_pc},
                                     PairingInput{c, &k.delta_g2_neg_pc}});
  Fq12 test;
  if (!final_exponentiation(ml, &test)) return 0;   // UnexpectedIdentity: not a valid proof
  *valid = fq12_eq(test, k.alpha_g1_beta_g2);
  return 0;
}

}  // namespace host
}  // namespace b2z
