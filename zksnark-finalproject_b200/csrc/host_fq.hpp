// Host-side Fq / Fq2 arithmetic (64-bit limbs, unsigned __int128) for the O(1)
// epilogue of a proof: the two scalar multiplications s*A and r*B1 (arkworks'
// `mul_bigint` on freshly computed points: ~380 dependent point operations each, 0.2 ms
// on a CPU core, 1.5 ms even with lane-cooperative arithmetic on the GPU -- they run here
// WHILE the GPU is still busy with the remaining MSMs), a few point additions, three
// affine normalisations (one field inversion each) and the zcash-format serialization --
// what ark-groth16 does with `into_affine()` and ark-serialize's `serialize_compressed`
// after its MSMs (SURVEY.md section 7, step 9 keeps this on the host).  A 381-bit inversion is
// ~25 us here versus ~0.6 ms for a lone GPU thread, and the three results have to
// cross PCIe anyway.  Nothing proportional to the circuit size runs on the host.
#pragma once
#include <cstdint>
#include <cstring>

namespace b2z {
namespace host {

typedef unsigned __int128 u128;

struct Fq {
  uint64_t l[6];
};

static const uint64_t kQ[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                               0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
static const uint64_t kQInv = 0x89f3fffcfffcfffdull;   // -q^-1 mod 2^64
// R mod q (Montgomery one)
static const uint64_t kOne[6] = {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull,
                                 0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull};

inline bool fq_geq_q(const uint64_t* a) {
  for (int i = 5; i >= 0; i--)
    if (a[i] != kQ[i]) return a[i] > kQ[i];
  return true;
}
inline void fq_sub_q(uint64_t* a) {
  uint64_t borrow = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a[i] - kQ[i] - borrow;
    a[i] = (uint64_t)t;
    borrow = (uint64_t)(t >> 64) & 1;
  }
}
inline Fq fq_zero() { Fq r; std::memset(r.l, 0, sizeof(r.l)); return r; }
inline Fq fq_one() { Fq r; std::memcpy(r.l, kOne, sizeof(r.l)); return r; }
inline bool fq_is_zero(const Fq& a) { uint64_t z = 0; for (int i = 0; i < 6; i++) z |= a.l[i]; return z == 0; }
inline bool fq_eq(const Fq& a, const Fq& b) { return std::memcmp(a.l, b.l, sizeof(a.l)) == 0; }
// canonicalise a lazily reduced device value (< 2q)
inline Fq fq_from_lazy(const uint32_t* limbs32) {
  Fq r;
  std::memcpy(r.l, limbs32, 48);
  if (fq_geq_q(r.l)) fq_sub_q(r.l);
  return r;
}
inline Fq fq_add(const Fq& a, const Fq& b) {
  Fq r;
  uint64_t c = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a.l[i] + b.l[i] + c;
    r.l[i] = (uint64_t)t;
    c = (uint64_t)(t >> 64);
  }
  if (c || fq_geq_q(r.l)) fq_sub_q(r.l);
  return r;
}
inline Fq fq_sub(const Fq& a, const Fq& b) {
  Fq r;
  uint64_t borrow = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a.l[i] - b.l[i] - borrow;
    r.l[i] = (uint64_t)t;
    borrow = (uint64_t)(t >> 64) & 1;
  }
  if (borrow) {
    uint64_t c = 0;
    for (int i = 0; i < 6; i++) {
      u128 t = (u128)r.l[i] + kQ[i] + c;
      r.l[i] = (uint64_t)t;
      c = (uint64_t)(t >> 64);
    }
  }
  return r;
}
inline Fq fq_neg(const Fq& a) { return fq_is_zero(a) ? a : fq_sub(fq_zero(), a); }
inline Fq fq_dbl(const Fq& a) { return fq_add(a, a); }
inline Fq fq_mul(const Fq& a, const Fq& b) {   // CIOS Montgomery product
  uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 6; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 6; j++) {
      u128 x = (u128)a.l[j] * b.l[i] + t[j] + c;
      t[j] = (uint64_t)x;
      c = (uint64_t)(x >> 64);
    }
    u128 x = (u128)t[6] + c;
    t[6] = (uint64_t)x;
    t[7] = (uint64_t)(x >> 64);
    const uint64_t m = t[0] * kQInv;
    x = (u128)m * kQ[0] + t[0];
    c = (uint64_t)(x >> 64);
    for (int j = 1; j < 6; j++) {
      x = (u128)m * kQ[j] + t[j] + c;
      t[j - 1] = (uint64_t)x;
      c = (uint64_t)(x >> 64);
    }
    x = (u128)t[6] + c;
    t[5] = (uint64_t)x;
    t[6] = t[7] + (uint64_t)(x >> 64);
  }
  Fq r;
  std::memcpy(r.l, t, sizeof(r.l));
  if (t[6] || fq_geq_q(r.l)) fq_sub_q(r.l);
  return r;
}
inline Fq fq_sqr(const Fq& a) { return fq_mul(a, a); }
inline Fq fq_inv(const Fq& a) {   // a^(q-2)
  uint64_t e[6];
  std::memcpy(e, kQ, sizeof(e));
  e[0] -= 2;
  Fq r = fq_one();
  for (int i = 6 * 64 - 1; i >= 0; i--) {
    r = fq_sqr(r);
    if ((e[i >> 6] >> (i & 63)) & 1) r = fq_mul(r, a);
  }
  return r;
}
inline void fq_to_canonical(const Fq& a, uint64_t out[6]) {
  Fq o = fq_zero();
  o.l[0] = 1;
  Fq c = fq_mul(a, o);
  std::memcpy(out, c.l, 48);
}
// canonical value > (q - 1) / 2 ?
inline bool fq_is_larger(const Fq& a) {
  uint64_t c[6], d[6];
  fq_to_canonical(a, c);
  uint64_t carry = 0;
  for (int i = 0; i < 6; i++) { d[i] = (c[i] << 1) | carry; carry = c[i] >> 63; }
  for (int i = 5; i >= 0; i--)
    if (d[i] != kQ[i]) return d[i] > kQ[i];
  return false;
}
inline void fq_write_be(uint8_t* dst, const Fq& a) {
  uint64_t c[6];
  fq_to_canonical(a, c);
  for (int i = 0; i < 6; i++)
    for (int b = 0; b < 8; b++) dst[8 * i + b] = (uint8_t)(c[5 - i] >> (56 - 8 * b));
}

struct Fq2 {
  Fq c0, c1;
};
inline Fq2 fq2_mul(const Fq2& a, const Fq2& b) {
  const Fq v0 = fq_mul(a.c0, b.c0), v1 = fq_mul(a.c1, b.c1);
  const Fq s = fq_mul(fq_add(a.c0, a.c1), fq_add(b.c0, b.c1));
  return Fq2{fq_sub(v0, v1), fq_sub(fq_sub(s, v0), v1)};
}
inline Fq2 fq2_sqr(const Fq2& a) { return fq2_mul(a, a); }
inline Fq2 fq2_inv(const Fq2& a) {
  const Fq d = fq_inv(fq_add(fq_sqr(a.c0), fq_sqr(a.c1)));
  return Fq2{fq_mul(a.c0, d), fq_neg(fq_mul(a.c1, d))};
}
inline bool fq2_is_zero(const Fq2& a) { return fq_is_zero(a.c0) && fq_is_zero(a.c1); }

// ---- G1 in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ), identity: ZZ = 0
struct G1Xyzz {
  Fq x, y, zz, zzz;
};
struct G2Xyzz {
  Fq2 x, y, zz, zzz;
};

inline G1Xyzz g1_dbl(const G1Xyzz& p) {
  if (fq_is_zero(p.zz)) return p;
  const Fq u = fq_dbl(p.y), v = fq_sqr(u), w = fq_mul(u, v), s = fq_mul(p.x, v), xx = fq_sqr(p.x);
  const Fq m = fq_add(fq_dbl(xx), xx);
  G1Xyzz r;
  r.x = fq_sub(fq_sqr(m), fq_dbl(s));
  r.y = fq_sub(fq_mul(m, fq_sub(s, r.x)), fq_mul(w, p.y));
  r.zz = fq_mul(v, p.zz);
  r.zzz = fq_mul(w, p.zzz);
  return r;
}
inline G1Xyzz g1_add(const G1Xyzz& a, const G1Xyzz& b) {
  if (fq_is_zero(a.zz)) return b;
  if (fq_is_zero(b.zz)) return a;
  const Fq u1 = fq_mul(a.x, b.zz), u2 = fq_mul(b.x, a.zz), s1 = fq_mul(a.y, b.zzz), s2 = fq_mul(b.y, a.zzz);
  const Fq p = fq_sub(u2, u1), r = fq_sub(s2, s1);
  if (fq_is_zero(p)) {
    if (fq_is_zero(r)) return g1_dbl(a);
    G1Xyzz id;
    id.x = id.y = id.zz = id.zzz = fq_zero();
    return id;
  }
  const Fq pp = fq_sqr(p), ppp = fq_mul(p, pp), q = fq_mul(u1, pp);
  G1Xyzz o;
  o.x = fq_sub(fq_sub(fq_sqr(r), ppp), fq_dbl(q));
  o.y = fq_sub(fq_mul(r, fq_sub(q, o.x)), fq_mul(s1, ppp));
  o.zz = fq_mul(fq_mul(a.zz, b.zz), pp);
  o.zzz = fq_mul(fq_mul(a.zzz, b.zzz), ppp);
  return o;
}

inline Fq2 fq2_add(const Fq2& a, const Fq2& b) { return Fq2{fq_add(a.c0, b.c0), fq_add(a.c1, b.c1)}; }
inline Fq2 fq2_sub(const Fq2& a, const Fq2& b) { return Fq2{fq_sub(a.c0, b.c0), fq_sub(a.c1, b.c1)}; }
inline Fq2 fq2_dbl(const Fq2& a) { return fq2_add(a, a); }
inline G2Xyzz g2_dbl(const G2Xyzz& p) {
  if (fq2_is_zero(p.zz)) return p;
  const Fq2 u = fq2_dbl(p.y), v = fq2_sqr(u), w = fq2_mul(u, v), s = fq2_mul(p.x, v), xx = fq2_sqr(p.x);
  const Fq2 m = fq2_add(fq2_dbl(xx), xx);
  G2Xyzz r;
  r.x = fq2_sub(fq2_sqr(m), fq2_dbl(s));
  r.y = fq2_sub(fq2_mul(m, fq2_sub(s, r.x)), fq2_mul(w, p.y));
  r.zz = fq2_mul(v, p.zz);
  r.zzz = fq2_mul(w, p.zzz);
  return r;
}
inline G2Xyzz g2_add(const G2Xyzz& a, const G2Xyzz& b) {
  if (fq2_is_zero(a.zz)) return b;
  if (fq2_is_zero(b.zz)) return a;
  const Fq2 u1 = fq2_mul(a.x, b.zz), u2 = fq2_mul(b.x, a.zz), s1 = fq2_mul(a.y, b.zzz), s2 = fq2_mul(b.y, a.zzz);
  const Fq2 p = fq2_sub(u2, u1), r = fq2_sub(s2, s1);
  if (fq2_is_zero(p)) {
    if (fq2_is_zero(r)) return g2_dbl(a);
    G2Xyzz id;
    std::memset(&id, 0, sizeof(id));
    return id;
  }
  const Fq2 pp = fq2_sqr(p), ppp = fq2_mul(p, pp), q = fq2_mul(u1, pp);
  G2Xyzz o;
  o.x = fq2_sub(fq2_sub(fq2_sqr(r), ppp), fq2_dbl(q));
  o.y = fq2_sub(fq2_mul(r, fq2_sub(q, o.x)), fq2_mul(s1, ppp));
  o.zz = fq2_mul(fq2_mul(a.zz, b.zz), pp);
  o.zzz = fq2_mul(fq2_mul(a.zzz, b.zzz), ppp);
  return o;
}

// k * p for a canonical 256-bit little-endian scalar (8 x u32): arkworks' `mul_bigint`
// (double-and-add, MSB first).  ~0.2 ms on one core.
inline G1Xyzz g1_mul_scalar(const G1Xyzz& p, const uint32_t k[8]) {
  G1Xyzz acc;
  std::memset(&acc, 0, sizeof(acc));
  int top = 255;
  while (top >= 0 && !((k[top >> 5] >> (top & 31)) & 1)) top--;
  for (int i = top; i >= 0; i--) {
    acc = g1_dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) acc = g1_add(acc, p);
  }
  return acc;
}
inline void g1_to_device_layout(const G1Xyzz& p, uint32_t* w) { std::memcpy(w, &p, sizeof(p)); }

// affine normalisation (Montgomery coordinates); returns true for the identity
inline bool g1_to_affine(const G1Xyzz& p, Fq* x, Fq* y) {
  if (fq_is_zero(p.zz)) { *x = fq_zero(); *y = fq_zero(); return true; }
  const Fq i = fq_inv(p.zzz), i2 = fq_sqr(i);
  *x = fq_mul(fq_mul(p.x, fq_sqr(p.zz)), i2);
  *y = fq_mul(p.y, i);
  return false;
}

// ark-bls12-381 0.4 `serialize_compressed` (zcash encoding)
inline void g1_serialize(uint8_t* dst, const G1Xyzz& p) {
  Fq x, y;
  if (g1_to_affine(p, &x, &y)) {
    std::memset(dst, 0, 48);
    dst[0] = 0xC0;
    return;
  }
  fq_write_be(dst, x);
  dst[0] |= 0x80;
  if (fq_is_larger(y)) dst[0] |= 0x20;
}
inline bool g2_to_affine(const G2Xyzz& p, Fq2* x, Fq2* y) {
  if (fq2_is_zero(p.zz)) { x->c0 = x->c1 = y->c0 = y->c1 = fq_zero(); return true; }
  const Fq2 i = fq2_inv(p.zzz), i2 = fq2_sqr(i);
  *x = fq2_mul(fq2_mul(p.x, fq2_sqr(p.zz)), i2);
  *y = fq2_mul(p.y, i);
  return false;
}
inline void g2_serialize(uint8_t* dst, const G2Xyzz& p) {
  Fq2 x, y;
  if (g2_to_affine(p, &x, &y)) {
    std::memset(dst, 0, 96);
    dst[0] = 0xC0;
    return;
  }
  fq_write_be(dst, x.c1);
  fq_write_be(dst + 48, x.c0);
  dst[0] |= 0x80;
  const bool larger = fq_is_zero(y.c1) ? fq_is_larger(y.c0) : fq_is_larger(y.c1);
  if (larger) dst[0] |= 0x20;
}

// Last step of the MSM bucket reduction, done here because it is strictly serial (msm.hpp,
// MsmHostPlanes): planes[0] + 2^chunk_log * sum_{k >= 1} 2^(k-1) planes[k].
inline G1Xyzz g1_identity() { G1Xyzz r; r.x = r.y = r.zz = r.zzz = fq_zero(); return r; }
inline G2Xyzz g2_identity() { G2Xyzz r; r.x = r.y = r.zz = r.zzz = Fq2{fq_zero(), fq_zero()}; return r; }
template <class P, class FromDev, class Dbl, class Add>
inline P planes_horner_generic(const uint32_t* dev_planes, uint32_t words_per_point, uint32_t nplanes, uint32_t chunk_log,
                               P identity, FromDev from_dev, Dbl dbl, Add add) {
  if (nplanes == 0) return identity;
  P acc = from_dev(dev_planes);
  if (nplanes > 1) {
    P h = from_dev(dev_planes + (size_t)(nplanes - 1) * words_per_point);
    for (uint32_t k = nplanes - 1; k-- > 1;) h = add(dbl(h), from_dev(dev_planes + (size_t)k * words_per_point));
    for (uint32_t i = 0; i < chunk_log; i++) h = dbl(h);
    acc = add(acc, h);
  }
  return acc;
}

// device XYZZ (32-bit limbs, lazily reduced) -> host structs
inline G1Xyzz g1_from_device(const uint32_t* w) {
  G1Xyzz p;
  p.x = fq_from_lazy(w);
  p.y = fq_from_lazy(w + 12);
  p.zz = fq_from_lazy(w + 24);
  p.zzz = fq_from_lazy(w + 36);
  return p;
}
inline G2Xyzz g2_from_device(const uint32_t* w) {
  G2Xyzz p;
  Fq* f = reinterpret_cast<Fq*>(&p);
  for (int i = 0; i < 8; i++) f[i] = fq_from_lazy(w + 12 * i);
  return p;
}

inline G1Xyzz g1_planes_horner(const uint32_t* planes, uint32_t nplanes, uint32_t chunk_log) {
  return planes_horner_generic<G1Xyzz>(planes, 48, nplanes, chunk_log, g1_identity(), g1_from_device, g1_dbl, g1_add);
}
inline G2Xyzz g2_planes_horner(const uint32_t* planes, uint32_t nplanes, uint32_t chunk_log) {
  return planes_horner_generic<G2Xyzz>(planes, 96, nplanes, chunk_log, g2_identity(), g2_from_device, g2_dbl, g2_add);
}
inline void g2_to_device_layout(const G2Xyzz& p, uint32_t* w) { std::memcpy(w, &p, sizeof(p)); }

}  // namespace host
}  // namespace b2z
