// BLS12-381 G1 / G2 group arithmetic for the bucket MSM (a = 0 short Weierstrass).
//
// Replaces, on the device, ark_ec::short_weierstrass::{Affine, Projective} as
// used by VariableBaseMSM::msm_bigint (ark-ec ^0.4.2, /root/reference/Cargo.toml:14).
// The curve template is instantiated with Fq (G1) and Fq2 (G2).
//
// Accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2): a mixed addition of an affine base costs 8M + 2S and needs no
// field inversion, which is what the bucket accumulation loop is made of.  The
// identity is ZZ = 0.  All special cases (identity operand, P + P, P + (-P)) are
// handled, because pk queries do contain repeated and identity points.
#pragma once
#include "mont.cuh"

namespace b2z {

// ---------------------------------------------------------------- Fq2 = Fq[u]/(u^2+1)
struct Fq2El {
  FqEl c0, c1;
};

struct Fq2 {
  using El = Fq2El;
  static B2Z_HD El zero() { return El{Fq::zero(), Fq::zero()}; }
  static B2Z_HD El one() { return El{Fq::one(), Fq::zero()}; }
  static B2Z_HD El add(const El& a, const El& b) { return El{Fq::add(a.c0, b.c0), Fq::add(a.c1, b.c1)}; }
  static B2Z_HD El sub(const El& a, const El& b) { return El{Fq::sub(a.c0, b.c0), Fq::sub(a.c1, b.c1)}; }
  static B2Z_HD El dbl(const El& a) { return El{Fq::dbl(a.c0), Fq::dbl(a.c1)}; }
  static B2Z_HD El neg(const El& a) { return El{Fq::neg(a.c0), Fq::neg(a.c1)}; }
  static B2Z_HD El mul(const El& a, const El& b) {   // Karatsuba, 3 base products
    const FqEl v0 = Fq::mul(a.c0, b.c0);
    const FqEl v1 = Fq::mul(a.c1, b.c1);
    const FqEl s = Fq::mul(Fq::add(a.c0, a.c1), Fq::add(b.c0, b.c1));
    return El{Fq::sub(v0, v1), Fq::sub(Fq::sub(s, v0), v1)};
  }
  static B2Z_HD El sqr(const El& a) {                // (a0+a1)(a0-a1), 2 a0 a1
    const FqEl t = Fq::mul(Fq::add(a.c0, a.c1), Fq::sub(a.c0, a.c1));
    const FqEl m = Fq::mul(a.c0, a.c1);
    return El{t, Fq::dbl(m)};
  }
  static B2Z_HD El reduce(const El& a) { return El{Fq::reduce(a.c0), Fq::reduce(a.c1)}; }
  static B2Z_HD bool is_zero(const El& a) { return Fq::is_zero(a.c0) && Fq::is_zero(a.c1); }
  static B2Z_HD bool eq(const El& a, const El& b) { return Fq::eq(a.c0, b.c0) && Fq::eq(a.c1, b.c1); }
  static B2Z_HD El inv(const El& a) {                // conj(a) / (a0^2 + a1^2)
    const FqEl d = Fq::inv(Fq::add(Fq::sqr(a.c0), Fq::sqr(a.c1)));
    return El{Fq::mul(a.c0, d), Fq::neg(Fq::mul(a.c1, d))};
  }
};

// ---------------------------------------------------------------- points
template <class F>
struct AffinePoint {
  typename F::El x, y;
};

template <class F>
struct XyzzPoint {
  typename F::El x, y, zz, zzz;
};

template <class F>
struct Curve {
  using Fld = F;
  using El = typename F::El;
  using Affine = AffinePoint<F>;
  using Xyzz = XyzzPoint<F>;

  static B2Z_HD Xyzz identity() { return Xyzz{F::zero(), F::zero(), F::zero(), F::zero()}; }
  static B2Z_HD bool is_identity(const Xyzz& p) { return F::is_zero(p.zz); }
  static B2Z_HD Xyzz from_affine(const Affine& p) { return Xyzz{p.x, p.y, F::one(), F::one()}; }
  static B2Z_HD Affine neg(const Affine& p) { return Affine{p.x, F::neg(p.y)}; }
  static B2Z_HD Xyzz neg(const Xyzz& p) { return Xyzz{p.x, F::neg(p.y), p.zz, p.zzz}; }

  // 2 * (affine p), p != identity
  static B2Z_HD Xyzz dbl_affine_inl(const Affine& p) {
    const El u = F::dbl(p.y);
    const El v = F::sqr(u);
    const El w = F::mul(u, v);
    const El s = F::mul(p.x, v);
    const El xx = F::sqr(p.x);
    const El m = F::add(F::dbl(xx), xx);
    Xyzz r;
    r.x = F::sub(F::sqr(m), F::dbl(s));
    r.y = F::sub(F::mul(m, F::sub(s, r.x)), F::mul(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
  }

  static B2Z_HD Xyzz dbl_inl(const Xyzz& p) {
    if (is_identity(p)) return p;
    const El u = F::dbl(p.y);
    const El v = F::sqr(u);
    const El w = F::mul(u, v);
    const El s = F::mul(p.x, v);
    const El xx = F::sqr(p.x);
    const El m = F::add(F::dbl(xx), xx);
    Xyzz r;
    r.x = F::sub(F::sqr(m), F::dbl(s));
    r.y = F::sub(F::mul(m, F::sub(s, r.x)), F::mul(w, p.y));
    r.zz = F::mul(v, p.zz);
    r.zzz = F::mul(w, p.zzz);
    return r;
  }

  // acc + (affine q), q != identity.  8M + 2S on the common path.
  static B2Z_HD Xyzz madd_inl(const Xyzz& a, const Affine& q) {
    if (is_identity(a)) return from_affine(q);
    const El u2 = F::mul(q.x, a.zz);
    const El s2 = F::mul(q.y, a.zzz);
    const El p = F::sub(u2, a.x);
    const El r = F::sub(s2, a.y);
    if (F::is_zero(p)) {
      if (F::is_zero(r)) return dbl_affine(q);   // out-of-line: rare path
      return identity();
    }
    const El pp = F::sqr(p);
    const El ppp = F::mul(p, pp);
    const El q1 = F::mul(a.x, pp);
    Xyzz o;
    o.x = F::sub(F::sub(F::sqr(r), ppp), F::dbl(q1));
    o.y = F::sub(F::mul(r, F::sub(q1, o.x)), F::mul(a.y, ppp));
    o.zz = F::mul(a.zz, pp);
    o.zzz = F::mul(a.zzz, ppp);
    return o;
  }

  // The same for the accumulation hot loop: the (rare) P + P branch re-reads the base from `src` inside an out-of-line
  // function instead of passing `q` by reference -- a by-reference argument makes the compiler park q in local
  // memory on EVERY iteration (6 / 12 STL.128 per mixed addition in round 1's SASS), whether the branch is taken or not.
  static B2Z_HD_NOINLINE Xyzz dbl_affine_at(const Affine* src, bool negated) {
    Affine p = *src;
    if (negated) p = neg(p);
    return dbl_affine_inl(p);
  }
  static B2Z_HD Xyzz madd_inl_at(const Xyzz& a, const Affine& q, const Affine* src, bool negated) {
    if (is_identity(a)) return from_affine(q);
    const El u2 = F::mul(q.x, a.zz);
    const El s2 = F::mul(q.y, a.zzz);
    const El p = F::sub(u2, a.x);
    const El r = F::sub(s2, a.y);
    if (F::is_zero(p)) {
      if (F::is_zero(r)) return dbl_affine_at(src, negated);
      return identity();
    }
    const El pp = F::sqr(p);
    const El ppp = F::mul(p, pp);
    const El q1 = F::mul(a.x, pp);
    Xyzz o;
    o.x = F::sub(F::sub(F::sqr(r), ppp), F::dbl(q1));
    o.y = F::sub(F::mul(r, F::sub(q1, o.x)), F::mul(a.y, ppp));
    o.zz = F::mul(a.zz, pp);
    o.zzz = F::mul(a.zzz, ppp);
    return o;
  }

  // general addition, 12M + 2S
  static B2Z_HD Xyzz add_inl(const Xyzz& a, const Xyzz& b) {
    if (is_identity(a)) return b;
    if (is_identity(b)) return a;
    const El u1 = F::mul(a.x, b.zz);
    const El u2 = F::mul(b.x, a.zz);
    const El s1 = F::mul(a.y, b.zzz);
    const El s2 = F::mul(b.y, a.zzz);
    const El p = F::sub(u2, u1);
    const El r = F::sub(s2, s1);
    if (F::is_zero(p)) {
      if (F::is_zero(r)) return dbl(a);           // out-of-line: rare path
      return identity();
    }
    const El pp = F::sqr(p);
    const El ppp = F::mul(p, pp);
    const El q1 = F::mul(u1, pp);
    Xyzz o;
    o.x = F::sub(F::sub(F::sqr(r), ppp), F::dbl(q1));
    o.y = F::sub(F::mul(r, F::sub(q1, o.x)), F::mul(s1, ppp));
    o.zz = F::mul(F::mul(a.zz, b.zz), pp);
    o.zzz = F::mul(F::mul(a.zzz, b.zzz), ppp);
    return o;
  }

  // Out-of-line entry points (one shared copy per curve and translation unit).
  static B2Z_HD_NOINLINE Xyzz dbl_affine(const Affine& p) { return dbl_affine_inl(p); }
  static B2Z_HD_NOINLINE Xyzz dbl(const Xyzz& p) { return dbl_inl(p); }
  static B2Z_HD_NOINLINE Xyzz madd(const Xyzz& a, const Affine& q) { return madd_inl(a, q); }
  static B2Z_HD_NOINLINE Xyzz add(const Xyzz& a, const Xyzz& b) { return add_inl(a, b); }

  // k * p for a 256-bit little-endian scalar (8 x u32).
  static B2Z_HD_NOINLINE Xyzz mul_scalar(const Xyzz& p, const uint32_t* k) {
    Xyzz acc = identity();
    for (int i = 255; i >= 0; i--) {
      acc = dbl(acc);
      if ((k[i >> 5] >> (i & 31)) & 1) acc = add(acc, p);
    }
    return acc;
  }

  // Affine normalisation; *is_inf set for the identity.  x = X ZZ^2 / ZZZ^2,
  // y = Y / ZZZ  (ZZ^3 = ZZZ^2), one inversion.
  static B2Z_HD_NOINLINE Affine to_affine(const Xyzz& p, bool* is_inf) {
    if (is_identity(p)) {
      *is_inf = true;
      return Affine{F::zero(), F::zero()};
    }
    *is_inf = false;
    const El i = F::inv(p.zzz);
    const El i2 = F::sqr(i);
    Affine r;
    r.x = F::reduce(F::mul(F::mul(p.x, F::sqr(p.zz)), i2));
    r.y = F::reduce(F::mul(p.y, i));
    return r;
  }
};

using G1 = Curve<Fq>;
using G2 = Curve<Fq2>;

}  // namespace b2z
