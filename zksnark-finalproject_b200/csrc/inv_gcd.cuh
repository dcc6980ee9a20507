// Fq inversion by division steps (Bernstein-Yang "safegcd", half-delta variant) on 30-bit signed limbs.
//
// The batched-affine bucket accumulation (msm_affine.inc) shares ONE field inversion among all the point additions a
// thread performs in a round (Montgomery's trick); Fermat's a^(q-2) costs ~570 dependent Fq products, this costs
// about 40 product-equivalents: per batch of 30 division steps a branch-free scalar loop on the low words of (f, g)
// builds a 2x2 transition matrix, which is then applied to the full-width pairs (f, g) and (d, e) with signed
// 32x32+64 multiply-adds.  Invariants: d * x = f, e * x = g (mod q); g reaches 0 with f = +-1, so d = +-1/x.
// Every lane of a warp runs its own inversion; a lane that is done waits for the slowest (batches vary by ~+-3).
//
// Replaces (for that kernel) ark_ff's Field::inverse / batch_inversion as ark-ec's batched normalisation uses them
// (ark-ff 0.4, /root/reference/Cargo.toml:11).  Host-testable like the rest of the limb code (hostcheck.cu op 6).
#pragma once
#include "mont.cuh"

namespace b2z {
namespace gcdinv {

constexpr int L = 13;                       // 13 x 30 = 390 bits >= 382 (values up to 2q in magnitude)
constexpr int32_t M30 = 0x3fffffff;
constexpr int kMaxBatches = 40;             // 1200 division steps; the proven bound for 382-bit inputs is 1104

struct Trans {
  int32_t u, v, q, r;
};

// 12 x u32 -> 13 limbs of 30 bits (all indices are compile-time constants after unrolling)
B2Z_HD void to_s30(const uint32_t (&a)[12], int32_t (&o)[L]) {
#pragma unroll
  for (int i = 0; i < L; i++) {
    const int bit = 30 * i, w = bit >> 5, s = bit & 31;
    uint32_t v = a[w] >> s;
    if (s > 2 && w + 1 < 12) v |= a[w + 1] << (32 - s);
    o[i] = (int32_t)(v & (uint32_t)M30);
  }
}

// 13 limbs in [0, 2^30), value < 2^384 -> 12 x u32
B2Z_HD void from_s30(const int32_t (&d)[L], uint32_t (&o)[12]) {
#pragma unroll
  for (int w = 0; w < 12; w++) {
    const int bit = 32 * w, i0 = bit / 30, s = bit - 30 * i0;
    uint32_t v = (uint32_t)d[i0] >> s;
    if (i0 + 1 < L) v |= (uint32_t)d[i0 + 1] << (30 - s);
    if (60 - s < 32 && i0 + 2 < L) v |= (uint32_t)d[i0 + 2] << (60 - s);
    o[w] = v;
  }
}

// 30 division steps on the low words; zeta = -(delta + 1/2).  Branch-free.
B2Z_HD int32_t divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, Trans& t) {
  uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#pragma unroll
  for (int i = 0; i < 30; i++) {
    uint32_t m1 = (uint32_t)(zeta >> 31);            // all ones iff zeta < 0
    const uint32_t m2 = 0u - (g & 1u);               // all ones iff g odd
    const uint32_t x = (f ^ m1) - m1, y = (u ^ m1) - m1, z = (v ^ m1) - m1;   // -f, -u, -v when zeta < 0
    g += x & m2; q += y & m2; r += z & m2;
    m1 &= m2;
    zeta = (int32_t)((uint32_t)zeta ^ m1) - 1;       // -zeta - 2 when swapping, zeta - 1 otherwise
    f += g & m1; u += q & m1; v += r & m1;
    g >>= 1; u <<= 1; v <<= 1;
  }
  t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
  return zeta;
}

// (f, g) <- t * (f, g) / 2^30   (exact)
B2Z_HD void update_fg(int32_t (&f)[L], int32_t (&g)[L], const Trans& t) {
  const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
  int64_t cf = u * f[0] + v * g[0];
  int64_t cg = q * f[0] + r * g[0];
  cf >>= 30; cg >>= 30;
#pragma unroll
  for (int i = 1; i < L; i++) {
    cf += u * f[i] + v * g[i];
    cg += q * f[i] + r * g[i];
    f[i - 1] = (int32_t)cf & M30; cf >>= 30;
    g[i - 1] = (int32_t)cg & M30; cg >>= 30;
  }
  f[L - 1] = (int32_t)cf;
  g[L - 1] = (int32_t)cg;
}

// (d, e) <- t * (d, e) / 2^30 mod q ; d, e stay in (-2q, q)
B2Z_HD void update_de(int32_t (&d)[L], int32_t (&e)[L], const Trans& t) {
  const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
  const int32_t sd = d[L - 1] >> 31, se = e[L - 1] >> 31;
  int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
  int64_t cd = (int64_t)u * d[0] + (int64_t)v * e[0];
  int64_t ce = (int64_t)q * d[0] + (int64_t)r * e[0];
  md -= (int32_t)((FqParams::PINV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
  me -= (int32_t)((FqParams::PINV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
  cd += (int64_t)FqParams::p30(0) * md;
  ce += (int64_t)FqParams::p30(0) * me;
  cd >>= 30; ce >>= 30;
#pragma unroll
  for (int i = 1; i < L; i++) {
    cd += (int64_t)u * d[i] + (int64_t)v * e[i];
    ce += (int64_t)q * d[i] + (int64_t)r * e[i];
    cd += (int64_t)FqParams::p30(i) * md;
    ce += (int64_t)FqParams::p30(i) * me;
    d[i - 1] = (int32_t)cd & M30; cd >>= 30;
    e[i - 1] = (int32_t)ce & M30; ce >>= 30;
  }
  d[L - 1] = (int32_t)cd;
  e[L - 1] = (int32_t)ce;
}

// d in (-2q, q), negate when sign < 0 -> [0, q)
B2Z_HD void normalize(int32_t (&d)[L], int32_t sign) {
  int32_t add = d[L - 1] >> 31;
#pragma unroll
  for (int i = 0; i < L; i++) d[i] += FqParams::p30(i) & add;
  const int32_t ng = sign >> 31;
#pragma unroll
  for (int i = 0; i < L; i++) d[i] = (d[i] ^ ng) - ng;
#pragma unroll
  for (int i = 0; i < L - 1; i++) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
  add = d[L - 1] >> 31;
#pragma unroll
  for (int i = 0; i < L; i++) d[i] += FqParams::p30(i) & add;
#pragma unroll
  for (int i = 0; i < L - 1; i++) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
}

// Montgomery inverse: x = a R (lazy) -> a^-1 R (lazy, < 2q); inv(0) = 0.  *batches: division-step batches used.
B2Z_HD FqEl inv(const FqEl& x_in, int* batches = nullptr) {
  const FqEl x = Fq::reduce(x_in);
  int32_t d[L], e[L], f[L], g[L];
#pragma unroll
  for (int i = 0; i < L; i++) { d[i] = 0; e[i] = 0; f[i] = FqParams::p30(i); }
  e[0] = 1;
  to_s30(x.l, g);
  int32_t zeta = -1;
  int it = 0;
  for (; it < kMaxBatches; it++) {
    int32_t nz = 0;
#pragma unroll
    for (int i = 0; i < L; i++) nz |= g[i];
    if (nz == 0) break;
    Trans t;
    zeta = divsteps_30(zeta, (uint32_t)f[0], (uint32_t)g[0], t);
    update_de(d, e, t);
    update_fg(f, g, t);
  }
  if (batches) *batches = it;
  normalize(d, f[L - 1]);
  FqEl plain, r3;
  from_s30(d, plain.l);
#pragma unroll
  for (int i = 0; i < 12; i++) r3.l[i] = FqParams::r3(i);
  return Fq::mul(plain, r3);                 // (aR)^-1 * R^3 * R^-1 = a^-1 R
}

}  // namespace gcdinv
}  // namespace b2z
