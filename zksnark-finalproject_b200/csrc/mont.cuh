// Montgomery-form prime-field arithmetic on 32-bit limbs (Fr: 8 limbs, Fq: 12).
//
// Replaces, on the device, ark_ff::Fp<MontBackend<_, N>> (ark-ff 0.4,
// /root/reference/Cargo.toml:11) -- same Montgomery radix (2^256 / 2^384), so
// elements cross the C ABI in arkworks' in-memory form without conversion.
//
// Representation: every value is kept *lazily reduced* in [0, 2p); `reduce()`
// brings it to the canonical representative for output.  The raw Montgomery
// product `mul(a, b)` skips the final subtraction and returns a value < 2p
// when the running sum (T + a*b_i + m*p < (a + 2p) * 2^32) fits N+1 limbs:
//   * Fq (381 bits in 384): 4q < 2^384, so any two lazy values qualify;
//   * Fr (255 bits in 256): only 2r < 2^256, so the FIRST operand must be
//     canonical (< r) -- the twiddle / constant in every NTT product; the
//     second may be lazy.  Use `mul_safe` when the first operand may be lazy.
// a + b of two lazy Fr values can carry out of the limb array; `add` handles it.
//
// Multiplication is an operand-scanning (CIOS-style) loop over the limbs of b
// that keeps TWO accumulators: `E` aligned at the current limb position and `O`
// one limb above it.  The even limbs of a multiply into E, the odd limbs into O,
// so that every 32x32 product lands on an aligned (lo, hi) register pair and the
// (mad.lo.cc, madc.hi.cc) pair becomes one IMAD.WIDE.U32.X.  After a row the
// lowest limb of E is zero by construction; the accumulators then swap roles
// (the old O is aligned at the next position, the old E shifted down by two
// limbs is the new O) at the cost of a single add.cc whose carry is consumed by
// the next chain.  2N^2 wide multiply-adds per product, ~4 adds per row.
#pragma once
#include "field_constants.cuh"

namespace b2z {

template <class P>
struct alignas(16) Fp {
  static constexpr int N = P::N;
  uint32_t l[N];
};

namespace detail {

// Modulus limb as the multiplier of the reduction step's multiply-adds.  For Fr the limbs come from the constant
// bank on the device instead of being immediates: ptxas splits a wide multiply-add whose multiplier is a 32-bit
// immediate and whose destination pair differs from its addend pair into IMAD.X + IMAD.HI.U32.X (two trips through
// the multiplier; half of all multiply-adds of the transform kernels in round 1's SASS), while a constant-bank
// operand keeps the single IMAD.WIDE.U32.X.  (Fq's rows are allocated in place and fuse as they are.)
#if defined(__CUDACC__)
static __constant__ uint32_t kFrModulusLimbs[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                                   0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
#endif
template <class P>
B2Z_HD uint32_t mod_limb(int j) { return P::p(j); }
template <>
B2Z_HD uint32_t mod_limb<FrParams>(int j) {
#if defined(__CUDA_ARCH__)
  return kFrModulusLimbs[j];
#else
  return FrParams::p(j);
#endif
}

template <class P>
B2Z_HD uint32_t mont_inv() { return P::INV; }
#if defined(__CUDACC__)
static __constant__ uint32_t kFrInv = 0xffffffffu;
#endif
template <>
B2Z_HD uint32_t mont_inv<FrParams>() {
#if defined(__CUDA_ARCH__)
  return kFrInv;
#else
  return FrParams::INV;
#endif
}

// One row of the product: acc += a * bi ; acc += m * p ; (acc's low limb is 0).
// On entry (unless FIRST) E holds the previous row's O and O the previous row's E.
template <class P, bool FIRST>
B2Z_HD void mont_row(uint32_t (&E)[P::N], uint32_t (&O)[P::N], const uint32_t (&a)[P::N], uint32_t bi) {
  constexpr int N = P::N;
  ptx::CF cf;
  if (FIRST) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      uint64_t te = (uint64_t)a[j] * bi;
      uint64_t to = (uint64_t)a[j + 1] * bi;
      E[j] = (uint32_t)te; E[j + 1] = (uint32_t)(te >> 32);
      O[j] = (uint32_t)to; O[j + 1] = (uint32_t)(to >> 32);
    }
  } else {
    // fold the stray limb of the old E (one position above its vanished low limb)
    E[0] = ptx::add_cc(cf, E[0], O[1]);
    // new O = (old E >> 2 limbs) + odd limbs of a * bi, carry-in from the fold
#pragma unroll
    for (int j = 0; j < N - 2; j += 2)
      ptx::madc_wide_cc3(cf, O[j], O[j + 1], a[j + 1], bi, O[j + 2], O[j + 3]);
    O[N - 2] = ptx::madc_lo_cc(cf, a[N - 1], bi, 0u);
    O[N - 1] = ptx::madc_hi(cf, a[N - 1], bi, 0u);
    // E += even limbs of a * bi
    ptx::mad_wide_cc(cf, E[0], E[1], a[0], bi);
#pragma unroll
    for (int j = 2; j < N; j += 2) ptx::madc_wide_cc(cf, E[j], E[j + 1], a[j], bi);
    O[N - 1] += ptx::addc(cf, 0u, 0u);
  }
  const uint32_t m = E[0] * mont_inv<P>();
  ptx::mad_wide_cc(cf, E[0], E[1], m, mod_limb<P>(0));
#pragma unroll
  for (int j = 2; j < N; j += 2) ptx::madc_wide_cc(cf, E[j], E[j + 1], m, mod_limb<P>(j));
  const uint32_t c = ptx::addc(cf, 0u, 0u);
  ptx::mad_wide_cc(cf, O[0], O[1], m, mod_limb<P>(1));
#pragma unroll
  for (int j = 2; j < N; j += 2) ptx::madc_wide_cc(cf, O[j], O[j + 1], m, mod_limb<P>(j + 1));
  O[N - 1] += c;
}

}  // namespace detail

template <class P>
struct Field {
  static constexpr int N = P::N;
  using El = Fp<P>;

  static B2Z_HD El zero() { El r; for (int i = 0; i < N; i++) r.l[i] = 0; return r; }
  static B2Z_HD El one() { El r; for (int i = 0; i < N; i++) r.l[i] = P::one(i); return r; }

  // a * b * R^-1; precondition: a canonical unless P::HEADROOM (see header); result < 2p.
  static B2Z_HD El mul(const El& a, const El& b) {
    uint32_t e[N], o[N];
    detail::mont_row<P, true>(e, o, a.l, b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
      detail::mont_row<P, false>(o, e, a.l, b.l[i]);
      if (i + 1 < N) detail::mont_row<P, false>(e, o, a.l, b.l[i + 1]);
    }
    // N is even: the last row ran with (E, O) = (o, e); e's low limb is dead,
    // result limb k = e[k] + o[k+1] ... expressed for that role assignment:
    // value = o[1..] (aligned one above) + e (shifted): r[k] = e[k] + o[k+1].
    El r;
    ptx::CF cf;
    r.l[0] = ptx::add_cc(cf, e[0], o[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.l[k] = ptx::addc_cc(cf, e[k], o[k + 1]);
    r.l[N - 1] = ptx::addc(cf, e[N - 1], 0u);
    return r;
  }
  // product of two arbitrary lazy values
  static B2Z_HD El mul_safe(const El& a, const El& b) {
    if (P::HEADROOM) return mul(a, b);
    return mul(reduce(a), b);
  }
  static B2Z_HD El sqr(const El& a) {
    if (P::HEADROOM) return mul(a, a);
    El t = reduce(a);
    return mul(t, t);
  }

  // (a + b) mod 2p
  static B2Z_HD El add(const El& a, const El& b) {
    El r, t;
    ptx::CF cf;
    r.l[0] = ptx::add_cc(cf, a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(cf, a.l[i], b.l[i]);
    r.l[N - 1] = ptx::addc_cc(cf, a.l[N - 1], b.l[N - 1]);
    const uint32_t carry = ptx::addc(cf, 0u, 0u);     // only Fr can carry out
    t.l[0] = ptx::sub_cc(cf, r.l[0], P::p2(0));
#pragma unroll
    for (int i = 1; i < N; i++) t.l[i] = ptx::subc_cc(cf, r.l[i], P::p2(i));
    const uint32_t borrow = ptx::subc(cf, 0u, 0u);   // 0xffffffff iff (low limbs of) r < 2p
    const bool keep = (borrow != 0u) && (carry == 0u);
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = keep ? r.l[i] : t.l[i];
    return r;
  }
  // (a - b) mod 2p
  static B2Z_HD El sub(const El& a, const El& b) {
    El r;
    ptx::CF cf;
    r.l[0] = ptx::sub_cc(cf, a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.l[i] = ptx::subc_cc(cf, a.l[i], b.l[i]);
    const uint32_t borrow = ptx::subc(cf, 0u, 0u);
    r.l[0] = ptx::add_cc(cf, r.l[0], P::p2(0) & borrow);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(cf, r.l[i], P::p2(i) & borrow);
    r.l[N - 1] = ptx::addc(cf, r.l[N - 1], P::p2(N - 1) & borrow);
    return r;
  }
  static B2Z_HD El dbl(const El& a) { return add(a, a); }
  static B2Z_HD El neg(const El& a) {   // 2p - a, in (0, 2p]; fold 2p -> handled by callers via is_zero/reduce
    return sub(zero(), a);
  }
  // canonical representative in [0, p)
  static B2Z_HD El reduce(const El& a) {
    El t;
    ptx::CF cf;
    t.l[0] = ptx::sub_cc(cf, a.l[0], P::p(0));
#pragma unroll
    for (int i = 1; i < N; i++) t.l[i] = ptx::subc_cc(cf, a.l[i], P::p(i));
    const uint32_t borrow = ptx::subc(cf, 0u, 0u);
    El r;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = borrow ? a.l[i] : t.l[i];
    return r;
  }
  static B2Z_HD bool is_zero(const El& a) {
    uint32_t z = 0, zp = 0;
#pragma unroll
    for (int i = 0; i < N; i++) { z |= a.l[i]; zp |= a.l[i] ^ P::p(i); }
    return z == 0 || zp == 0;
  }
  static B2Z_HD bool eq(const El& a, const El& b) { return is_zero(sub(a, b)); }

  static B2Z_HD El to_mont(const El& a) { El r2; for (int i = 0; i < N; i++) r2.l[i] = P::r2(i); return mul_safe(a, r2); }
  static B2Z_HD El from_mont(const El& a) { El o = zero(); o.l[0] = 1; return reduce(mul(o, a)); }

  // a^e for a little-endian 32-bit-limb exponent (host-side constants, inversion).
  static B2Z_HD_NOINLINE El pow(const El& a, const uint32_t* e, int nlimbs) {
    El r = one();
    bool started = false;
    for (int i = nlimbs * 32 - 1; i >= 0; i--) {
      if (started) r = sqr(r);
      if ((e[i >> 5] >> (i & 31)) & 1) { r = started ? mul_safe(r, a) : a; started = true; }
    }
    return r;
  }
  static B2Z_HD El pow_u64(const El& a, uint64_t e) {
    uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
    return pow(a, w, 2);
  }
  // Fermat inversion a^(p-2); inv(0) = 0.
  static B2Z_HD El inv(const El& a) {
    uint32_t e[N];
    uint32_t borrow = 2;
    for (int i = 0; i < N; i++) {
      uint32_t pi = P::p(i);
      e[i] = pi - borrow;
      borrow = pi < borrow ? 1u : 0u;
    }
    return pow(a, e, N);
  }
};

using Fr = Field<FrParams>;
using Fq = Field<FqParams>;
using FrEl = Fp<FrParams>;
using FqEl = Fp<FqParams>;

}  // namespace b2z
