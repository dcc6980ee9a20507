// Montgomery product on 30-bit limbs with carry-free 64-bit accumulation (device only).
//
// Measured on B200 (tools/micro/imad_probe.cu): a wide multiply-add that consumes and produces
// the carry flag (IMAD.WIDE.U32.X, the instruction a 32-bit-limb product is made of) issues
// once every 4.3 cycles per SM sub-partition; the same multiply-add with a plain 64-bit addend
// and no flags (IMAD.WIDE) issues every 2.1 cycles.  So this file trades the carry chains for
// headroom: operands are re-sliced from N 32-bit words into L = ceil(32N/30) limbs of 30 bits,
// every limb product (< 2^60) is added to a signed 64-bit column with one flag-free IMAD.WIDE,
// and carries are moved with shifts on the (otherwise idle) ALU pipe.
//
// It computes the SAME function as Field<P>::mul -- a * b * 2^-32N mod p, lazily reduced -- so
// elements keep arkworks' Montgomery form in memory and nothing outside this file changes:
//
//   * L rounds of 30 bits divide by 2^30L, which is 2^pre too much (pre = 30L - 32N: 6 bits
//     for Fq, 14 for Fr); the left operand is therefore sliced as a * 2^pre -- a different
//     funnel-shift offset, no extra work -- and (a 2^pre) b / 2^30L = a b / 2^32N exactly;
//   * word-serial reduction SUBTRACTS m * p (m = col0 * p^-1 mod 2^30), so a column holds
//     (at most 7 products) - (at most 7 products) and stays inside a signed 64-bit integer;
//     after 7 rounds every column is split once into its low 30 bits and a carry for its
//     neighbour (no ripple);
//   * the result (T - M p) / 2^30L lies in (-p, a b / 2^32N); adding p once puts it in
//     (0, 2p): for Fq (4q < 2^384) any two lazy operands qualify, for Fr the first operand
//     must be canonical -- the same contracts as the integer path (mont.cuh header).
//
// 2 L^2 + L multiplies: 351 for Fq against 288 carry-chained ones at twice the issue cost.
#pragma once
#include "field_constants.cuh"

namespace b2z {
namespace w30 {

constexpr uint32_t M30 = (1u << 30) - 1;

template <class P>
struct Consts {
  static constexpr int N = P::N;
  static constexpr int L = (32 * N + 29) / 30;
  static constexpr int PRE = 30 * L - 32 * N;
  // bits [30k, 30k + 30) of p
  static B2Z_HD constexpr uint32_t p30(int k) {
    const int bit = 30 * k, w = bit >> 5, s = bit & 31;
    const uint64_t lo = w < N ? P::p(w) : 0u;
    const uint64_t hi = w + 1 < N ? P::p(w + 1) : 0u;
    return (uint32_t)(((lo | (hi << 32)) >> s) & M30);
  }
  static B2Z_HD constexpr uint32_t p_inv30() {       // p^-1 mod 2^30
    const uint32_t p0 = P::p(0);
    uint32_t x = 1;
    for (int i = 0; i < 6; i++) x *= 2u - p0 * x;
    return x & M30;
  }
};

#ifdef __CUDACC__
// limbs of (a << PRE_SHIFT): limb k = bits [30k - PRE_SHIFT, 30k - PRE_SHIFT + 30) of a
template <class P, int PRE_SHIFT>
__device__ __forceinline__ void slice(const uint32_t (&w)[P::N], int32_t (&l)[Consts<P>::L]) {
  constexpr int N = P::N, L = Consts<P>::L;
#pragma unroll
  for (int k = 0; k < L; k++) {
    const int bit = 30 * k - PRE_SHIFT;
    uint32_t v;
    if (bit < 0) {
      v = w[0] << (-bit);
    } else {
      const int i = bit >> 5, s = bit & 31;
      const uint32_t lo = w[i];
      const uint32_t hi = i + 1 < N ? w[i + 1] : 0u;
      v = s == 0 ? lo : __funnelshift_r(lo, hi, s);
    }
    l[k] = (int32_t)(v & M30);
  }
}

// a6: limbs of a << PRE, b: limbs of b.  Returns the 32-bit words of a b 2^-32N + p.
template <class P>
__device__ __forceinline__ void mul_limbs(const int32_t (&a)[Consts<P>::L], const int32_t (&b)[Consts<P>::L],
                                          uint32_t (&out)[P::N]) {
  using K = Consts<P>;
  constexpr int N = P::N, L = K::L;
  long long col[L + 1];
#pragma unroll
  for (int j = 0; j <= L; j++) col[j] = 0;
#pragma unroll
  for (int i = 0; i < L; i++) {
#pragma unroll
    for (int j = 0; j < L; j++) col[j] += (long long)a[j] * (long long)b[i];
    const int32_t m = (int32_t)(((uint32_t)col[0] * K::p_inv30()) & M30);
#pragma unroll
    for (int j = 0; j < L; j++) col[j] += (long long)m * (long long)(-(int32_t)K::p30(j));
    const long long carry = col[0] >> 30;
#pragma unroll
    for (int j = 0; j < L; j++) col[j] = col[j + 1];
    col[L] = 0;
    col[0] += carry;
    if (i % 7 == 6 && i + 1 < L) {
      // one-level split: every column keeps its low 30 bits and hands the rest to its neighbour
#pragma unroll
      for (int j = L; j >= 1; j--) col[j] = (col[j] & (long long)M30) + (col[j - 1] >> 30);
      col[0] &= (long long)M30;
    }
  }
  // + p, ripple to 30-bit limbs (the value is positive and below 2p: the top carry is zero)
  uint32_t limb[L + 1];
  long long carry = 0;
#pragma unroll
  for (int j = 0; j < L; j++) {
    const long long t = col[j] + (long long)K::p30(j) + carry;
    limb[j] = (uint32_t)t & M30;
    carry = t >> 30;
  }
  limb[L] = (uint32_t)(col[L] + carry);
#pragma unroll
  for (int w = 0; w < N; w++) {
    const int bit = 32 * w, k = bit / 30, s = bit % 30;
    uint32_t v = limb[k] >> s;
    v |= limb[k + 1] << (30 - s);                      // s <= 28: always a real shift
    if (30 - s + 30 < 32 && k + 2 <= L) v |= limb[k + 2] << (60 - s);
    out[w] = v;
  }
}

template <class P>
__device__ __forceinline__ void mul_words(const uint32_t (&a)[P::N], const uint32_t (&b)[P::N], uint32_t (&out)[P::N]) {
  int32_t la[Consts<P>::L], lb[Consts<P>::L];
  slice<P, Consts<P>::PRE>(a, la);
  slice<P, 0>(b, lb);
  mul_limbs<P>(la, lb, out);
}
#endif  // __CUDACC__

}  // namespace w30
}  // namespace b2z
