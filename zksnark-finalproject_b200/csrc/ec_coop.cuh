// Lane-cooperative XYZZ point arithmetic for the latency-bound tails of the MSM.
//
// One field product keeps a lone warp busy for ~2000 cycles (the carry-chained wide
// multiply pipe issues one warp instruction every ~5 cycles), so a G2 addition done by a
// single thread costs ~119k cycles (62 us) and every SERIAL point operation in the
// bucket reduction is that expensive.  Here one "logical thread" is a group of lanes
// of a warp -- 4 for G1, 16 for G2 -- that executes ONE point operation together: the
// 12-14 field products of an addition are arranged in 4 phases of up to four
// independent products, each Fq product (three per Fq2 product, Karatsuba) on its own
// lane.  Operands and intermediates live in memory (the accumulator in global/shared
// memory, temporaries in a per-group shared scratch), so no lane holds more than two
// field elements in registers.  G2 addition: 4 product phases instead of 42 serial
// Fq products.
//
// All lanes of a group run the same control flow (every branch depends on values all
// lanes read from the same memory), and synchronise with __syncwarp(group mask).
#pragma once
#include "ec.cuh"

namespace b2z {

template <class F> struct CoopTraits;
template <> struct CoopTraits<Fq> { static constexpr int NL = 1, GROUP = 4; };    // lanes per product, group size
template <> struct CoopTraits<Fq2> { static constexpr int NL = 3, GROUP = 16; };

template <int GROUP>
struct Lanes {
  unsigned mask;
  int lane;
  __device__ __forceinline__ Lanes() {
    const int l = threadIdx.x & 31;
    lane = l % GROUP;
    mask = GROUP == 32 ? 0xffffffffu : (((1u << GROUP) - 1u) << (l - lane));
  }
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};

template <class F>
struct Coop {
  using C = Curve<F>;
  using El = typename F::El;
  using Xyzz = typename C::Xyzz;
  static constexpr int NL = CoopTraits<F>::NL;
  static constexpr int GROUP = CoopTraits<F>::GROUP;
  using LG = Lanes<GROUP>;

  enum { U1, U2, S1, S2, P, R, PP, RR, ZZ12, ZZZ12, PPP, Q, T1, T2, NV };
  struct alignas(16) Scratch {
    El v[NV];
    FqEl raw[NL == 1 ? 1 : 4][3];   // Karatsuba parts of the products in flight (Fq2 only)
  };

  // ---- memory helpers (16-byte vector accesses; every El is 16-byte aligned)
  template <class T>
  static __device__ __forceinline__ T ld(const T* p) {
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
    return r;
  }
  template <class T>
  static __device__ __forceinline__ void st(T* p, const T& v) {
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
  }
  static __device__ __forceinline__ void copy(const LG& g, Xyzz* dst, const Xyzz* src) {
    if (dst != src) {
      const uint4* s = reinterpret_cast<const uint4*>(src);
      uint4* d = reinterpret_cast<uint4*>(dst);
      for (int i = g.lane; i < (int)(sizeof(Xyzz) / 16); i += GROUP) d[i] = s[i];
    }
    g.sync();
  }
  static __device__ __forceinline__ void set_identity(const LG& g, Xyzz* dst) {
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = g.lane; i < (int)(sizeof(Xyzz) / 16); i += GROUP) d[i] = make_uint4(0, 0, 0, 0);
    g.sync();
  }
  static __device__ __forceinline__ bool is_identity(const Xyzz* p) { return F::is_zero(ld(&p->zz)); }

  // ---- one phase: up to four independent products OUT_i = A_i * B_i
  static __device__ __forceinline__ const El* pick(int i, const El* p0, const El* p1, const El* p2, const El* p3) {
    return i == 0 ? p0 : (i == 1 ? p1 : (i == 2 ? p2 : p3));
  }
  static __device__ __forceinline__ void mul_phase(const LG& g, Scratch* sc, int n, const El* a0, const El* b0, El* o0,
                                                   const El* a1, const El* b1, El* o1, const El* a2, const El* b2,
                                                   El* o2, const El* a3, const El* b3, El* o3) {
    const int i = g.lane / NL, k = g.lane % NL;
    const bool active = i < n;
    const El* a = pick(i, a0, a1, a2, a3);
    const El* b = pick(i, b0, b1, b2, b3);
    El* o = const_cast<El*>(pick(i, o0, o1, o2, o3));
    if (NL == 1) {
      FqEl r;
      if (active) r = Fq::mul(ld(reinterpret_cast<const FqEl*>(a)), ld(reinterpret_cast<const FqEl*>(b)));
      g.sync();                       // every operand has been read before any output is written
      if (active) st(reinterpret_cast<FqEl*>(o), r);
      g.sync();
    } else {
      if (active) {
        const FqEl* af = reinterpret_cast<const FqEl*>(a);
        const FqEl* bf = reinterpret_cast<const FqEl*>(b);
        FqEl x, y;
        if (k == 0) { x = ld(af); y = ld(bf); }
        else if (k == 1) { x = ld(af + 1); y = ld(bf + 1); }
        else { x = Fq::add(ld(af), ld(af + 1)); y = Fq::add(ld(bf), ld(bf + 1)); }
        st(&sc->raw[i][k], Fq::mul(x, y));
      }
      g.sync();
      if (active && k < 2) {
        const FqEl v0 = ld(&sc->raw[i][0]), v1 = ld(&sc->raw[i][1]);
        FqEl* of = reinterpret_cast<FqEl*>(o);
        if (k == 0) st(of, Fq::sub(v0, v1));
        else st(of + 1, Fq::sub(Fq::sub(ld(&sc->raw[i][2]), v0), v1));
      }
      g.sync();
    }
  }

  // out = 2 p   (out may alias p)
  static __device__ void dbl(const LG& g, Scratch* sc, const Xyzz* p, Xyzz* out) {
    if (is_identity(p)) { copy(g, out, p); return; }
    El* v = sc->v;
    if (g.lane == 0) st(&v[U1], F::dbl(ld(&p->y)));                                  // U = 2y
    g.sync();
    mul_phase(g, sc, 2, &v[U1], &v[U1], &v[PP], &p->x, &p->x, &v[RR], nullptr, nullptr, nullptr, nullptr, nullptr,
              nullptr);                                                              // V = U^2, XX = x^2
    if (g.lane == 0) { const El xx = ld(&v[RR]); st(&v[P], F::add(F::dbl(xx), xx)); }  // M = 3 XX
    g.sync();
    mul_phase(g, sc, 3, &v[U1], &v[PP], &v[PPP], &p->x, &v[PP], &v[Q], &v[P], &v[P], &v[U2], nullptr, nullptr,
              nullptr);                                                              // W = U V, S = x V, MM = M^2
    if (g.lane == 0) {
      const El s = ld(&v[Q]);
      const El x3 = F::sub(ld(&v[U2]), F::dbl(s));
      st(&v[S1], x3);
      st(&v[S2], F::sub(s, x3));
    }
    g.sync();
    mul_phase(g, sc, 4, &v[P], &v[S2], &v[T1], &v[PPP], &p->y, &v[T2], &v[PP], &p->zz, &out->zz, &v[PPP], &p->zzz,
              &out->zzz);                                                            // M (S - X3), W y, V zz, W zzz
    if (g.lane == 0) {
      st(&out->x, ld(&v[S1]));
      st(&out->y, F::sub(ld(&v[T1]), ld(&v[T2])));
    }
    g.sync();
  }

  // out = a + b   (out may alias a or b)
  static __device__ void add(const LG& g, Scratch* sc, const Xyzz* a, const Xyzz* b, Xyzz* out) {
    if (is_identity(a)) { copy(g, out, b); return; }
    if (is_identity(b)) { copy(g, out, a); return; }
    El* v = sc->v;
    mul_phase(g, sc, 4, &a->x, &b->zz, &v[U1], &b->x, &a->zz, &v[U2], &a->y, &b->zzz, &v[S1], &b->y, &a->zzz, &v[S2]);
    if (g.lane == 0) st(&v[P], F::sub(ld(&v[U2]), ld(&v[U1])));
    if (g.lane == 1) st(&v[R], F::sub(ld(&v[S2]), ld(&v[S1])));
    g.sync();
    if (F::is_zero(ld(&v[P]))) {
      if (F::is_zero(ld(&v[R]))) dbl(g, sc, a, out);
      else set_identity(g, out);
      return;
    }
    mul_phase(g, sc, 4, &v[P], &v[P], &v[PP], &v[R], &v[R], &v[RR], &a->zz, &b->zz, &v[ZZ12], &a->zzz, &b->zzz,
              &v[ZZZ12]);
    mul_phase(g, sc, 3, &v[P], &v[PP], &v[PPP], &v[U1], &v[PP], &v[Q], &v[ZZ12], &v[PP], &out->zz, nullptr, nullptr,
              nullptr);
    if (g.lane == 0) {
      const El q = ld(&v[Q]);
      const El x3 = F::sub(F::sub(ld(&v[RR]), ld(&v[PPP])), F::dbl(q));
      st(&out->x, x3);
      st(&v[T1], F::sub(q, x3));
    }
    g.sync();
    mul_phase(g, sc, 3, &v[R], &v[T1], &v[U2], &v[S1], &v[PPP], &v[T2], &v[ZZZ12], &v[PPP], &out->zzz, nullptr, nullptr,
              nullptr);
    if (g.lane == 0) st(&out->y, F::sub(ld(&v[U2]), ld(&v[T2])));
    g.sync();
  }
};

}  // namespace b2z
