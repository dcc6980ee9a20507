// CPU restatement of the arkworks 0.4 prove path -- TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library; the product (zksnark-finalproject_b200/)
// never does.  PARITY STATUS: "parity unpinned" -- the reference's arithmetic
// lives in un-vendored crates (/root/reference/Cargo.toml:11-18,43), cannot be
// built here (no Rust toolchain) and its tests pin no bytes; this file restates
// the published algorithms and is itself pinned against oracle/*.py (big-int
// Python, pairing-verified) by tests/test_oracle_cpu.py.
//
// What is restated, with the algorithmic structure of the
// crates so that it also serves as the "arkworks-equivalent" timed CPU baseline:
//   ark-ff 0.4       Fp<MontBackend<_, N>>: 64-bit-limb CIOS Montgomery arithmetic
//   ark-poly 0.4.2   Radix2EvaluationDomain: in-place radix-2 FFT, coset by
//                    distributing powers of the offset, threads over butterflies
//   ark-ec 0.4.2     VariableBaseMSM::msm_bigint (signed-digit windows,
//                    c = 3 | ceil(log2 n)*69/100 + 2, 2^c buckets, running sums,
//                    Horner), parallel over windows only -- as rayon does;
//                    Jacobian add / mixed add / double
//   ark-groth16 0.4  LibsnarkReduction::witness_map (after row evaluation) and
//                    create_proof_with_assignment
//   ark-bls12-381    zcash-format compressed serialization
//
// Constants are derived at start-up from the two moduli only.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

typedef uint64_t u64;
typedef unsigned __int128 u128;

namespace {

// ------------------------------------------------------------------ prime fields
template <int N>
struct FpParams {
  u64 p[N];
  u64 inv;       // -p^-1 mod 2^64
  u64 one[N];    // R mod p
  u64 r2[N];     // R^2 mod p
};

template <int N>
static bool geq(const u64* a, const u64* b) {
  for (int i = N - 1; i >= 0; i--) {
    if (a[i] != b[i]) return a[i] > b[i];
  }
  return true;
}
template <int N>
static u64 sub_n(u64* r, const u64* a, const u64* b) {
  u64 borrow = 0;
  for (int i = 0; i < N; i++) {
    u128 t = (u128)a[i] - b[i] - borrow;
    r[i] = (u64)t;
    borrow = (u64)(t >> 64) & 1;
  }
  return borrow;
}
template <int N>
static u64 add_n(u64* r, const u64* a, const u64* b) {
  u64 carry = 0;
  for (int i = 0; i < N; i++) {
    u128 t = (u128)a[i] + b[i] + carry;
    r[i] = (u64)t;
    carry = (u64)(t >> 64);
  }
  return carry;
}

template <int N, const FpParams<N>* (*PP)()>
struct Fp {
  u64 l[N];
  static const FpParams<N>& P() { return *PP(); }
  static Fp zero() { Fp r; memset(r.l, 0, sizeof(r.l)); return r; }
  static Fp one() { Fp r; memcpy(r.l, P().one, sizeof(r.l)); return r; }
  bool is_zero() const { u64 z = 0; for (int i = 0; i < N; i++) z |= l[i]; return z == 0; }
  bool operator==(const Fp& o) const { return memcmp(l, o.l, sizeof(l)) == 0; }
  bool operator!=(const Fp& o) const { return !(*this == o); }
  Fp operator+(const Fp& o) const {
    Fp r;
    u64 c = add_n<N>(r.l, l, o.l);
    if (c || geq<N>(r.l, P().p)) sub_n<N>(r.l, r.l, P().p);
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r;
    if (sub_n<N>(r.l, l, o.l)) add_n<N>(r.l, r.l, P().p);
    return r;
  }
  Fp neg() const { return is_zero() ? *this : (zero() - *this); }
  Fp dbl() const { return *this + *this; }
  Fp operator*(const Fp& o) const {   // CIOS Montgomery product
    const FpParams<N>& pp = P();
    u64 t[N + 2];
    memset(t, 0, sizeof(t));
    for (int i = 0; i < N; i++) {
      u64 c = 0;
      for (int j = 0; j < N; j++) {
        u128 x = (u128)l[j] * o.l[i] + t[j] + c;
        t[j] = (u64)x;
        c = (u64)(x >> 64);
      }
      u128 x = (u128)t[N] + c;
      t[N] = (u64)x;
      t[N + 1] = (u64)(x >> 64);
      const u64 m = t[0] * pp.inv;
      x = (u128)m * pp.p[0] + t[0];
      c = (u64)(x >> 64);
      for (int j = 1; j < N; j++) {
        x = (u128)m * pp.p[j] + t[j] + c;
        t[j - 1] = (u64)x;
        c = (u64)(x >> 64);
      }
      x = (u128)t[N] + c;
      t[N - 1] = (u64)x;
      t[N] = t[N + 1] + (u64)(x >> 64);
    }
    Fp r;
    if (t[N] || geq<N>(t, pp.p)) sub_n<N>(r.l, t, pp.p);
    else memcpy(r.l, t, sizeof(r.l));
    return r;
  }
  Fp sqr() const { return *this * *this; }
  Fp pow(const u64* e, int n) const {
    Fp r = one();
    for (int i = n * 64 - 1; i >= 0; i--) {
      r = r.sqr();
      if ((e[i >> 6] >> (i & 63)) & 1) r = r * *this;
    }
    return r;
  }
  Fp inverse() const {   // Fermat
    u64 e[N];
    memcpy(e, P().p, sizeof(e));
    u64 two[N] = {2};
    sub_n<N>(e, e, two);
    return pow(e, N);
  }
  static Fp from_u64(u64 v) {
    Fp x = zero();
    x.l[0] = v;
    Fp r2;
    memcpy(r2.l, P().r2, sizeof(r2.l));
    return x * r2;
  }
  void to_canonical(u64* out) const {   // leave Montgomery form
    Fp o = zero();
    o.l[0] = 1;
    Fp c = *this * o;
    memcpy(out, c.l, sizeof(c.l));
  }
};

template <int N>
static void init_params(FpParams<N>& P, const u64* modulus) {
  memcpy(P.p, modulus, sizeof(P.p));
  u64 inv = 1;
  for (int i = 0; i < 6; i++) inv *= 2 - modulus[0] * inv;   // Newton: p^-1 mod 2^64
  P.inv = (u64)0 - inv;
  // R mod p and R^2 mod p by repeated doubling of 1
  u64 x[N];
  memset(x, 0, sizeof(x));
  x[0] = 1;
  auto dbl_mod = [&](u64* v) {
    u64 c = add_n<N>(v, v, v);
    if (c || geq<N>(v, P.p)) sub_n<N>(v, v, P.p);
  };
  for (int i = 0; i < 64 * N; i++) dbl_mod(x);
  memcpy(P.one, x, sizeof(x));
  for (int i = 0; i < 64 * N; i++) dbl_mod(x);
  memcpy(P.r2, x, sizeof(x));
}

static FpParams<4> g_fr;
static FpParams<6> g_fq;
static const FpParams<4>* fr_params() { return &g_fr; }
static const FpParams<6>* fq_params() { return &g_fq; }
typedef Fp<4, fr_params> Fr;
typedef Fp<6, fq_params> Fq;

static const u64 kFrModulus[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull,
                                  0x73eda753299d7d48ull};
static const u64 kFqModulus[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                                  0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};

static Fr g_root32;      // 2^32-th root of unity 7^((r-1)/2^32)
static Fr g_generator;   // 7

struct Init {
  Init() {
    init_params<4>(g_fr, kFrModulus);
    init_params<6>(g_fq, kFqModulus);
    g_generator = Fr::from_u64(7);
    u64 e[4];
    memcpy(e, kFrModulus, sizeof(e));
    e[0] -= 1;                                   // r - 1
    // >> 32
    for (int i = 0; i < 4; i++) e[i] = (e[i] >> 32) | (i + 1 < 4 ? e[i + 1] << 32 : 0);
    g_root32 = g_generator.pow(e, 4);
  }
} g_init;

// ------------------------------------------------------------------ Fq2
struct Fq2 {
  Fq c0, c1;
  static Fq2 zero() { return Fq2{Fq::zero(), Fq::zero()}; }
  static Fq2 one() { return Fq2{Fq::one(), Fq::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool operator!=(const Fq2& o) const { return !(*this == o); }
  Fq2 operator+(const Fq2& o) const { return Fq2{c0 + o.c0, c1 + o.c1}; }
  Fq2 operator-(const Fq2& o) const { return Fq2{c0 - o.c0, c1 - o.c1}; }
  Fq2 neg() const { return Fq2{c0.neg(), c1.neg()}; }
  Fq2 dbl() const { return Fq2{c0.dbl(), c1.dbl()}; }
  Fq2 operator*(const Fq2& o) const {
    Fq v0 = c0 * o.c0, v1 = c1 * o.c1;
    return Fq2{v0 - v1, (c0 + c1) * (o.c0 + o.c1) - v0 - v1};
  }
  Fq2 sqr() const { return Fq2{(c0 + c1) * (c0 - c1), (c0 * c1).dbl()}; }
  Fq2 inverse() const {
    Fq d = (c0.sqr() + c1.sqr()).inverse();
    return Fq2{c0 * d, (c1 * d).neg()};
  }
};

// ------------------------------------------------------------------ curves (a = 0), Jacobian
template <class F>
struct Affine {
  F x, y;
  bool inf;
};
template <class F>
struct Jac {
  F x, y, z;
  static Jac identity() { return Jac{F::one(), F::one(), F::zero()}; }
  bool is_identity() const { return z.is_zero(); }
  Jac dbl() const {   // dbl-2009-l
    if (is_identity()) return *this;
    F a = x.sqr(), b = y.sqr(), c = b.sqr();
    F d = ((x + b).sqr() - a - c).dbl();
    F e = a + a.dbl(), f = e.sqr();
    Jac r;
    r.z = (y * z).dbl();
    r.x = f - d.dbl();
    r.y = e * (d - r.x) - c.dbl().dbl().dbl();
    return r;
  }
  Jac add_mixed(const Affine<F>& q) const {   // madd-2007-bl
    if (q.inf) return *this;
    if (is_identity()) return Jac{q.x, q.y, F::one()};
    F z1z1 = z.sqr(), u2 = q.x * z1z1, s2 = (q.y * z) * z1z1;
    if (x == u2) {
      if (y == s2) return dbl();
      return identity();
    }
    F h = u2 - x, hh = h.sqr(), i = hh.dbl().dbl(), j = h * i, r = (s2 - y).dbl(), v = x * i;
    Jac o;
    o.x = r.sqr() - j - v.dbl();
    o.y = r * (v - o.x) - (y * j).dbl();
    o.z = (z + h).sqr() - z1z1 - hh;
    return o;
  }
  Jac add(const Jac& q) const {   // add-2007-bl
    if (is_identity()) return q;
    if (q.is_identity()) return *this;
    F z1z1 = z.sqr(), z2z2 = q.z.sqr(), u1 = x * z2z2, u2 = q.x * z1z1;
    F s1 = (y * q.z) * z2z2, s2 = (q.y * z) * z1z1;
    if (u1 == u2) {
      if (s1 == s2) return dbl();
      return identity();
    }
    F h = u2 - u1, i = h.dbl().sqr(), j = h * i, r = (s2 - s1).dbl(), v = u1 * i;
    Jac o;
    o.x = r.sqr() - j - v.dbl();
    o.y = r * (v - o.x) - (s1 * j).dbl();
    o.z = ((z + q.z).sqr() - z1z1 - z2z2) * h;
    return o;
  }
  Jac neg() const { return Jac{x, y.neg(), z}; }
  Jac mul_bigint(const u64* k, int n) const {   // double-and-add, MSB first
    Jac r = identity();
    for (int i = n * 64 - 1; i >= 0; i--) {
      r = r.dbl();
      if ((k[i >> 6] >> (i & 63)) & 1) r = r.add(*this);
    }
    return r;
  }
  Affine<F> to_affine() const {
    if (is_identity()) return Affine<F>{F::zero(), F::zero(), true};
    F zi = z.inverse(), zi2 = zi.sqr();
    return Affine<F>{x * zi2, y * zi2 * zi, false};
  }
};

// ------------------------------------------------------------------ threads
static int g_threads = 1;
template <class Fn>
static void parallel_for(size_t n, Fn fn) {   // fn(begin, end)
  int t = std::min<size_t>(g_threads, n ? n : 1);
  if (t <= 1) { fn(0, n); return; }
  std::vector<std::thread> pool;
  size_t chunk = (n + t - 1) / t;
  for (int i = 0; i < t; i++) {
    size_t b = std::min(n, i * chunk), e = std::min(n, b + chunk);
    if (b < e) pool.emplace_back([=] { fn(b, e); });
  }
  for (auto& th : pool) th.join();
}

// ------------------------------------------------------------------ ark-poly radix-2 domain
struct Domain {
  uint32_t log_n;
  size_t n;
  Fr gen, gen_inv, size_inv;
  explicit Domain(uint32_t lg) : log_n(lg), n((size_t)1 << lg) {
    gen = g_root32;
    for (uint32_t i = lg; i < 32; i++) gen = gen.sqr();
    gen_inv = gen.inverse();
    size_inv = Fr::from_u64((u64)1 << lg).inverse();
  }
};

static void bit_reverse(Fr* v, uint32_t log_n) {
  size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) {
    size_t j = 0;
    for (uint32_t b = 0; b < log_n; b++) j |= ((i >> b) & 1) << (log_n - 1 - b);
    if (i < j) std::swap(v[i], v[j]);
  }
}

// in-place, natural order in and out, root w
static void fft_core(Fr* v, uint32_t log_n, const Fr& w) {
  const size_t n = (size_t)1 << log_n;
  bit_reverse(v, log_n);
  std::vector<Fr> roots(n / 2 ? n / 2 : 1);
  roots[0] = Fr::one();
  for (size_t i = 1; i < n / 2; i++) roots[i] = roots[i - 1] * w;
  for (uint32_t s = 0; s < log_n; s++) {
    const size_t half = (size_t)1 << s, m = half << 1, stride = n / m;
    parallel_for(n / 2, [&](size_t b, size_t e) {
      for (size_t idx = b; idx < e; idx++) {
        const size_t blk = idx / half, k = idx % half;
        Fr* lo = v + blk * m + k;
        Fr* hi = lo + half;
        const Fr t = *hi * roots[k * stride];
        *hi = *lo - t;
        *lo = *lo + t;
      }
    });
  }
}

static void distribute_powers(Fr* v, size_t n, const Fr& g, const Fr& c) {   // v[i] *= c * g^i
  parallel_for(n, [&](size_t b, size_t e) {
    u64 eb[1] = {(u64)b};
    Fr p = c * g.pow(eb, 1);
    for (size_t i = b; i < e; i++) {
      v[i] = v[i] * p;
      p = p * g;
    }
  });
}

static void domain_fft(const Domain& d, Fr* v, const Fr* offset) {
  if (offset) distribute_powers(v, d.n, *offset, Fr::one());
  fft_core(v, d.log_n, d.gen);
}
static void domain_ifft(const Domain& d, Fr* v, const Fr* offset) {
  fft_core(v, d.log_n, d.gen_inv);
  if (offset) distribute_powers(v, d.n, offset->inverse(), d.size_inv);
  else parallel_for(d.n, [&](size_t b, size_t e) { for (size_t i = b; i < e; i++) v[i] = v[i] * d.size_inv; });
}

// LibsnarkReduction::witness_map_from_matrices after the row evaluations; h left in a
static void witness_map(Fr* a, Fr* b, Fr* c, uint32_t log_n) {
  Domain d(log_n);
  const Fr g = g_generator;
  domain_ifft(d, a, nullptr);
  domain_ifft(d, b, nullptr);
  domain_fft(d, a, &g);
  domain_fft(d, b, &g);
  parallel_for(d.n, [&](size_t s, size_t e) { for (size_t i = s; i < e; i++) a[i] = a[i] * b[i]; });
  domain_ifft(d, c, nullptr);
  domain_fft(d, c, &g);
  Fr gn = g;
  for (uint32_t i = 0; i < log_n; i++) gn = gn.sqr();
  const Fr zinv = (gn - Fr::one()).inverse();
  parallel_for(d.n, [&](size_t s, size_t e) { for (size_t i = s; i < e; i++) a[i] = (a[i] - c[i]) * zinv; });
  domain_ifft(d, a, &g);
}

// ------------------------------------------------------------------ ark-ec msm_bigint (wnaf variant)
static size_t ark_window(size_t n) {
  if (n < 32) return 3;
  size_t lg = 0;
  while (((size_t)1 << lg) < n) lg++;
  return lg * 69 / 100 + 2;
}

static void make_digits(const u64* k, size_t w, size_t num_bits, int64_t* out, size_t count) {
  const u64 radix = (u64)1 << w, mask = radix - 1;
  u64 carry = 0;
  for (size_t i = 0; i < count; i++) {
    const size_t bit = i * w, idx = bit / 64, sh = bit % 64;
    u64 buf = 0;
    if (idx < 4) {
      buf = k[idx] >> sh;
      if (sh + w > 64 && idx + 1 < 4) buf |= k[idx + 1] << (64 - sh);
    }
    const u64 coef = carry + (buf & mask);
    carry = (coef + radix / 2) >> w;
    out[i] = (int64_t)coef - (int64_t)(carry << w);
  }
  out[count - 1] += (int64_t)(carry << w);
  (void)num_bits;
}

template <class F>
static Jac<F> msm_bigint(const Affine<F>* bases, const u64* scalars, size_t size) {
  typedef Jac<F> J;
  if (size == 0) return J::identity();
  const size_t c = ark_window(size), num_bits = 255, digits_count = (num_bits + c - 1) / c;
  std::vector<int64_t> digits(size * digits_count);
  parallel_for(size, [&](size_t b, size_t e) {
    for (size_t i = b; i < e; i++) make_digits(scalars + 4 * i, c, num_bits, &digits[i * digits_count], digits_count);
  });
  std::vector<J> window_sums(digits_count, J::identity());
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    for (;;) {
      const size_t w = next.fetch_add(1);
      if (w >= digits_count) return;
      std::vector<J> buckets((size_t)1 << c, J::identity());
      for (size_t i = 0; i < size; i++) {
        const int64_t d = digits[i * digits_count + w];
        if (d > 0) buckets[d - 1] = buckets[d - 1].add_mixed(bases[i]);
        else if (d < 0) {
          Affine<F> nb = bases[i];
          nb.y = nb.y.neg();
          buckets[-d - 1] = buckets[-d - 1].add_mixed(nb);
        }
      }
      J running = J::identity(), res = J::identity();
      for (size_t b = buckets.size(); b-- > 0;) {
        running = running.add(buckets[b]);
        res = res.add(running);
      }
      window_sums[w] = res;
    }
  };
  const int t = std::min<size_t>(g_threads, digits_count);   // rayon: parallel over windows only
  std::vector<std::thread> pool;
  for (int i = 1; i < t; i++) pool.emplace_back(worker);
  worker();
  for (auto& th : pool) th.join();
  J total = window_sums[digits_count - 1];
  for (size_t w = digits_count - 1; w-- > 0;) {
    for (size_t i = 0; i < c; i++) total = total.dbl();
    total = total.add(window_sums[w]);
  }
  return total;
}

// ------------------------------------------------------------------ layouts
static void load_g1(std::vector<Affine<Fq>>& out, const u64* limbs, const uint8_t* inf, size_t n) {
  out.resize(n);
  for (size_t i = 0; i < n; i++) {
    memcpy(out[i].x.l, limbs + 12 * i, 48);
    memcpy(out[i].y.l, limbs + 12 * i + 6, 48);
    out[i].inf = inf && ((inf[i >> 3] >> (i & 7)) & 1);
  }
}
static void load_g2(std::vector<Affine<Fq2>>& out, const u64* limbs, const uint8_t* inf, size_t n) {
  out.resize(n);
  for (size_t i = 0; i < n; i++) {
    memcpy(out[i].x.c0.l, limbs + 24 * i, 48);
    memcpy(out[i].x.c1.l, limbs + 24 * i + 6, 48);
    memcpy(out[i].y.c0.l, limbs + 24 * i + 12, 48);
    memcpy(out[i].y.c1.l, limbs + 24 * i + 18, 48);
    out[i].inf = inf && ((inf[i >> 3] >> (i & 7)) & 1);
  }
}

static void fq_to_be(uint8_t* dst, const Fq& v) {
  u64 c[6];
  v.to_canonical(c);
  for (int i = 0; i < 6; i++)
    for (int b = 0; b < 8; b++) dst[8 * i + b] = (uint8_t)(c[5 - i] >> (56 - 8 * b));
}
static bool fq_larger(const Fq& v) {   // canonical v > (q-1)/2
  u64 c[6], d[6];
  v.to_canonical(c);
  u64 carry = 0;
  for (int i = 0; i < 6; i++) { d[i] = (c[i] << 1) | carry; carry = c[i] >> 63; }
  for (int i = 5; i >= 0; i--) if (d[i] != kFqModulus[i]) return d[i] > kFqModulus[i];
  return false;
}
static void ser_g1(uint8_t* dst, const Affine<Fq>& p) {
  if (p.inf) { memset(dst, 0, 48); dst[0] = 0xC0; return; }
  fq_to_be(dst, p.x);
  dst[0] |= 0x80;
  if (fq_larger(p.y)) dst[0] |= 0x20;
}
static void ser_g2(uint8_t* dst, const Affine<Fq2>& p) {
  if (p.inf) { memset(dst, 0, 96); dst[0] = 0xC0; return; }
  fq_to_be(dst, p.x.c1);
  fq_to_be(dst + 48, p.x.c0);
  dst[0] |= 0x80;
  const bool larger = p.y.c1.is_zero() ? fq_larger(p.y.c0) : fq_larger(p.y.c1);
  if (larger) dst[0] |= 0x20;
}

template <class F>
static Jac<F> calculate_coeff(const Jac<F>& initial, const std::vector<Affine<F>>& query, const Affine<F>& vk_param,
                              const u64* assignment /* z[1..] canonical */) {
  Jac<F> acc = msm_bigint<F>(query.data() + 1, assignment, query.size() - 1);
  Jac<F> res = initial.add_mixed(query[0]);
  res = res.add(acc);
  return res.add_mixed(vk_param);
}


// ------------------------------------------------------------------ ark-ec FixedBase::msm (key generation)
// out[i] = k_i * G for canonical scalars k_i: unsigned 8-bit windows over per-window tables
// table[w][d] = d * 2^(8w) G, Jacobian mixed additions, one batched normalisation per thread chunk
// (ark-groth16's generator calls FixedBase::msm for every query array; reference call sites
// src/arkworks/backend/matrix_proof.rs:128-131, fibbonaci_handler.rs:107).
template <class F>
struct FixedBaseTable {
  std::vector<Affine<F>> t;   // 32 x 256
  explicit FixedBaseTable(const Affine<F>& g) : t(32 * 256) {
    std::vector<Jac<F>> j(32 * 256);
    Jac<F> base{g.x, g.y, F::one()};
    for (int w = 0; w < 32; w++) {
      Jac<F> acc = Jac<F>::identity();
      for (int d = 0; d < 256; d++) {
        j[w * 256 + d] = acc;
        acc = acc.add(base);
      }
      base = acc;   // 256 * base
    }
    parallel_for(j.size(), [&](size_t b, size_t e) { for (size_t i = b; i < e; i++) t[i] = j[i].to_affine(); });
  }
};

template <class F>
static void batch_to_affine(const Jac<F>* in, Affine<F>* out, size_t n) {
  // Montgomery's trick on the z coordinates
  std::vector<F> pref(n);
  F acc = F::one();
  for (size_t i = 0; i < n; i++) {
    pref[i] = acc;
    if (!in[i].is_identity()) acc = acc * in[i].z;
  }
  F inv = acc.inverse();
  for (size_t i = n; i-- > 0;) {
    if (in[i].is_identity()) { out[i] = Affine<F>{F::zero(), F::zero(), true}; continue; }
    const F zi = inv * pref[i];
    inv = inv * in[i].z;
    const F zi2 = zi.sqr();
    out[i] = Affine<F>{in[i].x * zi2, in[i].y * zi2 * zi, false};
  }
}

template <class F>
static void fixed_base_msm(const FixedBaseTable<F>& tab, const u64* scalars, size_t n, Affine<F>* out) {
  parallel_for(n, [&](size_t b, size_t e) {
    const size_t kChunk = 1024;
    std::vector<Jac<F>> tmp(kChunk);
    for (size_t s0 = b; s0 < e; s0 += kChunk) {
      const size_t cnt = std::min(kChunk, e - s0);
      for (size_t i = 0; i < cnt; i++) {
        const u64* k = scalars + 4 * (s0 + i);
        Jac<F> acc = Jac<F>::identity();
        for (int w = 0; w < 32; w++) {
          const unsigned d = (unsigned)(k[w >> 3] >> ((w & 7) * 8)) & 0xff;
          if (d) acc = acc.add_mixed(tab.t[w * 256 + d]);
        }
        tmp[i] = acc;
      }
      batch_to_affine<F>(tmp.data(), out + s0, cnt);
    }
  });
}

static void store_g1(u64* limbs, uint8_t* inf, const std::vector<Affine<Fq>>& v) {
  if (inf) memset(inf, 0, (v.size() + 7) / 8);
  for (size_t i = 0; i < v.size(); i++) {
    memcpy(limbs + 12 * i, v[i].x.l, 48);
    memcpy(limbs + 12 * i + 6, v[i].y.l, 48);
    if (v[i].inf && inf) inf[i >> 3] |= (uint8_t)(1u << (i & 7));
  }
}
static void store_g2(u64* limbs, uint8_t* inf, const std::vector<Affine<Fq2>>& v) {
  if (inf) memset(inf, 0, (v.size() + 7) / 8);
  for (size_t i = 0; i < v.size(); i++) {
    memcpy(limbs + 24 * i, v[i].x.c0.l, 48);
    memcpy(limbs + 24 * i + 6, v[i].x.c1.l, 48);
    memcpy(limbs + 24 * i + 12, v[i].y.c0.l, 48);
    memcpy(limbs + 24 * i + 18, v[i].y.c1.l, 48);
    if (v[i].inf && inf) inf[i >> 3] |= (uint8_t)(1u << (i & 7));
  }
}

static Fr fr_from_canonical(const u64* c) {
  Fr x, r2;
  memcpy(x.l, c, 32);
  memcpy(r2.l, g_fr.r2, 32);
  return x * r2;
}

struct Csr {
  const u64* rp;
  const uint32_t* ci;
  const u64* cf;   // Montgomery limbs, 4 per entry
};

}  // namespace

// ------------------------------------------------------------------ C entry points (ctypes)
extern "C" {

void ark_cpu_set_threads(int t) { g_threads = t < 1 ? 1 : t; }
int ark_cpu_hardware_threads() { int t = (int)std::thread::hardware_concurrency(); return t < 1 ? 1 : t; }

// data: n Fr elements (Montgomery limbs), in place; coset_gen NULL or 4 limbs (Montgomery)
void ark_cpu_ntt(u64* data, uint32_t log_n, int inverse, const u64* coset_gen) {
  Domain d(log_n);
  Fr g;
  if (coset_gen) memcpy(g.l, coset_gen, 32);
  if (inverse) domain_ifft(d, reinterpret_cast<Fr*>(data), coset_gen ? &g : nullptr);
  else domain_fft(d, reinterpret_cast<Fr*>(data), coset_gen ? &g : nullptr);
}

// a, b, c: n Montgomery elements each (clobbered); h_out: n elements
void ark_cpu_witness_map(u64* a, u64* b, u64* c, uint32_t log_n, u64* h_out) {
  witness_map(reinterpret_cast<Fr*>(a), reinterpret_cast<Fr*>(b), reinterpret_cast<Fr*>(c), log_n);
  memcpy(h_out, a, ((size_t)32) << log_n);
}

// out: affine x, y (12 limbs); returns 1 if the result is the identity
int ark_cpu_msm_g1(const u64* bases, const uint8_t* inf, const u64* scalars, u64 n, u64* out_affine) {
  std::vector<Affine<Fq>> pts;
  load_g1(pts, bases, inf, n);
  Affine<Fq> r = msm_bigint<Fq>(pts.data(), scalars, n).to_affine();
  memcpy(out_affine, r.x.l, 48);
  memcpy(out_affine + 6, r.y.l, 48);
  return r.inf ? 1 : 0;
}
int ark_cpu_msm_g2(const u64* bases, const uint8_t* inf, const u64* scalars, u64 n, u64* out_affine) {
  std::vector<Affine<Fq2>> pts;
  load_g2(pts, bases, inf, n);
  Affine<Fq2> r = msm_bigint<Fq2>(pts.data(), scalars, n).to_affine();
  memcpy(out_affine, r.x.c0.l, 48);
  memcpy(out_affine + 6, r.x.c1.l, 48);
  memcpy(out_affine + 12, r.y.c0.l, 48);
  memcpy(out_affine + 18, r.y.c1.l, 48);
  return r.inf ? 1 : 0;
}

struct ark_cpu_pk {
  u64 m, l;
  uint32_t log_n;
  std::vector<Affine<Fq>> a, b1, h, lq;
  std::vector<Affine<Fq2>> b2;
  Affine<Fq> alpha, beta1, delta1;
  Affine<Fq2> beta2, delta2;
};

ark_cpu_pk* ark_cpu_pk_new(u64 m, u64 l, uint32_t log_n, const u64* a, const uint8_t* a_inf, const u64* b1,
                           const uint8_t* b1_inf, const u64* b2, const uint8_t* b2_inf, const u64* h,
                           const uint8_t* h_inf, const u64* lq, const uint8_t* l_inf, const u64* alpha,
                           const u64* beta1, const u64* delta1, const u64* beta2, const u64* delta2) {
  ark_cpu_pk* pk = new ark_cpu_pk();
  pk->m = m; pk->l = l; pk->log_n = log_n;
  load_g1(pk->a, a, a_inf, m);
  load_g1(pk->b1, b1, b1_inf, m);
  load_g2(pk->b2, b2, b2_inf, m);
  load_g1(pk->h, h, h_inf, ((size_t)1 << log_n) - 1);
  load_g1(pk->lq, lq, l_inf, m - l);
  std::vector<Affine<Fq>> t;
  load_g1(t, alpha, nullptr, 1); pk->alpha = t[0];
  load_g1(t, beta1, nullptr, 1); pk->beta1 = t[0];
  load_g1(t, delta1, nullptr, 1); pk->delta1 = t[0];
  std::vector<Affine<Fq2>> t2;
  load_g2(t2, beta2, nullptr, 1); pk->beta2 = t2[0];
  load_g2(t2, delta2, nullptr, 1); pk->delta2 = t2[0];
  return pk;
}
void ark_cpu_pk_free(ark_cpu_pk* pk) { delete pk; }

// create_proof_with_reduction after synthesis.  a/b/c evals and z in Montgomery form
// (a/b/c clobbered); r, s Montgomery.  proof_out: 192 bytes.
void ark_cpu_prove(const ark_cpu_pk* pk, u64* a, u64* b, u64* c, const u64* z, const u64* r, const u64* s,
                   uint8_t* proof_out) {
  const size_t n = (size_t)1 << pk->log_n, m = pk->m, l = pk->l;
  witness_map(reinterpret_cast<Fr*>(a), reinterpret_cast<Fr*>(b), reinterpret_cast<Fr*>(c), pk->log_n);
  std::vector<u64> hbig(4 * n), zbig(4 * m);
  parallel_for(n, [&](size_t s0, size_t e) {
    for (size_t i = s0; i < e; i++) reinterpret_cast<const Fr*>(a)[i].to_canonical(&hbig[4 * i]);
  });
  parallel_for(m, [&](size_t s0, size_t e) {
    for (size_t i = s0; i < e; i++) reinterpret_cast<const Fr*>(z)[i].to_canonical(&zbig[4 * i]);
  });
  Fr rm, sm;
  memcpy(rm.l, r, 32);
  memcpy(sm.l, s, 32);
  u64 rb[4], sb[4];
  rm.to_canonical(rb);
  sm.to_canonical(sb);
  typedef Jac<Fq> J1;
  typedef Jac<Fq2> J2;
  const J1 h_acc = msm_bigint<Fq>(pk->h.data(), hbig.data(), std::min(pk->h.size(), n));
  const J1 l_acc = msm_bigint<Fq>(pk->lq.data(), zbig.data() + 4 * l, m - l);
  const J1 delta1 = J1{pk->delta1.x, pk->delta1.y, Fq::one()};
  const J1 rs_delta = delta1.mul_bigint(rb, 4).mul_bigint(sb, 4);
  const J1 g_a = calculate_coeff<Fq>(delta1.mul_bigint(rb, 4), pk->a, pk->alpha, zbig.data() + 4);
  const J1 s_g_a = g_a.mul_bigint(sb, 4);
  J1 g1_b = J1::identity();
  if (!rm.is_zero()) g1_b = calculate_coeff<Fq>(delta1.mul_bigint(sb, 4), pk->b1, pk->beta1, zbig.data() + 4);
  const J2 delta2 = J2{pk->delta2.x, pk->delta2.y, Fq2::one()};
  const J2 g2_b = calculate_coeff<Fq2>(delta2.mul_bigint(sb, 4), pk->b2, pk->beta2, zbig.data() + 4);
  const J1 r_g1_b = g1_b.mul_bigint(rb, 4);
  J1 g_c = s_g_a.add(r_g1_b).add(rs_delta.neg()).add(l_acc).add(h_acc);
  ser_g1(proof_out, g_a.to_affine());
  ser_g2(proof_out + 48, g2_b.to_affine());
  ser_g1(proof_out + 144, g_c.to_affine());
}


// ark-groth16 generator.rs generate_parameters_with_qap::<LibsnarkReduction> with caller-supplied toxic waste and
// generators (Groth16::setup / circuit_specific_setup draw these from the rng: matrix_proof.rs:128-131).
//   matrices: CSR (row_ptr u64[nc + 1], cols u32, coeffs Montgomery limbs); toxic: alpha, beta, gamma, delta, tau as
//   canonical 4-limb integers; g1_gen / g2_gen: affine generators, Montgomery limbs.
//   outputs (caller-allocated): the ProvingKey arrays in the layout of b2z_pk_desc plus gamma_g2 and gamma_abc_g1.
int ark_cpu_groth16_setup(u64 nc, u64 l, u64 m, const u64* a_rp, const uint32_t* a_ci, const u64* a_cf, const u64* b_rp,
                          const uint32_t* b_ci, const u64* b_cf, const u64* c_rp, const uint32_t* c_ci, const u64* c_cf,
                          const u64* toxic, const u64* g1_gen, const u64* g2_gen, u64* a_q, uint8_t* a_inf, u64* b1_q,
                          uint8_t* b1_inf, u64* b2_q, uint8_t* b2_inf, u64* h_q, uint8_t* h_inf, u64* l_q, uint8_t* l_inf,
                          u64* alpha_g1, u64* beta_g1, u64* delta_g1, u64* beta_g2, u64* gamma_g2, u64* delta_g2,
                          u64* gamma_abc, uint8_t* gamma_abc_inf) {
  uint32_t log_n = 0;
  while (((u64)1 << log_n) < nc + l) log_n++;
  if (log_n > 32) return 2;
  const size_t n = (size_t)1 << log_n;
  const bool verbose = getenv("ARK_CPU_VERBOSE") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    auto now = std::chrono::steady_clock::now();
    if (verbose) fprintf(stderr, "[ark_cpu setup] %-12s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
    t_last = now;
  };
  const Fr alpha = fr_from_canonical(toxic), beta = fr_from_canonical(toxic + 4), gamma = fr_from_canonical(toxic + 8),
           delta = fr_from_canonical(toxic + 12), tau = fr_from_canonical(toxic + 16);
  Domain d(log_n);
  // evaluate_all_lagrange_coefficients(tau): L_i = Z(tau)/n * w^i / (tau - w^i)
  Fr tn = tau;
  for (uint32_t i = 0; i < log_n; i++) tn = tn.sqr();
  const Fr zt = tn - Fr::one();
  if (zt.is_zero()) return 1;
  std::vector<Fr> lag(n), den(n);
  parallel_for(n, [&](size_t b, size_t e) {
    u64 eb[1] = {(u64)b};
    Fr w = d.gen.pow(eb, 1);
    for (size_t i = b; i < e; i++) { lag[i] = w; den[i] = tau - w; w = w * d.gen; }
  });
  parallel_for(n, [&](size_t b, size_t e) {   // batch inversion per chunk
    std::vector<Fr> pref(e - b);
    Fr acc = Fr::one();
    for (size_t i = b; i < e; i++) { pref[i - b] = acc; acc = acc * den[i]; }
    Fr inv = acc.inverse();
    const Fr zn = zt * d.size_inv;
    for (size_t i = e; i-- > b;) {
      const Fr di = inv * pref[i - b];
      inv = inv * den[i];
      lag[i] = zn * lag[i] * di;
    }
  });
  lap("lagrange");
  // instance_map_with_evaluation: a_i(tau), b_i(tau), c_i(tau) per variable
  std::vector<Fr> at(m, Fr::zero()), bt(m, Fr::zero()), ct(m, Fr::zero());
  for (u64 j = 0; j < l; j++) at[j] = lag[nc + j];
  const Csr mats[3] = {{a_rp, a_ci, a_cf}, {b_rp, b_ci, b_cf}, {c_rp, c_ci, c_cf}};
  std::vector<Fr>* outs[3] = {&at, &bt, &ct};
  {
    std::vector<std::thread> pool;   // one thread per matrix: column scatter is not row-parallel
    for (int k = 0; k < 3; k++)
      pool.emplace_back([&, k] {
        const Csr& M = mats[k];
        std::vector<Fr>& o = *outs[k];
        for (u64 i = 0; i < nc; i++)
          for (u64 e = M.rp[i]; e < M.rp[i + 1]; e++) {
            Fr cf;
            memcpy(cf.l, M.cf + 4 * e, 32);
            o[M.ci[e]] = o[M.ci[e]] + lag[i] * cf;
          }
      });
    for (auto& th : pool) th.join();
  }
  lap("qap");
  const Fr ginv = gamma.inverse(), dinv = delta.inverse();
  auto canon = [](const std::vector<Fr>& v) {
    std::vector<u64> out(4 * v.size());
    parallel_for(v.size(), [&](size_t b, size_t e) { for (size_t i = b; i < e; i++) v[i].to_canonical(&out[4 * i]); });
    return out;
  };
  Affine<Fq> g1;
  memcpy(g1.x.l, g1_gen, 48); memcpy(g1.y.l, g1_gen + 6, 48); g1.inf = false;
  Affine<Fq2> g2;
  memcpy(g2.x.c0.l, g2_gen, 48); memcpy(g2.x.c1.l, g2_gen + 6, 48);
  memcpy(g2.y.c0.l, g2_gen + 12, 48); memcpy(g2.y.c1.l, g2_gen + 18, 48); g2.inf = false;
  const FixedBaseTable<Fq> t1(g1);
  const FixedBaseTable<Fq2> t2(g2);
  lap("tables");
  auto mul1 = [&](const std::vector<Fr>& v, u64* out, uint8_t* inf) {
    std::vector<u64> k = canon(v);
    std::vector<Affine<Fq>> pts(v.size());
    fixed_base_msm<Fq>(t1, k.data(), v.size(), pts.data());
    store_g1(out, inf, pts);
  };
  auto mul2 = [&](const std::vector<Fr>& v, u64* out, uint8_t* inf) {
    std::vector<u64> k = canon(v);
    std::vector<Affine<Fq2>> pts(v.size());
    fixed_base_msm<Fq2>(t2, k.data(), v.size(), pts.data());
    store_g2(out, inf, pts);
  };
  mul1(at, a_q, a_inf);
  lap("a_query");
  mul1(bt, b1_q, b1_inf);
  lap("b_g1_query");
  mul2(bt, b2_q, b2_inf);
  lap("b_g2_query");
  {
    std::vector<Fr> hs(n - 1);
    parallel_for(n - 1, [&](size_t b, size_t e) {
      u64 eb[1] = {(u64)b};
      Fr t = zt * dinv * tau.pow(eb, 1);
      for (size_t i = b; i < e; i++) { hs[i] = t; t = t * tau; }
    });
    mul1(hs, h_q, h_inf);
    lap("h_query");
  }
  std::vector<Fr> abc(m);
  parallel_for(m, [&](size_t b, size_t e) { for (size_t i = b; i < e; i++) abc[i] = beta * at[i] + alpha * bt[i] + ct[i]; });
  {
    std::vector<Fr> lq(m - l), gq(l);
    for (u64 i = l; i < m; i++) lq[i - l] = abc[i] * dinv;
    for (u64 i = 0; i < l; i++) gq[i] = abc[i] * ginv;
    mul1(lq, l_q, l_inf);
    mul1(gq, gamma_abc, gamma_abc_inf);
  }
  uint8_t dummy[1];
  mul1(std::vector<Fr>{alpha}, alpha_g1, dummy);
  mul1(std::vector<Fr>{beta}, beta_g1, dummy);
  mul1(std::vector<Fr>{delta}, delta_g1, dummy);
  mul2(std::vector<Fr>{beta}, beta_g2, dummy);
  mul2(std::vector<Fr>{gamma}, gamma_g2, dummy);
  mul2(std::vector<Fr>{delta}, delta_g2, dummy);
  lap("l, vk");
  return 0;
}

// constraint-row evaluation (the evaluate_constraint loop of witness_map_from_matrices), threads over rows
void ark_cpu_constraint_evals(u64 nc, u64 l, uint32_t log_n, const u64* a_rp, const uint32_t* a_ci, const u64* a_cf,
                              const u64* b_rp, const uint32_t* b_ci, const u64* b_cf, const u64* c_rp,
                              const uint32_t* c_ci, const u64* c_cf, const u64* z, u64* a_out, u64* b_out, u64* c_out) {
  const size_t n = (size_t)1 << log_n;
  memset(a_out, 0, 32 * n); memset(b_out, 0, 32 * n); memset(c_out, 0, 32 * n);
  const Csr mats[3] = {{a_rp, a_ci, a_cf}, {b_rp, b_ci, b_cf}, {c_rp, c_ci, c_cf}};
  u64* outs[3] = {a_out, b_out, c_out};
  const Fr* zz = reinterpret_cast<const Fr*>(z);
  for (int k = 0; k < 3; k++)
    parallel_for(nc, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; i++) {
        Fr acc = Fr::zero();
        for (u64 t = mats[k].rp[i]; t < mats[k].rp[i + 1]; t++) {
          Fr cf;
          memcpy(cf.l, mats[k].cf + 4 * t, 32);
          acc = acc + cf * zz[mats[k].ci[t]];
        }
        memcpy(outs[k] + 4 * i, acc.l, 32);
      }
    });
  memcpy(a_out + 4 * nc, z, 32 * l);
}

}  // extern "C"
