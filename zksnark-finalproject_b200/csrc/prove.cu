// Groth16 prover assembly -- replaces ark-groth16 0.4's create_proof_with_reduction /
// create_proof_with_assignment minus circuit synthesis (reference call sites:
// src/arkworks/backend/fibbonaci_handler.rs:110, matrix_proof.rs:139-140,
// prime_snark.rs:119; equations: SURVEY.md A.4):
//
//   A  = alpha_1 + sum z_i a_i + r delta_1
//   B  = beta_2  + sum z_i b2_i + s delta_2
//   C  = s A + r B1 - r s delta_1 + sum_{i>=l} z_i l_i + sum h_i hq_i ,   B1 = beta_1 + sum z_i b1_i + s delta_1
//
// arkworks evaluates C with two 255-bit scalar multiplications of the freshly
// computed A and B1.  A serial double-and-add is the worst thing to run on a GPU
// (one field product occupies a lone warp for ~2000 cycles), so C is expanded instead:
//
//   C  = sum (s z_i) a_i + sum (r z_i) b1_i + sum_{i>=l} z_i l_i + s alpha_1 + r beta_1 + (r s) delta_1   [C_z]
//      + sum h_i hq_i                                                                                   [C_h]
//
// i.e. the same group element as ONE more multi-scalar multiplication over bases
// that are already on the device.  The whole proof is then four MSMs
//   A   : G1 bases [a | b1 | l | alpha beta delta], scalars [z | 0 | 0 | 1 0 r]
//   C_z : same bases,                              scalars [s z | r z | z_{>=l} | s r rs]
//   C_h : G1 h_query in bit-reversed order (the order the DIF witness map leaves h in)
//   B   : G2 bases [b2 | beta_2 delta_2],           scalars [z | 1 s]
// on four streams; A, C_z and B only need z and overlap the witness map.  The host
// finishes with one point addition, three affine normalisations and the serialization
// (host_fq.hpp).  r = 0 needs no special case: r B1 vanishes in the expansion exactly
// as arkworks' `if r.is_zero()` branch makes it vanish.
#include <cstring>

#include "api_glue.hpp"
#include "host_fq.hpp"
#include "msm.hpp"

namespace b2z {

namespace {

// dst[p] = src[bitrev(p)] when that index exists, else identity (flagged)
__global__ void permute_bitrev_g1_kernel(const G1::Affine* __restrict__ src, const uint32_t* __restrict__ src_inf,
                                         uint32_t n_src, uint32_t log_n, G1::Affine* dst, uint32_t* dst_flags) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (1u << log_n)) return;
  const uint32_t j = log_n == 0 ? 0 : (__brev(p) >> (32 - log_n));
  bool inf = j >= n_src;
  if (!inf && src_inf != nullptr) inf = (src_inf[j >> 5] >> (j & 31)) & 1;
  if (inf) {
    G1::Affine z;
    z.x = Fq::zero();
    z.y = Fq::zero();
    dst[p] = z;
  } else {
    dst[p] = src[j];
  }
  dst_flags[p] = inf ? 1u : 0u;
}

__global__ void pack_flags_kernel2(const uint32_t* flags, uint32_t n, uint32_t* words) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= (n + 31) / 32) return;
  uint32_t v = 0;
  for (uint32_t j = 0; j < 32; j++) {
    const uint32_t i = w * 32 + j;
    if (i < n && flags[i]) v |= 1u << j;
  }
  words[w] = v;
}

struct ScalarPrep {
  const FrEl* z;     // m Montgomery elements
  FrEl* out_a;       // n1 canonical scalars for A
  FrEl* out_c;       // n1 canonical scalars for C_z
  uint32_t m, l;
  FrEl r, s, rs;     // canonical integers (NOT Montgomery): mont_mul(k, z*R) = k*z
};

// index space: [0,m) a-part, [m,2m) b1-part, [2m, 3m-l) l-part, then alpha, beta, delta
__global__ void scalar_prep_kernel(ScalarPrep a) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n1 = 3 * a.m - a.l + 3;
  if (i >= n1) return;
  FrEl one = Fr::zero();
  one.l[0] = 1;
  FrEl sa = Fr::zero(), sc;
  if (i < a.m) {
    const FrEl z = a.z[i];
    sa = Fr::reduce(Fr::mul(one, z));
    sc = Fr::reduce(Fr::mul(a.s, z));
  } else if (i < 2 * a.m) {
    sc = Fr::reduce(Fr::mul(a.r, a.z[i - a.m]));
  } else if (i < 3 * a.m - a.l) {
    sc = Fr::reduce(Fr::mul(one, a.z[i - 2 * a.m + a.l]));
  } else {
    const uint32_t k = i - (3 * a.m - a.l);
    sa = k == 0 ? one : (k == 2 ? a.r : Fr::zero());
    sc = k == 0 ? a.s : (k == 1 ? a.r : a.rs);
  }
  a.out_a[i] = sa;
  a.out_c[i] = sc;
}

inline uint32_t nblk(uint64_t n, uint32_t t) { return (uint32_t)((n + t - 1) / t); }

FrEl fr_load_host(const uint64_t v[4]) {
  FrEl x;
  std::memcpy(x.l, v, 32);
  return x;
}

}  // namespace

// ---------------------------------------------------------------------------
// Proving key
// ---------------------------------------------------------------------------
struct PkImpl {
  uint32_t log_n = 0;
  uint64_t m = 0, l = 0;
  uint32_t n1 = 0;             // 3m - l + 3 bases in g1_all
  MsmBases<G1> g1_all;         // [a_query | b_g1_query | l_query | alpha_1 beta_1 delta_1]
  MsmBases<G1> h;              // h_query, bit-reversed order, padded to n
  MsmBases<G2> g2;             // [b_g2_query | beta_2 delta_2]
  // per-proof device scratch
  DevBuf<FrEl> ea, eb, ec, z, scal_a, scal_c, hc, tail;
  DevBuf<G1::Xyzz> g1_out;     // A, C_z, C_h
  DevBuf<G2::Xyzz> g2_out;     // B
  uint32_t* h_out = nullptr;   // pinned: 3 G1 XYZZ + 1 G2 XYZZ
  cudaEvent_t ev_z = nullptr, ev_done[3] = {nullptr, nullptr, nullptr};
  ~PkImpl() {
    if (h_out) cudaFreeHost(h_out);
    if (ev_z) cudaEventDestroy(ev_z);
    for (auto& e : ev_done)
      if (e) cudaEventDestroy(e);
  }
};

namespace {

// copies `count` points (+ identity bits) from host arrays into dst[at ...)
template <class Affine>
void stage_points(Affine* d_dst, std::vector<uint32_t>& words, uint64_t at, const uint64_t* pts, const uint8_t* inf,
                  uint64_t count, cudaStream_t st) {
  if (count == 0) return;
  B2Z_CUDA(cudaMemcpyAsync(d_dst + at, pts, count * sizeof(Affine), cudaMemcpyHostToDevice, st));
  if (inf != nullptr)
    for (uint64_t i = 0; i < count; i++)
      if ((inf[i >> 3] >> (i & 7)) & 1) words[(at + i) >> 5] |= 1u << ((at + i) & 31);
}

size_t pk_precompute_bytes(const b2z_pk_desc* d) {
  auto cost = [](uint64_t n, size_t pt) {
    if (n == 0) return (size_t)0;
    const uint32_t c = msm_pick_c(n, true);
    return (size_t)n * pt * msm_windows(c);
  };
  const uint64_t n = 1ull << d->log_domain;
  return cost(3 * d->num_variables - d->num_instance + 3, 96) + cost(d->num_variables + 2, 192) + cost(n, 96);
}

template <class C>
void msm_entry(Ctx& c, const uint64_t* bases, const uint8_t* inf_bitmap, const uint64_t* scalars, uint64_t n,
               uint64_t* out_xyz) {
  using Affine = typename C::Affine;
  using Xyzz = typename C::Xyzz;
  constexpr int kLimbs = sizeof(Affine) / 16;   // u64 limbs per coordinate (6 or 12)
  B2Z_REQUIRE(out_xyz != nullptr, B2Z_EINVAL, "msm: out is NULL");
  B2Z_REQUIRE(n == 0 || (bases != nullptr && scalars != nullptr), B2Z_EINVAL, "msm: NULL input");
  B2Z_REQUIRE(n < (1ull << 28), B2Z_ESIZE, "msm: more than 2^28 points in one call");
  cudaStream_t st = c.stream;
  // identity as arkworks represents it: (1, 1, 0), Montgomery limbs
  std::memset(out_xyz, 0, 3 * kLimbs * 8);
  std::memcpy(out_xyz, host::kOne, 48);
  std::memcpy(out_xyz + kLimbs, host::kOne, 48);
  if (n == 0) return;
  MsmBases<C> B;
  {
    DevBuf<Affine> d(n);
    std::vector<uint32_t> words((n + 31) / 32, 0u);
    stage_points<Affine>(d.p, words, 0, bases, inf_bitmap, n, st);
    DevBuf<uint32_t> dinf;
    if (inf_bitmap != nullptr) {
      dinf.alloc(words.size());
      B2Z_CUDA(cudaMemcpyAsync(dinf.p, words.data(), words.size() * 4, cudaMemcpyHostToDevice, st));
    }
    msm_bases_build<C>(&c, B, d.p, dinf.p, (uint32_t)n, /*precompute=*/false, 0, st);
    B2Z_CUDA(cudaStreamSynchronize(st));
  }
  DevBuf<FrEl> ds(n);
  B2Z_CUDA(cudaMemcpyAsync(ds.p, scalars, n * sizeof(FrEl), cudaMemcpyHostToDevice, st));
  DevBuf<Xyzz> res(1);
  msm_run<C>(&c, 0, B, ds.p, (uint32_t)n, nullptr, res.p, st);
  uint32_t h_res[sizeof(Xyzz) / 4];
  B2Z_CUDA(cudaMemcpyAsync(h_res, res.p, sizeof(Xyzz), cudaMemcpyDeviceToHost, st));
  B2Z_CUDA(cudaStreamSynchronize(st));
  // affine normalisation on the host (one inversion); Z = 1
  if (std::is_same<C, G1>::value) {
    host::Fq x, y;
    if (!host::g1_to_affine(host::g1_from_device(h_res), &x, &y)) {
      std::memcpy(out_xyz, x.l, 48);
      std::memcpy(out_xyz + 6, y.l, 48);
      std::memcpy(out_xyz + 12, host::kOne, 48);
    }
  } else {
    host::Fq2 x, y;
    if (!host::g2_to_affine(host::g2_from_device(h_res), &x, &y)) {
      std::memcpy(out_xyz, &x, 96);
      std::memcpy(out_xyz + 12, &y, 96);
      std::memset(out_xyz + 24, 0, 96);
      std::memcpy(out_xyz + 24, host::kOne, 48);
    }
  }
}

template <class C>
void fixed_base_entry(Ctx& c, const uint64_t* scalars, uint64_t n, uint64_t* out_points, uint8_t* out_inf) {
  using Affine = typename C::Affine;
  B2Z_REQUIRE(n == 0 || (scalars && out_points && out_inf), B2Z_EINVAL, "fixed_base_mul: NULL argument");
  B2Z_REQUIRE(n < (1ull << 28), B2Z_ESIZE, "fixed_base_mul: more than 2^28 scalars in one call");
  if (n == 0) return;
  cudaStream_t st = c.stream;
  DevBuf<FrEl> ds(n);
  DevBuf<Affine> dp(n);
  DevBuf<uint32_t> dw((n + 31) / 32);
  B2Z_CUDA(cudaMemcpyAsync(ds.p, scalars, n * sizeof(FrEl), cudaMemcpyHostToDevice, st));
  if (std::is_same<C, G1>::value)
    g1_fixed_base_mul_device(&c, ds.p, reinterpret_cast<G1::Affine*>(dp.p), dw.p, (uint32_t)n, st);
  else
    g2_fixed_base_mul_device(&c, ds.p, reinterpret_cast<G2::Affine*>(dp.p), dw.p, (uint32_t)n, st);
  std::vector<uint32_t> words((n + 31) / 32);
  B2Z_CUDA(cudaMemcpyAsync(out_points, dp.p, n * sizeof(Affine), cudaMemcpyDeviceToHost, st));
  B2Z_CUDA(cudaMemcpyAsync(words.data(), dw.p, words.size() * 4, cudaMemcpyDeviceToHost, st));
  B2Z_CUDA(cudaStreamSynchronize(st));
  std::memcpy(out_inf, words.data(), (n + 7) / 8);
}

void prove_device(Ctx& c, PkImpl& pk, FrEl* d_a, FrEl* d_b, FrEl* d_c, const FrEl* d_z, const uint64_t r[4],
                  const uint64_t s[4], uint8_t proof_out[192]) {
  cudaStream_t st = c.stream;
  const uint32_t m = (uint32_t)pk.m, l = (uint32_t)pk.l;
  // host-side scalars: r, s, r*s as canonical integers
  const FrEl r_m = Fr::reduce(fr_load_host(r)), s_m = Fr::reduce(fr_load_host(s));
  ScalarPrep sp;
  sp.z = d_z;
  sp.out_a = pk.scal_a.p;
  sp.out_c = pk.scal_c.p;
  sp.m = m;
  sp.l = l;
  sp.r = Fr::from_mont(r_m);
  sp.s = Fr::from_mont(s_m);
  sp.rs = Fr::from_mont(Fr::reduce(Fr::mul(r_m, s_m)));
  FrEl one_c = Fr::zero();
  one_c.l[0] = 1;
  const FrEl tail_h[2] = {one_c, sp.s};                    // B: {beta_2: 1, delta_2: s}
  B2Z_CUDA(cudaMemcpyAsync(pk.tail.p, tail_h, sizeof(tail_h), cudaMemcpyHostToDevice, c.aux[0]));
  scalar_prep_kernel<<<nblk(pk.n1, 256), 256, 0, c.aux[0]>>>(sp);
  B2Z_LAUNCHED(&c);
  B2Z_CUDA(cudaEventRecord(pk.ev_z, c.aux[0]));
  G1::Xyzz* g1o = pk.g1_out.p;
  // A (aux0), B (aux1), C_z (aux2)
  msm_run<G1>(&c, 1, pk.g1_all, pk.scal_a.p, pk.n1, nullptr, g1o + 0, c.aux[0]);
  B2Z_CUDA(cudaEventRecord(pk.ev_done[0], c.aux[0]));
  B2Z_CUDA(cudaStreamWaitEvent(c.aux[1], pk.ev_z, 0));
  msm_run<G2>(&c, 2, pk.g2, pk.scal_a.p, m, pk.tail.p, pk.g2_out.p, c.aux[1]);
  B2Z_CUDA(cudaEventRecord(pk.ev_done[1], c.aux[1]));
  B2Z_CUDA(cudaStreamWaitEvent(c.aux[2], pk.ev_z, 0));
  msm_run<G1>(&c, 3, pk.g1_all, pk.scal_c.p, pk.n1, nullptr, g1o + 1, c.aux[2]);
  B2Z_CUDA(cudaEventRecord(pk.ev_done[2], c.aux[2]));
  // witness map + C_h on the main stream
  witness_map_device(&c, d_a, d_b, d_c, pk.log_n, /*natural_out=*/false, st);
  fr_from_mont_device(&c, d_a, pk.hc.p, (size_t)1 << pk.log_n, st);
  msm_run<G1>(&c, 0, pk.h, pk.hc.p, 1u << pk.log_n, nullptr, g1o + 2, st);
  for (auto& e : pk.ev_done) B2Z_CUDA(cudaStreamWaitEvent(st, e, 0));
  constexpr size_t kG1 = sizeof(G1::Xyzz), kG2 = sizeof(G2::Xyzz);
  B2Z_CUDA(cudaMemcpyAsync(pk.h_out, g1o, 3 * kG1, cudaMemcpyDeviceToHost, st));
  B2Z_CUDA(cudaMemcpyAsync(pk.h_out + 3 * kG1 / 4, pk.g2_out.p, kG2, cudaMemcpyDeviceToHost, st));
  B2Z_CUDA(cudaStreamSynchronize(st));
  // host epilogue: C = C_z + C_h, three normalisations, serialization
  const host::G1Xyzz A = host::g1_from_device(pk.h_out);
  const host::G1Xyzz C = host::g1_add(host::g1_from_device(pk.h_out + kG1 / 4), host::g1_from_device(pk.h_out + 2 * kG1 / 4));
  const host::G2Xyzz B = host::g2_from_device(pk.h_out + 3 * kG1 / 4);
  host::g1_serialize(proof_out, A);
  host::g2_serialize(proof_out + 48, B);
  host::g1_serialize(proof_out + 144, C);
}

}  // namespace
}  // namespace b2z

using namespace b2z;

struct b2z_pk {
  PkImpl impl;
};

extern "C" {

b2z_status b2z_msm_g1(b2z_ctx* ctx, const uint64_t* bases, const uint8_t* inf_bitmap, const uint64_t* scalars,
                      uint64_t n, uint64_t out_xyz[18]) {
  return guarded(ctx, [&](Ctx& c) { msm_entry<G1>(c, bases, inf_bitmap, scalars, n, out_xyz); });
}

b2z_status b2z_msm_g2(b2z_ctx* ctx, const uint64_t* bases, const uint8_t* inf_bitmap, const uint64_t* scalars,
                      uint64_t n, uint64_t out_xyz[36]) {
  return guarded(ctx, [&](Ctx& c) { msm_entry<G2>(c, bases, inf_bitmap, scalars, n, out_xyz); });
}

b2z_status b2z_fixed_base_mul_g1(b2z_ctx* ctx, const uint64_t* scalars, uint64_t n, uint64_t* out_points,
                                 uint8_t* out_inf) {
  return guarded(ctx, [&](Ctx& c) { fixed_base_entry<G1>(c, scalars, n, out_points, out_inf); });
}

b2z_status b2z_fixed_base_mul_g2(b2z_ctx* ctx, const uint64_t* scalars, uint64_t n, uint64_t* out_points,
                                 uint8_t* out_inf) {
  return guarded(ctx, [&](Ctx& c) { fixed_base_entry<G2>(c, scalars, n, out_points, out_inf); });
}

b2z_status b2z_pk_upload(b2z_ctx* ctx, const b2z_pk_desc* d, b2z_pk** out) {
  if (out) *out = nullptr;
  return guarded(ctx, [&](Ctx& c) {
    B2Z_REQUIRE(d != nullptr && out != nullptr, B2Z_EINVAL, "b2z_pk_upload: NULL argument");
    B2Z_REQUIRE(d->log_domain <= 32, B2Z_ESIZE, "b2z_pk_upload: domain larger than 2^32");
    B2Z_REQUIRE(d->log_domain <= 26, B2Z_ENOMEM, "b2z_pk_upload: domain does not fit this build's single-GPU plan");
    B2Z_REQUIRE(d->num_instance >= 1 && d->num_variables >= d->num_instance, B2Z_EINVAL,
                "b2z_pk_upload: need 1 <= num_instance <= num_variables");
    const uint64_t m = d->num_variables, l = d->num_instance, n = 1ull << d->log_domain;
    B2Z_REQUIRE(m < (1ull << 26), B2Z_ESIZE, "b2z_pk_upload: too many variables");
    B2Z_REQUIRE(d->a_query && d->b_g1_query && d->b_g2_query && (n == 1 || d->h_query) && (m == l || d->l_query) &&
                    d->alpha_g1 && d->beta_g1 && d->delta_g1 && d->beta_g2 && d->delta_g2,
                B2Z_EINVAL, "b2z_pk_upload: NULL query array");
    std::unique_ptr<b2z_pk> pk(new b2z_pk());
    PkImpl& P = pk->impl;
    P.log_n = d->log_domain;
    P.m = m;
    P.l = l;
    P.n1 = (uint32_t)(3 * m - l + 3);
    cudaStream_t st = c.stream;
    size_t free_b = 0, total_b = 0;
    B2Z_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const bool pre = pk_precompute_bytes(d) < free_b / 2;
    {
      // G1: [a | b1 | l | alpha beta delta]
      DevBuf<G1::Affine> dpts(P.n1);
      std::vector<uint32_t> words((P.n1 + 31) / 32, 0u);
      stage_points<G1::Affine>(dpts.p, words, 0, d->a_query, d->a_inf, m, st);
      stage_points<G1::Affine>(dpts.p, words, m, d->b_g1_query, d->b_g1_inf, m, st);
      stage_points<G1::Affine>(dpts.p, words, 2 * m, d->l_query, d->l_inf, m - l, st);
      stage_points<G1::Affine>(dpts.p, words, 3 * m - l, d->alpha_g1, nullptr, 1, st);
      stage_points<G1::Affine>(dpts.p, words, 3 * m - l + 1, d->beta_g1, nullptr, 1, st);
      stage_points<G1::Affine>(dpts.p, words, 3 * m - l + 2, d->delta_g1, nullptr, 1, st);
      DevBuf<uint32_t> dinf(words.size());
      B2Z_CUDA(cudaMemcpyAsync(dinf.p, words.data(), words.size() * 4, cudaMemcpyHostToDevice, st));
      msm_bases_build<G1>(&c, P.g1_all, dpts.p, dinf.p, P.n1, pre, 0, st);
      B2Z_CUDA(cudaStreamSynchronize(st));
    }
    {
      // G2: [b2 | beta_2 delta_2]
      const uint32_t n2 = (uint32_t)(m + 2);
      DevBuf<G2::Affine> dpts(n2);
      std::vector<uint32_t> words((n2 + 31) / 32, 0u);
      stage_points<G2::Affine>(dpts.p, words, 0, d->b_g2_query, d->b_g2_inf, m, st);
      stage_points<G2::Affine>(dpts.p, words, m, d->beta_g2, nullptr, 1, st);
      stage_points<G2::Affine>(dpts.p, words, m + 1, d->delta_g2, nullptr, 1, st);
      DevBuf<uint32_t> dinf(words.size());
      B2Z_CUDA(cudaMemcpyAsync(dinf.p, words.data(), words.size() * 4, cudaMemcpyHostToDevice, st));
      msm_bases_build<G2>(&c, P.g2, dpts.p, dinf.p, n2, pre, 0, st);
      B2Z_CUDA(cudaStreamSynchronize(st));
    }
    {
      // h_query in bit-reversed order, padded to n with a flagged identity
      DevBuf<G1::Affine> src(n), dst(n);
      DevBuf<uint32_t> src_inf, flags(n), words((n + 31) / 32);
      if (n > 1) B2Z_CUDA(cudaMemcpyAsync(src.p, d->h_query, (n - 1) * sizeof(G1::Affine), cudaMemcpyHostToDevice, st));
      std::vector<uint32_t> hw;
      if (d->h_inf != nullptr && n > 1) {
        hw.assign((n + 31) / 32, 0u);
        std::memcpy(hw.data(), d->h_inf, (n - 1 + 7) / 8);
        src_inf.alloc(hw.size());
        B2Z_CUDA(cudaMemcpyAsync(src_inf.p, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice, st));
      }
      permute_bitrev_g1_kernel<<<nblk(n, 256), 256, 0, st>>>(src.p, src_inf.p, (uint32_t)(n - 1), P.log_n, dst.p,
                                                             flags.p);
      B2Z_LAUNCHED(&c);
      pack_flags_kernel2<<<nblk((n + 31) / 32, 128), 128, 0, st>>>(flags.p, (uint32_t)n, words.p);
      B2Z_LAUNCHED(&c);
      msm_bases_build<G1>(&c, P.h, dst.p, words.p, (uint32_t)n, pre, 0, st);
      B2Z_CUDA(cudaStreamSynchronize(st));
    }
    P.ea.alloc(n); P.eb.alloc(n); P.ec.alloc(n); P.hc.alloc(n);
    P.z.alloc(m); P.scal_a.alloc(P.n1); P.scal_c.alloc(P.n1); P.tail.alloc(2);
    P.g1_out.alloc(3); P.g2_out.alloc(1);
    B2Z_CUDA(cudaMallocHost(&P.h_out, 3 * sizeof(G1::Xyzz) + sizeof(G2::Xyzz)));
    B2Z_CUDA(cudaEventCreateWithFlags(&P.ev_z, cudaEventDisableTiming));
    for (auto& e : P.ev_done) B2Z_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *out = pk.release();
  });
}

void b2z_pk_free(b2z_ctx* ctx, b2z_pk* pk) {
  if (pk == nullptr) return;
  if (ctx != nullptr) {
    std::lock_guard<std::mutex> lock(ctx->impl.mu);
    cudaSetDevice(ctx->impl.device);
    cudaDeviceSynchronize();
    delete pk;
  } else {
    delete pk;
  }
}

b2z_status b2z_groth16_prove(b2z_ctx* ctx, const b2z_pk* pk_c, const uint64_t* a_evals, const uint64_t* b_evals,
                             const uint64_t* c_evals, const uint64_t* z, const uint64_t r[4], const uint64_t s[4],
                             uint8_t proof_out[192]) {
  return guarded(ctx, [&](Ctx& c) {
    B2Z_REQUIRE(pk_c && a_evals && b_evals && c_evals && z && r && s && proof_out, B2Z_EINVAL,
                "b2z_groth16_prove: NULL argument");
    PkImpl& P = const_cast<b2z_pk*>(pk_c)->impl;
    const size_t n = (size_t)1 << P.log_n;
    B2Z_CUDA(cudaMemcpyAsync(P.z.p, z, P.m * sizeof(FrEl), cudaMemcpyHostToDevice, c.aux[0]));
    B2Z_CUDA(cudaMemcpyAsync(P.ea.p, a_evals, n * sizeof(FrEl), cudaMemcpyHostToDevice, c.stream));
    B2Z_CUDA(cudaMemcpyAsync(P.eb.p, b_evals, n * sizeof(FrEl), cudaMemcpyHostToDevice, c.stream));
    B2Z_CUDA(cudaMemcpyAsync(P.ec.p, c_evals, n * sizeof(FrEl), cudaMemcpyHostToDevice, c.stream));
    prove_device(c, P, P.ea.p, P.eb.p, P.ec.p, P.z.p, r, s, proof_out);
  });
}

b2z_status b2z_groth16_prove_device(b2z_ctx* ctx, const b2z_pk* pk_c, uint64_t* d_a, uint64_t* d_b, uint64_t* d_c,
                                    const uint64_t* d_z, const uint64_t r[4], const uint64_t s[4],
                                    uint8_t proof_out[192]) {
  return guarded(ctx, [&](Ctx& c) {
    B2Z_REQUIRE(pk_c && d_a && d_b && d_c && d_z && r && s && proof_out, B2Z_EINVAL,
                "b2z_groth16_prove_device: NULL argument");
    PkImpl& P = const_cast<b2z_pk*>(pk_c)->impl;
    // the caller's buffers were produced on some other stream: order after everything issued so far
    B2Z_CUDA(cudaDeviceSynchronize());
    prove_device(c, P, reinterpret_cast<FrEl*>(d_a), reinterpret_cast<FrEl*>(d_b), reinterpret_cast<FrEl*>(d_c),
                 reinterpret_cast<const FrEl*>(d_z), r, s, proof_out);
  });
}

}  // extern "C"
