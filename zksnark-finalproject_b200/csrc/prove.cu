// Groth16 prover assembly on the device -- replaces ark-groth16 0.4's
// create_proof_with_reduction / create_proof_with_assignment minus circuit
// synthesis (reference call sites: src/arkworks/backend/fibbonaci_handler.rs:110,
// matrix_proof.rs:139-140, prime_snark.rs:119; equations: SURVEY.md A.4):
//
//   A  = alpha_1 + sum z_i a_i + r delta_1
//   B  = beta_2  + sum z_i b2_i + s delta_2        (B1 likewise in G1)
//   C  = s A + r B1 - r s delta_1 + sum_{i>=l} z_i l_i + sum h_i hq_i
//
// Differences from the CPU formulation, none of which change the result:
//   * alpha/beta/delta terms ride inside the MSMs as two extra (point, scalar)
//     pairs {vk_param: 1, delta: r or s}; query[0] is covered by z_0 = 1;
//   * h stays in the bit-reversed order the DIF witness map leaves it in; the
//     h_query bases were permuted once at upload (an MSM is order-agnostic);
//   * the five MSMs run on separate streams; the four that only need z overlap
//     the witness map, and s*A / r*B1 overlap the h MSM.
#include <cstring>

#include "api_glue.hpp"
#include "msm.hpp"

namespace b2z {

namespace {

template <class T>
__device__ __forceinline__ T ld_any(const T* p) { return *p; }

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

// out = k * in  (one thread; 255-bit double-and-add)
template <class C>
__global__ void scalar_mul_kernel(const typename C::Xyzz* in, FrEl k, typename C::Xyzz* out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  *out = C::mul_scalar(*in, k.l);
}

struct FinalizeArgs {
  const G1::Xyzz* A;       // alpha + sum z a + r delta
  const G1::Xyzz* sA;      // s * A
  const G1::Xyzz* rB1;     // r * B1
  const G1::Xyzz* L;
  const G1::Xyzz* H;
  const G1::Xyzz* rsD;     // (r s) * delta_1
  const G2::Xyzz* B2;
  uint8_t* out;            // 192 bytes (device)
};

__device__ void write_fq_be(uint8_t* dst, const FqEl& mont) {
  const FqEl c = Fq::from_mont(mont);
  for (int i = 0; i < 12; i++) {
    const uint32_t w = c.l[11 - i];
    dst[4 * i + 0] = (uint8_t)(w >> 24);
    dst[4 * i + 1] = (uint8_t)(w >> 16);
    dst[4 * i + 2] = (uint8_t)(w >> 8);
    dst[4 * i + 3] = (uint8_t)w;
  }
}

// canonical (non-Montgomery) y > (q-1)/2 ?
__device__ bool fq_is_larger(const FqEl& mont) {
  const FqEl c = Fq::from_mont(mont);
  // 2c > q  <=>  c > (q-1)/2 ; compare 2c with q limb-wise from the top (2c < 2^384)
  uint32_t carry = 0;
  uint32_t d[12];
  for (int i = 0; i < 12; i++) {
    d[i] = (c.l[i] << 1) | carry;
    carry = c.l[i] >> 31;
  }
  for (int i = 11; i >= 0; i--) {
    const uint32_t q = FqParams::p(i);
    if (d[i] != q) return d[i] > q;
  }
  return false;
}

__device__ bool fq_canon_is_zero(const FqEl& mont) { return Fq::is_zero(mont); }

// ark-bls12-381 0.4 serialize_compressed (zcash encoding)
__device__ void serialize_g1(uint8_t* dst, const G1::Xyzz& p) {
  bool inf;
  const G1::Affine a = G1::to_affine(p, &inf);
  if (inf) {
    for (int i = 0; i < 48; i++) dst[i] = 0;
    dst[0] = 0xC0;
    return;
  }
  write_fq_be(dst, a.x);
  dst[0] |= 0x80;
  if (fq_is_larger(a.y)) dst[0] |= 0x20;
}

__device__ void serialize_g2(uint8_t* dst, const G2::Xyzz& p) {
  bool inf;
  const G2::Affine a = G2::to_affine(p, &inf);
  if (inf) {
    for (int i = 0; i < 96; i++) dst[i] = 0;
    dst[0] = 0xC0;
    return;
  }
  write_fq_be(dst, a.x.c1);
  write_fq_be(dst + 48, a.x.c0);
  dst[0] |= 0x80;
  const bool larger = fq_canon_is_zero(a.y.c1) ? fq_is_larger(a.y.c0) : fq_is_larger(a.y.c1);
  if (larger) dst[0] |= 0x20;
}

// threads 0..2 of one block: 0 -> A, 1 -> B, 2 -> C
__global__ void groth16_finalize_kernel(FinalizeArgs a) {
  if (blockIdx.x != 0) return;
  if (threadIdx.x == 0) {
    serialize_g1(a.out, *a.A);
  } else if (threadIdx.x == 1) {
    serialize_g2(a.out + 48, *a.B2);
  } else if (threadIdx.x == 2) {
    G1::Xyzz c = G1::add(*a.sA, *a.rB1);
    c = G1::add(c, G1::neg(*a.rsD));
    c = G1::add(c, *a.L);
    c = G1::add(c, *a.H);
    serialize_g1(a.out + 144, c);
  }
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
  MsmBases<G1> a, b1, lq, h;
  MsmBases<G2> b2;
  G1::Affine delta_g1;
  // per-proof device scratch
  DevBuf<FrEl> ea, eb, ec, z, zc, hc, tail;
  DevBuf<G1::Xyzz> g1_out;   // A, B1, L, H, sA, rB1, rs*delta, delta
  DevBuf<G2::Xyzz> g2_out;   // B2
  DevBuf<uint8_t> proof;
  uint8_t* h_proof = nullptr;   // pinned
  cudaEvent_t ev_z = nullptr, ev_done[4] = {nullptr, nullptr, nullptr, nullptr};
  ~PkImpl() {
    if (h_proof) cudaFreeHost(h_proof);
    if (ev_z) cudaEventDestroy(ev_z);
    for (auto& e : ev_done)
      if (e) cudaEventDestroy(e);
  }
};

template <class C>
static void upload_query(Ctx* ctx, MsmBases<C>& out, const uint64_t* pts, const uint8_t* inf, uint64_t count,
                         const uint64_t* extra0, const uint64_t* extra1, bool precompute, cudaStream_t st) {
  using Affine = typename C::Affine;
  const uint64_t total = count + (extra0 ? 1 : 0) + (extra1 ? 1 : 0);
  DevBuf<Affine> d(total ? total : 1);
  if (count) B2Z_CUDA(cudaMemcpyAsync(d.p, pts, count * sizeof(Affine), cudaMemcpyHostToDevice, st));
  uint64_t k = count;
  if (extra0) B2Z_CUDA(cudaMemcpyAsync(d.p + k++, extra0, sizeof(Affine), cudaMemcpyHostToDevice, st));
  if (extra1) B2Z_CUDA(cudaMemcpyAsync(d.p + k++, extra1, sizeof(Affine), cudaMemcpyHostToDevice, st));
  DevBuf<uint32_t> dinf;
  std::vector<uint32_t> words;
  if (inf != nullptr && count) {
    words.assign((total + 31) / 32, 0u);
    std::memcpy(words.data(), inf, (count + 7) / 8);
    // clear bits past `count` in the last copied byte
    for (uint64_t i = count; i < (((count + 7) / 8) * 8); i++) words[i >> 5] &= ~(1u << (i & 31));
    dinf.alloc(words.size());
    B2Z_CUDA(cudaMemcpyAsync(dinf.p, words.data(), words.size() * 4, cudaMemcpyHostToDevice, st));
  }
  msm_bases_build<C>(ctx, out, d.p, dinf.p, (uint32_t)total, precompute, 0, st);
  B2Z_CUDA(cudaStreamSynchronize(st));
}

static size_t pk_precompute_bytes(const b2z_pk_desc* d) {
  auto cost = [](uint64_t n, size_t pt) {
    if (n == 0) return (size_t)0;
    const uint32_t c = msm_pick_c(n, true);
    return (size_t)n * pt * msm_windows(c);
  };
  const uint64_t n = 1ull << d->log_domain;
  return cost(d->num_variables + 2, 96) * 2 + cost(d->num_variables + 2, 192) + cost(n, 96) +
         cost(d->num_variables - d->num_instance, 96);
}

}  // namespace b2z

using namespace b2z;

struct b2z_pk {
  PkImpl impl;
};

namespace {

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
  std::memset(out_xyz, 0, 3 * kLimbs * 8);
  // identity as arkworks prints it: (1, 1, 0) in Montgomery form
  FqEl one = Fq::one();
  std::memcpy(out_xyz, one.l, 48);
  std::memcpy(out_xyz + kLimbs, one.l, 48);
  if (n == 0) return;
  MsmBases<C> B;
  {
    DevBuf<Affine> d(n);
    B2Z_CUDA(cudaMemcpyAsync(d.p, bases, n * sizeof(Affine), cudaMemcpyHostToDevice, st));
    DevBuf<uint32_t> dinf;
    std::vector<uint32_t> words;
    if (inf_bitmap != nullptr) {
      words.assign((n + 31) / 32, 0u);
      std::memcpy(words.data(), inf_bitmap, (n + 7) / 8);
      for (uint64_t i = n; i < (((n + 7) / 8) * 8); i++) words[i >> 5] &= ~(1u << (i & 31));
      dinf.alloc(words.size());
      B2Z_CUDA(cudaMemcpyAsync(dinf.p, words.data(), words.size() * 4, cudaMemcpyHostToDevice, st));
    }
    msm_bases_build<C>(&c, B, d.p, dinf.p, (uint32_t)n, /*precompute=*/false, 0, st);
    B2Z_CUDA(cudaStreamSynchronize(st));
  }
  DevBuf<FrEl> ds(n);
  B2Z_CUDA(cudaMemcpyAsync(ds.p, scalars, n * sizeof(FrEl), cudaMemcpyHostToDevice, st));
  DevBuf<Xyzz> res(1);
  DevBuf<Affine> aff(1);
  DevBuf<uint32_t> flag(1);
  msm_run<C>(&c, 0, B, ds.p, (uint32_t)n, nullptr, res.p, st);
  xyzz_to_affine_device<C>(&c, res.p, aff.p, flag.p, 1, st);
  Affine h_aff;
  uint32_t h_flag = 0;
  B2Z_CUDA(cudaMemcpyAsync(&h_aff, aff.p, sizeof(Affine), cudaMemcpyDeviceToHost, st));
  B2Z_CUDA(cudaMemcpyAsync(&h_flag, flag.p, 4, cudaMemcpyDeviceToHost, st));
  B2Z_CUDA(cudaStreamSynchronize(st));
  if (!h_flag) {
    std::memcpy(out_xyz, &h_aff, sizeof(Affine));                 // X, Y
    std::memset(out_xyz + 2 * kLimbs, 0, kLimbs * 8);
    std::memcpy(out_xyz + 2 * kLimbs, one.l, 48);                 // Z = 1
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
  const uint64_t m = pk.m, l = pk.l;
  // host-side scalars: r, s canonical, r*s
  const FrEl r_m = Fr::reduce(fr_load_host(r)), s_m = Fr::reduce(fr_load_host(s));
  const FrEl r_c = Fr::from_mont(r_m), s_c = Fr::from_mont(s_m);
  const FrEl rs_c = Fr::from_mont(Fr::reduce(Fr::mul(r_m, s_m)));
  FrEl one_c = Fr::zero();
  one_c.l[0] = 1;
  FrEl tail_h[4] = {one_c, r_c, one_c, s_c};   // {1, r} for A ; {1, s} for B1, B2
  B2Z_CUDA(cudaMemcpyAsync(pk.tail.p, tail_h, sizeof(tail_h), cudaMemcpyHostToDevice, c.aux[0]));
  // z -> canonical
  fr_from_mont_device(&c, d_z, pk.zc.p, m, c.aux[0]);
  B2Z_CUDA(cudaEventRecord(pk.ev_z, c.aux[0]));
  G1::Xyzz* g1o = pk.g1_out.p;
  // A (aux0), B1 (aux1), B2 (aux2), L (aux3)
  msm_run<G1>(&c, 1, pk.a, pk.zc.p, (uint32_t)m, pk.tail.p, g1o + 0, c.aux[0]);
  {
    ProfileScope ps(&c, PH_FINALIZE, c.aux[0], 1);
    scalar_mul_kernel<G1><<<1, 32, 0, c.aux[0]>>>(g1o + 0, s_c, g1o + 4);
    B2Z_LAUNCHED(&c);
  }
  B2Z_CUDA(cudaEventRecord(pk.ev_done[0], c.aux[0]));
  B2Z_CUDA(cudaStreamWaitEvent(c.aux[1], pk.ev_z, 0));
  msm_run<G1>(&c, 2, pk.b1, pk.zc.p, (uint32_t)m, pk.tail.p + 2, g1o + 1, c.aux[1]);
  {
    ProfileScope ps(&c, PH_FINALIZE, c.aux[1], 1);
    scalar_mul_kernel<G1><<<1, 32, 0, c.aux[1]>>>(g1o + 1, r_c, g1o + 5);
    B2Z_LAUNCHED(&c);
  }
  B2Z_CUDA(cudaEventRecord(pk.ev_done[1], c.aux[1]));
  B2Z_CUDA(cudaStreamWaitEvent(c.aux[2], pk.ev_z, 0));
  msm_run<G2>(&c, 3, pk.b2, pk.zc.p, (uint32_t)m, pk.tail.p + 2, pk.g2_out.p, c.aux[2]);
  B2Z_CUDA(cudaEventRecord(pk.ev_done[2], c.aux[2]));
  B2Z_CUDA(cudaStreamWaitEvent(c.aux[3], pk.ev_z, 0));
  scalar_mul_kernel<G1><<<1, 32, 0, c.aux[3]>>>(g1o + 7, rs_c, g1o + 6);
  B2Z_LAUNCHED(&c);
  msm_run<G1>(&c, 4, pk.lq, pk.zc.p + l, (uint32_t)(m - l), nullptr, g1o + 2, c.aux[3]);
  B2Z_CUDA(cudaEventRecord(pk.ev_done[3], c.aux[3]));
  // witness map + H on the main stream
  witness_map_device(&c, d_a, d_b, d_c, pk.log_n, /*natural_out=*/false, st);
  fr_from_mont_device(&c, d_a, pk.hc.p, (size_t)1 << pk.log_n, st);
  msm_run<G1>(&c, 0, pk.h, pk.hc.p, 1u << pk.log_n, nullptr, g1o + 3, st);
  for (auto& e : pk.ev_done) B2Z_CUDA(cudaStreamWaitEvent(st, e, 0));
  FinalizeArgs fa;
  fa.A = g1o + 0; fa.sA = g1o + 4; fa.rB1 = g1o + 5; fa.L = g1o + 2; fa.H = g1o + 3;
  fa.rsD = g1o + 6;
  fa.B2 = pk.g2_out.p;
  fa.out = pk.proof.p;
  {
    ProfileScope ps(&c, PH_FINALIZE, st, 1);
    groth16_finalize_kernel<<<1, 32, 0, st>>>(fa);
    B2Z_LAUNCHED(&c);
  }
  B2Z_CUDA(cudaMemcpyAsync(pk.h_proof, pk.proof.p, 192, cudaMemcpyDeviceToHost, st));
  B2Z_CUDA(cudaStreamSynchronize(st));
  std::memcpy(proof_out, pk.h_proof, 192);
}

}  // namespace

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
    B2Z_REQUIRE(m < (1ull << 27), B2Z_ESIZE, "b2z_pk_upload: too many variables");
    B2Z_REQUIRE(d->a_query && d->b_g1_query && d->b_g2_query && (n == 1 || d->h_query) && (m == l || d->l_query) &&
                    d->alpha_g1 && d->beta_g1 && d->delta_g1 && d->beta_g2 && d->delta_g2,
                B2Z_EINVAL, "b2z_pk_upload: NULL query array");
    std::unique_ptr<b2z_pk> pk(new b2z_pk());
    PkImpl& P = pk->impl;
    P.log_n = d->log_domain;
    P.m = m;
    P.l = l;
    cudaStream_t st = c.stream;
    size_t free_b = 0, total_b = 0;
    B2Z_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const bool pre = pk_precompute_bytes(d) < free_b / 2;
    upload_query<G1>(&c, P.a, d->a_query, d->a_inf, m, d->alpha_g1, d->delta_g1, pre, st);
    upload_query<G1>(&c, P.b1, d->b_g1_query, d->b_g1_inf, m, d->beta_g1, d->delta_g1, pre, st);
    upload_query<G2>(&c, P.b2, d->b_g2_query, d->b_g2_inf, m, d->beta_g2, d->delta_g2, pre, st);
    upload_query<G1>(&c, P.lq, d->l_query, d->l_inf, m - l, nullptr, nullptr, pre, st);
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
      pack_flags_kernel2<<<nblk((n + 31) / 32, 128), 128, 0, st>>>(flags.p, (uint32_t)n, words.p);
      c.launches += 2;
      B2Z_CUDA(cudaGetLastError());
      msm_bases_build<G1>(&c, P.h, dst.p, words.p, (uint32_t)n, pre, 0, st);
      B2Z_CUDA(cudaStreamSynchronize(st));
    }
    std::memcpy(&P.delta_g1, d->delta_g1, sizeof(G1::Affine));
    P.ea.alloc(n); P.eb.alloc(n); P.ec.alloc(n); P.hc.alloc(n);
    P.z.alloc(m); P.zc.alloc(m); P.tail.alloc(4);
    P.g1_out.alloc(8); P.g2_out.alloc(1); P.proof.alloc(192);
    {
      const G1::Xyzz dx = G1::from_affine(P.delta_g1);
      B2Z_CUDA(cudaMemcpy(P.g1_out.p + 7, &dx, sizeof(dx), cudaMemcpyHostToDevice));
    }
    B2Z_CUDA(cudaMallocHost(&P.h_proof, 192));
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
