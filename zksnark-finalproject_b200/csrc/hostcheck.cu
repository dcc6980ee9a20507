// Host-side self checks: the device limb algorithms (mont.cuh, ec.cuh, the MSM
// digit recoding) run here through the carry-flag emulation of ptx.cuh, so the
// CPU-only test suite can compare them with the oracle.  None of this is used
// by the proving path.
#include <cstring>

#include "host_fq.hpp"
#include "msm.hpp"

using namespace b2z;

namespace {

template <class F>
int field_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  typename F::El x, y, r;
  std::memcpy(x.l, a, sizeof(x.l));
  if (b) std::memcpy(y.l, b, sizeof(y.l));
  switch (op) {
    case 0: r = F::mul_safe(x, y); break;
    case 1: r = F::add(x, y); break;
    case 2: r = F::sub(x, y); break;
    case 3: r = F::reduce(x); break;
    case 4: r = F::reduce(F::inv(x)); break;
    case 5: r = F::mul(x, y); break;   // raw product: caller honours the operand precondition
    default: return -1;
  }
  std::memcpy(out, r.l, sizeof(r.l));
  return 0;
}

template <class C>
int point_sum(const uint32_t* points, const uint8_t* neg, uint32_t n, uint32_t* out_affine) {
  using Affine = typename C::Affine;
  typename C::Xyzz acc = C::identity();
  typename C::Xyzz acc2 = C::identity();
  for (uint32_t i = 0; i < n; i++) {
    Affine p;
    std::memcpy(&p, points + (size_t)i * (sizeof(Affine) / 4), sizeof(Affine));
    if (neg && neg[i]) p = C::neg(p);
    acc = C::madd(acc, p);                       // mixed addition path
    acc2 = C::add(acc2, C::from_affine(p));      // general addition path
  }
  bool inf = false, inf2 = false;
  const Affine r = C::to_affine(acc, &inf);
  const Affine r2 = C::to_affine(acc2, &inf2);
  if (inf != inf2 || std::memcmp(&r, &r2, sizeof(Affine)) != 0) return -1;   // the two paths must agree
  std::memcpy(out_affine, &r, sizeof(Affine));
  return inf ? 1 : 0;
}

}  // namespace

extern "C" {

int b2z_host_field_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  if (a == nullptr || out == nullptr) return -1;
  return field == 0 ? field_op<Fr>(op, a, b, out) : field_op<Fq>(op, a, b, out);
}

int b2z_host_point_sum(int group, const uint32_t* points, const uint8_t* neg, uint32_t n, uint32_t* out_affine) {
  if (out_affine == nullptr || (n && points == nullptr)) return -1;
  return group == 1 ? point_sum<G1>(points, neg, n, out_affine) : point_sum<G2>(points, neg, n, out_affine);
}

uint32_t b2z_host_msm_digits(const uint32_t scalar[8], uint32_t c, int32_t* digits) {
  DigitCfg cfg;
  cfg.c = c;
  cfg.windows = msm_windows(c);
  cfg.nb = 1u << (c - 1);
  cfg.n = 1;
  cfg.precomputed = 0;
  FrEl k;
  std::memcpy(k.l, scalar, 32);
  for (uint32_t w = 0; w < cfg.windows; w++) digits[w] = 0;
  for_each_digit(k, cfg, [&](uint32_t w, uint32_t v, bool neg) { digits[w] = neg ? -(int32_t)v : (int32_t)v; });
  return cfg.windows;
}

int b2z_host_planes_horner(int group, const uint32_t* planes_xyzz, uint32_t nplanes, uint32_t chunk_log, uint8_t* out) {
  if (out == nullptr || (nplanes && planes_xyzz == nullptr)) return -1;
  if (group == 1) host::g1_serialize(out, host::g1_planes_horner(planes_xyzz, nplanes, chunk_log));
  else host::g2_serialize(out, host::g2_planes_horner(planes_xyzz, nplanes, chunk_log));
  return 0;
}

uint32_t b2z_host_msm_window_bits(uint64_t n, int precomputed) { return msm_pick_c(n, precomputed != 0); }

}  // extern "C"
