// extern "C" entry points of the host-side verifier (include/b200zk.h, row f4): no GPU, no context.
#include <cstring>
#include <new>

#include "../../include/b200zk.h"
#include "host_pairing.hpp"

using namespace b2z::host;

namespace {

bool load_g1(const uint64_t* limbs, bool inf, G1Aff* out) {
  out->inf = inf;
  if (inf) { out->x = out->y = fq_zero(); return true; }
  std::memcpy(out->x.l, limbs, 48);
  std::memcpy(out->y.l, limbs + 6, 48);
  if (fq_geq_q(out->x.l) || fq_geq_q(out->y.l)) return false;
  // y^2 = x^3 + 4
  return fq_eq(fq_sqr(out->y), fq_add(fq_mul(fq_sqr(out->x), out->x), fq_from_u64(4)));
}
bool load_g2(const uint64_t* limbs, G2Aff* out) {
  out->inf = false;
  std::memcpy(out->x.c0.l, limbs, 48);
  std::memcpy(out->x.c1.l, limbs + 6, 48);
  std::memcpy(out->y.c0.l, limbs + 12, 48);
  std::memcpy(out->y.c1.l, limbs + 18, 48);
  for (const Fq* f : {&out->x.c0, &out->x.c1, &out->y.c0, &out->y.c1})
    if (fq_geq_q(f->l)) return false;
  const Fq four = fq_from_u64(4);
  return fq2_eq(fq2_sqr(out->y), fq2_add(fq2_mul(fq2_sqr(out->x), out->x), Fq2{four, four}));
}

}  // namespace

extern "C" {

b2z_status b2z_groth16_prepare_verifying_key(const b2z_vk_desc* vk, uint8_t* pvk_out, uint64_t capacity,
                                             uint64_t* pvk_len) {
  if (vk == nullptr || pvk_len == nullptr || vk->alpha_g1 == nullptr || vk->beta_g2 == nullptr ||
      vk->gamma_g2 == nullptr || vk->delta_g2 == nullptr || (vk->num_instance && vk->gamma_abc_g1 == nullptr))
    return B2Z_EINVAL;
  try {
    PreparedVk k;
    if (!load_g1(vk->alpha_g1, false, &k.alpha_g1) || !load_g2(vk->beta_g2, &k.beta_g2) ||
        !load_g2(vk->gamma_g2, &k.gamma_g2) || !load_g2(vk->delta_g2, &k.delta_g2))
      return B2Z_EINVAL;
    k.gamma_abc_g1.resize(vk->num_instance);
    for (uint64_t i = 0; i < vk->num_instance; i++) {
      const bool inf = vk->gamma_abc_inf != nullptr && ((vk->gamma_abc_inf[i >> 3] >> (i & 7)) & 1);
      if (!load_g1(vk->gamma_abc_g1 + 12 * i, inf, &k.gamma_abc_g1[i])) return B2Z_EINVAL;
    }
    if (!prepare_vk(&k)) return B2Z_EINVAL;
    const std::vector<uint8_t> bytes = pvk_serialize(k);
    *pvk_len = bytes.size();
    if (pvk_out == nullptr) return B2Z_OK;                 // size query
    if (capacity < bytes.size()) return B2Z_ESIZE;
    std::memcpy(pvk_out, bytes.data(), bytes.size());
    return B2Z_OK;
  } catch (const std::bad_alloc&) {
    return B2Z_ENOMEM;
  }
}

b2z_status b2z_groth16_verify_with_processed_vk(const uint8_t* pvk, uint64_t pvk_len, const uint64_t* public_inputs,
                                                uint64_t num_inputs, const uint8_t proof[192], int32_t* valid) {
  if (pvk == nullptr || proof == nullptr || valid == nullptr || (num_inputs && public_inputs == nullptr)) return B2Z_EINVAL;
  *valid = 0;
  try {
    PreparedVk k;
    if (!pvk_deserialize(pvk, pvk_len, &k)) return B2Z_EINVAL;
    // inputs arrive as Fr elements (Montgomery limbs, arkworks' in-memory form); the scalar multiplication
    // takes into_bigint()
    std::vector<uint64_t> canon(4 * num_inputs);
    for (uint64_t i = 0; i < num_inputs; i++)
      if (!fr_mont_to_canonical(public_inputs + 4 * i, &canon[4 * i])) return B2Z_EINVAL;
    bool ok = false;
    if (verify_with_processed_vk(k, canon.data(), num_inputs, proof, &ok) != 0) return B2Z_EINVAL;
    *valid = ok ? 1 : 0;
    return B2Z_OK;
  } catch (const std::bad_alloc&) {
    return B2Z_ENOMEM;
  }
}

}  // extern "C"
