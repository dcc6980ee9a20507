// Key-generation scalars on the host, multithreaded (include/b200zk.h, row f3): no GPU, no context.
//
// ark_groth16::generator::generate_parameters_with_qap (reached from Groth16::setup / circuit_specific_setup,
// /root/reference/src/arkworks/backend/matrix_proof.rs:128-131, fibbonaci_handler.rs:107, prime_snark.rs:111-115) spends
// its time in two places: O(n) scalar preparation -- the Lagrange coefficients of LibsnarkReduction::
// instance_map_with_evaluation, the powers of tau for the h query, the beta A_i + alpha B_i + C_i combinations -- and the
// fixed-base multiplications.  The second part is b2z_fixed_base_mul_g1/g2 and b2z_spmv_fr on the GPU; this file is the
// first part, which the host mirror used to do in Python integers (20 s for a 2^22 domain).  Element-wise vector
// operations over Fr, split over std::threads; every function is exact (canonical Montgomery limbs in and out).
#include <algorithm>
#include <new>
#include <system_error>
#include <thread>
#include <vector>

#include "../../include/b200zk.h"
#include "host_fr.hpp"

using namespace b2z::hostfr;

namespace {

uint32_t pick_threads(uint32_t want, uint64_t items, uint64_t min_per_thread) {
  if (want == 0) {
    want = std::thread::hardware_concurrency();
    if (want == 0) want = 4;
  }
  want = std::min<uint32_t>(want, 256);
  const uint64_t cap = std::max<uint64_t>(1, items / std::max<uint64_t>(1, min_per_thread));
  if ((uint64_t)want > cap) want = (uint32_t)cap;
  return std::max<uint32_t>(want, 1);
}

// fn(lo, hi) over [0, count) in nt contiguous chunks
template <class F>
void parallel_chunks(uint64_t count, uint32_t nt, F fn) {
  if (nt <= 1 || count == 0) {
    fn((uint64_t)0, count);
    return;
  }
  std::vector<std::thread> pool;
  for (uint32_t t = 1; t < nt; t++) pool.emplace_back(fn, count * t / nt, count * (t + 1) / nt);
  fn((uint64_t)0, count / nt);
  for (auto& th : pool) th.join();
}

// vector elements may arrive lazily reduced (any value below 2^256 < 3 r): bring them into [0, r)
inline Fr load_reduced(const Fr& v) {
  Fr o = v;
  while (geq_mod(o.l)) sub_mod(o.l);
  return o;
}

constexpr uint64_t kBatch = 1024;   // elements per shared inversion

}  // namespace

extern "C" {

b2z_status b2z_fr_lagrange_at(uint32_t log_n, const uint64_t tau[4], uint64_t count, uint32_t threads, uint64_t* out) {
  if (tau == nullptr || (count && out == nullptr) || log_n > 32 || !fr_is_canonical(tau)) return B2Z_EINVAL;
  const uint64_t n = 1ull << log_n;
  if (count > n) return B2Z_EINVAL;
  try {
    const Fr t = fr_load(tau);
    // w = root^(2^(32 - log_n)); Z(tau) = tau^n - 1
    Fr w = fr_two_adic_root();
    for (uint32_t i = log_n; i < 32; i++) w = fr_mul(w, w);
    Fr tn = t;
    for (uint32_t i = 0; i < log_n; i++) tn = fr_mul(tn, tn);
    const Fr zt = fr_sub(tn, fr_one());
    if (fr_is_zero(zt)) return B2Z_EINVAL;                       // tau lies in the domain
    const Fr zn = fr_mul(zt, fr_inv(fr_from_u64(n)));            // Z(tau) / n
    Fr* o = reinterpret_cast<Fr*>(out);
    const uint32_t nt = pick_threads(threads, count, 4096);
    parallel_chunks(count, nt, [&](uint64_t lo, uint64_t hi) {
      std::vector<Fr> wi(kBatch), pref(kBatch);
      Fr cur = fr_pow_u64(w, lo);
      for (uint64_t base = lo; base < hi; base += kBatch) {
        const uint64_t len = std::min<uint64_t>(kBatch, hi - base);
        // d_i = tau - w^i; prefix products; one inversion; L_i = zn * w^i / d_i
        Fr acc = fr_one();
        for (uint64_t k = 0; k < len; k++) {
          wi[k] = cur;
          pref[k] = acc;
          acc = fr_mul(acc, fr_sub(t, cur));                      // non-zero: tau^n != 1
          cur = fr_mul(cur, w);
        }
        Fr inv = fr_inv(acc);
        for (uint64_t k = len; k-- > 0;) {
          const Fr dinv = fr_mul(inv, pref[k]);
          inv = fr_mul(inv, fr_sub(t, wi[k]));
          o[base + k] = fr_mul(fr_mul(zn, wi[k]), dinv);
        }
      }
    });
    return B2Z_OK;
  } catch (const std::bad_alloc&) {
    return B2Z_ENOMEM;
  } catch (const std::system_error&) {
    return B2Z_ENOMEM;
  }
}

b2z_status b2z_fr_geometric(const uint64_t base[4], const uint64_t scale[4], uint64_t count, uint32_t threads,
                            uint64_t* out) {
  if (base == nullptr || scale == nullptr || (count && out == nullptr) || !fr_is_canonical(base) || !fr_is_canonical(scale))
    return B2Z_EINVAL;
  try {
    const Fr b = fr_load(base), s = fr_load(scale);
    Fr* o = reinterpret_cast<Fr*>(out);
    const uint32_t nt = pick_threads(threads, count, 4096);
    parallel_chunks(count, nt, [&](uint64_t lo, uint64_t hi) {
      Fr cur = fr_mul(s, fr_pow_u64(b, lo));
      for (uint64_t i = lo; i < hi; i++) {
        o[i] = cur;
        cur = fr_mul(cur, b);
      }
    });
    return B2Z_OK;
  } catch (const std::bad_alloc&) {
    return B2Z_ENOMEM;
  } catch (const std::system_error&) {
    return B2Z_ENOMEM;
  }
}

b2z_status b2z_fr_lincomb3(uint64_t count, const uint64_t a[4], const uint64_t* x, const uint64_t b[4], const uint64_t* y,
                           const uint64_t c[4], const uint64_t* z, uint32_t threads, uint64_t* out) {
  if ((count && out == nullptr) || (x && a == nullptr) || (y && b == nullptr) || (z && c == nullptr)) return B2Z_EINVAL;
  if ((x && !fr_is_canonical(a)) || (y && !fr_is_canonical(b)) || (z && !fr_is_canonical(c))) return B2Z_EINVAL;
  try {
    const Fr ca = x ? fr_load(a) : fr_zero(), cb = y ? fr_load(b) : fr_zero(), cc = z ? fr_load(c) : fr_zero();
    const Fr one = fr_one();
    const bool a1 = x && std::memcmp(ca.l, one.l, 32) == 0, b1 = y && std::memcmp(cb.l, one.l, 32) == 0,
               c1 = z && std::memcmp(cc.l, one.l, 32) == 0;
    const Fr* xs = reinterpret_cast<const Fr*>(x);
    const Fr* ys = reinterpret_cast<const Fr*>(y);
    const Fr* zs = reinterpret_cast<const Fr*>(z);
    Fr* o = reinterpret_cast<Fr*>(out);
    const uint32_t nt = pick_threads(threads, count, 8192);
    parallel_chunks(count, nt, [&](uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; i++) {
        Fr acc = fr_zero();
        if (x) {
          const Fr v = load_reduced(xs[i]);
          acc = a1 ? v : fr_mul(ca, v);
        }
        if (y) {
          const Fr v = load_reduced(ys[i]);
          acc = fr_add(acc, b1 ? v : fr_mul(cb, v));
        }
        if (z) {
          const Fr v = load_reduced(zs[i]);
          acc = fr_add(acc, c1 ? v : fr_mul(cc, v));
        }
        o[i] = acc;
      }
    });
    return B2Z_OK;
  } catch (const std::bad_alloc&) {
    return B2Z_ENOMEM;
  } catch (const std::system_error&) {
    return B2Z_ENOMEM;
  }
}

b2z_status b2z_fr_into_bigint(uint64_t count, const uint64_t* in, uint32_t threads, uint64_t* out) {
  if (count && (in == nullptr || out == nullptr)) return B2Z_EINVAL;
  try {
    const Fr* is = reinterpret_cast<const Fr*>(in);
    Fr* o = reinterpret_cast<Fr*>(out);
    const uint32_t nt = pick_threads(threads, count, 16384);
    parallel_chunks(count, nt, [&](uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; i++) o[i] = fr_into_bigint(load_reduced(is[i]));
    });
    return B2Z_OK;
  } catch (const std::bad_alloc&) {
    return B2Z_ENOMEM;
  } catch (const std::system_error&) {
    return B2Z_ENOMEM;
  }
}

}  // extern "C"
