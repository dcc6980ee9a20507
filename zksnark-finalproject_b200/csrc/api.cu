// extern "C" entry points of libb200zk (see include/b200zk.h for the contract).
// Every entry point catches StatusError so nothing unwinds across the boundary.
#include <cstring>

#include "api_glue.hpp"
#include "msm.hpp"

using namespace b2z;

extern "C" {

b2z_status b2z_ctx_create(int device_id, b2z_ctx** out) {
  if (out == nullptr) return B2Z_EINVAL;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return B2Z_ECUDA;   // no CPU fallback
  if (device_id < 0 || device_id >= count) return B2Z_EINVAL;
  b2z_ctx* ctx = new (std::nothrow) b2z_ctx();
  if (ctx == nullptr) return B2Z_ENOMEM;
  ctx->impl.device = device_id;
  try {
    B2Z_CUDA(cudaSetDevice(device_id));
    // Priorities.  The main stream carries the row evaluation, the witness map and the H sum -- the chain the
    // LAST accumulation waits for -- and gets the highest priority: its integer-bound transform CTAs then take
    // every SM slot a retiring sort CTA frees, so the (latency / atomics bound) sorts of the z-only MSMs on the
    // auxiliary streams run beside the transforms instead of in front of them (the sorts' thousands of CTAs
    // otherwise fill every thread slot first: 5 ms of a 2^22 proof with nothing else running).
    int prio_lo = 0, prio_hi = 0;
    B2Z_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    // The auxiliary streams carry the z-only MSMs; their accumulations are chained A (aux 0) -> B1 (aux 2) -> B
    // (aux 1) -> L (aux 3), and the sorts get the priorities that make them finish in that order.
    auto below = [&](int p) { return p < prio_lo ? p + 1 : p; };
    const int p1 = below(prio_hi), p2 = below(p1), p3 = below(p2);
    const int aux_prio[4] = {p1, p3, p2, p3};
    B2Z_CUDA(cudaStreamCreateWithPriority(&ctx->impl.stream, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < 4; i++)
      B2Z_CUDA(cudaStreamCreateWithPriority(&ctx->impl.aux[i], cudaStreamNonBlocking, aux_prio[i]));
  } catch (const StatusError& e) {
    delete ctx;
    return e.code;
  }
  *out = ctx;
  return B2Z_OK;
}

void b2z_ctx_destroy(b2z_ctx* ctx) {
  if (ctx == nullptr) return;
  cudaSetDevice(ctx->impl.device);
  cudaDeviceSynchronize();
  ctx->impl.domains.clear();
  msm_release_scratch(&ctx->impl);
  for (auto& sp : ctx->impl.spans) {
    cudaEventDestroy(sp.start);
    cudaEventDestroy(sp.stop);
  }
  ctx->impl.spans.clear();
  for (auto& e : ctx->impl.event_pool) cudaEventDestroy(e);
  ctx->impl.event_pool.clear();
  if (ctx->impl.units_pool) cudaFreeHost(ctx->impl.units_pool);
  if (ctx->impl.stream) cudaStreamDestroy(ctx->impl.stream);
  for (auto& s : ctx->impl.aux)
    if (s) cudaStreamDestroy(s);
  delete ctx;
}

b2z_status b2z_profile_enable(b2z_ctx* ctx, int on) {
  return guarded(ctx, [&](Ctx& c) {
    if (on && c.units_pool == nullptr) {
      // everything the spans need is created here, outside any timed region
      B2Z_CUDA(cudaMallocHost(&c.units_pool, kUnitsPool * sizeof(uint32_t)));
      for (int i = 0; i < 2048; i++) {
        cudaEvent_t e;
        B2Z_CUDA(cudaEventCreate(&e));
        c.event_pool.push_back(e);
      }
    }
    c.profile = on != 0;
  });
}

b2z_status b2z_profile_read(b2z_ctx* ctx, double* ms, uint64_t* launches, uint64_t* units, int reset) {
  return guarded(ctx, [&](Ctx& c) {
    B2Z_REQUIRE(ms && launches && units, B2Z_EINVAL, "b2z_profile_read: NULL argument");
    B2Z_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < PH_COUNT; i++) { ms[i] = 0; launches[i] = 0; units[i] = 0; }
    std::lock_guard<std::mutex> lock(c.span_mu);
    for (auto& sp : c.spans) {
      float t = 0;
      B2Z_CUDA(cudaEventElapsedTime(&t, sp.start, sp.stop));
      ms[sp.phase] += t;
      launches[sp.phase] += 1;
      units[sp.phase] += sp.units_pinned ? *sp.units_pinned : sp.units;
    }
    if (reset) {
      for (auto& sp : c.spans) {
        c.event_pool.push_back(sp.start);
        c.event_pool.push_back(sp.stop);
      }
      c.spans.clear();
      c.units_used = 0;
    }
  });
}

int b2z_profile_spans(b2z_ctx* ctx, int max_spans, int* phase, double* start_ms, double* stop_ms) {
  int n = 0;
  guarded(ctx, [&](Ctx& c) {
    B2Z_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lock(c.span_mu);
    if (c.spans.empty()) return;
    for (auto& sp : c.spans) {
      if (n >= max_spans) break;
      float t0 = 0, t1 = 0;
      B2Z_CUDA(cudaEventElapsedTime(&t0, c.spans[0].start, sp.start));
      B2Z_CUDA(cudaEventElapsedTime(&t1, c.spans[0].start, sp.stop));
      phase[n] = sp.phase;
      start_ms[n] = t0;
      stop_ms[n] = t1;
      n++;
    }
  });
  return n;
}

uint64_t b2z_kernel_launches(const b2z_ctx* ctx) { return ctx ? ctx->impl.launches : 0; }

b2z_status b2z_measure_int_peak(b2z_ctx* ctx, double* imad_per_s, double* imad_wide_per_s) {
  return guarded(ctx, [&](Ctx& c) {
    B2Z_REQUIRE(imad_per_s && imad_wide_per_s, B2Z_EINVAL, "b2z_measure_int_peak: NULL argument");
    measure_int_peak(&c, imad_per_s, imad_wide_per_s);
  });
}

b2z_status b2z_host_register(b2z_ctx* ctx, void* ptr, uint64_t bytes) {
  return guarded(ctx, [&](Ctx&) {
    B2Z_REQUIRE(ptr != nullptr && bytes > 0, B2Z_EINVAL, "b2z_host_register: empty range");
    B2Z_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  });
}

b2z_status b2z_host_unregister(b2z_ctx* ctx, void* ptr) {
  return guarded(ctx, [&](Ctx&) {
    B2Z_REQUIRE(ptr != nullptr, B2Z_EINVAL, "b2z_host_unregister: NULL");
    B2Z_CUDA(cudaHostUnregister(ptr));
  });
}

const char* b2z_last_error(const b2z_ctx* ctx) { return ctx ? ctx->impl.last_error.c_str() : "null context"; }

b2z_status b2z_ntt_fr(b2z_ctx* ctx, uint64_t* data, uint32_t log_n, int inverse, const uint64_t coset_gen[4]) {
  return guarded(ctx, [&](Ctx& c) {
    B2Z_REQUIRE(data != nullptr, B2Z_EINVAL, "b2z_ntt_fr: data is NULL");
    B2Z_REQUIRE(log_n <= 32, B2Z_ESIZE, "b2z_ntt_fr: domain larger than 2^32 (PolynomialDegreeTooLarge)");
    B2Z_REQUIRE(log_n <= 28, B2Z_ENOMEM, "b2z_ntt_fr: domain does not fit this build's single-GPU plan");
    const size_t n = (size_t)1 << log_n;
    cudaStream_t st = c.stream;
    DevBuf<FrEl> d(n);
    B2Z_CUDA(cudaMemcpyAsync(d.p, data, n * sizeof(FrEl), cudaMemcpyHostToDevice, st));
    DevBuf<FrEl> pw_lo, pw_hi;
    const bool coset = coset_gen != nullptr;
    if (coset) {
      // g^i (forward) or g^-i (inverse) as lo[i & 1023] * hi[i >> 10]
      FrEl g;
      std::memcpy(g.l, coset_gen, sizeof(g.l));
      g = Fr::reduce(g);
      if (inverse) g = Fr::reduce(Fr::inv(g));
      FrEl g1024 = g;
      for (int i = 0; i < 10; i++) g1024 = Fr::sqr(g1024);
      g1024 = Fr::reduce(g1024);
      const uint32_t nhi = (uint32_t)((n + 1023) >> 10);
      pw_lo.alloc(1024);
      pw_hi.alloc(nhi);
      fr_pow_table(&c, pw_lo.p, g, 1024, st);
      fr_pow_table(&c, pw_hi.p, g1024, nhi, st);
    }
    if (!inverse) {
      if (coset) ntt_scale_powers(&c, d.p, log_n, pw_lo.p, pw_hi.p, st);
      ntt_dif(&c, ntt_twiddles(&c, log_n, TW_FWD, st), d.p, log_n, st);
      ntt_bitrev(&c, d.p, log_n, nullptr, nullptr, nullptr, st);
    } else {
      ntt_dif(&c, ntt_twiddles(&c, log_n, TW_INV, st), d.p, log_n, st);
      const NttDomain& dom = ntt_domain(&c, log_n);
      ntt_bitrev(&c, d.p, log_n, &dom.n_inv, coset ? pw_lo.p : nullptr, coset ? pw_hi.p : nullptr, st);
    }
    B2Z_CUDA(cudaMemcpyAsync(data, d.p, n * sizeof(FrEl), cudaMemcpyDeviceToHost, st));
    B2Z_CUDA(cudaStreamSynchronize(st));
  });
}

b2z_status b2z_witness_map(b2z_ctx* ctx, const uint64_t* a, const uint64_t* b, const uint64_t* cc, uint32_t log_n,
                           uint64_t* h_out) {
  return guarded(ctx, [&](Ctx& c) {
    B2Z_REQUIRE(a && b && cc && h_out, B2Z_EINVAL, "b2z_witness_map: NULL buffer");
    B2Z_REQUIRE(log_n <= 32, B2Z_ESIZE, "b2z_witness_map: domain larger than 2^32 (PolynomialDegreeTooLarge)");
    B2Z_REQUIRE(log_n <= 28, B2Z_ENOMEM, "b2z_witness_map: domain does not fit this build's single-GPU plan");
    const size_t n = (size_t)1 << log_n;
    cudaStream_t st = c.stream;
    DevBuf<FrEl> da(n), db(n), dc(n);
    B2Z_CUDA(cudaMemcpyAsync(da.p, a, n * sizeof(FrEl), cudaMemcpyHostToDevice, st));
    B2Z_CUDA(cudaMemcpyAsync(db.p, b, n * sizeof(FrEl), cudaMemcpyHostToDevice, st));
    B2Z_CUDA(cudaMemcpyAsync(dc.p, cc, n * sizeof(FrEl), cudaMemcpyHostToDevice, st));
    witness_map_device(&c, da.p, db.p, dc.p, log_n, /*natural_out=*/true, st);
    B2Z_CUDA(cudaMemcpyAsync(h_out, da.p, n * sizeof(FrEl), cudaMemcpyDeviceToHost, st));
    B2Z_CUDA(cudaStreamSynchronize(st));
  });
}

}  // extern "C"
