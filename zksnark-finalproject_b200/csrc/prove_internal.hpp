// Entry points of prove.cu used by r1cs.cu.
#pragma once
#include "common.hpp"

struct b2z_pk;

namespace b2z {
bool pk_matches(const b2z_pk* pk, uint32_t log_n, uint64_t m, uint64_t l);
// whole-key proof from device-resident a/b/c evaluations (clobbered) and assignment
void prove_on_device_buffers(Ctx& c, const b2z_pk* pk, FrEl* d_a, FrEl* d_b, FrEl* d_c, const FrEl* d_z,
                             const uint64_t r[4], const uint64_t s[4], uint8_t proof_out[192]);
// the same for one shard of a key: B2Z_PARTIAL_BYTES of partial sums
void prove_partial_on_device_buffers(Ctx& c, const b2z_pk* pk, FrEl* d_a, FrEl* d_b, FrEl* d_c, const FrEl* d_z,
                                     const uint64_t r[4], const uint64_t s[4], uint8_t* partial_out);
// the same in two calls (sharded prover with a distributed witness map): z-only work first (returns while the
// GPU is busy), then quotient + H + host epilogue from the three COSET evaluation vectors (d_a is clobbered)
void prove_begin_on(Ctx& c, const b2z_pk* pk, const FrEl* d_z, const uint64_t r[4], const uint64_t s[4]);
// tile-sharded prover (r1cs.cu: b2z_dist_*): z-only sorts; the H sort from this shard's chunk of h; the chained
// accumulations + host epilogue once `ready` has fired
void prove_begin_sorts_on(Ctx& c, const b2z_pk* pk, const FrEl* d_z, const uint64_t r[4], const uint64_t s[4]);
uint32_t pk_h_chunk(const b2z_pk* pk, uint32_t* h_lo);
void prove_dist_h_sort_on(Ctx& c, const b2z_pk* pk, const FrEl* d_h, cudaStream_t st);
void prove_dist_finish_on(Ctx& c, const b2z_pk* pk, cudaEvent_t wm_done, cudaStream_t h_st, uint8_t* partial_out);
void combine_partials_host(const uint8_t* partials, uint32_t world, uint8_t proof_out[192]);
// remembers where the "assignment uploaded" flag of the r1cs used by shard_begin lives; prove_finish_on clears it
void pk_bind_assignment_flag(const b2z_pk* pk, bool* flag);
void prove_finish_on(Ctx& c, const b2z_pk* pk, FrEl* d_a, const FrEl* d_b, const FrEl* d_c, uint8_t* partial_out);
}  // namespace b2z
