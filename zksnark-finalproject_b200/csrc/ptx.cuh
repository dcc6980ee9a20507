// Carry-chain primitives for wide-integer arithmetic on sm_100a.
//
// On the device every function is one PTX instruction (add.cc / addc.cc /
// mad.lo.cc / madc.hi.cc ...) that reads or writes the PTX carry flag CC.CF.
// ptxas fuses an adjacent (mad.lo.cc , madc.hi.cc) pair on an aligned register
// pair into one IMAD.WIDE.U32(.X) -- the 32x32+64 multiply-add of the integer
// pipe -- which is the instruction the MSM / NTT roofline is counted in.
//
// On the host the same functions are emulated with an explicit carry object
// (CF) so that the limb algorithms built on top (mont.cuh, ec.cuh) can be
// unit-tested on a CPU-only box and reused for the O(log n) host-side constants
// of a transform plan (roots of unity, n^-1, ...).  On the device CF is an
// empty tag: the carry lives in the hardware flag.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define B2Z_HD __host__ __device__ __forceinline__
// Out-of-line variant for the big curve operations: every kernel outside the
// accumulation hot loop calls ONE shared copy per curve instead of inlining
// several multi-thousand-instruction bodies (ptxas time is superlinear in them).
#define B2Z_HD_NOINLINE __host__ __device__ __noinline__
#else
#define B2Z_HD inline
#define B2Z_HD_NOINLINE __attribute__((noinline))
#endif

namespace b2z {
namespace ptx {

struct CF {
#if !defined(__CUDA_ARCH__)
  uint32_t v = 0;
#endif
};

B2Z_HD uint32_t add_cc(CF& cf, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  (void)cf; uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
  uint64_t t = (uint64_t)a + b; cf.v = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
B2Z_HD uint32_t addc_cc(CF& cf, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  (void)cf; uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
  uint64_t t = (uint64_t)a + b + cf.v; cf.v = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
B2Z_HD uint32_t addc(CF& cf, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  (void)cf; uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
  return a + b + cf.v;
#endif
}
B2Z_HD uint32_t sub_cc(CF& cf, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  (void)cf; uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
  uint64_t t = (uint64_t)a - b; cf.v = (uint32_t)((t >> 32) & 1); return (uint32_t)t;
#endif
}
B2Z_HD uint32_t subc_cc(CF& cf, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  (void)cf; uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
  uint64_t t = (uint64_t)a - b - cf.v; cf.v = (uint32_t)((t >> 32) & 1); return (uint32_t)t;
#endif
}
B2Z_HD uint32_t subc(CF& cf, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  (void)cf; uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
  return a - b - cf.v;
#endif
}
B2Z_HD uint32_t mad_lo_cc(CF& cf, uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  (void)cf; uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
  return add_cc(cf, (uint32_t)((uint64_t)a * b), c);
#endif
}
B2Z_HD uint32_t madc_lo_cc(CF& cf, uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  (void)cf; uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
  return addc_cc(cf, (uint32_t)((uint64_t)a * b), c);
#endif
}
B2Z_HD uint32_t mad_hi_cc(CF& cf, uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  (void)cf; uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
  return add_cc(cf, (uint32_t)(((uint64_t)a * b) >> 32), c);
#endif
}
B2Z_HD uint32_t madc_hi_cc(CF& cf, uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  (void)cf; uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
  return addc_cc(cf, (uint32_t)(((uint64_t)a * b) >> 32), c);
#endif
}
B2Z_HD uint32_t madc_hi(CF& cf, uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  (void)cf; uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
  return addc(cf, (uint32_t)(((uint64_t)a * b) >> 32), c);
#endif
}

// One multiply-add column: (lo, hi) += a*b with the carry chain running through
// both halves; a single asm block keeps the pair adjacent for IMAD.WIDE fusion.
B2Z_HD void mad_wide_cc(CF& cf, uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  (void)cf;
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
#else
  lo = mad_lo_cc(cf, a, b, lo); hi = madc_hi_cc(cf, a, b, hi);
#endif
}
B2Z_HD void madc_wide_cc(CF& cf, uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  (void)cf;
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
#else
  lo = madc_lo_cc(cf, a, b, lo); hi = madc_hi_cc(cf, a, b, hi);
#endif
}
// (lo, hi) = a*b + (clo, chi): three-address form used when the accumulator is
// shifted down by two limbs in the same step.
B2Z_HD void madc_wide_cc3(CF& cf, uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
#if defined(__CUDA_ARCH__)
  (void)cf;
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;"
               : "=&r"(lo), "=&r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
#else
  lo = madc_lo_cc(cf, a, b, clo); hi = madc_hi_cc(cf, a, b, chi);
#endif
}
B2Z_HD void mad_wide_cc3(CF& cf, uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
#if defined(__CUDA_ARCH__)
  (void)cf;
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;"
               : "=&r"(lo), "=&r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
#else
  lo = mad_lo_cc(cf, a, b, clo); hi = madc_hi_cc(cf, a, b, chi);
#endif
}

}  // namespace ptx
}  // namespace b2z
