// Development microbenchmark: latency of the curve operations for ONE warp / one
// thread, inlined vs out-of-line (decides how the serial tails of the MSM are built).
#include <cstdio>
#include <cuda_runtime.h>
#include "ec.cuh"
using namespace b2z;

template <class C, int MODE>
__global__ void lat_kernel(typename C::Xyzz* io, typename C::Affine* aff, int iters, long long* cycles) {
  typename C::Xyzz a = io[threadIdx.x], b = io[threadIdx.x + 32];
  typename C::Affine q = aff[threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if constexpr (MODE == 0) a = C::add(a, b);
    if constexpr (MODE == 1) a = C::add_inl(a, b);
    if constexpr (MODE == 2) a = C::dbl(a);
    if constexpr (MODE == 3) a = C::dbl_inl(a);
    if constexpr (MODE == 4) a = C::madd(a, q);
    if constexpr (MODE == 5) a = C::madd_inl(a, q);
    if constexpr (MODE == 6) { a.x = Fq::mul(a.x, b.x); }     // one field mul (G1 only)
  }
  long long t1 = clock64();
  io[threadIdx.x] = a;
  if (threadIdx.x == 0) *cycles = (t1 - t0) / iters;
}

template <class C>
__global__ void init_kernel(typename C::Xyzz* io, typename C::Affine* aff) {
  // any on-curve-ish data is fine for latency; use small multiples of a fake point via doubling of (x,y)
  typename C::Affine g;
  for (int i = 0; i < 12; i++) {
    if constexpr (sizeof(typename C::Affine) == 96) { g.x.l[i] = FqParams::g1_gen_x(i); g.y.l[i] = FqParams::g1_gen_y(i); }
    else { g.x.c0.l[i] = FqParams::g2_gen_x0(i); g.x.c1.l[i] = FqParams::g2_gen_x1(i); g.y.c0.l[i] = FqParams::g2_gen_y0(i); g.y.c1.l[i] = FqParams::g2_gen_y1(i); }
  }
  typename C::Xyzz p = C::from_affine(g);
  for (int i = 0; i < 64; i++) { p = C::dbl(p); io[i] = p; p = C::madd(p, g); }
  for (int i = 0; i < 32; i++) aff[i] = g;
}

template <class C, int MODE>
void run(const char* name, int threads) {
  typename C::Xyzz* io; typename C::Affine* aff; long long* cyc;
  cudaMalloc(&io, 64 * sizeof(typename C::Xyzz)); cudaMalloc(&aff, 32 * sizeof(typename C::Affine)); cudaMalloc(&cyc, 8);
  init_kernel<C><<<1, 1>>>(io, aff);
  lat_kernel<C, MODE><<<1, threads>>>(io, aff, 4, cyc);
  lat_kernel<C, MODE><<<1, threads>>>(io, aff, 64, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-28s threads=%2d  %8lld cycles/op  (%s)\n", name, threads, h, cudaGetErrorString(e));
  cudaFree(io); cudaFree(aff); cudaFree(cyc);
}

int main() {
  run<G1, 6>("G1 Fq::mul", 1);
  run<G1, 6>("G1 Fq::mul", 32);
  run<G1, 0>("G1 add (call)", 1);  run<G1, 0>("G1 add (call)", 32);
  run<G1, 1>("G1 add (inline)", 1); run<G1, 1>("G1 add (inline)", 32);
  run<G1, 2>("G1 dbl (call)", 1);  run<G1, 3>("G1 dbl (inline)", 1);
  run<G1, 4>("G1 madd (call)", 1); run<G1, 5>("G1 madd (inline)", 1); run<G1, 5>("G1 madd (inline)", 32);
  run<G2, 0>("G2 add (call)", 1);  run<G2, 1>("G2 add (inline)", 1);
  run<G2, 2>("G2 dbl (call)", 1);  run<G2, 5>("G2 madd (inline)", 1);
  return 0;
}
