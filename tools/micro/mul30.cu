// Development microbenchmark (negative result, kept as evidence for DESIGN.md 4): a 30-bit-limb
// carry-free product (mont30.cuh, next to this file) against the shipped 32-bit-limb carry-chained
// one (csrc/mont.cuh) -- agreement on random and edge inputs for Fq and Fr, then throughput of both.
// B200: 20.0 vs 29.5 G Fq products/s -- dispatch, not the multiplier pipe, is the limit.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "mont.cuh"
#include "mont30.cuh"
using namespace b2z;

template <class P> __device__ __forceinline__ Fp<P> mul30(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r; w30::mul_words<P>(a.l, b.l, r.l); return r;
}

template <class P, int MODE>
__global__ void chain(const Fp<P>* in, Fp<P>* out, int iters) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fp<P> x = in[2 * t], y = in[2 * t + 1];
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) x = Field<P>::mul(y, x);
    if (MODE == 1) x = mul30<P>(y, x);
  }
  out[t] = x;
}

template <class P>
__global__ void check(const Fp<P>* in, int n, int* bad) {
  using F = Field<P>;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  Fp<P> x = F::reduce(in[2 * t]), y = in[2 * t + 1];          // first operand canonical (needed for Fr only)
  for (int i = 0; i < 64; i++) {
    Fp<P> r0 = F::reduce(F::mul(x, y)), r1 = mul30<P>(x, y);
    bool lt = false;                                            // lazy result must be below 2p
    for (int k = P::N - 1; k >= 0; k--) { if (r1.l[k] != P::p2(k)) { lt = r1.l[k] < P::p2(k); break; } }
    r1 = F::reduce(r1);
    bool ok = lt;
    for (int k = 0; k < P::N; k++) ok = ok && r0.l[k] == r1.l[k];
    if (!ok) atomicAdd(bad, 1);
    Fp<P> nx = F::reduce(F::add(F::mul(x, y), y));
    y = F::sub(F::mul(x, y), x);                            // lazy
    x = nx;
  }
}

static unsigned long long rng = 88172645463325252ull;
static unsigned int next32() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (unsigned int)(rng >> 16); }

template <class P>
void run(const char* name, unsigned top_mask) {
  const int n = 1 << 17, N = P::N;
  Fp<P>* h = (Fp<P>*)malloc(sizeof(Fp<P>) * 2 * n);
  for (int i = 0; i < 2 * n; i++) {
    for (int k = 0; k < N; k++) h[i].l[k] = next32();
    h[i].l[N - 1] &= top_mask;                       // below 2p
  }
  for (int k = 0; k < N; k++) {                      // edge values: 0, p (lazy 0), 2p-1, p-1, 1
    h[0].l[k] = 0; h[1].l[k] = P::p(k);
    h[2].l[k] = P::p(k); h[3].l[k] = P::p2(k);
    h[4].l[k] = P::p(k); h[5].l[k] = k == 0;
    h[6].l[k] = P::p(k); h[7].l[k] = P::p(k);
    h[8].l[k] = k == 0; h[9].l[k] = P::p2(k);
  }
  h[2].l[0] -= 1; h[3].l[0] -= 1; h[4].l[0] -= 1; h[9].l[0] -= 1;
  Fp<P>*din, *dout; int* dbad;
  cudaMalloc(&din, sizeof(Fp<P>) * 2 * n); cudaMalloc(&dout, sizeof(Fp<P>) * n); cudaMalloc(&dbad, 4); cudaMemset(dbad, 0, 4);
  cudaMemcpy(din, h, sizeof(Fp<P>) * 2 * n, cudaMemcpyHostToDevice);
  check<P><<<n / 128, 128>>>(din, n, dbad);
  int bad = -1; cudaMemcpy(&bad, dbad, 4, cudaMemcpyDeviceToHost);
  printf("%s check: %d mismatches over %d products (%s)\n", name, bad, n * 64, cudaGetErrorString(cudaDeviceSynchronize()));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int bps = 1; bps <= 4; bps *= 2) {
    for (int mode = 0; mode < 2; mode++) {
      const int blocks = sms * bps, iters = 2000;
      float ms = 0;
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) chain<P, 0><<<blocks, 128>>>(din, dout, iters);
        if (mode == 1) chain<P, 1><<<blocks, 128>>>(din, dout, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("%s %-9s %d warp/SMSP: %8.3f ms  %7.2f G mul/s  %6.0f cycles per warp product per SMSP (%s)\n", name,
             mode ? "30-bit" : "32-bit", bps, ms, (double)blocks * 128 * iters / ms / 1e6,
             ms * 1e-3 * 1.9e9 / ((double)iters * bps), cudaGetErrorString(cudaGetLastError()));
    }
  }
}

int main() {
  run<FqParams>("Fq", 0x1fffffffu);
  run<FrParams>("Fr", 0x7fffffffu);
  return 0;
}
