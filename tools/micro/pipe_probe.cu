// Development microbenchmark: issue rates of the instruction mixes the FP64 field product uses.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;

template <int MODE>
__global__ void probe(u64* out, int iters) {
  u64 acc[8];
  double d[8];
  unsigned int w[8];
  for (int k = 0; k < 8; k++) { acc[k] = out[threadIdx.x + k]; d[k] = (double)(acc[k] & 0xffffffffffffull); w[k] = (unsigned)acc[k]; }
  const double y = (double)(out[threadIdx.x + 9] & 0xffffffffffffull);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (MODE == 0) {            // 64-bit add, 2 addends (IADD3 + IADD3.X with two carries)
        acc[k] += acc[(k + 1) & 7] + acc[(k + 2) & 7];
      } else if (MODE == 1) {     // 3 FP64 + the 2 integer adds of a limb product
        double h = __fma_rz(d[k], y, 0x1p100);
        double l = __fma_rz(d[k], y, (0x1p100 + 0x1p52) - h);
        acc[k] += (u64)__double_as_longlong(h);
        acc[(k + 1) & 7] += (u64)__double_as_longlong(l);
        d[k] = l - 0x1p52;
      } else if (MODE == 2) {     // 3 FP64 only
        double h = __fma_rz(d[k], y, 0x1p100);
        double l = __fma_rz(d[k], y, (0x1p100 + 0x1p52) - h);
        d[k] = l - 0x1p52;
      } else if (MODE == 3) {     // 32-bit 3-input add, no carry
        w[k] += w[(k + 1) & 7] + w[(k + 2) & 7];
      } else if (MODE == 4) {     // all-FP variant: chained hi, lo accumulated by DADD
        double prev = d[(k + 1) & 7];
        double nw = __fma_rz(d[k], y, prev);
        double l = __fma_rz(d[k], y, prev - nw);
        d[(k + 1) & 7] = nw;
        d[(k + 2) & 7] += l;
      } else if (MODE == 5) {     // 3 FP64 + 2 32-bit adds without carries
        double h = __fma_rz(d[k], y, 0x1p100);
        double l = __fma_rz(d[k], y, (0x1p100 + 0x1p52) - h);
        w[k] += (unsigned)__double2loint(h) + (unsigned)__double2hiint(l);
        w[(k + 1) & 7] += (unsigned)__double2loint(l) + (unsigned)__double2hiint(h);
        d[k] = l - 0x1p52;
      }
    }
  }
  u64 s = 0;
  for (int k = 0; k < 8; k++) s += acc[k] + (u64)__double_as_longlong(d[k]) + w[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  u64* d; cudaMalloc(&d, 8 * sms * 8 * 256 + 1024); cudaMemset(d, 1, 8 * sms * 8 * 256 + 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[6] = {"u64 add3", "4 FP64 + 2 u64 add", "4 FP64", "u32 add3", "4 FP64 (all-FP)", "4 FP64 + 2 u32 add3"};
  for (int wps = 1; wps <= 4; wps *= 2) {
    for (int mode = 0; mode < 6; mode++) {
      const int iters = 4000;
      float ms = 0;
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        const int blocks = sms, threads = 128 * wps;
        if (mode == 0) probe<0><<<blocks, threads>>>(d, iters);
        if (mode == 1) probe<1><<<blocks, threads>>>(d, iters);
        if (mode == 2) probe<2><<<blocks, threads>>>(d, iters);
        if (mode == 3) probe<3><<<blocks, threads>>>(d, iters);
        if (mode == 4) probe<4><<<blocks, threads>>>(d, iters);
        if (mode == 5) probe<5><<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      // cycles per "item" (one k-iteration) per warp per SMSP at 1.9 GHz
      const double items = 8.0 * iters * wps;          // per SMSP: wps warps
      printf("%-22s %d warp/SMSP: %7.3f ms  %6.2f cycles per item per SMSP\n", names[mode], wps, ms, ms * 1e-3 * 1.9e9 / items);
    }
  }
  return 0;
}
