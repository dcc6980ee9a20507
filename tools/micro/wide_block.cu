// Development microbenchmark: 13x13 blocks of flag-free wide multiply-adds with register
// operands as the 30-bit-limb product issues them, with and without ALU work beside them.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(const int* in, long long* out, int iters) {
  int a[13], b[13];
  long long col[13];
  for (int j = 0; j < 13; j++) { a[j] = in[threadIdx.x * 26 + j] & 0x3fffffff; b[j] = in[threadIdx.x * 26 + 13 + j] & 0x3fffffff; col[j] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 13; i++) {
#pragma unroll
      for (int j = 0; j < 13; j++) col[j] += (long long)a[j] * (long long)b[i];
      if (MODE >= 1) {          // per-round: mask / shift / add on the ALU pipe (5 instructions)
        const long long carry = col[0] >> 30;
        col[1] += carry;
        col[0] &= 0x3fffffff;
      }
      if (MODE >= 2) {          // more ALU work: 13 cheap ops per round
#pragma unroll
        for (int j = 0; j < 13; j++) a[j] = (a[j] ^ (int)col[j]) & 0x3fffffff;
      }
    }
  }
  long long s = 0;
  for (int j = 0; j < 13; j++) s += col[j] + a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int* in; long long* out;
  cudaMalloc(&in, 4 * 26 * 512); cudaMemset(in, 0x5a, 4 * 26 * 512); cudaMalloc(&out, 8 * sms * 512);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int wps = 1; wps <= 4; wps *= 2)
    for (int mode = 0; mode < 3; mode++) {
      const int iters = 1000; float ms = 0;
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<sms, 128 * wps>>>(in, out, iters);
        if (mode == 1) k<1><<<sms, 128 * wps>>>(in, out, iters);
        if (mode == 2) k<2><<<sms, 128 * wps>>>(in, out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("mode %d  %d warp/SMSP: %7.3f ms  %5.2f cycles per wide multiply-add per SMSP\n", mode, wps, ms,
             ms * 1e-3 * 1.9e9 / (169.0 * iters * wps));
    }
  return 0;
}
