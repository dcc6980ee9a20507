// Development microbenchmark: issue cost of the integer multiply-add forms.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;

template <int MODE>
__global__ void probe(u64* out, int iters) {
  u64 acc[8];
  unsigned w[8], x[8];
  for (int k = 0; k < 8; k++) { acc[k] = out[threadIdx.x + k]; w[k] = (unsigned)(acc[k] >> 7); x[k] = (unsigned)acc[k] | 1; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (MODE == 0) {          // IMAD.WIDE.U32 with 64-bit accumulate, no carry flags
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(w[k]));
      } else if (MODE == 1) {   // IMAD.LO 32-bit
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(w[k]) : "r"(x[k]), "r"(x[(k + 1) & 7]));
      } else if (MODE == 4) {   // one flag-free wide multiply-add + one 3-input add on the ALU pipe
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(w[0]));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(x[(k + 3) & 7]));
      } else if (MODE == 5) {   // one wide multiply-add + two ALU instructions
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(w[0]));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(x[(k + 3) & 7]));
        asm volatile("xor.b32 %0, %0, %1;" : "+r"(w[k]) : "r"(x[(k + 5) & 7]));
      } else if (MODE == 6) {   // ALU only: two instructions
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(x[(k + 3) & 7]));
        asm volatile("xor.b32 %0, %0, %1;" : "+r"(w[k]) : "r"(x[(k + 5) & 7]));
      } else if (MODE == 2) {   // IMAD.HI 32-bit
        asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(w[k]) : "r"(x[k]), "r"(x[(k + 1) & 7]));
      }
    }
    if (MODE == 3) {            // carry chain of 8 wide multiply-adds (what the field product issues)
      unsigned lo[8], hi[8];
#pragma unroll
      for (int k = 0; k < 8; k++) { lo[k] = (unsigned)acc[k]; hi[k] = (unsigned)(acc[k] >> 32); }
      asm volatile(
          "mad.lo.cc.u32 %0, %16, %24, %0;\n\tmadc.hi.cc.u32 %1, %16, %24, %1;\n\t"
          "madc.lo.cc.u32 %2, %17, %24, %2;\n\tmadc.hi.cc.u32 %3, %17, %24, %3;\n\t"
          "madc.lo.cc.u32 %4, %18, %24, %4;\n\tmadc.hi.cc.u32 %5, %18, %24, %5;\n\t"
          "madc.lo.cc.u32 %6, %19, %24, %6;\n\tmadc.hi.cc.u32 %7, %19, %24, %7;\n\t"
          "madc.lo.cc.u32 %8, %20, %24, %8;\n\tmadc.hi.cc.u32 %9, %20, %24, %9;\n\t"
          "madc.lo.cc.u32 %10, %21, %24, %10;\n\tmadc.hi.cc.u32 %11, %21, %24, %11;\n\t"
          "madc.lo.cc.u32 %12, %22, %24, %12;\n\tmadc.hi.cc.u32 %13, %22, %24, %13;\n\t"
          "madc.lo.cc.u32 %14, %23, %24, %14;\n\tmadc.hi.u32 %15, %23, %24, %15;"
          : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]),
            "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
          : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(w[0]));
#pragma unroll
      for (int k = 0; k < 8; k++) acc[k] = ((u64)hi[k] << 32) | lo[k];
    }
  }
  u64 s = 0;
  for (int k = 0; k < 8; k++) s += acc[k] + w[k] + x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  u64* d; cudaMalloc(&d, 8 * sms * 512 + 1024); cudaMemset(d, 3, 8 * sms * 512 + 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[7] = {"IMAD.WIDE acc64", "IMAD.LO", "IMAD.HI", "IMAD.WIDE.X chain", "WIDE + 1 ALU", "WIDE + 2 ALU", "2 ALU"};
  for (int wps = 1; wps <= 4; wps *= 2) {
    for (int mode = 0; mode < 7; mode++) {
      const int iters = 8000;
      float ms = 0;
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) probe<0><<<sms, 128 * wps>>>(d, iters);
        if (mode == 1) probe<1><<<sms, 128 * wps>>>(d, iters);
        if (mode == 2) probe<2><<<sms, 128 * wps>>>(d, iters);
        if (mode == 3) probe<3><<<sms, 128 * wps>>>(d, iters);
        if (mode == 4) probe<4><<<sms, 128 * wps>>>(d, iters);
        if (mode == 5) probe<5><<<sms, 128 * wps>>>(d, iters);
        if (mode == 6) probe<6><<<sms, 128 * wps>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("%-18s %d warp/SMSP: %7.3f ms  %5.2f cycles per multiply-add per SMSP\n", names[mode], wps, ms,
             ms * 1e-3 * 1.9e9 / (8.0 * iters * wps));
    }
  }
  return 0;
}
