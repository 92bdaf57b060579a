// Issue-rate probe for the FP32 pipes on sm_100a: cycles per warp instruction per SM
// sub-partition for scalar and packed (f32x2) forms.  Development aid.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_pipe probe_pipe.cu && ./probe_pipe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 2048
#define NACC 8
template <int OP>
__global__ void __launch_bounds__(256) k(u64 *out, u64 seed, float fa, float fb) {
  u64 acc[NACC];
  float f[2 * NACC];
  int ii[NACC];
  for (int i = 0; i < NACC; ++i) {
    acc[i] = seed + i * 0x3f8000003f800000ull + threadIdx.x;
    f[2 * i] = fa + i + threadIdx.x; f[2 * i + 1] = fb + i;
    ii[i] = threadIdx.x + i;
  }
  u64 A, B;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(fa), "f"(fb));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(fb), "f"(fa));
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (OP == 0) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[2*i]) : "f"(fa), "f"(fb));
                     asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[2*i+1]) : "f"(fa), "f"(fb)); }
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(A), "l"(B));
      if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[i]) : "l"(A));
      if (OP == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(acc[i]) : "l"(A));
      if (OP == 4) asm volatile("mul.rz.f32x2 %0, %0, %1;" : "+l"(acc[i]) : "l"(A));
      if (OP == 5) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(ii[i]) : "r"((int)seed), "r"(it)); }
      if (OP == 6) { asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(ii[i]) : "f"(f[2*i]));
                     asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[2*i]) : "r"(ii[i])); }
      if (OP == 7) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(A), "l"(B));
                     asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[2*i]) : "f"(fa), "f"(fb)); }
      if (OP == 8) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(A), "l"(B));
                     asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(ii[i]) : "r"((int)seed), "r"(it)); }
      if (OP == 9) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[2*i]) : "f"(fa));
                     asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[2*i+1]) : "f"(fb)); }
      if (OP == 10) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(A), "l"(B));
                      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ii[i]) : "r"((int)seed), "r"(it)); }
    }
  }
  u64 r = 0;
  for (int i = 0; i < NACC; ++i) r += acc[i] + (u64)__float_as_uint(f[2*i]) + __float_as_uint(f[2*i+1]) + ii[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int OP> void run(const char *name, int per_iter, int blocks_per_sm) {
  int dev_sms = 148;
  int blocks = dev_sms * blocks_per_sm;
  u64 *out; cudaMalloc(&out, (size_t)blocks * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<blocks, 256>>>(out, 1, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<OP><<<blocks, 256>>>(out, 1, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cycles = ms * 1e-3 * clk * 1e3;
  double warps_per_smsp = blocks_per_sm * 8 / 4.0;
  double instr = (double)ITERS * NACC * per_iter * warps_per_smsp;
  printf("%-28s blocks/SM %d: %.3f ms, %.2f cycles per warp-instr per SMSP (nominal clock %d kHz)\n",
         name, blocks_per_sm, ms, cycles / instr, clk);
  cudaFree(out);
}
int main() {
  for (int b : {2, 4}) {
    if (b == 2) {
    run<0>("fma.f32 x2 (scalar)", 2, 2); run<1>("fma.f32x2", 1, 2); run<2>("add.f32x2", 1, 2);
    run<3>("mul.f32x2", 1, 2); run<4>("mul.rz.f32x2", 1, 2); run<5>("mad.lo.s32", 1, 2);
    run<6>("cvt f2i + i2f", 2, 2); run<7>("fma.f32x2 + fma.f32", 2, 2); run<8>("fma.f32x2 + imad", 2, 2);
    run<9>("add.f32 x2 (scalar)", 2, 2); run<10>("fma.f32x2 + lop3", 2, 2);
    } else {
    run<0>("fma.f32 x2 (scalar)", 2, 4); run<1>("fma.f32x2", 1, 4); run<2>("add.f32x2", 1, 4);
    run<7>("fma.f32x2 + fma.f32", 2, 4); run<8>("fma.f32x2 + imad", 2, 4);
    }
  }
  return 0;
}
