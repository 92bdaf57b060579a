// probe_fp.cu — device facts needed by the kernel design (run once on a B200):
//  * does ptxas' FFMA2 fusion of mul.rn.f32x2 + add.rn.f32x2 change results?
//  * is the Markstein 3-op division bit-identical to IEEE division for a given divisor?
//  * device properties.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}

__global__ void probe_fusion(unsigned long long *mismatch_packed, unsigned long long *mismatch_scalar, int iters) {
  uint32_t s = hash32(blockIdx.x * blockDim.x + threadIdx.x + 1);
  unsigned long long mp = 0, ms = 0;
  for (int i = 0; i < iters; ++i) {
    s = hash32(s + i);
    float a = __uint_as_float((s & 0x007fffffu) | 0x3f800000u);            // [1,2)
    s = hash32(s);
    float b = __uint_as_float((s & 0x007fffffu) | 0x3f800000u);
    s = hash32(s);
    float c = -__uint_as_float((s & 0x007fffffu) | 0x3f800000u);
    // reference: two roundings
    float ref = __fadd_rn(__fmul_rn(a, b), c);
    float fused = __fmaf_rn(a, b, c);
    // packed path
    float2 A = make_float2(a, a), B = make_float2(b, b), Cc = make_float2(c, c);
    float2 P = __fadd2_rn(__fmul2_rn(A, B), Cc);
    if (P.x != ref) mp++;
    float sc = __fadd_rn(__fmul_rn(a, b), c);
    if (sc != ref || ref == fused) ms += (sc != ref);
  }
  atomicAdd(mismatch_packed, mp);
  atomicAdd(mismatch_scalar, ms);
}

__global__ void probe_div(float res, unsigned long long *mismatch, unsigned long long *idx_mismatch, int iters) {
  uint32_t s = hash32(blockIdx.x * blockDim.x + threadIdx.x + 77);
  const float r = __frcp_rn(res);
  unsigned long long mm = 0, im = 0;
  for (int i = 0; i < iters; ++i) {
    s = hash32(s + i);
    // a in [-2, 512): the range of (m - origin)
    float a = (float)(s >> 8) * (514.0f / 16777216.0f) - 2.0f;
    if (i & 1) { // near cell boundaries
      int cell = (s >> 8) % 10000;
      float edge = __fmul_rn((float)cell, res);
      int ulps = (int)(s & 7) - 3;
      a = __int_as_float(__float_as_int(edge) + ulps);
    }
    float q_ref = __fdiv_rn(a, res);
    float q = __fmul_rn(a, r);
    float rem = __fmaf_rn(-q, res, a);
    float q2 = __fmaf_rn(rem, r, q);
    if (q2 != q_ref) mm++;
    if (__float2int_rz(q2) != __float2int_rz(q_ref)) im++;
  }
  atomicAdd(mismatch, mm);
  atomicAdd(idx_mismatch, im);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device: %s sm_%d%d SMs=%d L2=%d MB smem/SM=%zu KB smem/block(optin)=%zu KB regs/SM=%d clock=%d kHz mem=%zu MB\n",
         p.name, p.major, p.minor, p.multiProcessorCount, p.l2CacheSize >> 20, p.sharedMemPerMultiprocessor >> 10,
         p.sharedMemPerBlockOptin >> 10, p.regsPerMultiprocessor, p.clockRate, p.totalGlobalMem >> 20);
  unsigned long long *d; cudaMalloc(&d, 4 * sizeof(unsigned long long)); cudaMemset(d, 0, 32);
  probe_fusion<<<296, 256>>>(d, d + 1, 4096);
  float ress[] = {0.05f, 0.025f, 0.1f, 0.5f, 0.03f, 0.0123f};
  unsigned long long h[4];
  cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("fusion probe: packed mul2+add2 mismatches vs 2-rounding reference: %llu of %llu ; scalar mismatches: %llu\n",
         h[0], 296ull * 256 * 4096, h[1]);
  for (float res : ress) {
    cudaMemset(d, 0, 32);
    probe_div<<<296, 256>>>(res, d + 2, d + 3, 8192);
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("markstein division res=%g: quotient mismatches %llu, index mismatches %llu of %llu\n", res, h[2], h[3],
           296ull * 256 * 8192);
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
