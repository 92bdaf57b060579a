// probe_gather.cu — measured gather peaks for the sensor model's roofline
// ("if the field is L2-resident, also report against a measured L2
// random-4 B-gather peak").
//
// Every thread issues independent gathers into a table of `bytes` bytes (1-byte
// or 4-byte elements) and sums what it reads.  Two address patterns:
//   random : every lane of a warp reads an unrelated element (32 lines per
//            warp instruction) — the textbook random-gather bound;
//   contour: the 32 lanes of a warp read inside one 16 x 8-cell neighbourhood
//            plus a short walk (3-4 lines per warp instruction) — what K7's
//            beam-major mapping produces on the strip layouts.
// The table sizes are K7's: 64 KB (the part of a 4 MB map a tracking cloud
// touches, L1-resident), 4 MB (C2 palette map), 16 MB (C2 float field / C3
// palette map), 64 MB (C3 float field), 256 MB (beyond L2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_gather probe_gather.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

template <typename T, int PATTERN>
__global__ void __launch_bounds__(256) gather_kernel(const T *__restrict__ table, uint32_t mask,
                                                     int iters, unsigned long long *sink) {
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31, warp = tid >> 5;
  uint32_t s = (PATTERN == 0 ? tid : warp) * 2654435761u + 12345u;
  unsigned acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  for (int i = 0; i < iters; ++i) {
    uint32_t idx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s ^= s << 13;
      s ^= s >> 17;
      s ^= s << 5;
      if (PATTERN == 0) {
        idx[k] = s & mask;
      } else {
        // warp-uniform anchor, lanes spread over a 512-element neighbourhood
        idx[k] = ((s & mask & ~511u) + ((lane * 13u + (s >> 20)) & 511u)) & mask;
      }
    }
    acc0 += static_cast<unsigned>(__ldg(table + idx[0]));
    acc1 += static_cast<unsigned>(__ldg(table + idx[1]));
    acc2 += static_cast<unsigned>(__ldg(table + idx[2]));
    acc3 += static_cast<unsigned>(__ldg(table + idx[3]));
  }
  const unsigned acc = acc0 + acc1 + acc2 + acc3;
  if (acc == 0xdeadbeefu) *sink = acc;
}

#define CK(x)                                                              \
  do {                                                                     \
    cudaError_t e = (x);                                                   \
    if (e != cudaSuccess) {                                                \
      printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));     \
      return 1;                                                            \
    }                                                                      \
  } while (0)

template <typename T, int PATTERN>
static int run(const void *table, size_t bytes, const char *name, unsigned long long *sink,
               FILE *json, bool *first) {
  const size_t elems = bytes / sizeof(T);
  const uint32_t mask = static_cast<uint32_t>(elems - 1);
  const int blocks = 148 * 8, iters = 512;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(a));
    gather_kernel<T, PATTERN><<<blocks, 256>>>(static_cast<const T *>(table), mask, iters, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (rep && ms < best) best = ms;
  }
  const double gathers = static_cast<double>(blocks) * 256 * iters * 4;
  const double rate = gathers / (best * 1e-3);
  printf("%-8s %zu-byte elements, table %7.2f MB: %8.1f G gathers/s = %7.1f GB/s useful, %7.1f GB/s at 4 B/gather\n",
         name, sizeof(T), bytes / 1048576.0, rate * 1e-9, rate * sizeof(T) * 1e-9, rate * 4e-9);
  fprintf(json, "%s\n  {\"pattern\": \"%s\", \"elem_bytes\": %zu, \"table_bytes\": %zu, \"gathers_per_s\": %.4e}",
          *first ? "" : ",", name, sizeof(T), bytes, rate);
  *first = false;
  return 0;
}

int main(int argc, char **argv) {
  const size_t max_bytes = size_t(256) << 20;
  void *table;
  CK(cudaMalloc(&table, max_bytes));
  CK(cudaMemset(table, 1, max_bytes));
  unsigned long long *sink;
  CK(cudaMalloc(&sink, 8));
  FILE *json = fopen(argc > 1 ? argv[1] : "/dev/null", "w");
  if (!json) return 1;
  fprintf(json, "{\"what\": \"tools/probe_gather.cu: independent gathers per second, 148 x 8 blocks x 256 threads\", \"results\": [");
  bool first = true;
  for (size_t bytes : {size_t(64) << 10, size_t(4) << 20, size_t(16) << 20, size_t(64) << 20, size_t(256) << 20}) {
    if (run<uint8_t, 0>(table, bytes, "random", sink, json, &first)) return 1;
    if (run<float, 0>(table, bytes, "random", sink, json, &first)) return 1;
    if (run<uint8_t, 1>(table, bytes, "contour", sink, json, &first)) return 1;
    if (run<float, 1>(table, bytes, "contour", sink, json, &first)) return 1;
  }
  fprintf(json, "\n]}\n");
  fclose(json);
  return 0;
}
