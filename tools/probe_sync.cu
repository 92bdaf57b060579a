// probe_sync.cu — what a dependent launch and a grid-wide barrier cost on this GPU.
//
// Decides how the non-sensor part of an update is structured (DESIGN.md 4,
// "fused tail"): ~25 dependent kernels of 3-10 us each against one cooperative
// kernel whose phases are separated by grid barriers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_sync probe_sync.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

namespace cg = cooperative_groups;

__global__ void empty_kernel(int *p) {
  if (p && threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1;
}

__global__ void __launch_bounds__(256) sync_kernel(int iters, unsigned *sink) {
  cg::grid_group grid = cg::this_grid();
  unsigned acc = 0;
  for (int i = 0; i < iters; ++i) {
    acc += i;
    grid.sync();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

// Hand-rolled sense-reversing barrier on one global counter (all blocks resident).
__global__ void __launch_bounds__(256) spin_kernel(int iters, unsigned *counter, unsigned *sink) {
  unsigned acc = 0;
  unsigned target = 0;
  for (int i = 0; i < iters; ++i) {
    acc += i;
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counter, 1u);
      while (*reinterpret_cast<volatile unsigned *>(counter) < target) {
      }
      __threadfence();
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

#define CK(x)                                                              \
  do {                                                                     \
    cudaError_t e = (x);                                                   \
    if (e != cudaSuccess) {                                                \
      printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));     \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main() {
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  int *d;
  CK(cudaMalloc(&d, 64));
  CK(cudaMemset(d, 0, 64));
  unsigned *sink = reinterpret_cast<unsigned *>(d) + 4, *counter = reinterpret_cast<unsigned *>(d) + 8;
  float ms;
  // 1. back-to-back launches (each depends on the previous one through the stream)
  for (int blocks : {1, 148, 592}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(a, s));
      for (int i = 0; i < 2000; ++i) empty_kernel<<<blocks, 256, 0, s>>>(d);
      CK(cudaEventRecord(b, s));
      CK(cudaEventSynchronize(b));
      CK(cudaEventElapsedTime(&ms, a, b));
    }
    printf("stream launches  grid %4d x 256: %.2f us per dependent launch\n", blocks, ms * 1e3 / 2000);
  }
  // 2. the same 25 launches as a CUDA graph
  {
    cudaGraph_t g;
    cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < 25; ++i) empty_kernel<<<148, 256, 0, s>>>(d);
    CK(cudaStreamEndCapture(s, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(a, s));
      for (int i = 0; i < 100; ++i) CK(cudaGraphLaunch(ge, s));
      CK(cudaEventRecord(b, s));
      CK(cudaEventSynchronize(b));
      CK(cudaEventElapsedTime(&ms, a, b));
    }
    printf("graph of 25 kernels (148 x 256): %.2f us per graph, %.2f us per node\n", ms * 1e3 / 100,
           ms * 1e3 / 2500);
  }
  // 3. cooperative grid.sync and the hand-rolled barrier
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  for (int per_sm : {1, 2, 4, 8}) {
    int iters = 2000;
    int blocks = sms * per_sm;
    void *args[] = {&iters, &sink};
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(a, s));
      CK(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(sync_kernel), dim3(blocks), dim3(256), args, 0, s));
      CK(cudaEventRecord(b, s));
      CK(cudaEventSynchronize(b));
      CK(cudaEventElapsedTime(&ms, a, b));
    }
    printf("grid.sync        grid %4d x 256: %.3f us per barrier\n", blocks, ms * 1e3 / iters);
    CK(cudaMemsetAsync(counter, 0, 4, s));
    void *args2[] = {&iters, &counter, &sink};
    CK(cudaEventRecord(a, s));
    CK(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(spin_kernel), dim3(blocks), dim3(256), args2, 0, s));
    CK(cudaEventRecord(b, s));
    CK(cudaEventSynchronize(b));
    CK(cudaEventElapsedTime(&ms, a, b));
    printf("atomic barrier   grid %4d x 256: %.3f us per barrier\n", blocks, ms * 1e3 / iters);
  }
  // 4. one cooperative launch, end to end (launch + 1 barrier)
  {
    int iters = 1;
    void *args[] = {&iters, &sink};
    CK(cudaEventRecord(a, s));
    for (int i = 0; i < 200; ++i)
      CK(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(sync_kernel), dim3(sms), dim3(256), args, 0, s));
    CK(cudaEventRecord(b, s));
    CK(cudaEventSynchronize(b));
    CK(cudaEventElapsedTime(&ms, a, b));
    printf("cooperative launch (148 x 256, 1 barrier): %.2f us each\n", ms * 1e3 / 200);
  }
  return 0;
}
