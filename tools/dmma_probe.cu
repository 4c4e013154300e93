// The north-star's only tensor-core question (BASELINE.json configs[3]): when many chains are
// evaluated together the basis x coefficient product is a dense contraction -- is fp64 DMMA worth it on B200?
// The contraction has 28 non-zeros out of 164 per row (7 dims x 4 taps), so the dense tensor-core form does
// 164 / 28 = 5.9x the flops of the sparse DFMA form; DMMA wins only if its throughput is > 5.9x the DFMA pipe's.
// This probe measures both peaks on the device: (a) independent DFMA chains, (b) mma.sync.aligned.m8n8k4.f64 with
// independent accumulators.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int blocks = prop.multiProcessorCount * 4, threads = 256, iters = 20000;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 16 * iters * (double)blocks * threads;
    if (rep) printf("{\"probe\": \"DFMA\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, flop / ms / 1e9);
    cudaEventRecord(e0);
    dmma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop2 = 2.0 * 8 * 8 * 4 * 8 * iters * (double)blocks * (threads / 32);
    if (rep) printf("{\"probe\": \"DMMA m8n8k4\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, flop2 / ms / 1e9);
  }
  printf("{\"device\": \"%s\", \"sms\": %d, \"error\": \"%s\"}\n", prop.name, prop.multiProcessorCount, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
