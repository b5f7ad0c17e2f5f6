// Throughput of the FP64 mma.sync shapes on this GPU (developer probe; build: nvcc -arch=sm_100a -O3 -o dmma_shapes dmma_shapes.cu)
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__global__ void probe(double* out, int iters) {
  double c[8][4];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = 0.0;
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  double b0 = 1.0 + threadIdx.x * 1e-6, b1 = b0 + 1, b2 = b0 + 2, b3 = b0 + 3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (SHAPE == 0)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a0), "d"(b0));
      else if (SHAPE == 1)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b0));
      else if (SHAPE == 2)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(a4), "d"(a5), "d"(a6), "d"(a7), "d"(b0), "d"(b1), "d"(b2), "d"(b3));
    }
  }
  double s = 0;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
void run(const char* name, double flops_per_mma) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 4, threads = 256, iters = 20000;
  double* out;
  cudaMalloc(&out, (size_t)blocks * threads * sizeof(double));
  probe<SHAPE><<<blocks, threads>>>(out, 100);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<SHAPE><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = (double)blocks * (threads / 32) * iters * 8.0;
  printf("%-10s %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, ms, mmas * flops_per_mma / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("m8n8k4", 2.0 * 8 * 8 * 4);
  run<1>("m16n8k4", 2.0 * 16 * 8 * 4);
  run<2>("m16n8k8", 2.0 * 16 * 8 * 8);
  run<3>("m16n8k16", 2.0 * 16 * 8 * 16);
  return 0;
}
