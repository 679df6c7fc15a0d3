// FP64 pipe microbenchmark for B200: dependent-issue latencies (DFMA, DMUL, sqrt, 1/x, x/y, rsqrt) with one
// warp, and DFMA throughput of the whole chip. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o
// tools/microbench_fp64 tools/microbench_fp64.cu ; run under gpurun. Numbers feed DESIGN.md §4.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void lat(double * out, long long * cyc, double x0, double c, int iters)
{
  double x = x0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++)
  {
#pragma unroll
    for (int k = 0; k < 16; k++)
    {
      if (OP == 0) x = fma(x, c, c);
      if (OP == 1) x = x * c;
      if (OP == 2) x = sqrt(x) + c;
      if (OP == 3) x = 1.0 / x + c;
      if (OP == 4) x = c / x;
      if (OP == 5) x = rsqrt(x) + c;
      if (OP == 6) x = 1.0 / sqrt(x) + c;
      if (OP == 7) x = __longlong_as_double(__shfl_sync(0xffffffffu, __double_as_longlong(x), 0)) ;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x] = x; cyc[blockIdx.x] = t1 - t0; }
}

__global__ void thr(double * out, double c, int iters)
{
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 1e-3 + k;
#pragma unroll 1
  for (int i = 0; i < iters; i++)
  {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int k = 0; k < 8; k++) a[k] = fma(a[k], c, c);
  }
  double s = 0;
  for (int k = 0; k < 8; k++) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
  double * out; long long * cyc;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 1 << 16);
  const char * names[] = {"DFMA", "DMUL", "sqrt(x)+c", "1/x+c", "c/x", "rsqrt(x)+c", "1/sqrt(x)+c", "shfl64"};
  for (int op = 0; op < 8; op++)
  {
    long long h = 0;
    for (int rep = 0; rep < 2; rep++)
    {
      switch (op) {
        case 0: lat<0><<<1, 32>>>(out, cyc, 1.3, 0.999, 1000); break;
        case 1: lat<1><<<1, 32>>>(out, cyc, 1.3, 0.9999, 1000); break;
        case 2: lat<2><<<1, 32>>>(out, cyc, 1.3, 0.5, 1000); break;
        case 3: lat<3><<<1, 32>>>(out, cyc, 1.3, 0.5, 1000); break;
        case 4: lat<4><<<1, 32>>>(out, cyc, 1.3, 1.7, 1000); break;
        case 5: lat<5><<<1, 32>>>(out, cyc, 1.3, 0.5, 1000); break;
        case 6: lat<6><<<1, 32>>>(out, cyc, 1.3, 0.5, 1000); break;
        case 7: lat<7><<<1, 32>>>(out, cyc, 1.3, 0.5, 1000); break;
      }
      cudaDeviceSynchronize();
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    }
    printf("latency %-14s %.1f cycles/op (op 2,3,5,6 include one dependent DADD)\n", names[op], h / 16000.0);
  }
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  for (int tpb = 128; tpb <= 1024; tpb *= 2)
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000, blocks = p.multiProcessorCount * (2048 / tpb);
    thr<<<blocks, tpb>>>(out, 0.999, 100);
    cudaEventRecord(e0);
    thr<<<blocks, tpb>>>(out, 0.999, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 32.0 * iters * (double)blocks * tpb;
    printf("DFMA throughput: %d blocks x %d threads: %.2f TFLOP/s (%.1f DFMA/clk/SM at %d MHz nominal)\n", blocks, tpb, fl / ms / 1e9,
           fl / 2 / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1000);
  }
  return 0;
}
