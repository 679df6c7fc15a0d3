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

// DMMA (mma.sync.aligned.m8n8k4.row.col.f64): D(8x8) += A(8x4) B(4x8), 256 FMA per warp instruction.
__device__ __forceinline__ void dmma(double & d0, double & d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_thr(double * out, long long * cyc, double c, int iters)
{
  double d[NACC][2];
#pragma unroll
  for (int k = 0; k < NACC; k++) d[k][0] = d[k][1] = threadIdx.x * 1e-3 + k;
  const double a = 1e-3 * (threadIdx.x & 7), b = c;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++)
  {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int k = 0; k < NACC; k++) dmma(d[k][0], d[k][1], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < NACC; k++) s += d[k][0] + d[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NACC>
static void run_dmma(double * out, long long * cyc, const cudaDeviceProp & p)
{
  // latency: one warp, NACC independent chains
  long long h = 0;
  for (int rep = 0; rep < 2; rep++)
  {
    dmma_thr<NACC><<<1, 32>>>(out, cyc, 0.999, 1000);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  }
  printf("DMMA m8n8k4, one warp, %d independent accumulators: %.1f cycles per mma (issue-to-issue)\n", NACC, h / (4000.0 * NACC));
  for (int tpb = 128; tpb <= 512; tpb *= 2)
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 5000, blocks = p.multiProcessorCount * (1024 / tpb);
    dmma_thr<NACC><<<blocks, tpb>>>(out, cyc, 0.999, 100);
    cudaEventRecord(e0);
    dmma_thr<NACC><<<blocks, tpb>>>(out, cyc, 0.999, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 256.0 * 4.0 * NACC * iters * (double)blocks * (tpb / 32);
    printf("DMMA throughput, %d acc/warp: %d blocks x %d threads (%d warps/SM): %.2f TFLOP/s (%.1f FMA/clk/SM)\n", NACC, blocks, tpb, 1024 / 32,
           fl / ms / 1e9, fl / 2 / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3));
  }
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
  run_dmma<1>(out, cyc, p);
  run_dmma<2>(out, cyc, p);
  run_dmma<4>(out, cyc, p);
  run_dmma<8>(out, cyc, p);
  return 0;
}
