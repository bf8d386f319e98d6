// dmma_bench.cu -- register-resident DMMA issue-rate microbenchmark (which mma.sync f64 shape / occupancy
// reaches the FP64 tensor peak on sm_100a).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int NACC>
__global__ void k(double *out, int iters, double a0, double b0)
{
  double acc[NACC][4];
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = a0 + threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = b0 + threadIdx.x * 1e-3 - i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (SHAPE == 0)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a[i & 1]), "d"(b[i & 3]));
      else if (SHAPE == 1)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[i & 3]));
      else if (SHAPE == 2)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[i & 1]), "d"(b[2 + (i & 1)]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                       "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int NACC>
void run(const char *name, double flop_per_mma, int warps)
{
  double *out;
  cudaMalloc(&out, 148 * 1024 * 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<SHAPE, NACC><<<148, warps * 32>>>(out, 100, 1.0, 2.0);
  cudaEventRecord(e0);
  k<SHAPE, NACC><<<148, warps * 32>>>(out, iters, 1.0, 2.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double fl = (double)148 * warps * iters * NACC * flop_per_mma;
  printf("%-10s nacc=%2d warps/SM=%2d : %7.2f TFLOP/s  (%s)\n", name, NACC, warps, fl / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main()
{
  for (int w : {4, 8, 16, 32}) {
    run<0, 16>("m8n8k4", 512, w);
    run<1, 16>("m16n8k4", 1024, w);
    run<2, 16>("m16n8k8", 2048, w);
    run<3, 16>("m16n8k16", 4096, w);
  }
  run<0, 4>("m8n8k4", 512, 8);
  run<0, 8>("m8n8k4", 512, 8);
  run<0, 32>("m8n8k4", 512, 8);
  run<3, 4>("m16n8k16", 4096, 8);
  run<3, 8>("m16n8k16", 4096, 8);
  return 0;
}
