// dmma_smem_bench.cu -- ceiling of the k_step main loop body: DMMA m8n8k4 fed from shared memory with exactly the
// fragment addressing of slab_mma<8, 0> (A slab [128][16] k-swizzled, field slab [16][68]), no TMA, no epilogue.
// Variants: CTAs per SM (1 or 2 by dynamic smem size), a per-slab warp sync, operand tiles per slab distinct or reused.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_smem_bench dmma_smem_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define KB 16
#define SB 68

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NSTAGE, int SYNC>
__global__ void __launch_bounds__(256, 2) k(double *out, int slabs)
{
  extern __shared__ __align__(128) unsigned char smem[];
  const int stage_bytes = 128 * KB * 8 + KB * SB * 8;
  const int tid = threadIdx.x, lane = tid & 31, wr = tid >> 5, gq = lane >> 2, tq = lane & 3;
  for (int i = tid; i < NSTAGE * stage_bytes / 8; i += blockDim.x) reinterpret_cast<double *>(smem)[i] = 1e-3 * (i % 97);
  __syncthreads();
  double acc[2][8][2];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  const int swz = 4 * (gq & 3);
  for (int s = 0; s < slabs; ++s) {
    const unsigned char *sp = smem + (s % NSTAGE) * stage_bytes;
    const double *a = reinterpret_cast<const double *>(sp) + (wr * 16 + gq) * KB;
    const double *b = reinterpret_cast<const double *>(sp + 128 * KB * 8) + tq * SB + gq;
#pragma unroll
    for (int ks4 = 0; ks4 < 4; ++ks4) {
      const int kc = (ks4 * 4 + tq) ^ swz;
      const double a0 = a[kc], a1 = a[8 * KB + kc];
      double bv[8];
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) bv[ni] = b[ks4 * 4 * SB + ni * 8];
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) { dmma(acc[0][ni][0], acc[0][ni][1], a0, bv[ni]); dmma(acc[1][ni][0], acc[1][ni][1], a1, bv[ni]); }
    }
    if (SYNC) __syncwarp();
  }
  double sum = 0;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) sum += acc[mi][ni][0] + acc[mi][ni][1];
  out[blockIdx.x * blockDim.x + tid] = sum;
}

template <int NSTAGE, int SYNC>
void run(int ctas_per_sm)
{
  double *out;
  cudaMalloc(&out, 148 * 2 * 256 * 8);
  const int slabs = 20000;
  const int smem = ctas_per_sm == 2 ? 100 * 1024 : 200 * 1024;   // occupancy selected by the shared memory request
  cudaFuncSetAttribute(k<NSTAGE, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NSTAGE, SYNC><<<148 * ctas_per_sm, 256, smem>>>(out, 100);
  cudaEventRecord(e0);
  k<NSTAGE, SYNC><<<148 * ctas_per_sm, 256, smem>>>(out, slabs);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fl = (double)148 * ctas_per_sm * 8 * slabs * 64 * 512;
  printf("stages=%d warpsync=%d CTAs/SM=%d (%2d warps/SM): %6.2f TFLOP/s (%s)\n", NSTAGE, SYNC, ctas_per_sm, 8 * ctas_per_sm,
         fl / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main()
{
  run<1, 0>(1); run<1, 0>(2);
  run<3, 0>(1); run<3, 0>(2);
  run<3, 1>(2);
  return 0;
}
