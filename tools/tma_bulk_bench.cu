// tma_bulk_bench.cu -- per-SM throughput of non-tensor TMA bulk copies (cp.async.bulk global -> shared, SASS UBLKCP) as
// a function of copy size and copies in flight, L2-resident source.  Question it answers: the step kernel
// moves ~10 B/clk/SM through UBLKCP while its DMMA pipe idles 30 % -- is that the engine's ceiling for 1-D bulk copies?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bulk_bench tma_bulk_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int DEPTH>
__global__ void __launch_bounds__(128, 1) k(const char *src, size_t src_bytes, int copy_bytes, int iters, unsigned long long *cycles)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ unsigned long long bar[DEPTH];
  if (threadIdx.x == 0) {
    for (int d = 0; d < DEPTH; ++d) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bar + d)));
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const size_t span = (src_bytes / gridDim.x) & ~(size_t)1023;                   // each CTA walks its own L2-resident window
  const char *base = src + (size_t)blockIdx.x * span;
  const size_t nwin = span / copy_bytes;
  auto issue = [&](int i) {
    const int d = i % DEPTH;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar + d)), "r"(copy_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem + (size_t)d * copy_bytes)),
                 "l"(base + (size_t)(i % nwin) * copy_bytes), "r"(copy_bytes), "r"(smem_u32(bar + d))
                 : "memory");
  };
  const long long t0 = clock64();
  for (int i = 0; i < DEPTH - 1 && i < iters; ++i) issue(i);
  for (int i = 0; i < iters; ++i) {
    if (i + DEPTH - 1 < iters) issue(i + DEPTH - 1);
    const int d = i % DEPTH;
    const unsigned parity = (i / DEPTH) & 1;
    unsigned ok;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(smem_u32(bar + d)), "r"(parity) : "memory");
    } while (!ok);
  }
  cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

template <int DEPTH> void run(const char *src, size_t src_bytes, int copy_bytes, int ctas_per_sm_hint)
{
  const int grid = 148 * ctas_per_sm_hint;
  const int iters = 4000;
  unsigned long long *cyc;
  cudaMalloc(&cyc, grid * sizeof(unsigned long long));
  const int smem = DEPTH * copy_bytes;
  if (smem > 200 * 1024) { cudaFree(cyc); return; }
  cudaFuncSetAttribute(k<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<DEPTH><<<grid, 128, smem>>>(src, src_bytes, copy_bytes, 200, cyc);      // warm L2
  k<DEPTH><<<grid, 128, smem>>>(src, src_bytes, copy_bytes, iters, cyc);
  cudaDeviceSynchronize();
  unsigned long long h[148 * 2];
  cudaMemcpy(h, cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? (double)h[i] : mx;
  printf("copy %6d B  in flight %d  CTAs %3d : %6.2f B/clk per CTA, %6.2f B/clk per SM (%s)\n", copy_bytes, DEPTH - 1, grid,
         (double)iters * copy_bytes / mx, (double)iters * copy_bytes / mx * ctas_per_sm_hint, cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc);
}

int main()
{
  const size_t src_bytes = (size_t)64 << 20;                     // fits the 126 MB L2
  char *src;
  cudaMalloc(&src, src_bytes);
  cudaMemset(src, 1, src_bytes);
  for (int cb : {1024, 4096, 8704, 16384, 32768}) {
    run<2>(src, src_bytes, cb, 1);
    run<3>(src, src_bytes, cb, 1);
    run<5>(src, src_bytes, cb, 1);
    run<9>(src, src_bytes, cb, 1);
  }
  return 0;
}
