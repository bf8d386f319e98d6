// sosgpu_host.h -- host-side definitions shared by the translation units of libsosgpu.so
#pragma once
#include "../../include/sosgpu.h"
#include "sosgpu_internal.h"
#include <string>

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                               \
      return SOSGPU_ERR_CUDA;                                                                      \
    }                                                                                              \
  } while (0)

struct sosgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  long long launches = 0;
  size_t field_budget = (size_t)48 << 30;
  int max_wave_orders = 0;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr; float last_kernel_ms = 0.f;   // device time of the last glitter / synthesis kernel
  int *h_count = nullptr;            // pinned word pair for the active-count readback of the wave loop
  char *cache_field = nullptr; size_t cache_field_bytes = 0;   // wave pools parked by the last freed batch
  char *cache_kpool = nullptr; size_t cache_kpool_bytes = 0;
  double *grec_cache = nullptr; size_t grec_cache_bytes = 0;   // group-sum buffer parked by the last freed batch
  cudaMemPool_t pool = nullptr;      // stream-ordered pool (release threshold = never) behind sos_dmalloc / sos_dfree
};

template <class T> static inline cudaError_t sos_dmalloc(sosgpu_ctx *ctx, T **p, size_t bytes)
{
  return cudaMallocFromPoolAsync((void **)p, bytes ? bytes : 8, ctx->pool, ctx->stream);
}
static inline void sos_dfree(sosgpu_ctx *ctx, void *p)
{
  if (!p) return;
  if (ctx) cudaFreeAsync(p, ctx->stream);
  else cudaFree(p);
}
