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
  // device pools handed back by freed batches (cudaMalloc/cudaFree of multi-GB pools cost tens of ms per call)
  char *cache_field = nullptr; size_t cache_field_bytes = 0;
  char *cache_kpool = nullptr; size_t cache_kpool_bytes = 0;
};
