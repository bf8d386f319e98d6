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
};
