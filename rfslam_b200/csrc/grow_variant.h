// grow_variant.h -- one build of the tree kernel (grow_kernel.cuh at some CTA width) behind plain function pointers,
// so that grow.cu can pick the width per launch.  Each variant lives in its own translation unit: CTA width is a
// compile-time constant of the kernel (warp roles, staging per thread, launch bounds).
#pragma once
#include <cstdlib>
#include "internal.h"

namespace rfslam {
struct DevData; struct GrowParams; struct GrowWork;
struct GrowVariant {
  int threads;
  int ctasPerSm;                                                          // resident CTAs per SM the build is compiled for (launch bounds)
  cudaError_t (*prepare)(size_t smem, int* ctasPerSm);                    // opt in to `smem` bytes of dynamic shared memory; resident CTAs per SM
  void (*launch)(const DevData& d, const GrowParams& g, const GrowWork& w, int grid, size_t smem);
};
const GrowVariant* grow_variant_t128();
const GrowVariant* grow_variant_t256();
const GrowVariant* grow_variant_t256f();   // the FULL build: factors, nsplit, wide covariates, staged subtrees, emit mode, profile
const GrowVariant* grow_variant_t320();
const GrowVariant* grow_variant_t512();
}  // namespace rfslam

#define RFSLAM_DEFINE_VARIANT(tag, kernel)                                                                        \
  namespace rfslam {                                                                                              \
  static cudaError_t prepare_##tag(size_t smem, int* ctasPerSm) {                                                 \
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    if (e != cudaSuccess) return e;                                                                               \
    if (const char* cv = getenv("RFSLAM_CARVEOUT")) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv)); \
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctasPerSm, kernel, kThreads, smem);                      \
  }                                                                                                               \
  static void launch_##tag(const DevData& d, const GrowParams& g, const GrowWork& w, int grid, size_t smem) {     \
    kernel<<<grid, kThreads, smem>>>(d, g, w);                                                                    \
  }                                                                                                               \
  const GrowVariant* grow_variant_##tag() {                                                                       \
    static const GrowVariant v = {kThreads, RF_MIN_CTAS, prepare_##tag, launch_##tag};                                         \
    return &v;                                                                                                    \
  }                                                                                                               \
  }
