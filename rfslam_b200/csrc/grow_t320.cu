// grow_t320.cu -- the tree kernel with ten warps per CTA (two CTAs per SM): one warp per candidate when
// mtry is 9 or 10 (configs[2]: mtry = ceil(sqrt(100)) = 10), where the eight-warp build scores two rounds.
#define RF_THREADS 320
#define RF_MIN_CTAS 2
#define RF_KERNEL_NAME rfslam_grow_kernel_t320
#include "grow_kernel.cuh"
#include "grow_variant.h"
RFSLAM_DEFINE_VARIANT(t320, rfslam_grow_kernel_t320)
