// grow_t128.cu -- the tree kernel with four warps per CTA (four CTAs per SM): twice the trees in flight per SM.  A tree
// is a sequential chain of nodes whose serial sections (draws, tracker, records) occupy one warp while the rest of
// its CTA waits, so narrow CTAs leave fewer warps idle per serial section; used when a launch has enough trees.
#define RF_THREADS 128
#define RF_MIN_CTAS 4
#define RF_KERNEL_NAME rfslam_grow_kernel_t128
#include "grow_kernel.cuh"
#include "grow_variant.h"
RFSLAM_DEFINE_VARIANT(t128, rfslam_grow_kernel_t128)
