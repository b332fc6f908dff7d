// grow_t256f.cu -- the tree kernel with unordered-factor (MWCP) support compiled in: level-subset cuts (deterministic and
// random complementary pairs, rf.c:14452-14518), routing by level mask (rf.c:1607-1616).  Tables without factor covariates
// run the all-numeric builds, whose once-per-node code is smaller.
#define RF_THREADS 256
#define RF_MIN_CTAS 2
#define RF_FACTORS 1
#define RF_KERNEL_NAME rfslam_grow_kernel_t256f
#include "grow_kernel.cuh"
#include "grow_variant.h"
RFSLAM_DEFINE_VARIANT(t256f, rfslam_grow_kernel_t256f)
