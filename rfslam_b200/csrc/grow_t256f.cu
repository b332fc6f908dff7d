// grow_t256f.cu -- the FULL build of the tree kernel: unordered-factor (MWCP) cuts (rf.c:14452-14518, routing rf.c:1607-1616),
// the sorted-walk path of covariates with more than 256 distinct values and the single-node emit mode
// (rfslam_b200_score_node) on top of what every build has.  Tables that need none of these run the lean builds, whose
// once-per-node code is smaller (see kFull in grow_kernel.cuh).
#define RF_THREADS 256
#define RF_MIN_CTAS 2
#define RF_FULL 1
#define RF_KERNEL_NAME rfslam_grow_kernel_t256f
#include "grow_kernel.cuh"
#include "grow_variant.h"
RFSLAM_DEFINE_VARIANT(t256f, rfslam_grow_kernel_t256f)
