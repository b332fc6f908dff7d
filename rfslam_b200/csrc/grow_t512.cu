// grow_t512.cu -- the tree kernel with sixteen warps per CTA (one CTA per SM): used when a launch holds no
// more trees than the GPU has SMs (a 1000-tree forest sharded over 8 GPUs), where half of every SM would idle
// and the passes over large nodes are the per-tree critical path.
#define RF_THREADS 512
#define RF_MIN_CTAS 1
#define RF_KERNEL_NAME rfslam_grow_kernel_t512
#include "grow_kernel.cuh"
#include "grow_variant.h"
RFSLAM_DEFINE_VARIANT(t512, rfslam_grow_kernel_t512)
