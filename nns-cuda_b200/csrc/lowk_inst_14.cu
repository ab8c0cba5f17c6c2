// lowk_inst_14.cu -- instantiates the low-k search kernels for k = 29..30 (split for parallel builds)
#define LOWK_K_LO 29
#define LOWK_K_HI 30
#define LOWK_RANGE_FN lowk_launch_range_14
#include "lowk_inst.cuh"
