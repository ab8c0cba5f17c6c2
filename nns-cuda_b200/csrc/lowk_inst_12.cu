// lowk_inst_12.cu -- instantiates the low-k search kernels for k = 25..26 (split for parallel builds)
#define LOWK_K_LO 25
#define LOWK_K_HI 26
#define LOWK_RANGE_FN lowk_launch_range_12
#include "lowk_inst.cuh"
