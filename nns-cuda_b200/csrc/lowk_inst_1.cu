// lowk_inst_1.cu -- instantiates the low-k search kernels for k = 5..8 (split for parallel builds)
#define LOWK_K_LO 5
#define LOWK_K_HI 8
#define LOWK_RANGE_FN lowk_launch_range_1
#include "lowk_inst.cuh"
