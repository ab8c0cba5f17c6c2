// lowk_inst_1.cu -- instantiates the low-k search kernels for k = 3..4 (split for parallel builds)
#define LOWK_K_LO 3
#define LOWK_K_HI 4
#define LOWK_RANGE_FN lowk_launch_range_1
#include "lowk_inst.cuh"
