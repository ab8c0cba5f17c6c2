// lowk_inst_3.cu -- instantiates the low-k search kernels for k = 7..8 (split for parallel builds)
#define LOWK_K_LO 7
#define LOWK_K_HI 8
#define LOWK_RANGE_FN lowk_launch_range_3
#include "lowk_inst.cuh"
