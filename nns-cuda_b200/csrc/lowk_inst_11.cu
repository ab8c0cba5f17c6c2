// lowk_inst_11.cu -- instantiates the low-k search kernels for k = 23..24 (split for parallel builds)
#define LOWK_K_LO 23
#define LOWK_K_HI 24
#define LOWK_RANGE_FN lowk_launch_range_11
#include "lowk_inst.cuh"
