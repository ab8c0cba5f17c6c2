// lowk_inst_0.cu -- instantiates the low-k search kernels for k = 1..2 (split for parallel builds)
#define LOWK_K_LO 1
#define LOWK_K_HI 2
#define LOWK_RANGE_FN lowk_launch_range_0
#include "lowk_inst.cuh"
