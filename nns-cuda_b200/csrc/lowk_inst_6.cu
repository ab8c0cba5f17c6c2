// lowk_inst_6.cu -- instantiates the low-k search kernels for k = 25..28 (split for parallel builds)
#define LOWK_K_LO 25
#define LOWK_K_HI 28
#define LOWK_RANGE_FN lowk_launch_range_6
#include "lowk_inst.cuh"
