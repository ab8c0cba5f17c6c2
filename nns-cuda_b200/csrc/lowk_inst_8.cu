// lowk_inst_8.cu -- instantiates the low-k search kernels for k = 17..18 (split for parallel builds)
#define LOWK_K_LO 17
#define LOWK_K_HI 18
#define LOWK_RANGE_FN lowk_launch_range_8
#include "lowk_inst.cuh"
