// lowk_inst_5.cu -- instantiates the low-k search kernels for k = 11..12 (split for parallel builds)
#define LOWK_K_LO 11
#define LOWK_K_HI 12
#define LOWK_RANGE_FN lowk_launch_range_5
#include "lowk_inst.cuh"
