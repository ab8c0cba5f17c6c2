// lowk_inst_2.cu -- instantiates the low-k search kernels for k = 5..6 (split for parallel builds)
#define LOWK_K_LO 5
#define LOWK_K_HI 6
#define LOWK_RANGE_FN lowk_launch_range_2
#include "lowk_inst.cuh"
