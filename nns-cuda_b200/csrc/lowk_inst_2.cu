// lowk_inst_2.cu -- instantiates the low-k search kernels for k = 9..12 (split for parallel builds)
#define LOWK_K_LO 9
#define LOWK_K_HI 12
#define LOWK_RANGE_FN lowk_launch_range_2
#include "lowk_inst.cuh"
