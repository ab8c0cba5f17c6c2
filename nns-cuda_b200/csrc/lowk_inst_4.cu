// lowk_inst_4.cu -- instantiates the low-k search kernels for k = 9..10 (split for parallel builds)
#define LOWK_K_LO 9
#define LOWK_K_HI 10
#define LOWK_RANGE_FN lowk_launch_range_4
#include "lowk_inst.cuh"
