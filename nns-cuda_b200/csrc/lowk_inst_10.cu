// lowk_inst_10.cu -- instantiates the low-k search kernels for k = 21..22 (split for parallel builds)
#define LOWK_K_LO 21
#define LOWK_K_HI 22
#define LOWK_RANGE_FN lowk_launch_range_10
#include "lowk_inst.cuh"
