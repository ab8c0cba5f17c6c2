// lowk_inst_7.cu -- instantiates the low-k search kernels for k = 15..16 (split for parallel builds)
#define LOWK_K_LO 15
#define LOWK_K_HI 16
#define LOWK_RANGE_FN lowk_launch_range_7
#include "lowk_inst.cuh"
