// lowk_inst_4.cu -- instantiates the low-k search kernels for k = 17..20 (split for parallel builds)
#define LOWK_K_LO 17
#define LOWK_K_HI 20
#define LOWK_RANGE_FN lowk_launch_range_4
#include "lowk_inst.cuh"
