// lowk_inst_9.cu -- instantiates the low-k search kernels for k = 19..20 (split for parallel builds)
#define LOWK_K_LO 19
#define LOWK_K_HI 20
#define LOWK_RANGE_FN lowk_launch_range_9
#include "lowk_inst.cuh"
