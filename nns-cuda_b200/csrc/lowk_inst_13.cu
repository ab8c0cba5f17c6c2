// lowk_inst_13.cu -- instantiates the low-k search kernels for k = 27..28 (split for parallel builds)
#define LOWK_K_LO 27
#define LOWK_K_HI 28
#define LOWK_RANGE_FN lowk_launch_range_13
#include "lowk_inst.cuh"
