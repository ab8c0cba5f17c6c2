// lowk_inst_15.cu -- instantiates the low-k search kernels for k = 31..32 (split for parallel builds)
#define LOWK_K_LO 31
#define LOWK_K_HI 32
#define LOWK_RANGE_FN lowk_launch_range_15
#include "lowk_inst.cuh"
