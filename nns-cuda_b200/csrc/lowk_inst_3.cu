// lowk_inst_3.cu -- instantiates the low-k search kernels for k = 13..16 (split for parallel builds)
#define LOWK_K_LO 13
#define LOWK_K_HI 16
#define LOWK_RANGE_FN lowk_launch_range_3
#include "lowk_inst.cuh"
