// lowk_inst_6.cu -- instantiates the low-k search kernels for k = 13..14 (split for parallel builds)
#define LOWK_K_LO 13
#define LOWK_K_HI 14
#define LOWK_RANGE_FN lowk_launch_range_6
#include "lowk_inst.cuh"
