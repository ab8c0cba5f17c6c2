// nns_b200.hpp -- the binding a maintainer of sty-hhh/NNS-CUDA adds to use libnns_b200.so:
// one more namespace with the reference's callback signature (core.cu:23-29), assignable to the
// driver's function pointer `void (*func)(int, int, int, float *, float *, int **)` (main.cu:7)
// exactly like &v9::cudaCall (main.cu:116-117).  See INTEGRATION.md.
#pragma once
#include "nns_b200.h"

namespace b200 {
inline void cudaCall(int k, int m, int n, float *s_points, float *r_points, int **results)
{
    nns_b200_cudaCall(k, m, n, s_points, r_points, results);
}
}  // namespace b200
