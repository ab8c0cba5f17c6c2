// lowk_inst.cuh -- instantiation helper: each lowk_inst_N.cu defines LOWK_K_LO / LOWK_K_HI and
// includes this file to emit lowk_launch_range_N covering k in [LO, HI].
#include "lowk_search.cuh"

namespace nns {

template <int K>
static cudaError_t lowk_launch_k(int q, int mode, const LowkArgs& a, int* occ)
{
    constexpr int QD = lowk_q_default(K), QA = lowk_q_alt(K);
    if (q == QD) return lowk_launch_t<K, QD>(mode, a, occ);
    if (q == QA && mode != LOWK_EXACT_V0) {
        if (mode == LOWK_FILTER)
            return lowk_launch_kernel(lowk_filter_kernel<K, QA, lowk_minb(K, QA), lowk_unroll(K), lowk_screen_quads(K)>, K, a, occ, true);
        return lowk_launch_kernel(lowk_exact_kernel<K, QA, false, lowk_minb(K, QA), lowk_unroll(K)>, K, a, occ, false);
    }
    return cudaErrorInvalidValue;
}

template <int K, int KHI>
struct LowkRange {
    static cudaError_t go(int k, int q, int mode, const LowkArgs& a, int* occ)
    {
        if (k == K) return lowk_launch_k<K>(q, mode, a, occ);
        if constexpr (K < KHI) return LowkRange<K + 1, KHI>::go(k, q, mode, a, occ);
        return cudaErrorInvalidValue;
    }
};

cudaError_t LOWK_RANGE_FN(int k, int q, int mode, const LowkArgs& a, int* occ)
{
    return LowkRange<LOWK_K_LO, LOWK_K_HI>::go(k, q, mode, a, occ);
}

}  // namespace nns
