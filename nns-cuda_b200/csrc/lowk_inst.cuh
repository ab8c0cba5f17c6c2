// lowk_inst.cuh -- instantiation helper: each lowk_inst_N.cu defines LOWK_K_LO / LOWK_K_HI and
// includes this file to emit lowk_launch_range_N covering k in [LO, HI].
#include "lowk_search.cuh"

namespace nns {

template <int K>
static cudaError_t lowk_launch_k(int q, bool exact, const LowkArgs& a, int* occ)
{
    constexpr int QD = lowk_q_default(K), QA = lowk_q_alt(K);
    if (q == QD) return exact ? lowk_launch_t<K, QD, true>(a, occ) : lowk_launch_t<K, QD, false>(a, occ);
    if (q == QA && !exact) return lowk_launch_t<K, QA, false>(a, occ);
    return cudaErrorInvalidValue;
}

template <int K, int KHI>
struct LowkRange {
    static cudaError_t go(int k, int q, bool exact, const LowkArgs& a, int* occ)
    {
        if (k == K) return lowk_launch_k<K>(q, exact, a, occ);
        if constexpr (K < KHI) return LowkRange<K + 1, KHI>::go(k, q, exact, a, occ);
        return cudaErrorInvalidValue;
    }
};

cudaError_t LOWK_RANGE_FN(int k, int q, bool exact, const LowkArgs& a, int* occ)
{
    return LowkRange<LOWK_K_LO, LOWK_K_HI>::go(k, q, exact, a, occ);
}

}  // namespace nns
