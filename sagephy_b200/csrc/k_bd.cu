#include "kernels_bd.cuh"
#include "launch.hpp"
#include <cstdlib>
namespace sgp {
// `blocks` is sized for 4 resident blocks per SM (persistent kernel: the grid is the residency); blocks_per_sm scales it.
void launch_bd_find(const BdArgs& a, int blocks, int blocks_per_sm, cudaStream_t st) {
    static const int variant = [] { const char* e = std::getenv("SGP_FIND_VARIANT"); return e ? std::atoi(e) : 0; }();
    auto scaled = [&](int per_sm) { const int b = (int)((long long)blocks * per_sm / 4); return b > 0 ? b : 1; };
    if (!(a.H.nV == 1 && a.H.vtime[0] == 0.0)) { k_bd_find<<<scaled(blocks_per_sm), kBdFindThreads, 0, st>>>(a); return; }
    switch (variant) {
        default: k_bd_find_single<false, 1, true, 4><<<scaled(blocks_per_sm), kBdFindThreads, 0, st>>>(a); break;      // local-memory stack, one table copy, control words via PTX
        case 1: k_bd_find_single<false, 4, false, 4><<<scaled(blocks_per_sm), kBdFindThreads, 0, st>>>(a); break;
        case 2: k_bd_find_single<true, 1, false, 4><<<scaled(blocks_per_sm), kBdFindThreads, 0, st>>>(a); break;
        case 3: k_bd_find_single<false, 1, false, 4><<<scaled(blocks_per_sm), kBdFindThreads, 0, st>>>(a); break;
    }
}
void launch_bd_build(const BdArgs& a, int blocks, cudaStream_t st) { k_bd_build<<<blocks, kBdBuildThreads, 0, st>>>(a); }
}  // namespace sgp
