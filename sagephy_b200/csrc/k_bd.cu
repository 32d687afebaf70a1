#include "kernels_bd.cuh"
#include "launch.hpp"
namespace sgp {
void launch_bd_find(const BdArgs& a, int blocks, cudaStream_t st) {
    if (a.H.nV == 1 && a.H.vtime[0] == 0.0) k_bd_find_single<<<blocks, kBdFindThreads, 0, st>>>(a); else k_bd_find<<<blocks, kBdFindThreads, 0, st>>>(a);
}
void launch_bd_build(const BdArgs& a, int blocks, cudaStream_t st) { k_bd_build<<<blocks, 256, 0, st>>>(a); }
}  // namespace sgp
