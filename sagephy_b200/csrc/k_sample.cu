#include "kernels_sample.cuh"
#include "launch.hpp"
namespace sgp {
void launch_find_general(bool replay, bool domain, const FindArgs& a, int blocks, cudaStream_t st) {
    if (domain) { if (replay) k_find_general<true, true><<<blocks, 128, 0, st>>>(a); else k_find_general<false, true><<<blocks, 128, 0, st>>>(a); }
    else { if (replay) k_find_general<true, false><<<blocks, 128, 0, st>>>(a); else k_find_general<false, false><<<blocks, 128, 0, st>>>(a); }
}
void launch_build_general(bool replay, bool domain, const BuildArgs& a, int blocks, cudaStream_t st) {
    if (domain) { if (replay) k_build_general<true, true><<<blocks, 128, 0, st>>>(a); else k_build_general<false, true><<<blocks, 128, 0, st>>>(a); }
    else { if (replay) k_build_general<true, false><<<blocks, 128, 0, st>>>(a); else k_build_general<false, false><<<blocks, 128, 0, st>>>(a); }
}
}  // namespace sgp
