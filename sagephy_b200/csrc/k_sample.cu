#include "kernels_sample.cuh"
#include "launch.hpp"
namespace sgp {
void launch_find_general(bool replay, const FindArgs& a, int blocks, cudaStream_t st) { if (replay) k_find_general<true><<<blocks, 128, 0, st>>>(a); else k_find_general<false><<<blocks, 128, 0, st>>>(a); }
void launch_build_general(bool replay, const BuildArgs& a, int blocks, cudaStream_t st) { if (replay) k_build_general<true><<<blocks, 128, 0, st>>>(a); else k_build_general<false><<<blocks, 128, 0, st>>>(a); }
}  // namespace sgp
