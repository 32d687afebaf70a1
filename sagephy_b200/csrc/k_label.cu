#include "kernels_label.cuh"
#include "launch.hpp"
namespace sgp {
void launch_label_prune(const LabelArgs& a, int blocks, cudaStream_t st) { k_label_prune<<<blocks, 128, 0, st>>>(a); }
}  // namespace sgp
