#include "kernels_emit.cuh"
#include "launch.hpp"
namespace sgp {
void launch_emit(bool write, const EmitArgs& a, unsigned blocks, cudaStream_t st) { if (write) k_emit<true><<<blocks, 128, 0, st>>>(a); else k_emit<false><<<blocks, 128, 0, st>>>(a); }
}  // namespace sgp
