#include "kernels_emit.cuh"
#include "launch.hpp"
namespace sgp {
template <bool W, int S> static void launch_one(const EmitArgs& a, unsigned blocks, cudaStream_t st) { k_emit<W, S><<<blocks, kEmitWarps * 32, 0, st>>>(a); }
template <bool W> static int launch_all(const EmitArgs& a, unsigned blocks, cudaStream_t st) {
    int n = 0;
    if (a.stream_mask & (1u << 0)) { launch_one<W, 0>(a, blocks, st); ++n; }
    if (a.stream_mask & (1u << 1)) { launch_one<W, 1>(a, blocks, st); ++n; }
    if (a.stream_mask & (1u << 2)) { launch_one<W, 2>(a, blocks, st); ++n; }
    if (a.stream_mask & (1u << 3)) { launch_one<W, 3>(a, blocks, st); ++n; }
    if (a.stream_mask & (1u << 4)) { launch_one<W, 4>(a, blocks, st); ++n; }
    if (a.stream_mask & (1u << 5)) { launch_one<W, 5>(a, blocks, st); ++n; }
    if (a.stream_mask & (1u << 6)) { launch_one<W, 6>(a, blocks, st); ++n; }
    if (a.stream_mask & (1u << 7)) { launch_one<W, 7>(a, blocks, st); ++n; }
    return n;
}
int launch_emit(bool write, const EmitArgs& a, unsigned blocks, cudaStream_t st) { return write ? launch_all<true>(a, blocks, st) : launch_all<false>(a, blocks, st); }
}  // namespace sgp
