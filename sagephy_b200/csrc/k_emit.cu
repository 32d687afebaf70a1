#include "kernels_emit.cuh"
#include "launch.hpp"
namespace sgp {
template <bool W, int S> static void launch_one(const EmitArgs& a, unsigned blocks, cudaStream_t st) { k_emit<W, S><<<blocks, kEmitWarps * 32, 0, st>>>(a); }
template <bool W> static int launch_all(const EmitArgs& a, unsigned blocks, cudaStream_t st) {
    int n = 0;
    if (a.stream_mask & (1u << 0)) { launch_one<W, 0>(a, blocks, st); ++n; }
    if (a.stream_mask & (1u << 1)) { launch_one<W, 1>(a, blocks, st); ++n; }
    // both .info files in one launch: a thread per (replicate, file)
    if (a.stream_mask & 0xCu) { k_emit_info<W><<<(unsigned)((2 * a.n + kInfoWarps * 32 - 1) / (kInfoWarps * 32)), kInfoWarps * 32, 0, st>>>(a); ++n; }
    if (a.stream_mask & (1u << 4)) { launch_one<W, 4>(a, blocks, st); ++n; }
    if (a.stream_mask & (1u << 5)) { launch_one<W, 5>(a, blocks, st); ++n; }
    if (a.stream_mask & (1u << 6)) { launch_one<W, 6>(a, blocks, st); ++n; }
    if (a.stream_mask & (1u << 7)) { launch_one<W, 7>(a, blocks, st); ++n; }
    return n;
}
int launch_emit(bool write, const EmitArgs& a, unsigned blocks, cudaStream_t st) { return write ? launch_all<true>(a, blocks, st) : launch_all<false>(a, blocks, st); }
}  // namespace sgp
