#include "kernels_emit.cuh"
#include "launch.hpp"
#include <cstdlib>
namespace sgp {
// size pass: finishes the records the label kernel left (tree and row streams) and measures the .info files
int launch_emit_sizes(const EmitArgs& a, cudaStream_t st) {
    int n = 0;
    if (a.stream_mask & 0xF3u) { k_finish_records<0><<<(unsigned)((a.n * 32 + 255) / 256), 256, 0, st>>>(a); ++n; }
    if (a.stream_mask & 0xCu) { k_emit_info<false><<<(unsigned)((2 * a.n + kInfoWarps * 32 - 1) / (kInfoWarps * 32)), kInfoWarps * 32, 0, st>>>(a); ++n; }
    return n;
}
// The streams of a batch are written by independent kernels: with side streams they run next to each other (fork after `st`, join
// back into `st`), so that one kernel's tail (its last, partly filled wave of blocks) is covered by the next kernel's blocks.
template <bool BULK> static int launch_write(const EmitArgs& a, cudaStream_t st, cudaStream_t* side, int n_side, cudaEvent_t fork, cudaEvent_t* join) {
    const unsigned blocks = (unsigned)((a.n * 32 + kEmitWarps * 32 - 1) / (kEmitWarps * 32)), threads = kEmitWarps * 32;
    const size_t smem = (size_t)kEmitWarps * 2 * a.stage_bytes;
    int n = 0;
    if (n_side > 0) cudaEventRecord(fork, st);
    auto on = [&](int k) { if (n_side <= 0 || k == 0) return st; cudaStream_t s = side[(k - 1) % n_side]; cudaStreamWaitEvent(s, fork, 0); return s; };
    if (a.stream_mask & (1u << ST_UNPRUNED_TREE)) { k_write_tree<0, BULK><<<blocks, threads, smem, on(0)>>>(a); ++n; }
    if (a.stream_mask & (1u << ST_PRUNED_TREE)) { k_write_tree<1, BULK><<<blocks, threads, smem, on(1)>>>(a); ++n; }
    // both .info files in one launch: a thread per (replicate, file)
    if (a.stream_mask & 0xCu) { k_emit_info<true><<<(unsigned)((2 * a.n + kInfoWarps * 32 - 1) / (kInfoWarps * 32)), kInfoWarps * 32, 0, on(2)>>>(a); ++n; }
    if (a.stream_mask & (1u << ST_UNPRUNED_G2H)) { k_write_rows<0, 0, BULK><<<blocks, threads, smem, on(1)>>>(a); ++n; }
    if (a.stream_mask & (1u << ST_PRUNED_G2H)) { k_write_rows<0, 1, BULK><<<blocks, threads, smem, on(2)>>>(a); ++n; }
    if (a.stream_mask & (1u << ST_UNPRUNED_LEAFMAP)) { k_write_rows<1, 0, BULK><<<blocks, threads, smem, on(1)>>>(a); ++n; }
    if (a.stream_mask & (1u << ST_PRUNED_LEAFMAP)) { k_write_rows<1, 1, BULK><<<blocks, threads, smem, on(2)>>>(a); ++n; }
    for (int k = 0; k < n_side; ++k) { cudaEventRecord(join[k], side[k]); cudaStreamWaitEvent(st, join[k], 0); }
    return n;
}
int launch_emit_write(const EmitArgs& a, cudaStream_t st, cudaStream_t* side, int n_side, cudaEvent_t fork, cudaEvent_t* join) {
    // SGP_EMIT_FLUSH=lsu: plain 16-byte stores instead of the bulk (TMA) copy, for A/B measurements
    static const bool lsu = [] { const char* e = std::getenv("SGP_EMIT_FLUSH"); return e && e[0] == 'l'; }();
    static const bool serial = std::getenv("SGP_EMIT_SERIAL") != nullptr;
    if (serial) n_side = 0;
    return lsu ? launch_write<false>(a, st, side, n_side, fork, join) : launch_write<true>(a, st, side, n_side, fork, join);
}
}  // namespace sgp
