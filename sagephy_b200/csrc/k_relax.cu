#include "kernels_relax.cuh"
#include "launch.hpp"
#include <algorithm>
namespace sgp {
void launch_relax_rates(bool replay, const RelaxArgs& a, int first_attempt, int only_flagged, cudaStream_t st) {
    const int blocks = (int)((a.n + 127) / 128);
    if (replay) k_relax_rates<true><<<blocks, 128, 0, st>>>(a, first_attempt, only_flagged); else k_relax_rates<false><<<blocks, 128, 0, st>>>(a, first_attempt, only_flagged);
}
void launch_relax_apply(const RelaxArgs& a, long long total_slots, cudaStream_t st) {
    k_relax_apply<<<(unsigned)((total_slots + 255) / 256), 256, 0, st>>>(a, total_slots);
    k_relax_finish_apply<<<(unsigned)((a.n + 255) / 256), 256, 0, st>>>(a);
}
void launch_relax_apply_per_rep(const RelaxArgs& a, cudaStream_t st) {
    int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long blocks = std::max<long long>(1, std::min<long long>((a.n + 7) / 8, (long long)sms * 8));
    k_relax_apply_per_rep<<<(unsigned)blocks, 256, 0, st>>>(a);
    k_relax_finish_apply<<<(unsigned)((a.n + 255) / 256), 256, 0, st>>>(a);
}
void launch_ingest_count(const IngestArgs& g, cudaStream_t st) { k_ingest_count<<<(unsigned)((g.n_src + 256) / 256), 256, 0, st>>>(g); }
void launch_ingest_fill(const IngestArgs& g, cudaStream_t st) { k_ingest_fill<<<(unsigned)((g.n_src * 32 + 255) / 256), 256, 0, st>>>(g); }
void launch_relax_emit(bool write, const RelaxArgs& a, cudaStream_t st) {
    const unsigned blocks = (unsigned)((a.n * 32 + kEmitWarps * 32 - 1) / (kEmitWarps * 32));
    if (write) k_relax_emit<true><<<blocks, kEmitWarps * 32, 0, st>>>(a); else k_relax_emit<false><<<blocks, kEmitWarps * 32, 0, st>>>(a);
}
}  // namespace sgp
