#include "kernels_label.cuh"
#include "launch.hpp"
#include "host_model.hpp"
#include <algorithm>
#include <cstdlib>
namespace sgp {
// One launch per size class so that the shared-memory stage of a warp is never much larger than its tree (the stage size sets
// how many trees an SM works on at a time); whatever exceeds the largest stage goes to the thread-per-tree kernel.
int launch_label_prune(LabelArgs a, int max_pool, int sm_count, cudaStream_t st) {
    static const int caps[] = {48, 64, 96, 128, 160, 192, 224, 256, 288, 320, 384, 448, 512, 640, 768, 1024, 1280, 1536, 2048, 2340};
    const int kMaxSmem = 227 * 1024;
    // the opt-in is per device (and per context of the calling thread): set it on every call, it costs nothing next to a launch
    if (cudaFuncSetAttribute(k_label_prune_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 1024) != cudaSuccess)
        throw SgpError("cannot opt in to large dynamic shared memory for the label kernel on this device");
    int launches = 0, lo = 0;
    const bool thread_only = getenv("SGP_LABEL_THREAD_KERNEL") != nullptr;      // test hook: exercise the fallback on small trees
    const bool global_only = getenv("SGP_LABEL_GLOBAL_KERNEL") != nullptr;      // test hook: every tree through the in-place warp kernel
    for (int cap : caps) {
        if (thread_only || global_only) break;
        a.min_pool = lo; a.cap = cap;
        const size_t smem = label_smem_bytes(cap);
        if (smem + 2048 > (size_t)kMaxSmem) break;
        const int per_sm = (int)std::min<size_t>(32, (size_t)kMaxSmem / (smem + 1536));     // + static shared memory and the per-block reservation
        const long long want = (a.n + 31) / 32;
        const int blocks = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count * per_sm));
        k_label_prune_warp<<<blocks, 32, smem, st>>>(a); ++launches;
        lo = cap;
        if (cap >= max_pool) break;
    }
    // trees above the largest shared-memory class: the warp algorithm in place, work arrays in the scratch the engine passed
    int warp_max = lo;
    if (max_pool > lo && !thread_only && a.work && a.work_blocks > 0 && a.work_stride >= 9ll * (max_pool + 2)) {
        LabelArgs g = a; g.min_pool = lo; g.cap = max_pool;
        k_label_prune_warp_global<<<a.work_blocks, 32, 0, st>>>(g); ++launches;
        warp_max = max_pool;
    }
    if (warp_max > 0) { LabelArgs b = a; b.cap = warp_max; k_info_sums<<<(unsigned)((4 * a.n + 127) / 128), 128, 0, st>>>(b); ++launches; }
    if (max_pool > warp_max) {
        if (!a.aux || !a.write_nodes) throw SgpError("internal: oversized trees need the order arrays");
        a.cap = warp_max; k_label_prune<<<(unsigned)((a.n + 127) / 128), 128, 0, st>>>(a); ++launches;
        k_records_from_nodes<<<(unsigned)((a.n * 32 + 127) / 128), 128, 0, st>>>(a); ++launches;
    }
    return launches;
}
int label_smem_max_pool() { return (getenv("SGP_LABEL_THREAD_KERNEL") || getenv("SGP_LABEL_GLOBAL_KERNEL")) ? 0 : 2340; }
}  // namespace sgp
