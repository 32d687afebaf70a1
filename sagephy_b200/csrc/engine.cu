// Host orchestration of the tree-generator pipeline on one GPU:
//   K0 tables -> K1 find (accepted attempt per replicate) -> scan -> K1 build (node arena) -> K3 label/prune
//   -> K4 emit size -> scan -> K4 emit write -> summary reduction.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_partition.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <cstdio>
#include <cstring>
#include "engine.hpp"
#include "hostfmt.hpp"
#include "launch.hpp"
#include "rng.cuh"
#include "jmath_ext.cuh"
#include "sim.cuh"

namespace sgp {

const unsigned long long* d2s_g_table(); int d2s_g_table_len(); const double* p10_table(); int p10_table_len();

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw SgpError(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call); } while (0)

namespace {

// Small device buffers come from the CUDA stream-ordered pool (release threshold = unlimited, see Context::Context), large ones from
// a per-device free list of cudaMalloc blocks (BlockCache below): steady-state batches make no driver allocation call.
thread_local cudaStream_t g_alloc_stream = nullptr;
// SGP_TRACE=1: host time stamps between the phases of a run (where does the host wait?), printed to stderr at the end of the run
struct HostMarks {
    bool on = getenv("SGP_TRACE") != nullptr; std::vector<std::pair<const char*, double>> v;
    void mark(const char* name) { if (on) v.emplace_back(name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count()); }
    void dump() { if (!on || v.empty()) return; std::string o = "sgp trace (host ms):"; for (size_t i = 1; i < v.size(); ++i) { char b[96]; snprintf(b, sizeof b, " %s %.2f", v[i].first, v[i].second - v[i - 1].second); o += b; } fprintf(stderr, "%s\n", o.c_str()); v.clear(); }
};
thread_local HostMarks g_marks;
thread_local unsigned long long g_h2d_bytes = 0;      // host->device bytes of the current run (tables, constant strings, templates)
// Size classes of the large buffers. The sizes of a run's vertex arena, records and output streams differ by a fraction of a percent
// from one batch to the next; rounded up to one of 64 classes per octave (< 1.6 % over-allocation) consecutive batches ask for the
// same sizes and find the previous batch's blocks in the free list.
inline size_t pool_size_class(size_t bytes) {
    if (bytes < (size_t)(1u << 20)) return bytes;
    size_t g = 1; while ((g << 6) < bytes) g <<= 1;          // bytes / 128 < g <= bytes / 64
    return (bytes + g - 1) / g * g;
}
// Large device buffers (vertex arena, records, output streams: everything from kBigBytes up) do not go through the stream-ordered
// pool at all. The pool serves a request that fits none of its free blocks by re-mapping physical memory, and that waits for ALL
// work in flight on the device - a caller's device-to-host copies of the previous batch included (measured: 3-100 ms per
// cudaMallocAsync inside the end-to-end pipeline, 0.01 ms without copies in flight); which
// requests fit depends on the history of the pool, and the stalls came and went with the order of the workloads. These buffers
// are plain cudaMalloc blocks kept in a per-device free list instead: a request takes the smallest free block of at least its
// size class (and at most 1.25 x), a free puts the block back - no driver call, no mapping, nothing to wait for. Reuse is safe in
// stream order: a block handed back while work may still be using it carries an event recorded on the stream that used it, and a
// taker on another stream (a second context on the same device) waits for that event first; blocks that were visible to the caller
// (a batch's outputs, which it may still be copying) are only put back after the device has gone idle, as cudaFree would have
// waited for. When the device runs out of memory every unused block (and the pool's reserve) goes back to the driver and the
// request is retried.
constexpr size_t kBigBytes = (size_t)32 << 20;
struct BlockCache {
    struct Block { void* p; size_t bytes; bool used; cudaEvent_t ev; cudaStream_t last; bool pending; };
    std::mutex mu; std::vector<Block> blocks;
    static BlockCache& of(int device) { static BlockCache caches[64]; return caches[device & 63]; }
    void* take(size_t bytes, cudaStream_t on) {
        const size_t want = pool_size_class(bytes);
        {
            std::lock_guard<std::mutex> g(mu);
            Block* best = nullptr;
            for (Block& b : blocks) if (!b.used && b.bytes >= want && b.bytes <= want + want / 4 && (!best || b.bytes < best->bytes)) best = &b;
            if (best) {
                best->used = true;
                if (best->pending && best->last != on) cudaStreamWaitEvent(on, best->ev, 0);
                best->pending = false;
                return best->p;
            }
        }
        void* q = nullptr;
        const bool trace = getenv("SGP_TRACE") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        cudaError_t e = cudaMalloc(&q, want);
        const bool oom = e == cudaErrorMemoryAllocation;
        if (oom) { cudaGetLastError(); drop_unused(); e = cudaMalloc(&q, want); }
        if (trace) fprintf(stderr, "sgp trace: new block of %zu MiB%s, %.2f ms\n", want >> 20, oom ? " after returning every unused block to the driver" : "", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        if (e != cudaSuccess) { cudaGetLastError(); throw SgpError(std::string("out of device memory (") + std::to_string(want >> 20) + " MiB requested): " + cudaGetErrorString(e)); }
        std::lock_guard<std::mutex> g(mu);
        blocks.push_back({q, want, true, nullptr, nullptr, false});
        return q;
    }
    // true if p was one of ours. `idle`: the device is known to have finished with the block; otherwise work queued on `on` may still use it
    bool give_back(void* p, cudaStream_t on, bool idle) {
        std::lock_guard<std::mutex> g(mu);
        for (Block& b : blocks) if (b.p == p) {
            b.used = false; b.pending = false;
            if (!idle) {
                if (!b.ev) cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming);
                if (b.ev && cudaEventRecord(b.ev, on) == cudaSuccess) { b.pending = true; b.last = on; } else { cudaGetLastError(); cudaDeviceSynchronize(); }
            }
            return true;
        }
        return false;
    }
    void drop_unused() {
        cudaDeviceSynchronize();
        {
            std::lock_guard<std::mutex> g(mu);
            for (size_t i = 0; i < blocks.size();) { if (!blocks[i].used) { cudaFree(blocks[i].p); if (blocks[i].ev) cudaEventDestroy(blocks[i].ev); blocks[i] = blocks.back(); blocks.pop_back(); } else ++i; }
        }
        cudaMemPool_t dflt; int dev = 0; cudaGetDevice(&dev);
        if (cudaDeviceGetDefaultMemPool(&dflt, dev) == cudaSuccess) cudaMemPoolTrimTo(dflt, 0);
    }
};
thread_local int g_device = 0;
inline cudaError_t malloc_async_retry(void** p, size_t bytes, cudaStream_t on) {
    cudaError_t e = cudaMallocAsync(p, bytes, on);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); BlockCache::of(g_device).drop_unused(); e = cudaMallocAsync(p, bytes, on); }
    return e;
}
template <class T> struct DevBuf {
    T* p = nullptr; size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { drop(); }
    void drop() { if (p && !BlockCache::of(g_device).give_back(p, g_alloc_stream, false)) cudaFreeAsync(p, g_alloc_stream); p = nullptr; }
    void alloc(size_t count) {
        drop(); n = count;
        if (!count) return;
        if (count * sizeof(T) >= kBigBytes) p = (T*)BlockCache::of(g_device).take(count * sizeof(T), g_alloc_stream);
        else CK(malloc_async_retry((void**)&p, count * sizeof(T), g_alloc_stream));
    }
    void upload(const std::vector<T>& v) { alloc(v.size()); if (!v.empty()) { CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, g_alloc_stream)); CK(cudaStreamSynchronize(g_alloc_stream)); g_h2d_bytes += v.size() * sizeof(T); } }
    T* release() { T* q = p; p = nullptr; n = 0; return q; }
};

struct HostOnDevice {
    DevBuf<int32_t> parent, left, right, host_tag, v2e, ep_off, ep_arcs, leaf_rank;
    DevBuf<double> vtime, ep_lo, ep_mid, ep_up, lca_time, ep_times;
    DevHost view{};
    void upload(const HostModel& h) {
        parent.upload(h.parent); left.upload(h.left); right.upload(h.right); host_tag.upload(h.host_tag); v2e.upload(h.v2e);
        ep_off.upload(h.ep_off); ep_arcs.upload(h.ep_arcs);
        std::vector<int32_t> lr(h.nV, -1); for (int i = 0; i < h.nLeaves; ++i) lr[h.leaves[i]] = i;
        leaf_rank.upload(lr);
        vtime.upload(h.vtime); ep_lo.upload(h.ep_lo); ep_mid.upload(h.ep_mid); ep_up.upload(h.ep_up); lca_time.upload(h.lca_time);
        std::vector<double> et(3 * (size_t)h.nE); for (int e = 0; e < h.nE; ++e) { et[3 * e] = h.ep_lo[e]; et[3 * e + 1] = h.ep_mid[e]; et[3 * e + 2] = h.ep_up[e]; }
        ep_times.upload(et);
        view.nV = h.nV; view.nE = h.nE; view.root = h.root; view.nLeaves = h.nLeaves; view.max_arcs = h.max_arcs; view.has_lca = !h.lca_time.empty();
        view.ep_sorted = 1; for (int e = 1; e < h.nE; ++e) if (!(h.ep_up[e] >= h.ep_up[e - 1])) view.ep_sorted = 0;
        view.parent = parent.p; view.left = left.p; view.right = right.p; view.host_tag = host_tag.p; view.v2e = v2e.p; view.ep_off = ep_off.p;
        view.ep_arcs = ep_arcs.p; view.leaf_rank = leaf_rank.p; view.vtime = vtime.p; view.ep_lo = ep_lo.p; view.ep_mid = ep_mid.p; view.ep_up = ep_up.p;
        view.lca_time = lca_time.p;
    }
};

__global__ void k_extract_pool(const RepInfo* info, long long n, long long* out, int* max_pool) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int v = 0;
    if (i < n) { v = info[i].status == REP_OK ? info[i].n_pool : 0; out[i] = v; }
    v = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(max_pool, v);
}
__global__ void k_collect_status(const RepInfo* info, long long n, int status, long long* list, unsigned long long* count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && info[i].status == status) { unsigned long long k = atomicAdd(count, 1ULL); list[k] = i; }
}

// build order: replicates whose tree is much larger than average go first (cub::DevicePartition keeps them in front)
struct PoolAbove { const long long* pool; long long thr; __device__ bool operator()(long long i) const { return pool[i] > thr; } };

struct DevSummary { unsigned long long n_ok, n_failed, attempts_total, vertices, vertices_abandoned, attempts_completed; unsigned long long ev[17]; unsigned long long leaf_hist[1025]; };
// Block-level reduction in shared memory first (a few hundred global atomics per block instead of ~8 per replicate on the same
// handful of addresses, which cost 5 ms per 10^6 replicates).
__global__ void __launch_bounds__(256) k_summary(const RepInfo* info, long long n, DevSummary* s) {
    __shared__ unsigned long long sh[sizeof(DevSummary) / sizeof(unsigned long long)];
    constexpr int kWords = (int)(sizeof(DevSummary) / sizeof(unsigned long long));
    for (int i = threadIdx.x; i < kWords; i += blockDim.x) sh[i] = 0ull;
    __syncthreads();
    DevSummary* b = reinterpret_cast<DevSummary*>(sh);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const RepInfo& x = info[i];
        atomicAdd(&b->attempts_total, (unsigned long long)x.attempts);
        atomicAdd(&b->vertices, (unsigned long long)x.vertices_sampled);
        atomicAdd(&b->vertices_abandoned, (unsigned long long)x.vertices_abandoned);
        atomicAdd(&b->attempts_completed, (unsigned long long)x.attempts_completed);
        if (x.status == REP_OK) {
            atomicAdd(&b->n_ok, 1ULL);
            for (int e = 0; e < 17; ++e) if (x.cnt_u[e]) atomicAdd(&b->ev[e], (unsigned long long)x.cnt_u[e]);
            int bin = x.sampled_leaves; if (bin > 1024) bin = 1024; if (bin < 0) bin = 0;
            atomicAdd(&b->leaf_hist[bin], 1ULL);
        } else atomicAdd(&b->n_failed, 1ULL);
    }
    __syncthreads();
    unsigned long long* g = reinterpret_cast<unsigned long long*>(s);
    for (int i = threadIdx.x; i < kWords; i += blockDim.x) if (sh[i]) atomicAdd(&g[i], sh[i]);
}

// probes
__global__ void k_probe_d2s(const double* v, long long n, char* out, MathTables T) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    char* o = out + i * 32; for (int k = 0; k < 32; ++k) o[k] = 0;
    MemSink m{o}; put_double(m, T, v[i]);
}
__global__ void k_probe_math(int fn, const double* v, long long n, double* out, MathTables T) {
    __shared__ double2 s_logtab[128];
    load_log_table(s_logtab);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = v[i];
    int err = 0;
    out[i] = fn == 0 ? jlog(x) : fn == 1 ? jlog10(x) : fn == 2 ? jexp(x) : fn == 3 ? round_sig(T, x, 8) : fn == 4 ? jnormal_quantile(x)
           : fn == 5 ? jpow(x, 0.37) : fn == 6 ? jchi2_quantile(x, 4.0, &err) : fn == 7 ? jchi2_quantile(x, 0.25, &err) : fn == 8 ? jpow(2.718281828459045, x) : log_unit(x, s_logtab);
}
__global__ void k_probe_mt(SeedSpec seed, long long offset, long long n, uint32_t* out, uint32_t* state) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    MtStream mt; mt.mt = state; mt.stride = 32;
    uint32_t key[4]; seed_words(seed, (unsigned long long)offset, key); mt.seed(key);
    for (long long i = 0; i < n; ++i) out[i] = mt.next_word();
}
__global__ void k_probe_philox(int variant, const uint32_t* ctr, const uint32_t* key, long long n, uint32_t* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* c = ctr + 4 * i; const uint32_t* k = key + 2 * i;
    const U4 r = variant == 0 ? philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1])
                              : philox4x32_10_rk(c[0], c[1], c[2], c[3], philox_keys((uint64_t)k[0] | ((uint64_t)k[1] << 32)));
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}
// issue-rate microbenchmarks: independent chains per thread, enough ILP to saturate the pipe
__global__ void k_peak_dfma(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7; const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_peak_imad(uint32_t* out, int iters) {
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7; const uint32_t b = 0xD2511F53u, c = 12345u;
    // a*a+c cannot be re-associated across iterations, so each statement is exactly one IMAD
    for (int i = 0; i < iters; ++i) { a0 = a0 * a0 + c; a1 = a1 * a1 + c; a2 = a2 * a2 + c; a3 = a3 * a3 + c; a4 = a4 * a4 + c; a5 = a5 * a5 + c; a6 = a6 * a6 + c; a7 = a7 * a7 + c; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ b;
}
// mixed integer issue: IMAD (fma pipe) interleaved with LOP3 / shifts (alu pipe), the mix Philox is made of
__global__ void k_peak_mixed(uint32_t* out, int iters) {
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 7 + 1, b1 = b0 + 1, b2 = b0 + 2, b3 = b0 + 3; const uint32_t c = 12345u;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * a0 + c; b0 = (b0 ^ a1) ^ (b0 >> 7); a1 = a1 * a1 + c; b1 = (b1 ^ a2) ^ (b1 >> 7);
        a2 = a2 * a2 + c; b2 = (b2 ^ a3) ^ (b2 >> 7); a3 = a3 * a3 + c; b3 = (b3 ^ a0) ^ (b3 >> 7);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
}

// Pipelined runs (see run_treegen): replicates per sub-batch, head room of the output buffers that are sized from the first
// sub-batch, and the mailbox regions of the two streams.
constexpr long long kSubBatch = 131072;
constexpr double kOutputMargin = 1.03;
constexpr size_t kMailFind = 16384, kMailEmit = 16384 + 512;

// offsets[s][i] = carry[s] + sum of sizes[s][0..i) for i in [0, m], one block per stream; carry[s] is advanced to offsets[s][m], which
// is where the next sub-batch of the run continues (slices of the [stream][stride] arrays of the whole run).
__global__ void __launch_bounds__(1024) k_scan_streams(const unsigned long long* sizes, unsigned long long* offsets, long long stride, long long m, unsigned long long* carry) {
    typedef cub::BlockScan<unsigned long long, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const int s = blockIdx.x;
    const unsigned long long* in = sizes + (size_t)s * stride; unsigned long long* outp = offsets + (size_t)s * stride;
    unsigned long long running = carry[s];
    for (long long base = 0; base < m; base += 1024 * 4) {
        unsigned long long v[4];
        const long long i0 = base + (long long)threadIdx.x * 4;
        for (int q = 0; q < 4; ++q) v[q] = i0 + q < m ? in[i0 + q] : 0ull;
        unsigned long long block_total;
        Scan(tmp).ExclusiveSum(v, v, block_total);
        for (int q = 0; q < 4; ++q) if (i0 + q < m) outp[i0 + q] = running + v[q];
        running += block_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) { outp[m] = running; carry[s] = running; }
}

struct EventTimer {
    cudaEvent_t a, b; cudaStream_t s;
    explicit EventTimer(cudaStream_t st) : s(st) { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~EventTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, s); }
    double stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
};

}  // namespace

struct Context::Impl {
    cudaStream_t stream = nullptr;
    cudaStream_t side[2] = {nullptr, nullptr}; cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};   // side streams of the write kernels
    cudaStream_t stream2 = nullptr;        // second compute stream of pipelined runs (build / label / emit of sub-batch k under find of k+1)
    cudaStream_t copy = nullptr;
    // Small device->host results (counts, totals, the summary block) travel through a mapped pinned mailbox written by a
    // kernel, not through cudaMemcpy: a DMA copy would queue behind a caller's bulk copies that are still in flight on the
    // copy stream (Batch::copy_stream_to_async) and stall the next batch's compute for their whole duration.
    unsigned char* mailbox = nullptr; static constexpr size_t kMailboxBytes = 32768;
    // Large scratch of the general find kernel (up to ~25 GB): kept for the life of the context and only ever grown. Going
    // through the stream-ordered pool for it made run times erratic near HBM capacity (a 15 GB request that does not fit the
    // pool's free blocks costs hundreds of milliseconds of unmap / map).
    void* scratch_buf[8] = {nullptr}; size_t scratch_bytes[8] = {0};
    void* scratch(int slot, size_t bytes);
    // Output streams: block cache for the large ones (see BlockCache), the stream-ordered pool for the rest.
    void alloc_output(char** p, size_t bytes, cudaStream_t on);
    void post(size_t off, const void* dsrc, size_t bytes, cudaStream_t on = nullptr);
    template <class T> T take(size_t off) const { T v; memcpy(&v, mailbox + off, sizeof(T)); return v; }
    DevBuf<unsigned long long> g_table; DevBuf<double> p10;
    // vertices per replicate seen in earlier runs of the same configuration (app + arguments without the seed value): sizes the arena
    // of the one-pass general engine without a pilot run
    std::map<std::string, double> nodes_per_replicate;
    MathTables T{};
    int sm_count = 148;
};

__global__ void k_post_words(uint32_t* dst, const uint32_t* src, int words) {
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
}
void Context::Impl::post(size_t off, const void* dsrc, size_t bytes, cudaStream_t on) {
    if (off + bytes > kMailboxBytes || (bytes & 3) || (off & 3)) throw SgpError("internal: mailbox overflow");
    k_post_words<<<1, 256, 0, on ? on : stream>>>(reinterpret_cast<uint32_t*>(mailbox + off), static_cast<const uint32_t*>(dsrc), (int)(bytes / 4));
}

void Context::Impl::alloc_output(char** p, size_t bytes, cudaStream_t on) {
    const size_t want = std::max<size_t>(bytes, 1);
    if (want >= kBigBytes) *p = (char*)BlockCache::of(g_device).take(want, on);
    else CK(malloc_async_retry((void**)p, want, on));
}

void* Context::Impl::scratch(int slot, size_t bytes) {
    if (bytes > scratch_bytes[slot]) {
        if (scratch_buf[slot]) { CK(cudaStreamSynchronize(stream)); CK(cudaFree(scratch_buf[slot])); scratch_buf[slot] = nullptr; scratch_bytes[slot] = 0; }
        cudaError_t e = cudaMalloc(&scratch_buf[slot], bytes);
        if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); BlockCache::of(g_device).drop_unused(); e = cudaMalloc(&scratch_buf[slot], bytes); }
        CK(e); scratch_bytes[slot] = bytes;
    }
    return scratch_buf[slot];
}

Context::Context(int dev) : device(dev), impl(new Impl()) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) throw SgpError(std::string("no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") + "); sagephy_b200 has no CPU fallback");
    if (dev < 0 || dev >= count) throw SgpError("invalid CUDA device index " + std::to_string(dev));
    CK(cudaSetDevice(dev));
    CK(cudaStreamCreate(&impl->stream));
    CK(cudaStreamCreateWithFlags(&impl->stream2, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) { CK(cudaStreamCreateWithFlags(&impl->side[k], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&impl->ev_join[k], cudaEventDisableTiming)); }
    CK(cudaEventCreateWithFlags(&impl->ev_fork, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&impl->copy, cudaStreamNonBlocking));
    CK(cudaHostAlloc((void**)&impl->mailbox, Impl::kMailboxBytes, cudaHostAllocMapped | cudaHostAllocPortable));
    g_alloc_stream = impl->stream;
    { cudaMemPool_t pool; CK(cudaDeviceGetDefaultMemPool(&pool, dev)); unsigned long long thr = ~0ull; CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr)); }
    std::vector<unsigned long long> g(d2s_g_table(), d2s_g_table() + d2s_g_table_len());
    std::vector<double> p(p10_table(), p10_table() + p10_table_len());
    impl->g_table.upload(g); impl->p10.upload(p);
    impl->T.g = impl->g_table.p; impl->T.p10 = impl->p10.p;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    impl->sm_count = prop.multiProcessorCount;
}
Context::~Context() {
    if (impl && impl->stream) {
        cudaSetDevice(device); g_alloc_stream = impl->stream;
        impl->g_table.alloc(0); impl->p10.alloc(0);       // release pool memory before the stream goes away
        cudaStreamSynchronize(impl->stream); cudaStreamDestroy(impl->stream); impl->stream = nullptr; g_alloc_stream = nullptr;
        if (impl->stream2) { cudaStreamSynchronize(impl->stream2); cudaStreamDestroy(impl->stream2); impl->stream2 = nullptr; }
        for (int k = 0; k < 2; ++k) { if (impl->side[k]) { cudaStreamSynchronize(impl->side[k]); cudaStreamDestroy(impl->side[k]); impl->side[k] = nullptr; } if (impl->ev_join[k]) { cudaEventDestroy(impl->ev_join[k]); impl->ev_join[k] = nullptr; } }
        if (impl->ev_fork) { cudaEventDestroy(impl->ev_fork); impl->ev_fork = nullptr; }
        if (impl->copy) { cudaStreamSynchronize(impl->copy); cudaStreamDestroy(impl->copy); impl->copy = nullptr; }
        if (impl->mailbox) { cudaFreeHost(impl->mailbox); impl->mailbox = nullptr; }
        for (int i = 0; i < 8; ++i) if (impl->scratch_buf[i]) { cudaFree(impl->scratch_buf[i]); impl->scratch_buf[i] = nullptr; }
        BlockCache::of(device).drop_unused();      // recycled buffers nobody holds go back to the driver with the context (live batches keep theirs)
    }
}

Batch::Batch() {}
Batch::~Batch() {
    cudaSetDevice(device);
    // plain cudaFree: the owning context (and its stream) may already be gone when a batch is released
    // cached blocks go back to the per-device free list once the device is idle (the caller may have copies of them in flight)
    bool synced = false;
    auto rel = [&](void* q) {
        if (!q) return;
        if (!synced) { cudaDeviceSynchronize(); synced = true; }
        if (!BlockCache::of(device).give_back(q, nullptr, true)) cudaFree(q);
    };
    for (int s = 0; s < 8; ++s) { rel(d_stream[s]); if (h_stream[s]) cudaFreeHost(h_stream[s]); }
    rel(d_offsets); rel(d_info); rel(d_values[0]); rel(d_values[1]); rel(d_nodes); rel(d_aux); rel(d_node_off);
}
const char* Batch::stream_host(int s) {
    if (s < 0 || s >= 8) return nullptr;
    if (!h_stream[s]) {
        cudaSetDevice(device);
        size_t bytes = stream_bytes[s];
        if (cudaMallocHost(&h_stream[s], bytes + 1) != cudaSuccess) return nullptr;
        if (bytes) cudaMemcpy(h_stream[s], d_stream[s], bytes, cudaMemcpyDeviceToHost);
        h_stream[s][bytes] = 0;
    }
    return h_stream[s];
}
bool Batch::copy_stream_to(int s, void* dst) {
    cudaSetDevice(device);
    if (!stream_bytes[s]) return true;
    return cudaMemcpy(dst, d_stream[s], stream_bytes[s], cudaMemcpyDeviceToHost) == cudaSuccess;
}
void* Context::copy_stream() { return (void*)impl->copy; }
void Context::wait_copies() { cudaSetDevice(device); cudaStreamSynchronize(impl->copy); }
bool Batch::copy_stream_to_async(int s, void* dst, void* cs) {
    cudaSetDevice(device);
    if (!stream_bytes[s]) return true;
    return cudaMemcpyAsync(dst, d_stream[s], stream_bytes[s], cudaMemcpyDeviceToHost, (cudaStream_t)cs) == cudaSuccess;
}
bool Batch::copy_offsets_to_async(int s, unsigned long long* dst, void* cs) {
    cudaSetDevice(device);
    return cudaMemcpyAsync(dst, d_offsets + (size_t)s * (n + 1), (size_t)(n + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, (cudaStream_t)cs) == cudaSuccess;
}
bool Batch::copy_offsets_to(int s, unsigned long long* dst) {
    cudaSetDevice(device);
    return cudaMemcpy(dst, d_offsets + (size_t)s * (n + 1), (size_t)(n + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess;
}
void Batch::fetch_offsets() {
    if (!h_offsets.empty() || !d_offsets) return;
    cudaSetDevice(device);
    h_offsets.resize((size_t)8 * (n + 1));
    cudaMemcpy(h_offsets.data(), d_offsets, h_offsets.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
}
// Per-replicate fields are gathered into compact arrays on the device first, so a caller that only wants the status codes of a
// large batch moves 12 bytes per replicate over PCIe instead of the 256-byte records.
__global__ void k_gather_info_small(const RepInfo* info, long long n, int32_t* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[i] = info[i].status; out[n + i] = info[i].attempts; out[2 * n + i] = info[i].sampled_leaves; }
}
__global__ void k_gather_info_stats(const RepInfo* info, long long n, double* times, int32_t* events) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        times[i] = info[i].total_u; times[n + i] = info[i].p_root_time;
        for (int e = 0; e < EV_COUNT; ++e) events[i * EV_COUNT + e] = info[i].cnt_u[e];
    }
}
void Batch::fetch_info() {
    if (have_info || !d_info) return;
    cudaSetDevice(device);
    status.resize(n); attempts.resize(n); sampled.resize(n);
    if (n > 0) {
        int32_t* tmp = nullptr;
        // stream-ordered allocation on the default stream: cudaMalloc / cudaFree would synchronise the whole device, including a
        // caller's copies that are still in flight on the copy stream
        if (cudaMallocAsync(&tmp, (size_t)n * 3 * sizeof(int32_t), 0) != cudaSuccess) throw SgpError("out of device memory while fetching replicate status");
        k_gather_info_small<<<(unsigned)((n + 255) / 256), 256>>>((const RepInfo*)d_info, n, tmp);
        cudaMemcpy(status.data(), tmp, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost);
        cudaMemcpy(attempts.data(), tmp + n, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost);
        cudaMemcpy(sampled.data(), tmp + 2 * n, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost);
        cudaFreeAsync(tmp, 0);
    }
    have_info = true;
}
void Batch::fetch_stats() {
    if (have_stats) return;
    if (app == 3) { root_time.assign((size_t)n, 0.0); total_time.assign((size_t)n, 0.0); events.assign((size_t)n * 17, 0); have_stats = true; return; }     // BranchRelaxer: no tree statistics
    if (!d_info) return;
    cudaSetDevice(device);
    root_time.assign(n, 0.0); total_time.resize(n); events.resize((size_t)n * 17);
    if (n > 0) {
        double* t = nullptr; int32_t* ev = nullptr;
        if (cudaMallocAsync(&t, (size_t)n * 2 * sizeof(double), 0) != cudaSuccess || cudaMallocAsync(&ev, (size_t)n * EV_COUNT * sizeof(int32_t), 0) != cudaSuccess) { if (t) cudaFreeAsync(t, 0); throw SgpError("out of device memory while fetching replicate statistics"); }
        k_gather_info_stats<<<(unsigned)((n + 255) / 256), 256>>>((const RepInfo*)d_info, n, t, ev);
        cudaMemcpy(total_time.data(), t, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
        cudaMemcpy(root_time.data(), t + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
        cudaMemcpy(events.data(), ev, (size_t)n * EV_COUNT * sizeof(int32_t), cudaMemcpyDeviceToHost);
        cudaFreeAsync(t, 0); cudaFreeAsync(ev, 0);
    }
    have_stats = true;
}

// device view of the run's seed; the leading decimal digits of a wide seed go into the constant-string blob
static SeedSpec make_seed_spec(const SeedHost& h, std::string& blob) {
    SeedSpec s{}; s.lo = h.lo; s.hi = h.hi; s.wide = h.wide ? 1 : 0; s.negative = h.negative ? 1 : 0; s.dec_lo19 = h.dec_lo19;
    for (int k = 0; k < 2; ++k) { s.hs_off[k] = (int)blob.size(); s.hs_len[k] = (int)h.hs[k].size(); blob += h.hs[k]; }
    return s;
}

static TinyHost make_tiny(const HostModel& h) {
    TinyHost t{}; t.nV = h.nV; t.root = h.root;
    for (int i = 0; i < h.nV; ++i) { t.left[i] = h.left[i]; t.right[i] = h.right[i]; t.parent[i] = h.parent[i]; t.v2e[i] = h.v2e[i]; t.vtime[i] = h.vtime[i]; }
    for (int e = 0; e < h.nE; ++e) { t.ep_lo[e] = h.ep_lo[e]; t.ep_mid[e] = h.ep_mid[e]; t.ep_up[e] = h.ep_up[e]; }
    t.tip = h.ep_up[h.nE - 1];
    return t;
}

void Context::run_treegen(const TreeGenConfig& cfg, const BatchOpts& opts, Batch& out) {
    CK(cudaSetDevice(device));
    cudaStream_t st = impl->stream;
    g_alloc_stream = st; g_device = device;
    const long long n = opts.n;
    if (n <= 0) throw SgpError("replicate count must be positive");
    const bool replay = opts.rng_mode == 1;
    out.stream = (void*)st;
    out.n = n; out.device = device; out.app = cfg.app; out.quiet = cfg.quiet; out.out_prefix = cfg.out_prefix;
    Timings tm; EventTimer total(st);
    total.start();
    g_h2d_bytes = 0;

    SimParams P{};
    P.lambda = cfg.P.lambda; P.mu = cfg.P.mu; P.tau = cfg.P.tau; P.theta = cfg.P.theta; P.rho = cfg.P.rho; P.bias_rate = cfg.P.bias_rate;
    P.gene_birth = cfg.P.gene_birth; P.bias = cfg.P.bias; P.do_gene_birth = cfg.P.do_gene_birth; P.ishost = cfg.P.ishost;
    P.min_leaves = cfg.P.min_leaves; P.max_leaves = cfg.P.max_leaves; P.minper = cfg.P.minper; P.maxper = cfg.P.maxper; P.max_attempts = cfg.P.max_attempts;
    P.n_sizes = (int)cfg.sizes.size();
    for (int k = 0; k < 2; ++k) {
        double sum = k == 0 ? P.lambda + P.mu : (P.lambda + P.mu) + P.tau; if (sum == 0.0) sum = 1e-48;
        P.sum[k] = sum; P.inv_sum[k] = 1.0 / sum;
        P.thr_dup[k] = k == 0 ? P.lambda / sum : (P.mu + P.tau) / sum;      // root arc: dup if U < lambda/sum; others: dup if U >= (mu+tau)/sum
        P.thr_loss[k] = P.mu / sum;
    }
    const unsigned long long seed = cfg.seed.philox_key;
    std::string seed_blob_dummy; const SeedSpec sseed = make_seed_spec(cfg.seed, seed_blob_dummy);      // MT keys need no blob

    g_marks.mark("start");
    HostOnDevice hd; hd.upload(cfg.host);
    g_marks.mark("host_upload");
    const bool domain = cfg.app == 2;
    std::vector<std::unique_ptr<HostOnDevice>> gene_dev; std::vector<DevHost> gene_views; int max_gene_nv = 1;
    for (const HostModel& g : cfg.genes) { gene_dev.emplace_back(new HostOnDevice()); gene_dev.back()->upload(g); gene_views.push_back(gene_dev.back()->view); max_gene_nv = std::max(max_gene_nv, g.nV); }
    max_gene_nv = std::max(max_gene_nv, cfg.host.nV);
    DevBuf<DevHost> d_genes; d_genes.upload(gene_views);
    // Inter-gene candidate index (DomainTreeInHostTreeCreator.sampleInterGeneArc :1004-1033 scans every vertex of every other gene
    // tree for "vertex time < t < upper time of the epoch directly above the vertex"). That interval is the vertex's first epoch,
    // and the epochs of a gene tree partition time, so the candidates of gene g at time t are the vertices filed under the one
    // epoch that contains t: per gene, epoch -> vertex ids in ascending order (the reference's scan order).
    std::vector<int32_t> ev_base, ev_off, ev_list;
    for (const HostModel& g : cfg.genes) {
        ev_base.push_back((int32_t)ev_off.size());
        std::vector<std::vector<int32_t>> by_epoch((size_t)g.nE);
        for (int v = 0; v < g.nV; ++v) {
            const int e = g.v2e[v];
            if (e < 0 || e >= g.nE || g.vtime[v] != g.ep_lo[e]) throw SgpError("internal: gene tree vertex does not start its epoch");
            by_epoch[(size_t)e].push_back(v);
        }
        for (int e = 0; e < g.nE; ++e) { ev_off.push_back((int32_t)ev_list.size()); ev_list.insert(ev_list.end(), by_epoch[(size_t)e].begin(), by_epoch[(size_t)e].end()); }
        ev_off.push_back((int32_t)ev_list.size());
    }
    DevBuf<int32_t> d_ev_base, d_ev_off, d_ev_list; d_ev_base.upload(ev_base); d_ev_off.upload(ev_off); d_ev_list.upload(ev_list);
    DevBuf<int32_t> scratch_used;
    DevBuf<int32_t> d_sizes; { std::vector<int32_t> s(cfg.sizes.begin(), cfg.sizes.end()); d_sizes.upload(s); }
    DevBuf<RepInfo> info; info.alloc((size_t)n);
    CK(cudaMemsetAsync(info.p, 0, (size_t)n * sizeof(RepInfo), st));
    const bool fast_bd = !domain && !replay && cfg.P.tau == 0.0 && !cfg.P.do_gene_birth && cfg.host.nV <= 3;
    const bool per_leaf = !(cfg.P.minper <= 0 && cfg.P.maxper >= cfg.P.max_leaves);
    const int sm = impl->sm_count;

    // ---- sub-batches ----------------------------------------------------------------------------------------------------------
    // A run can be cut into sub-batches of replicates. With one sub-batch (the default) everything runs on the context's stream in the order find ->
    // build -> label -> emit. With several (HostTreeGen in Philox mode: the find kernel is issue-bound, build / label / emit are
    // latency-bound) the find kernel of sub-batch k+1 runs on stream A while sub-batch k is built, labelled and emitted on stream B:
    // the two share the SMs (the find kernel then keeps 3 instead of 4 blocks per SM resident). Stream offsets carry over from one
    // sub-batch to the next, so the output is the same contiguous, index-ordered byte stream either way.
    // Opt-in (SGP_FLAG_PIPELINE or SGP_PIPELINE=1): measured on a B200 the two streams slow each other down (481 ms against 401 ms
    // per 10^6 C2 replicates), so the default is one sub-batch.
    const bool want_pipe = opts.pipeline || getenv("SGP_PIPELINE") != nullptr;
    long long sub_n = n;
    if (fast_bd && want_pipe && !opts.keep_trees && !opts.serial && n >= 2 * kSubBatch) sub_n = kSubBatch;
    const int K = (int)((n + sub_n - 1) / sub_n);
    const bool piped = K > 1;
    cudaStream_t stA = st, stB = piped ? impl->stream2 : st;
    struct Sub {
        long long r0 = 0, n = 0, total_nodes = 0, aux_stride = 1; int max_pool = 0;
        bool one_pass = false; unsigned long long arena_cap = 0; DevBuf<unsigned long long> cursor; std::string cfg_key;     // general engine, one-pass mode
        DevBuf<unsigned long long> counter; DevBuf<long long> pool, node_off; DevBuf<int> max_pool_d;
        DevBuf<Node> nodes; DevBuf<int32_t> aux; DevBuf<double> terms; DevBuf<TreeRec> trec; DevBuf<TreeRecX> trecx; DevBuf<RowRec> rows;
        cudaEvent_t ev_found = nullptr;
        ~Sub() { if (ev_found) cudaEventDestroy(ev_found); }
    };
    std::vector<std::unique_ptr<Sub>> subs;
    for (int k = 0; k < K; ++k) { subs.emplace_back(new Sub()); subs.back()->r0 = (long long)k * sub_n; subs.back()->n = std::min<long long>(sub_n, n - (long long)k * sub_n); CK(cudaEventCreateWithFlags(&subs.back()->ev_found, cudaEventDisableTiming)); }
    // phase times: events recorded around every phase of every sub-batch, read once at the end (no host sync inside the pipeline)
    struct Span { cudaEvent_t a = nullptr, b = nullptr; double* sink = nullptr; };
    std::vector<Span> spans;
    auto span_begin = [&](double* sink, cudaStream_t s) { Span sp; sp.sink = sink; cudaEventCreate(&sp.a); cudaEventCreate(&sp.b); cudaEventRecord(sp.a, s); spans.push_back(sp); return spans.size() - 1; };
    auto span_end = [&](size_t id, cudaStream_t s) { cudaEventRecord(spans[id].b, s); };
    struct SpanGuard { std::vector<Span>& v; ~SpanGuard() { for (auto& sp : v) { if (sp.a) cudaEventDestroy(sp.a); if (sp.b) cudaEventDestroy(sp.b); } } } span_guard{spans};
    // whatever happens, no buffer may be released while the second stream still works on it
    struct StreamFence { cudaStream_t a, b; ~StreamFence() { if (a != b) { cudaStreamSynchronize(b); } } } fence{stA, stB};

    DevBuf<uint32_t> mt_state; DevBuf<int32_t> marks, leafcnt; DevBuf<unsigned long long> arc_head;
    size_t scan_tmp_bytes = 0; cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp_bytes, (long long*)nullptr, (long long*)nullptr, (int)(sub_n + 1), stA);
    DevBuf<char> scan_tmp; scan_tmp.alloc(scan_tmp_bytes);

    // ---- K1 find (stream A) -------------------------------------------------------------------------------------------------
    auto enqueue_find = [&](Sub& S, int k) {
        const size_t sp = span_begin(&tm.find_ms, stA);
        S.counter.alloc(1); CK(cudaMemsetAsync(S.counter.p, 0, sizeof(unsigned long long), stA));
        if (fast_bd) {
            BdArgs a{}; a.P = P; a.H = make_tiny(cfg.host); a.T = impl->T; a.n = S.n; a.rep_base = opts.offset + S.r0; a.seed = seed; a.keys = philox_keys(seed); a.info = info.p + S.r0;
            a.work_counter = S.counter.p; a.sizes = d_sizes.p;
            int blocks = sm * 4; long long need = (S.n * 32 + 255) / 256; if (need < blocks) blocks = (int)need;
            launch_bd_find(a, blocks, piped ? 3 : 4, stA); ++tm.launches;
            CK(cudaGetLastError());
        } else {
            // node capacity of a worker's scratch pool: typical trees fit; the rare larger ones are re-run with 8x the capacity
            int ref_leaves = cfg.host.nLeaves; for (const HostModel& g : cfg.genes) ref_leaves = std::max(ref_leaves, g.nLeaves);
            int cap0 = 1024; while (cap0 < 6 * ref_leaves && cap0 < (1 << 16)) cap0 *= 2;
            const int mark_n = domain ? max_gene_nv : cfg.host.nV, cnt_n = domain ? max_gene_nv : std::max(1, cfg.host.nLeaves);
            auto general_find = [&](long long n_run, Node* arena, unsigned long long arena_cap, unsigned long long* arena_cursor, long long* node_off_out) {
            int cap = cap0; long long workers = std::min<long long>((long long)sm * 1024, (n_run + 127) / 128 * 128);
            DevBuf<long long> redo, keep; DevBuf<unsigned long long> redo_count;
            const long long* rep_ids = nullptr; long long todo = n_run;
            for (int round = 0;; ++round) {
                // bound scratch to ~24 GiB
                while ((double)workers * cap * (sizeof(Node) + 8.0) > 24.0 * (1ull << 30) && workers > 128) workers /= 2;
                workers = std::max<long long>(128, workers / 128 * 128);
                Node* s_nodes = (Node*)impl->scratch(0, (size_t)workers * cap * sizeof(Node));
                int32_t* s_queue = (int32_t*)impl->scratch(1, (size_t)workers * cap * sizeof(int32_t)), *s_pend = (int32_t*)impl->scratch(2, (size_t)workers * cap * sizeof(int32_t));
                // spare pools of 8x the size (~0.8 GB): trees that outgrow a worker's pool are redone at once inside this launch
                DevBuf<int> b_next;
                const int big_cap = (int)std::min<long long>((long long)cap * 8, 1 << 24);
                const int big_count = (int)std::max<long long>(0, std::min<long long>(1024, (long long)((800ull << 20) / ((unsigned long long)big_cap * (sizeof(Node) + 8)))));
                Node* b_nodes = nullptr; int32_t* b_queue = nullptr, *b_pend = nullptr;
                if (big_count > 0) {
                    b_nodes = (Node*)impl->scratch(3, (size_t)big_count * big_cap * sizeof(Node));
                    b_queue = (int32_t*)impl->scratch(4, (size_t)big_count * big_cap * sizeof(int32_t)); b_pend = (int32_t*)impl->scratch(5, (size_t)big_count * big_cap * sizeof(int32_t));
                }
                b_next.alloc(1); CK(cudaMemsetAsync(b_next.p, 0, sizeof(int), stA));
                marks.alloc((size_t)workers * mark_n); leafcnt.alloc((size_t)workers * cnt_n);
                // per-arc vertex lists of the victim search: GuestTreeGen on hosts where a guest tree is large enough for the pool scan to
                // cost more than keeping the lists (1000-leaf host: find 140 -> 95 ms per 8192 trees; 100-leaf host: 29 -> 31 ms per 2*10^5),
                // and not so large that the heads would take > 8 GB
                const char* al_env = getenv("SGP_ARC_LISTS");         // test hook: 1 forces the lists on small hosts, 0 switches them off
                const bool arc_lists = !domain && (al_env ? al_env[0] == '1' : cfg.host.nV >= 512) && (double)workers * cfg.host.nV * 8.0 <= 8.0 * (1ull << 30);
                if (arc_lists) { arc_head.alloc((size_t)workers * cfg.host.nV); CK(cudaMemsetAsync(arc_head.p, 0, (size_t)workers * cfg.host.nV * 8, stA)); } else arc_head.alloc(0);
                if (domain) scratch_used.alloc((size_t)workers * cfg.genes.size());
                if (replay) mt_state.alloc((size_t)(workers / 32 + 1) * 624 * 32);
                FindArgs a{}; a.P = P; a.H = hd.view; a.T = impl->T; a.n = todo; a.rep_ids = rep_ids; a.rep_base = opts.offset; a.seed = seed; a.sseed = sseed; a.info = info.p;
                a.work_counter = S.counter.p; a.scratch_nodes = s_nodes; a.scratch_queue = s_queue; a.scratch_pend = s_pend; a.cap = cap;
                a.scratch_marks = marks.p; a.scratch_leafcnt = leafcnt.p; a.scratch_arc_head = arc_head.p; a.mt_state = mt_state.p; a.sizes = d_sizes.p; a.per_leaf_check = per_leaf;
                a.genes = d_genes.p; a.n_genes = (int)cfg.genes.size(); a.max_gene_nv = max_gene_nv; a.all_genes = cfg.all_genes; a.inter_gene = cfg.inter_gene; a.inter_species = cfg.inter_species;
                a.scratch_used = scratch_used.p; a.ev_base = d_ev_base.p; a.ev_off = d_ev_off.p; a.ev_list = d_ev_list.p;
                a.big_nodes = b_nodes; a.big_queue = b_queue; a.big_pend = b_pend; a.big_cap = big_cap; a.big_count = big_count; a.big_next = b_next.p;
                a.arena = arena; a.arena_cap = arena_cap; a.arena_cursor = arena_cursor; a.node_off_out = node_off_out;
                CK(cudaMemsetAsync(S.counter.p, 0, sizeof(unsigned long long), stA));
                int blocks = (int)(workers / 128);
                launch_find_general(replay, domain, a, blocks, stA);
                ++tm.launches;
                CK(cudaGetLastError());
                // replicates that ran out of node capacity are re-run with a larger pool
                redo.alloc((size_t)n_run); redo_count.alloc(1);
                CK(cudaMemsetAsync(redo_count.p, 0, sizeof(unsigned long long), stA));
                k_collect_status<<<(unsigned)((n_run + 255) / 256), 256, 0, stA>>>(info.p, n_run, REP_NODE_OVERFLOW, redo.p, redo_count.p); ++tm.launches;
                impl->post(0, redo_count.p, sizeof(unsigned long long), stA);
                CK(cudaStreamSynchronize(stA));
                const unsigned long long cnt = impl->take<unsigned long long>(0);
                if (cnt == 0 || cap >= (1 << 24) || round > 8) break;
                keep.alloc((size_t)cnt);
                CK(cudaMemcpyAsync(keep.p, redo.p, cnt * sizeof(long long), cudaMemcpyDeviceToDevice, stA));
                CK(cudaStreamSynchronize(stA));
                rep_ids = keep.p; todo = (long long)cnt; cap *= 8; workers = std::min<long long>(workers, (todo + 127) / 128 * 128);
            }
            };      // general_find
            // ---- one-pass mode (large batches): accepted trees are copied into a bump-allocated arena by the find kernel itself. The
            // arena is sized from a pilot run over a few hundred replicates (their results are overwritten by the main run); if
            // the estimate turns out too small the vertex records are regenerated by the build kernel as in the two-pass mode.
            S.node_off.alloc((size_t)S.n + 1); CK(cudaMemsetAsync(S.node_off.p, 0, ((size_t)S.n + 1) * sizeof(long long), stA));
            // worth it when the arena can be sized without a pilot (this configuration ran before), or when the batch is large enough
            // for a pilot over n / 64 replicates to be cheap next to it
            const bool want_one_pass = getenv("SGP_TWO_PASS") == nullptr && (n >= 8192 || (n >= 256 && getenv("SGP_ARENA_MARGIN") == nullptr));
            std::string cfg_key = std::to_string(cfg.app) + (replay ? "r" : "p");
            for (size_t i = 0; i < cfg.args.size(); ++i) if ((int)i != cfg.seed_arg_index) { cfg_key += '\x1f'; cfg_key += cfg.args[i]; }
            S.cfg_key = cfg_key;
            const auto known = impl->nodes_per_replicate.find(cfg_key);
            if (want_one_pass && known != impl->nodes_per_replicate.end() && getenv("SGP_ARENA_MARGIN") == nullptr) {
                // this configuration ran before on this context: no pilot (a pilot costs one full wave of the latency-bound kernel)
                S.arena_cap = (unsigned long long)(known->second * (double)n * 1.25) + (1ull << 20);
                g_alloc_stream = stB; S.nodes.alloc((size_t)S.arena_cap); g_alloc_stream = stA;
                S.cursor.alloc(1); CK(cudaMemsetAsync(S.cursor.p, 0, sizeof(unsigned long long), stA));
                general_find(n, S.nodes.p, S.arena_cap, S.cursor.p, S.node_off.p);
                S.one_pass = true;
            } else if (want_one_pass && n >= 8192) {
                const long long n_pilot = std::max<long long>(256, n / 64);
                general_find(n_pilot, nullptr, 0, nullptr, nullptr);
                DevBuf<long long> ppool, poff; DevBuf<int> pmax; ppool.alloc((size_t)n_pilot + 1); poff.alloc((size_t)n_pilot + 1); pmax.alloc(1);
                CK(cudaMemsetAsync(ppool.p, 0, ((size_t)n_pilot + 1) * sizeof(long long), stA)); CK(cudaMemsetAsync(pmax.p, 0, sizeof(int), stA));
                k_extract_pool<<<(unsigned)((n_pilot + 255) / 256), 256, 0, stA>>>(info.p, n_pilot, ppool.p, pmax.p); ++tm.launches;
                size_t tb = scan_tmp_bytes;
                cub::DeviceScan::ExclusiveSum(scan_tmp.p, tb, ppool.p, poff.p, (int)(n_pilot + 1), stA); ++tm.launches;
                impl->post(0, poff.p + n_pilot, sizeof(long long), stA); impl->post(8, pmax.p, sizeof(int), stA);
                CK(cudaStreamSynchronize(stA));
                const long long pilot_nodes = impl->take<long long>(0); const int pilot_max = impl->take<int>(8);
                const double per_rep = (double)pilot_nodes / (double)n_pilot;
                const char* me = getenv("SGP_ARENA_MARGIN");            // test hook: a margin below 1 forces the two-pass fallback
                const double margin = me ? atof(me) : 1.25;
                S.arena_cap = (unsigned long long)(per_rep * (double)n * margin) + (me ? 0ull : (unsigned long long)pilot_max * 64ull + (1ull << 16));
                CK(cudaMemsetAsync(info.p, 0, (size_t)n * sizeof(RepInfo), stA));
                g_alloc_stream = stB; S.nodes.alloc((size_t)S.arena_cap); g_alloc_stream = stA;
                S.cursor.alloc(1); CK(cudaMemsetAsync(S.cursor.p, 0, sizeof(unsigned long long), stA));
                general_find(n, S.nodes.p, S.arena_cap, S.cursor.p, S.node_off.p);
                S.one_pass = true;
            } else general_find(n, nullptr, 0, nullptr, nullptr);
        }
        span_end(sp, stA);
        // node offsets of the sub-batch: exclusive scan of the accepted attempts' pool sizes
        S.pool.alloc((size_t)S.n + 1); if (!S.node_off.p) S.node_off.alloc((size_t)S.n + 1); S.max_pool_d.alloc(1);
        CK(cudaMemsetAsync(S.pool.p, 0, ((size_t)S.n + 1) * sizeof(long long), stA));
        CK(cudaMemsetAsync(S.max_pool_d.p, 0, sizeof(int), stA));
        k_extract_pool<<<(unsigned)((S.n + 255) / 256), 256, 0, stA>>>(info.p + S.r0, S.n, S.pool.p, S.max_pool_d.p); ++tm.launches;
        if (S.one_pass) {
            // the arena is filled already; did every accepted tree fit?
            impl->post(0, S.cursor.p, sizeof(unsigned long long), stA);
            CK(cudaStreamSynchronize(stA));
            const unsigned long long used = impl->take<unsigned long long>(0);
            if (used > S.arena_cap) { S.one_pass = false; g_alloc_stream = stB; S.nodes.alloc(0); g_alloc_stream = stA; }       // too small: two-pass (the find results stand, only the copies are incomplete)
            else CK(cudaMemcpyAsync(S.node_off.p + S.n, S.cursor.p, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stA));      // "total" slot
        }
        if (!S.one_pass) {
            size_t tb = scan_tmp_bytes;
            cub::DeviceScan::ExclusiveSum(scan_tmp.p, tb, S.pool.p, S.node_off.p, (int)(S.n + 1), stA); ++tm.launches;
        }
        impl->post(kMailFind + (size_t)(k & 15) * 16, S.node_off.p + S.n, sizeof(long long), stA); impl->post(kMailFind + (size_t)(k & 15) * 16 + 8, S.max_pool_d.p, sizeof(int), stA);
        CK(cudaEventRecord(S.ev_found, stA));
    };

    unsigned mask = opts.stream_mask;
    if (mask == 0) {
        if (cfg.quiet) mask = 1u << ST_PRUNED_TREE;
        else {
            mask = 0xFu;
            if (cfg.app != 0 && !cfg.exclude_meta) mask |= 0xF0u;
        }
    }
    // constant strings
    std::string blob; std::vector<int32_t> gname, leafname;
    EmitArgs e{};
    auto add = [&](const std::string& s, int& off, int& len) { off = (int)blob.size(); len = (int)s.size(); blob += s; };
    {
        std::string pre = "[", post;
        bool split = replay && cfg.seed_arg_index >= 0;
        for (size_t i = 0; i < cfg.args.size(); ++i) {
            std::string& dst = (split && (int)i > cfg.seed_arg_index) ? post : pre;
            if ((int)i == cfg.seed_arg_index && split) { if (i) pre += ", "; continue; }
            if (i) dst += ", ";
            dst += cfg.args[i];
        }
        if (split) post += "]"; else pre += "]";
        add(pre, e.args_pre_off, e.args_pre_len); add(post, e.args_post_off, e.args_post_len);
        e.seed_mode = split ? 1 : 0; e.sseed = make_seed_spec(cfg.seed, blob);
        e.args_pre_off16[0] = -1;
        if (pre.size() >= 512) for (int k = 0; k < 16; ++k) { while ((int)(blob.size() % 16) != k) blob.push_back(' '); e.args_pre_off16[k] = (int)blob.size(); blob += pre; }
    }
    add("# GUEST-TO-HOST MAP\nHost tree:\t" + cfg.host.header + "\nGuest vertex name:\tGuest vertex ID:\tGuest vertex type:\tGuest vertex time:\tHost vertex/arc ID:\tHost epoch ID:\n",
        e.g2h_head_off, e.g2h_head_len);
    {
        const std::string head = blob.substr(e.g2h_head_off, e.g2h_head_len);
        for (int k = 0; k < 16; ++k) { while ((int)(blob.size() % 16) != k) blob.push_back(' '); e.g2h_head_off16[k] = (int)blob.size(); blob += head; }
    }
    std::vector<int32_t> leaf_base, g2h_gene; std::vector<double> ep_times_all;
    if (!domain) {
        { int o, l; add(cfg.host.has_tree_name ? cfg.host.tree_name : std::string("null"), o, l); gname.push_back(o); gname.push_back(l); }
        for (int v = 0; v < cfg.host.nV; ++v) { int o, l; add(cfg.host.left[v] < 0 ? cfg.host.leaf_name[v] : std::string("null"), o, l); leafname.push_back(o); leafname.push_back(l); }
    } else {
        for (const HostModel& g : cfg.genes) {
            { int o, l; add(g.has_tree_name ? g.tree_name : std::string("null"), o, l); gname.push_back(o); gname.push_back(l); }
            leaf_base.push_back((int32_t)(leafname.size() / 2));
            for (int v = 0; v < g.nV; ++v) { int o, l; add(g.left[v] < 0 ? g.leaf_name[v] : std::string("null"), o, l); leafname.push_back(o); leafname.push_back(l); }
            { int o, l; add("# GUEST-TO-HOST MAP\nHost tree:\t" + g.header + "\nGuest vertex name:\tGuest vertex ID:\tGuest vertex type:\tGuest vertex time:\tHost vertex/arc ID:\tHost epoch ID:\n", o, l);
              g2h_gene.push_back(o); g2h_gene.push_back(l); }
        }
    }
    DevBuf<char> d_blob; { std::vector<char> b(blob.begin(), blob.end()); b.push_back(0); d_blob.upload(b); }
    std::vector<double> gene_ept; std::vector<int32_t> gene_epb;
    for (const HostModel& gm : cfg.genes) { gene_epb.push_back((int32_t)(gene_ept.size() / 3)); for (int q = 0; q < gm.nE; ++q) { gene_ept.push_back(gm.ep_lo[q]); gene_ept.push_back(gm.ep_mid[q]); gene_ept.push_back(gm.ep_up[q]); } }
    DevBuf<double> d_gene_ept; d_gene_ept.upload(gene_ept); DevBuf<int32_t> d_gene_epb; d_gene_epb.upload(gene_epb);
    DevBuf<int32_t> d_gname, d_leafname, d_leaf_base, d_g2h_gene; d_gname.upload(gname); d_leafname.upload(leafname); d_leaf_base.upload(leaf_base); d_g2h_gene.upload(g2h_gene);
    // sizes / offsets of all replicates, [stream][n + 1]; a sub-batch fills its slice, offsets carry over
    const long long stride = n + 1;
    DevBuf<unsigned long long> sizes, offsets, carry; sizes.alloc((size_t)ST_COUNT * stride); offsets.alloc((size_t)ST_COUNT * stride); carry.alloc(ST_COUNT);
    CK(cudaMemsetAsync(sizes.p, 0, (size_t)ST_COUNT * stride * sizeof(unsigned long long), stA));
    CK(cudaMemsetAsync(offsets.p, 0, (size_t)ST_COUNT * stride * sizeof(unsigned long long), stA));
    CK(cudaMemsetAsync(carry.p, 0, ST_COUNT * sizeof(unsigned long long), stA));
    DevBuf<int> overflow; overflow.alloc(1); CK(cudaMemsetAsync(overflow.p, 0, sizeof(int), stA));
    e.rep_base = opts.offset; e.T = impl->T;
    e.app = cfg.app; e.exclude_meta = cfg.exclude_meta; e.append_sigma = cfg.append_sigma; e.is_species = cfg.is_species; e.stream_mask = mask;
    e.prefix_len = (int)std::min<size_t>(cfg.vertex_prefix.size(), sizeof(e.prefix)); memcpy(e.prefix, cfg.vertex_prefix.data(), e.prefix_len);
    if (cfg.vertex_prefix.size() > sizeof(e.prefix)) throw SgpError("vertex prefix longer than 24 bytes is not supported");
    e.blob = d_blob.p; e.gname_off = d_gname.p; e.leafname_off = d_leafname.p; e.ep_times = hd.ep_times.p;
    e.leaf_base = d_leaf_base.p; e.g2h_gene_off = d_g2h_gene.p; e.gene_ep_times = d_gene_ept.p; e.gene_ep_base = d_gene_epb.p;
    e.stride = stride; e.overflow = overflow.p;
    cudaEvent_t ev_setup = nullptr; CK(cudaEventCreateWithFlags(&ev_setup, cudaEventDisableTiming));
    struct EvGuard { cudaEvent_t& e; ~EvGuard() { if (e) cudaEventDestroy(e); } } ev_guard{ev_setup};
    CK(cudaEventRecord(ev_setup, stA));
    if (piped) CK(cudaStreamWaitEvent(stB, ev_setup, 0));
    unsigned long long totals[ST_COUNT] = {0}, out_cap[ST_COUNT] = {0};
    double mean_piece = 0.0; e.stage_bytes = 3072;

    // ---- build -> label -> emit of one sub-batch (stream B) ---------------------------------------------------------------------
    auto emit_args_for = [&](Sub& S) {
        EmitArgs a = e; a.n = S.n; a.rep_base = opts.offset + S.r0; a.info = info.p + S.r0; a.node_off = S.node_off.p; a.nodes = S.nodes.p; a.aux = S.aux.p; a.aux_stride = S.aux_stride;
        a.trec = S.trec.p; a.trecx = S.trecx.p; a.rows = S.rows.p; a.sizes = sizes.p + S.r0; a.offsets = offsets.p + S.r0;
        for (int q = 0; q < ST_COUNT; ++q) { a.out[q] = out.d_stream[q]; a.cap[q] = out_cap[q]; }
        return a;
    };
    auto post_process = [&](Sub& S, int k) {
        g_marks.mark("enqueued_find");
        CK(cudaEventSynchronize(S.ev_found));
        g_marks.mark("find_done");
        S.total_nodes = impl->take<long long>(kMailFind + (size_t)(k & 15) * 16); S.max_pool = impl->take<int>(kMailFind + (size_t)(k & 15) * 16 + 8);
        if (piped) CK(cudaStreamWaitEvent(stB, S.ev_found, 0));
        g_alloc_stream = stB;
        if (!fast_bd && !S.cfg_key.empty() && S.n > 0) {
            if (impl->nodes_per_replicate.size() > 256) impl->nodes_per_replicate.clear();      // (a cache, not a history)
            double& seen = impl->nodes_per_replicate[S.cfg_key];
            seen = std::max(seen, (double)S.total_nodes / (double)S.n);
        }
        S.aux_stride = S.one_pass ? (long long)S.arena_cap : std::max<long long>(S.total_nodes, 1);
        // the order arrays and the labelled vertex records are only kept for chained stages and for trees too large for the
        // shared-memory label kernel; everything else leaves the label kernel as emission records
        // trees above the shared-memory label kernel's reach are labelled in place with global work arrays (bounded to ~2 GB);
        // only if even that does not fit do they go to the thread-per-tree kernel, which needs the order arrays
        int work_blocks = 0; long long work_stride = 0; int32_t* work = nullptr;
        const bool thread_label = getenv("SGP_LABEL_THREAD_KERNEL") != nullptr;
        if (S.max_pool > label_smem_max_pool() && !thread_label) {
            work_stride = 9ll * (S.max_pool + 2) + 16;
            work_blocks = (int)std::min<long long>((long long)sm * 32, (2ll << 30) / (work_stride * 4));      // (latency-bound: 32 one-warp blocks per SM measured best)
            if (const char* wb = getenv("SGP_LABEL_GLOBAL_BLOCKS")) work_blocks = std::max(16, atoi(wb));
            if (work_blocks >= 16) work = (int32_t*)impl->scratch(6, (size_t)work_blocks * work_stride * 4); else work_blocks = 0;
        }
        const bool keep_nodes = opts.keep_trees || (S.max_pool > label_smem_max_pool() && work_blocks == 0);
        const bool want_rows = (mask & 0xF0u) != 0, want_trees = (mask & 0x3u) != 0;
        if (!S.one_pass) S.nodes.alloc((size_t)S.aux_stride);
        if (keep_nodes || (!fast_bd && !S.one_pass)) S.aux.alloc((size_t)S.aux_stride * 4);      // (the general build kernel uses them as scratch)
        if (want_trees) { S.trec.alloc((size_t)2 * S.aux_stride); if (cfg.app != 0) S.trecx.alloc((size_t)2 * S.aux_stride); }
        if (want_rows) S.rows.alloc((size_t)2 * S.aux_stride);
        S.terms.alloc((size_t)(cfg.app == 0 ? 2 : 4) * S.aux_stride);
        g_alloc_stream = stA;
        size_t sp = span_begin(&tm.build_ms, stB);
        if (fast_bd) {
            BdArgs a{}; a.P = P; a.H = make_tiny(cfg.host); a.T = impl->T; a.n = S.n; a.rep_base = opts.offset + S.r0; a.seed = seed; a.keys = philox_keys(seed); a.info = info.p + S.r0;
            a.node_off = S.node_off.p; a.nodes = S.nodes.p;
            launch_bd_build(a, (int)((S.n + 255) / 256), stB); ++tm.launches;
        } else if (!S.one_pass) {
            const long long threads = std::max<long long>(128, std::min<long long>((long long)sm * 1024, (n + 127) / 128 * 128));
            marks.alloc((size_t)threads * (domain ? max_gene_nv : cfg.host.nV));
            {
                const char* al_env = getenv("SGP_ARC_LISTS");
                const bool arc_lists = !domain && (al_env ? al_env[0] == '1' : cfg.host.nV >= 512) && (double)threads * cfg.host.nV * 8.0 <= 8.0 * (1ull << 30);
                if (arc_lists) { arc_head.alloc((size_t)threads * cfg.host.nV); CK(cudaMemsetAsync(arc_head.p, 0, (size_t)threads * cfg.host.nV * 8, stB)); } else arc_head.alloc(0);
            }
            if (domain) scratch_used.alloc((size_t)threads * cfg.genes.size());
            if (replay) mt_state.alloc((size_t)(threads / 32 + 1) * 624 * 32);
            CK(cudaMemsetAsync(S.counter.p, 0, sizeof(unsigned long long), stB));
            BuildArgs a{}; a.P = P; a.H = hd.view; a.T = impl->T; a.n = n; a.rep_base = opts.offset; a.seed = seed; a.sseed = sseed; a.info = info.p; a.node_off = S.node_off.p; a.work_counter = S.counter.p;
            a.nodes = S.nodes.p; a.aux = S.aux.p; a.aux_stride = S.aux_stride; a.scratch_marks = marks.p; a.scratch_leafcnt = nullptr; a.scratch_arc_head = arc_head.p; a.mt_state = mt_state.p;
            a.genes = d_genes.p; a.n_genes = (int)cfg.genes.size(); a.max_gene_nv = max_gene_nv; a.all_genes = cfg.all_genes; a.inter_gene = cfg.inter_gene; a.inter_species = cfg.inter_species;
            a.scratch_used = scratch_used.p; a.ev_base = d_ev_base.p; a.ev_off = d_ev_off.p; a.ev_list = d_ev_list.p;
            // a lane builds its tree alone, so the largest trees set the kernel's tail: start them first
            DevBuf<long long> order, n_sel; DevBuf<char> ptmp;
            if (n >= 4096) {
                order.alloc((size_t)n); n_sel.alloc(1);
                cub::CountingInputIterator<long long> ids(0);
                PoolAbove big{S.pool.p, 4 * (S.total_nodes / n) + 64};
                size_t pb = 0; cub::DevicePartition::If(nullptr, pb, ids, order.p, n_sel.p, (int)n, big, stB);
                ptmp.alloc(pb);
                cub::DevicePartition::If(ptmp.p, pb, ids, order.p, n_sel.p, (int)n, big, stB); ++tm.launches;
                a.order = order.p; a.n_big = n_sel.p;
            }
            launch_build_general(replay, domain, a, (int)(threads / 128), stB);
            ++tm.launches;
        }
        CK(cudaGetLastError());
        span_end(sp, stB);
        sp = span_begin(&tm.label_ms, stB);
        {
            LabelArgs a{}; a.n = S.n; a.info = info.p + S.r0; a.node_off = S.node_off.p; a.nodes = S.nodes.p; a.aux = S.aux.p; a.aux_stride = S.aux_stride; a.T = impl->T;
            a.host_root_time = cfg.host.vtime[cfg.host.root]; a.has_stem_override = cfg.has_stem_override; a.stem_override = cfg.stem_override;
            a.pool = S.pool.p;
            a.write_nodes = keep_nodes ? 1 : 0; a.work = work; a.work_stride = work_stride; a.work_blocks = work_blocks; a.E = emit_args_for(S); a.terms = S.terms.p; a.want_beneath = cfg.app == 0 ? 0 : 1;
            tm.launches += launch_label_prune(a, S.max_pool, impl->sm_count, stB);
            CK(cudaGetLastError());
        }
        span_end(sp, stB);
        g_marks.mark("label_enqueued");
        sp = span_begin(&tm.emit_size_ms, stB);
        EmitArgs ea = emit_args_for(S);
        tm.launches += launch_emit_sizes(ea, stB);
        CK(cudaGetLastError());
        k_scan_streams<<<ST_COUNT, 1024, 0, stB>>>(sizes.p + S.r0, offsets.p + S.r0, stride, S.n, carry.p); ++tm.launches;
        span_end(sp, stB);
        if (k == 0) {
            // output buffers: exact for a single sub-batch, otherwise sized from the first sub-batch's bytes per replicate with a margin
            // (a later sub-batch that does not fit raises `overflow`; the run is then repeated unpipelined with exact sizes)
            impl->post(kMailEmit, carry.p, ST_COUNT * sizeof(unsigned long long), stB);
            g_marks.mark("sizes_enqueued");
            CK(cudaStreamSynchronize(stB));
            g_marks.mark("sizes_done");
            g_alloc_stream = stB;
            for (int q = 0; q < ST_COUNT; ++q) {
                const unsigned long long first = impl->take<unsigned long long>(kMailEmit + (size_t)q * 8);
                unsigned long long cap = first;
                if (piped) {
                    const char* me = getenv("SGP_OUTPUT_MARGIN");        // test hook: a margin below 1 forces the fallback
                    const double margin = me ? atof(me) : kOutputMargin;
                    const double per = (double)first / (double)S.n; cap = (unsigned long long)(per * (double)n * margin) + (me ? 0ull : (1ull << 20));
                }
                out_cap[q] = std::max<unsigned long long>(cap, 1);
                impl->alloc_output(&out.d_stream[q], out_cap[q], stB);
                // mean piece length of the Newick streams (a vertex is one piece; the rows of the other streams are shorter)
                if (q == ST_UNPRUNED_TREE || q == ST_PRUNED_TREE) mean_piece = std::max(mean_piece, (double)first / (double)std::max<long long>(S.total_nodes, 1));
            }
            // stage of the write kernels: room for 32 pieces of 1.2 x the mean length (3 KB for HostTreeGen's 57-byte pieces - small
            // stages keep 8 blocks per SM resident -, 5 KB for the 100-byte pieces of GuestTreeGen on a 1000-leaf host); a trip that
            // still does not fit is written with plain stores
            e.stage_bytes = (int)std::min<long long>(6144, std::max<long long>(3072, ((long long)(mean_piece * 38.0) + 256 + 127) / 128 * 128));
            g_alloc_stream = stA;
            ea = emit_args_for(S);
        }
        g_marks.mark("outputs_allocated");
        sp = span_begin(&tm.emit_write_ms, stB);
        tm.launches += launch_emit_write(ea, stB, impl->side, 2, impl->ev_fork, impl->ev_join);
        CK(cudaGetLastError());
        span_end(sp, stB);
        if (piped) {        // the vertex records of this sub-batch are not needed any more: hand them back in stream order
            g_alloc_stream = stB;
            S.nodes.alloc(0); S.aux.alloc(0); S.terms.alloc(0); S.trec.alloc(0); S.trecx.alloc(0); S.rows.alloc(0);
            g_alloc_stream = stA;
        }
    };

    struct StreamFence2 { cudaStream_t a, b; ~StreamFence2() { if (a != b) cudaStreamSynchronize(b); } } fence_last{stA, stB};      // destroyed first: before any buffer above
    enqueue_find(*subs[0], 0);
    for (int k = 0; k < K; ++k) {
        if (k + 1 < K) enqueue_find(*subs[(size_t)k + 1], k + 1);       // stream A always has the next find queued before the host waits
        post_process(*subs[(size_t)k], k);
    }
    impl->post(kMailEmit, carry.p, ST_COUNT * sizeof(unsigned long long), stB); impl->post(kMailEmit + 64, overflow.p, sizeof(int), stB);
    g_marks.mark("write_enqueued");
    CK(cudaStreamSynchronize(stB)); CK(cudaStreamSynchronize(stA));
    g_marks.mark("write_done");
    for (int q = 0; q < ST_COUNT; ++q) { totals[q] = impl->take<unsigned long long>(kMailEmit + (size_t)q * 8); out.stream_bytes[q] = totals[q]; }
    if (impl->take<int>(kMailEmit + 64) != 0) {
        // a stream outgrew its estimated buffer: start over without the pipeline (exact sizes)
        for (int q = 0; q < ST_COUNT; ++q) { if (out.d_stream[q]) { if (!BlockCache::of(device).give_back(out.d_stream[q], stB, false)) cudaFree(out.d_stream[q]); out.d_stream[q] = nullptr; } out.stream_bytes[q] = 0; }
        subs.clear();
        BatchOpts o2 = opts; o2.serial = true;
        run_treegen(cfg, o2, out);
        return;
    }
    for (auto& sp : spans) { float ms = 0; if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) *sp.sink += ms; }

    // ---- summary ---------------------------------------------------------------------------------------------------------------
    {
        DevBuf<DevSummary> ds; ds.alloc(1); CK(cudaMemsetAsync(ds.p, 0, sizeof(DevSummary), st));
        k_summary<<<(unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)impl->sm_count * 4)), 256, 0, st>>>(info.p, n, ds.p); ++tm.launches;
        impl->post(0, ds.p, sizeof(DevSummary), st); CK(cudaStreamSynchronize(st));
        g_marks.mark("summary_done");
        const DevSummary hs = impl->take<DevSummary>(0);
        Summary& S = out.summary; S.n_ok = (long long)hs.n_ok; S.n_failed = (long long)hs.n_failed; S.attempts_total = (long long)hs.attempts_total; S.vertices_sampled = (long long)hs.vertices; S.vertices_abandoned = (long long)hs.vertices_abandoned; S.attempts_completed = (long long)hs.attempts_completed;
        for (int i = 0; i < 17; ++i) S.event_counts[i] = (long long)hs.ev[i];
        for (int i = 0; i < 1025; ++i) S.leaf_hist[i] = (long long)hs.leaf_hist[i];
        for (int s = 0; s < ST_COUNT; ++s) S.out_bytes[s] = (long long)totals[s];
    }
    out.d_offsets = offsets.release(); out.d_info = info.release(); out.offsets_stride = stride;
    if (opts.keep_trees) {
        Sub& S = *subs[0];
        out.d_nodes = S.nodes.release(); out.d_aux = S.aux.release(); out.d_node_off = S.node_off.release(); out.aux_stride = S.aux_stride;
        out.name_prefix = cfg.vertex_prefix; out.name_append_sigma = cfg.append_sigma; out.name_exclude_meta = cfg.exclude_meta;
    }
    static const char* host_sfx[8] = {".unpruned.tree", ".pruned.tree", ".unpruned.info", ".pruned.info", ".unpruned.guest2host", ".pruned.guest2host", ".unpruned.leafmap", ".pruned.leafmap"};
    for (int s = 0; s < 8; ++s) out.suffix[s] = (mask & (1u << s)) ? host_sfx[s] : "";
    out.stream_mask = mask;
    tm.total_ms = total.stop(); tm.h2d_bytes = (long long)g_h2d_bytes;
    g_marks.mark("end"); g_marks.dump();
    tm.sub_batches = K;
    out.timings = tm;
    if (!opts.keep_on_device) {
        EventTimer d2h(st); d2h.start();
        for (int s = 0; s < 8; ++s) if (mask & (1u << s)) out.stream_host(s);
        out.fetch_offsets();
        out.timings.d2h_ms = d2h.stop();
    }
}

// ---- chained stage: trees of a generator batch as parsed input trees ---------------------------------------------------------------
// What io/NewickTreeReader.java:158-296 + io/NewickVertex.java:367-375 would make of the batch's <prefix>.pruned.tree files: vertex ids
// in post-order, lengths, names, the HOST= value of every vertex (0 without tags, -10 with tags that carry no HOST=). The device
// part (k_ingest_count / k_ingest_fill, the kernels of the chained BranchRelaxer) reads the kept vertex records and order arrays and
// leaves ~25 bytes per vertex, which is all that crosses PCIe: no Newick text is written or parsed.
std::vector<FlatTree> Context::ingest_trees(const Batch& src, long long first, long long count, bool pruned) {
    CK(cudaSetDevice(device));
    cudaStream_t st = impl->stream; g_alloc_stream = st; g_device = device;
    if (!src.d_nodes || !src.d_info || !src.d_aux || src.device != device) throw SgpError("source batch holds no trees on this device (run the generator with SGP_FLAG_KEEP_TREES)");
    if (first < 0 || count <= 0 || first + count > src.n) throw SgpError("tree range outside the source batch");
    if (count > (1 << 24)) throw SgpError("too many trees for one forest");
    const int T = (int)count;
    IngestArgs g{}; g.n_src = count; g.info = (const RepInfo*)src.d_info + first; g.node_off = src.d_node_off + first; g.nodes = (const Node*)src.d_nodes;
    g.aux = src.d_aux; g.aux_stride = src.aux_stride; g.pruned = pruned ? 1 : 0; g.keep_interior = 1;
    DevBuf<long long> nv, v0; nv.alloc((size_t)T + 1); v0.alloc((size_t)T + 1);
    g.nv = nv.p; launch_ingest_count(g, st);
    size_t tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, nv.p, v0.p, T + 1, st);
    DevBuf<char> tmp; tmp.alloc(tb);
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, nv.p, v0.p, T + 1, st);
    std::vector<long long> h_v0((size_t)T + 1);
    CK(cudaMemcpyAsync(h_v0.data(), v0.p, ((size_t)T + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    const size_t cnt = (size_t)std::max<long long>(h_v0[(size_t)T], 1);
    DevBuf<RelaxTreeView> views; DevBuf<double> orig; DevBuf<int16_t> chain; DevBuf<uint8_t> flags; DevBuf<int32_t> num, sig, topo, parent, post_id;
    views.alloc((size_t)T); orig.alloc(cnt); chain.alloc(cnt); flags.alloc(cnt); num.alloc(cnt); sig.alloc(cnt); topo.alloc(cnt); parent.alloc(cnt);
    post_id.alloc((size_t)std::max<long long>(src.aux_stride, 1));
    g.v0 = v0.p; g.views = views.p; g.orig = orig.p; g.chain = chain.p; g.vflags = flags.p; g.number = num.p; g.sigma = sig.p; g.topo = topo.p; g.parent = parent.p; g.post_id = post_id.p;
    launch_ingest_fill(g, st);
    CK(cudaGetLastError());
    std::vector<double> h_bl(cnt); std::vector<uint8_t> h_fl(cnt); std::vector<int32_t> h_num(cnt), h_sig(cnt), h_par(cnt);
    CK(cudaMemcpyAsync(h_bl.data(), orig.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(h_fl.data(), flags.p, cnt, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_num.data(), num.p, cnt * 4, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(h_sig.data(), sig.p, cnt * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_par.data(), parent.p, cnt * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::vector<FlatTree> out((size_t)T);
    for (int t = 0; t < T; ++t) {
        const long long b = h_v0[(size_t)t]; const int n = (int)(h_v0[(size_t)t + 1] - b);
        if (n == 0) throw SgpError("source replicate " + std::to_string(first + t) + " has no " + (pruned ? "pruned" : "unpruned") + " tree");
        FlatTree& f = out[(size_t)t];
        f.n = n; f.root = n - 1;          // the root is last in post-order
        f.parent.assign(n, -1); f.first_child.assign(n, -1); f.next_sibling.assign(n, -1); f.n_children.assign(n, 0);
        f.name.resize(n); f.has_name.assign(n, 1); f.bl.assign(n, 0.0); f.has_bl.assign(n, 1); f.meta.resize(n); f.has_meta.assign(n, src.name_exclude_meta ? 0 : 1);
        f.host_tag.assign(n, 0);
        for (int x = 0; x < n; ++x) {
            const size_t k = (size_t)(b + x);
            f.parent[x] = x == n - 1 ? -1 : h_par[k];
            f.bl[x] = h_bl[k];
            f.name[x] = src.name_prefix + std::to_string(h_num[k]) + (src.name_append_sigma ? "_" + std::to_string(h_sig[k]) : std::string());
            // HOST= capture of io/NewickTreeReader.java:215-225: the tag's value; 0 for a vertex without any tag; -10 for tags without HOST= (species trees)
            f.host_tag[x] = src.name_exclude_meta ? 0 : (src.app == 0 ? -10 : h_sig[k]);
        }
        for (int x = 0; x < n - 1; ++x) {          // children in id order: the left child of a vertex precedes the right one in post-order
            const int p = f.parent[x];
            if (p < 0 || p >= n) throw SgpError("internal: bad parent link in an ingested tree");
            if (f.first_child[p] < 0) f.first_child[p] = x; else { int c = f.first_child[p]; while (f.next_sibling[c] >= 0) c = f.next_sibling[c]; f.next_sibling[c] = x; }
            ++f.n_children[p];
        }
        f.bfs.push_back(f.root);
        for (size_t i = 0; i < f.bfs.size(); ++i) for (int x = f.first_child[f.bfs[i]]; x >= 0; x = f.next_sibling[x]) f.bfs.push_back(x);
    }
    return out;
}

// ---- BranchRelaxer -----------------------------------------------------------------------------------------------------------
const double* Batch::values_host(int which) {
    if (which < 0 || which > 1 || !d_values[which]) return nullptr;
    if (h_values[which].empty() && n_values) {
        cudaSetDevice(device);
        h_values[which].resize((size_t)n_values);
        cudaMemcpy(h_values[which].data(), d_values[which], (size_t)n_values * sizeof(double), cudaMemcpyDeviceToHost);
    }
    return h_values[which].data();
}

void Context::run_relax(const RelaxConfig& cfg, const BatchOpts& opts, Batch& out, const Batch* src, bool src_pruned) {
    CK(cudaSetDevice(device));
    cudaStream_t st = impl->stream; g_alloc_stream = st; g_device = device;
    const long long n = opts.n;
    if (n <= 0) throw SgpError("replicate count must be positive");
    const bool replay = opts.rng_mode == 1;
    out.stream = (void*)st; out.n = n; out.device = device; out.app = 3; out.quiet = !cfg.has_outfile; out.out_prefix = cfg.outfile;
    Timings tm; EventTimer total(st), ph(st); total.start();
    g_h2d_bytes = 0;

    // tree templates: flattened host templates, or (chained stage) ingested on the device from a generator batch
    const bool chained = src != nullptr;
    if (chained && (!src->d_nodes || !src->d_info || src->device != device)) throw SgpError("source batch holds no trees on this device");
    const int T = chained ? (int)src->n : (int)cfg.trees.size();
    std::vector<RelaxTreeView> views(chained ? 0 : T); std::vector<double> orig; std::vector<int16_t> chain; std::vector<uint8_t> flags; std::vector<int32_t> name_off; std::string blob;
    long long sum_v = 0;
    DevBuf<RelaxTreeView> d_views; DevBuf<double> d_orig; DevBuf<int16_t> d_chain; DevBuf<uint8_t> d_flags; DevBuf<int32_t> d_name, d_num, d_sig, d_topo, d_parent;
    const bool autocorr = cfg.model >= RM_ACTK98;
    std::vector<int32_t> topo, parent;
    if (chained) {
        IngestArgs g{}; g.n_src = src->n; g.info = (const RepInfo*)src->d_info; g.node_off = src->d_node_off; g.nodes = (const Node*)src->d_nodes;
        g.aux = src->d_aux; g.aux_stride = src->aux_stride; g.pruned = src_pruned ? 1 : 0; g.keep_interior = cfg.keep_interior ? 1 : 0;
        DevBuf<long long> nv, v0; nv.alloc((size_t)T + 1); v0.alloc((size_t)T + 1);
        g.nv = nv.p; launch_ingest_count(g, st); ++tm.launches;
        size_t tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, nv.p, v0.p, T + 1, st);
        DevBuf<char> tmp; tmp.alloc(tb);
        cub::DeviceScan::ExclusiveSum(tmp.p, tb, nv.p, v0.p, T + 1, st); ++tm.launches;
        impl->post(0, v0.p + T, sizeof(long long)); CK(cudaStreamSynchronize(st));
        sum_v = impl->take<long long>(0);
        const size_t cnt = (size_t)std::max<long long>(sum_v, 1);
        d_views.alloc((size_t)T); d_orig.alloc(cnt); d_chain.alloc(cnt); d_flags.alloc(cnt); d_num.alloc(cnt); d_sig.alloc(cnt);
        g.v0 = v0.p; g.views = d_views.p; g.orig = d_orig.p; g.chain = d_chain.p; g.vflags = d_flags.p; g.number = d_num.p; g.sigma = d_sig.p;
        DevBuf<int32_t> post_id;
        if (autocorr) { d_topo.alloc(cnt); d_parent.alloc(cnt); post_id.alloc((size_t)std::max<long long>(src->aux_stride, 1)); g.topo = d_topo.p; g.parent = d_parent.p; g.post_id = post_id.p; }
        launch_ingest_fill(g, st); ++tm.launches;
        CK(cudaGetLastError()); CK(cudaStreamSynchronize(st));          // nv / v0 go out of scope here
    }
    for (int t = 0; t < T && !chained; ++t) {
        const RelaxTreeTemplate& tt = cfg.trees[t];
        views[t].nV = tt.nV; views[t].root = tt.root; views[t].ignore = tt.ignore; views[t].v0 = sum_v; sum_v += tt.nV;
        orig.insert(orig.end(), tt.orig.begin(), tt.orig.end()); chain.insert(chain.end(), tt.chain.begin(), tt.chain.end()); flags.insert(flags.end(), tt.flags.begin(), tt.flags.end());
        if (autocorr) { topo.insert(topo.end(), tt.topo.begin(), tt.topo.end()); parent.insert(parent.end(), tt.parent.begin(), tt.parent.end()); }
        for (int v = 0; v < tt.nV; ++v) { name_off.push_back((int32_t)blob.size()); name_off.push_back((int32_t)tt.names[v].size()); blob += tt.names[v]; }
    }
    const long long rounds = (n + T - 1) / T; const long long slots = rounds * sum_v;
    RelaxArgs a{};
    auto add = [&](const std::string& s, int& off, int& len) { off = (int)blob.size(); len = (int)s.size(); blob += s; };
    {
        std::string pre = "[", post; const bool split = replay && cfg.seed_arg_index >= 0;
        for (size_t i = 0; i < cfg.args.size(); ++i) {
            std::string& dst = (split && (int)i > cfg.seed_arg_index) ? post : pre;
            if ((int)i == cfg.seed_arg_index && split) { if (i) pre += ", "; continue; }
            if (i) dst += ", ";
            dst += cfg.args[i];
        }
        if (split) post += "]"; else pre += "]";
        add(pre, a.args_pre_off, a.args_pre_len); add(post, a.args_post_off, a.args_post_len); a.seed_mode = split ? 1 : 0; a.sseed = make_seed_spec(cfg.seed, blob);
        std::string tail = "Model:\t" + cfg.model_name + "\nModel parameters:\n";
        for (auto& kv : cfg.params) tail += "\t" + kv.first + "\t" + kv.second + "\n";
        add(tail, a.model_tail_off, a.model_tail_len);
    }
    DevBuf<double> d_samples, rates, relaxed; d_samples.upload(cfg.samples);
    DevBuf<int32_t> d_rk_off; DevBuf<double> d_rk_k, d_rk_theta, d_rk_w;
    if (cfg.model == RM_RK11) {
        if (chained || T != 1) throw SgpError("IIDRK11 relaxes one guest tree against one host tree: chained and multi-tree input are not supported");
        d_rk_off.upload(cfg.rk_off); d_rk_k.upload(cfg.rk_k); d_rk_theta.upload(cfg.rk_theta); d_rk_w.upload(cfg.rk_w);
        a.rk_off = d_rk_off.p; a.rk_k = d_rk_k.p; a.rk_theta = d_rk_theta.p; a.rk_w = d_rk_w.p;
    }
    if (!chained) { d_views.upload(views); d_orig.upload(orig); d_chain.upload(chain); d_flags.upload(flags); d_name.upload(name_off); if (autocorr) { d_topo.upload(topo); d_parent.upload(parent); } }
    else {
        a.syn_names = 1; a.syn_number = d_num.p; a.syn_sigma = d_sig.p; a.syn_append_sigma = src->name_append_sigma ? 1 : 0;
        if (src->name_prefix.size() > sizeof(a.syn_prefix)) throw SgpError("vertex prefix longer than 24 bytes is not supported");
        a.syn_prefix_len = (int)src->name_prefix.size(); memcpy(a.syn_prefix, src->name_prefix.data(), src->name_prefix.size());
    }
    DevBuf<char> d_blob; { std::vector<char> b(blob.begin(), blob.end()); b.push_back(0); d_blob.upload(b); }
    rates.alloc((size_t)slots); relaxed.alloc((size_t)slots);
    DevBuf<int32_t> attempts, status; attempts.alloc((size_t)n); status.alloc((size_t)n);
    CK(cudaMemsetAsync(attempts.p, 0, (size_t)n * sizeof(int32_t), st)); CK(cudaMemsetAsync(status.p, 0, (size_t)n * sizeof(int32_t), st));
    DevBuf<uint32_t> mt_state; if (replay) mt_state.alloc((size_t)((n + 127) / 128 * 128 / 32 + 1) * 624 * 32);
    a.n = n; a.rep_base = opts.offset; a.seed = cfg.seed.philox_key; a.model = cfg.model; a.pa = cfg.pa; a.pb = cfg.pb; a.pc = cfg.pc; a.pd = cfg.pd; a.sigma = cfg.sigma; a.gamma_lng = jln_gamma(0.5 * (2.0 * cfg.pa)); a.ln2 = jlog(2.0); a.topo = d_topo.p; a.parent = d_parent.p;
    a.min_rate = cfg.min_rate; a.max_rate = cfg.max_rate; a.max_attempts = cfg.max_attempts; a.samples = d_samples.p; a.n_samples = (int)cfg.samples.size();
    a.n_trees = T; a.sum_v = sum_v; a.trees = d_views.p; a.orig = d_orig.p; a.chain = d_chain.p; a.vflags = d_flags.p; a.name_off = d_name.p; a.blob = d_blob.p;
    a.T = impl->T; a.rates = rates.p; a.relaxed = relaxed.p; a.attempts = attempts.p; a.status = status.p; a.mt_state = mt_state.p;
    a.do_meta = cfg.do_meta; a.has_outfile = cfg.has_outfile;

    // ---- K5 rates + lengths ------------------------------------------------------------------------------------------------
    ph.start();
    if (replay || autocorr) {
        // one thread per replicate from attempt 0 (the autocorrelated models are sequential along the tree in either RNG mode)
        launch_relax_rates(replay, a, 0, 0, st); ++tm.launches;
    } else {
        if (chained) launch_relax_apply_per_rep(a, st); else launch_relax_apply(a, slots, st);          // attempt 0 in parallel
        tm.launches += 2;
        launch_relax_rates(false, a, 1, 1, st); ++tm.launches;       // flagged replicates continue from attempt 1
    }
    CK(cudaGetLastError());
    tm.find_ms = ph.stop();

    // ---- emission ------------------------------------------------------------------------------------------------------------
    const unsigned mask = opts.stream_mask ? opts.stream_mask : (cfg.has_outfile ? 3u : 1u);
    DevBuf<unsigned long long> sizes, sizes_n, offsets;
    sizes.alloc((size_t)2 * (n + 1)); sizes_n.alloc((size_t)2 * n); offsets.alloc((size_t)8 * (n + 1));
    CK(cudaMemsetAsync(sizes.p, 0, (size_t)2 * (n + 1) * sizeof(unsigned long long), st)); CK(cudaMemsetAsync(sizes_n.p, 0, (size_t)2 * n * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(offsets.p, 0, (size_t)8 * (n + 1) * sizeof(unsigned long long), st));
    DevBuf<uint16_t> plen_tree, plen_info; plen_tree.alloc((size_t)slots); plen_info.alloc((size_t)(3 * slots + 8 * n));
    DevBuf<unsigned long long> rdig_f; DevBuf<int16_t> rdig_e; rdig_f.alloc((size_t)slots); rdig_e.alloc((size_t)slots);
    a.plen_tree = plen_tree.p; a.plen_info = plen_info.p; a.rdig_f = rdig_f.p; a.rdig_e = rdig_e.p; a.sizes = sizes_n.p; a.offsets = offsets.p;
    a.has_outfile = (mask & 2u) ? 1 : 0;
    unsigned long long totals[2] = {0, 0};
    if (mask & 3u) {
        ph.start();
        launch_relax_emit(false, a, st); ++tm.launches; CK(cudaGetLastError());
        size_t tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, sizes.p, offsets.p, (int)(n + 1), st);
        DevBuf<char> tmp; tmp.alloc(tb);
        for (int s = 0; s < 2; ++s) {
            CK(cudaMemcpyAsync(sizes.p + (size_t)s * (n + 1), sizes_n.p + (size_t)s * n, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
            cub::DeviceScan::ExclusiveSum(tmp.p, tb, sizes.p + (size_t)s * (n + 1), offsets.p + (size_t)s * (n + 1), (int)(n + 1), st); ++tm.launches;
        }
        for (int s = 0; s < 2; ++s) impl->post((size_t)s * 8, offsets.p + (size_t)s * (n + 1) + n, sizeof(unsigned long long));
        CK(cudaStreamSynchronize(st));
        for (int s = 0; s < 2; ++s) totals[s] = impl->take<unsigned long long>((size_t)s * 8);
        tm.emit_size_ms = ph.stop();
        for (int s = 0; s < 2; ++s) { out.stream_bytes[s] = totals[s]; impl->alloc_output(&out.d_stream[s], totals[s], st); a.out[s] = out.d_stream[s]; }
        ph.start();
        launch_relax_emit(true, a, st); ++tm.launches; CK(cudaGetLastError());
        tm.emit_write_ms = ph.stop();
    }
    out.stream_mask = mask & 3u;
    out.suffix[0] = ""; out.suffix[1] = ".info";
    // per-replicate scalars
    // (the per-replicate tree statistics of a relaxer batch are all zero: they are materialised only if somebody asks, see fetch_stats -
    // zero-filling 200 MB of host vectors per 2*10^6 replicates took half of a chained step)
    out.status.resize((size_t)n); out.attempts.resize((size_t)n); out.sampled.clear();
    CK(cudaMemcpyAsync(out.status.data(), status.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out.attempts.data(), attempts.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    out.have_info = true; out.have_stats = false;
    Summary& S = out.summary;
    for (long long i = 0; i < n; ++i) { if (out.status[i] == REP_OK) ++S.n_ok; else ++S.n_failed; S.attempts_total += out.attempts[i]; }
    S.vertices_sampled = slots; S.out_bytes[0] = (long long)totals[0]; S.out_bytes[1] = (long long)totals[1];
    out.d_offsets = offsets.release();
    out.d_values[0] = rates.release(); out.d_values[1] = relaxed.release(); out.n_values = slots;
    tm.total_ms = total.stop(); tm.h2d_bytes = (long long)g_h2d_bytes; out.timings = tm;
    if (!opts.keep_on_device) { for (int s = 0; s < 2; ++s) if (mask & (1u << s)) out.stream_host(s); out.fetch_offsets(); }
}

void Context::probe_d2s(const double* v, long long n, char* out) {
    CK(cudaSetDevice(device));
    DevBuf<double> dv; dv.alloc((size_t)n); DevBuf<char> dout; dout.alloc((size_t)n * 32);
    CK(cudaMemcpy(dv.p, v, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    k_probe_d2s<<<(unsigned)((n + 255) / 256), 256, 0, impl->stream>>>(dv.p, n, dout.p, impl->T);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout.p, (size_t)n * 32, cudaMemcpyDeviceToHost, impl->stream)); CK(cudaStreamSynchronize(impl->stream));
}
void Context::probe_math(int fn, const double* v, long long n, double* out) {
    CK(cudaSetDevice(device));
    DevBuf<double> dv, dout; dv.alloc((size_t)n); dout.alloc((size_t)n);
    CK(cudaMemcpy(dv.p, v, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    k_probe_math<<<(unsigned)((n + 255) / 256), 256, 0, impl->stream>>>(fn, dv.p, n, dout.p, impl->T);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, impl->stream)); CK(cudaStreamSynchronize(impl->stream));
}
void Context::probe_mt(const SeedHost& seed, long long offset, long long n_words, uint32_t* out) {
    CK(cudaSetDevice(device));
    DevBuf<uint32_t> state, dout; state.alloc(624 * 32); dout.alloc((size_t)n_words);
    std::string unused; k_probe_mt<<<1, 32, 0, impl->stream>>>(make_seed_spec(seed, unused), offset, n_words, dout.p, state.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout.p, (size_t)n_words * sizeof(uint32_t), cudaMemcpyDeviceToHost, impl->stream)); CK(cudaStreamSynchronize(impl->stream));
}
void Context::probe_philox(int variant, const uint32_t* ctr, const uint32_t* key, long long n, uint32_t* out) {
    CK(cudaSetDevice(device));
    if (n <= 0) return;
    DevBuf<uint32_t> dc, dk, dout; dc.alloc((size_t)n * 4); dk.alloc((size_t)n * 2); dout.alloc((size_t)n * 4);
    CK(cudaMemcpy(dc.p, ctr, (size_t)n * 16, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dk.p, key, (size_t)n * 8, cudaMemcpyHostToDevice));
    k_probe_philox<<<(unsigned)((n + 127) / 128), 128, 0, impl->stream>>>(variant, dc.p, dk.p, n, dout.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout.p, (size_t)n * 16, cudaMemcpyDeviceToHost, impl->stream)); CK(cudaStreamSynchronize(impl->stream));
}
void Context::probe_issue(double* dfma, double* imad, double* ddep) {
    CK(cudaSetDevice(device));
    const int blocks = impl->sm_count * 8, threads = 256, iters = 20000;
    DevBuf<double> od; od.alloc((size_t)blocks * threads); DevBuf<uint32_t> oi; oi.alloc((size_t)blocks * threads);
    EventTimer t(impl->stream);
    k_peak_dfma<<<blocks, threads, 0, impl->stream>>>(od.p, 100); CK(cudaStreamSynchronize(impl->stream));
    t.start(); k_peak_dfma<<<blocks, threads, 0, impl->stream>>>(od.p, iters); double ms = t.stop();
    *dfma = (double)blocks * threads / 32.0 * iters * 8.0 / (ms * 1e-3) * 1e-9;
    t.start(); k_peak_imad<<<blocks, threads, 0, impl->stream>>>(oi.p, iters); ms = t.stop();
    *imad = (double)blocks * threads / 32.0 * iters * 8.0 / (ms * 1e-3) * 1e-9;
    // 4 IMAD + 4 (SHF + LOP3) per iteration = 12 issued instructions per warp
    t.start(); k_peak_mixed<<<blocks, threads, 0, impl->stream>>>(oi.p, iters); ms = t.stop();
    *ddep = (double)blocks * threads / 32.0 * iters * 12.0 / (ms * 1e-3) * 1e-9;
    CK(cudaGetLastError());
}

}  // namespace sgp
