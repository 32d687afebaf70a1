// K4: text emission. Every output file ("stream") of a replicate is a concatenation of independent pieces whose byte
// offsets are an exclusive prefix sum of their lengths:
//   *.tree       one piece per vertex in post-order:  ['(' x run | ')'] name ':' length [tag] [','] [tree tag ';' '\n']
//   *.info       one piece per line
//   *.guest2host constant header + one row per vertex in breadth-first order
//   *.leafmap    one row per leaf in breadth-first order
// Sizes:  the label kernel (kernels_label.cuh) leaves one emission record per piece, in the order the piece is written, with
//         the piece's length and the shortest-decimal digits of its floating-point field, and the size of every tree / row stream
//         of the replicate; k_emit_info<false> adds the sizes of the .info files.
// Write:  k_write_tree / k_write_rows, one warp per replicate, 32 pieces per trip: the lanes load 32 consecutive records
//         (coalesced, the next trip's records are in flight while this trip is formatted) -> warp exclusive scan of the lengths ->
//         each lane formats its piece once into a shared-memory stage laid out congruent (mod 16) to the destination -> the
//         stage leaves as one bulk shared->global copy (cp.async.bulk, TMA) plus byte stores for the unaligned head and tail,
//         double-buffered so that the next trip is formatted while the copy drains.
//
// Formats: emit_format.cuh; HostTreeGen.getInfo (apps/SaGePhy/HostTreeGen.java:99-110,124-174),
// GuestTreeInHostTreeCreator.getInfo (:431-511) and the file writers in GuestTreeGen.java:66-101.
#pragma once
#include "emit_format.cuh"

namespace sgp {

template <class Sink> __device__ void put_kv_int(Sink& s, const char* k, long long v) { put_str(s, k); put_i64(s, v); s.put('\n'); }
template <class Sink> __device__ void put_kv_dbl(Sink& s, const EmitArgs& a, const char* k, double v) { put_str(s, k); put_double(s, a.T, v); s.put('\n'); }

__device__ __forceinline__ int info_lines(const EmitArgs& a, bool pruned) { return a.app == 0 ? (pruned ? 8 : 10) : a.app == 1 ? 18 : 30; }

// line `ln` of an .info file. Lanes format different lines in the same trip, so the line-specific part only SELECTS data
// (key string, integer or double operand); the expensive work (round8, shortest-digit conversion) runs on one shared path.
template <bool PRUNED, class Sink> __device__ void put_info_line(Sink& s, const EmitArgs& a, const RepInfo& inf, long long gid, int ln) {
    const int32_t* c = PRUNED ? inf.cnt_p : inf.cnt_u;
    if (ln == 0) { put_str(s, a.app == 0 ? "# HOSTTREEGEN\n" : a.app == 1 ? "# GUESTTREEGEN\n" : "# DOMAINTREEGEN\n"); return; }
    if (ln == 1) {
        put_lit(s, "Arguments:\t");
        put_blob(s, a.blob, a.args_pre_off, a.args_pre_len);
        if (a.seed_mode) { put_seed(s, a.sseed, a.blob, (unsigned long long)gid); put_blob(s, a.blob, a.args_post_off, a.args_post_len); }
        s.put('\n'); return;
    }
    const char* key = ""; bool is_double = false; long long iv = 0; double num = 0.0; bool ratio = false;
    if (ln == 2) { key = "Attempts:\t"; iv = inf.attempts; }
    else if (ln == 3) { key = "No. of vertices:\t"; iv = PRUNED ? inf.n_p : inf.n_u; }
    else if (ln == 4) { key = "No. of extant leaves:\t"; iv = c[EV_LEAF] + c[EV_UNSAMPLED_LEAF]; }
    else if (a.app == 0) {
        const int births = c[EV_DUPLICATION] + c[EV_SPECIATION], deaths = c[EV_LOSS];
        switch (ln) {
            case 5: key = "No. of births:\t"; iv = births; break;
            case 6: key = "No. of deaths:\t"; iv = deaths; break;
            case 7: key = "Total branch time:\t"; is_double = true; num = PRUNED ? inf.total_p : inf.total_u; break;
            case 8: key = "Birth ML estimate:\t"; is_double = true; ratio = true; num = (double)births; break;
            default: key = "Death ML estimate:\t"; is_double = true; ratio = true; num = (double)deaths; break;
        }
    } else if (a.app == 2) {
        // DomainTreeInHostTreeCreator.getInfo :1128-1252
        static const char* const kCnt[8] = {"No. of additive transfers intragene intraspecies:\t", "No. of additive transfers intragene interspecies:\t",
            "No. of additive transfers intergene intraspecies:\t", "No. of additive transfers intergene interspecies:\t",
            "No. of replacing transfers intragene intraspecies:\t", "No. of replacing transfers intragene interspecies:\t",
            "No. of replacing transfers intergene intraspecies:\t", "No. of replacing transfers intergene interspecies:\t"};
        static const char* const kMl[8] = {"Additive Transfer Intragene Intraspecies ML estimate:\t", "Additive Transfer Intragene Interspecies ML estimate:\t",
            "Additive Transfer Intergene Intraspecies ML estimate:\t", "Additive Transfer Intergene Interspecies ML estimate:\t",
            "Replacing Transfer Intragene Intraspecies ML estimate:\t", "Replacing Transfer Intragene Interspecies ML estimate:\t",
            "Replacing Transfer Intergene Intraspecies ML estimate:\t", "Replacing Transfer Intergene Interspecies ML estimate:\t"};
        if (ln == 5) { key = "No. of co-divergences:\t"; iv = c[EV_CODIVERGENCE]; }
        else if (ln == 6) { key = "No. of duplications:\t"; iv = c[EV_DUPLICATION]; }
        else if (ln == 7) { key = "No. of losses:\t"; iv = c[EV_LOSS]; }
        else if (ln == 8) { key = "No. of replacing losses:\t"; iv = c[EV_REPLACING_LOSS]; }
        else if (ln <= 16) { key = kCnt[ln - 9]; iv = c[EV_AT_GG_SS + ln - 9]; }
        else if (ln == 17) { key = "Total branch time:\t"; is_double = true; num = PRUNED ? inf.total_p : inf.total_u; }
        else if (ln == 18) { key = "Total branch time beneath host stem:\t"; is_double = true; num = PRUNED ? inf.beneath_p : inf.beneath_u; }
        else if (ln == 19) { key = "Duplication ML estimate:\t"; is_double = true; ratio = true; num = (double)c[EV_DUPLICATION]; }
        else if (ln == 20) { key = "Loss ML estimate:\t"; is_double = true; ratio = true; num = (double)c[EV_LOSS]; }
        else if (ln == 21) { key = "Replacing Loss ML estimate:\t"; is_double = true; ratio = true; num = (double)c[EV_REPLACING_LOSS]; }
        else { key = kMl[ln - 22]; is_double = true; ratio = true; num = (double)c[EV_AT_GG_SS + ln - 22]; }
    } else {
        switch (ln) {
            case 5: key = "No. of speciations:\t"; iv = c[EV_SPECIATION]; break;
            case 6: key = "No. of duplications:\t"; iv = c[EV_DUPLICATION]; break;
            case 7: key = "No. of losses:\t"; iv = c[EV_LOSS]; break;
            case 8: key = "No. of replacing losses:\t"; iv = c[EV_REPLACING_LOSS]; break;
            case 9: key = "No. of additive transfers:\t"; iv = c[EV_ADDITIVE_TRANSFER]; break;
            case 10: key = "No. of replacing transfers:\t"; iv = c[EV_REPLACING_TRANSFER]; break;
            case 11: key = "Total branch time:\t"; is_double = true; num = PRUNED ? inf.total_p : inf.total_u; break;
            case 12: key = "Total branch time beneath host stem:\t"; is_double = true; num = PRUNED ? inf.beneath_p : inf.beneath_u; break;
            case 13: key = "Duplication ML estimate:\t"; is_double = true; ratio = true; num = (double)c[EV_DUPLICATION]; break;
            case 14: key = "Loss ML estimate:\t"; is_double = true; ratio = true; num = (double)c[EV_LOSS]; break;
            case 15: key = "Replacing Loss ML estimate:\t"; is_double = true; ratio = true; num = (double)c[EV_REPLACING_LOSS]; break;
            case 16: key = "Additive Transfer ML estimate:\t"; is_double = true; ratio = true; num = (double)c[EV_ADDITIVE_TRANSFER]; break;
            default: key = "Replacing Transfer ML estimate:\t"; is_double = true; ratio = true; num = (double)c[EV_REPLACING_TRANSFER]; break;
        }
    }
    put_str(s, key);
    if (is_double) {
        if (ratio) num = XDIV(num, round_sig(a.T, PRUNED ? inf.total_p : inf.total_u, 8));
        put_double(s, a.T, round_sig(a.T, num, 8));
    } else put_i64(s, iv);
    s.put('\n');
}

__device__ __forceinline__ unsigned warp_excl_scan_u32(unsigned v, unsigned* total) {
    unsigned x = v; const int lane = threadIdx.x & 31;
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

constexpr int kEmitWarps = 4;               // warps per block
constexpr int kStageBytes = 3072;           // one staging buffer (32 pieces; longer groups fall back to direct stores); two per warp

// Flushes `total` staged bytes (stage[mis .. mis+total), mis = dst & 15) to dst with aligned 16-byte stores.
__device__ __forceinline__ void flush_stage(const char* stage, int mis, char* dst, unsigned total) {
    const int lane = threadIdx.x & 31;
    unsigned head = (16u - (unsigned)mis) & 15u; if (head > total) head = total;
    if ((unsigned)lane < head) dst[lane] = stage[mis + lane];
    const unsigned body = (total - head) >> 4;
    const uint4* src4 = reinterpret_cast<const uint4*>(stage + mis + head);
    uint4* dst4 = reinterpret_cast<uint4*>(dst + head);
    for (unsigned i = lane; i < body; i += 32) dst4[i] = src4[i];
    const unsigned done = head + (body << 4);
    if (done + lane < total) dst[done + lane] = stage[mis + done + lane];
}

// Pieces 0..n_pieces-1 of one stream of one replicate, lengths cached by the caller (BranchRelaxer emitter, kernels_relax.cuh).
//   SIZE pass : len(i) computed with a CountSink by `piece(CountSink&, i, true)`, stored through store_len(i, len).
//   WRITE pass: len(i) read through load_len(i); bytes produced by `piece(MemSink&, i, false)`.
template <bool WRITE, class F, class LL, class SL>
__device__ unsigned long long emit_pieces(int n_pieces, char* base, char* stage, F&& piece, LL&& load_len, SL&& store_len) {
    const int lane = threadIdx.x & 31;
    unsigned long long running = 0;
    for (int i0 = 0; i0 < n_pieces; i0 += 32) {
        const int i = i0 + lane;
        unsigned len = 0;
        if (i < n_pieces) {
            if (WRITE) {
                len = load_len(i);
                // the cache is 16 bits wide and saturates: a piece that long (a huge name or prefix) is measured again here
                if (len == 0xffffu) { CountSink c; piece(c, i); len = (unsigned)c.n; }
            } else { CountSink c; piece(c, i); len = (unsigned)c.n; store_len(i, len); }
        }
        unsigned total; const unsigned off = warp_excl_scan_u32(len, &total);
        if (WRITE && total) {
            char* dst = base + running;
            const int mis = (int)(reinterpret_cast<unsigned long long>(dst) & 15ull);
            if (total + 16 <= (unsigned)kStageBytes) {
                if (len) { SmemSink m{(uint32_t)__cvta_generic_to_shared(stage) + (uint32_t)mis + off}; piece(m, i); }
                __syncwarp();
                flush_stage(stage, mis, dst, total);
                __syncwarp();
            } else if (len) { MemSink m{dst + off}; piece(m, i); }
        }
        running += total;
    }
    return running;
}

// .info streams: one THREAD per (replicate, file). A file is ~10-30 short lines, so "lane per line" left most of a warp idle and
// cost ~8 000 warp instructions per replicate; here the 32 lanes of a warp format 32 files in lock step (same line number at the
// same time) into a shared-memory stage that the lanes share by a prefix sum of their file sizes (as many lanes per round as
// fit), and the warp then flushes the files one after the other with aligned 16-byte stores. A long "Arguments:" line (it may
// embed a whole Newick host tree) is not staged: the whole warp copies it straight from the constant blob, file by file.
constexpr int kInfoWarps = 4;
constexpr int kInfoStage = 12 * 1024;          // bytes of stage per warp
template <bool PRUNED, class Sink> __device__ __forceinline__ void put_info_file(Sink& s, const EmitArgs& a, const RepInfo& inf, long long gid, int first_line) {
    const int L = info_lines(a, PRUNED);
    for (int ln = first_line; ln < L; ++ln) put_info_line<PRUNED>(s, a, inf, gid, ln);
}
template <bool WRITE>
__global__ void __launch_bounds__(kInfoWarps * 32) k_emit_info(EmitArgs a) {
    __shared__ __align__(16) char s_stage[kInfoWarps][kInfoStage];
    const int lane = threadIdx.x & 31;
    char* stage = s_stage[threadIdx.x >> 5];
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = item >> 1; const bool pr = (item & 1) != 0;
    const int st = pr ? ST_PRUNED_INFO : ST_UNPRUNED_INFO;
    const bool active = r < a.n && (a.stream_mask & (1u << st));
    const long long gid = a.rep_base + (active ? r : 0);
    const RepInfo& inf = a.info[active ? r : 0];
    const bool ok = active && inf.status == REP_OK;
    const bool note = active && !ok && pr && inf.status == REP_MAX_ATTEMPTS;      // the reference writes only the failure note
    // header = lines 0-1 ("# APP" and "Arguments: [...]"): closed-form length; copied by the whole warp when it is long
    const bool long_head = a.args_pre_len > 192;
    const int h0len = a.app == 0 ? 25 : a.app == 1 ? 26 : 27;
    const int seedlen = (ok && a.seed_mode) ? seed_dec_len(a.sseed, a.blob, (unsigned long long)gid) : 0;
    const unsigned head = ok ? (unsigned)(h0len + a.args_pre_len + (a.seed_mode ? seedlen + a.args_post_len : 0) + 1) : 0u;
    unsigned body = 0;             // what this lane formats itself
    if (ok) { CountSink c; if (pr) put_info_file<true>(c, a, inf, gid, long_head ? 2 : 0); else put_info_file<false>(c, a, inf, gid, long_head ? 2 : 0); body = (unsigned)c.n; }
    else if (note) body = 65;
    const unsigned long long total = (unsigned long long)body + (long_head ? head : 0u);
    if (!WRITE) { if (active) a.sizes[(size_t)st * a.stride + r] = total; return; }
    const bool room = active && a.offsets[(size_t)st * a.stride + r + 1] <= a.cap[st];
    if (active && !room) { *a.overflow = 1; body = 0; }
    char* dst = room ? a.out[st] + a.offsets[(size_t)st * a.stride + r] : nullptr;
    if (long_head) {
        unsigned todo = __ballot_sync(0xffffffffu, ok && room);
        while (todo) {
            const int l = __ffs(todo) - 1; todo &= todo - 1;
            char* d = reinterpret_cast<char*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(dst), l));
            const long long g = __shfl_sync(0xffffffffu, gid, l);
            const char* h0 = a.app == 0 ? "# HOSTTREEGEN\nArguments:\t" : a.app == 1 ? "# GUESTTREEGEN\nArguments:\t" : "# DOMAINTREEGEN\nArguments:\t";
            for (int i = lane; i < h0len; i += 32) d[i] = h0[i];
            if (a.args_pre_off16[0] >= 0) {
                char* q = d + h0len; const int m = (int)(reinterpret_cast<unsigned long long>(q) & 15ull);
                flush_stage(a.blob + a.args_pre_off16[m] - m, m, q, (unsigned)a.args_pre_len);       // source congruent to destination mod 16
            } else for (int i = lane; i < a.args_pre_len; i += 32) d[h0len + i] = a.blob[a.args_pre_off + i];
            unsigned long long o = (unsigned long long)h0len + a.args_pre_len;
            if (a.seed_mode) {
                const int sl = __shfl_sync(0xffffffffu, seedlen, l);
                if (lane == 0) { MemSink ms{d + o}; put_seed(ms, a.sseed, a.blob, (unsigned long long)g); }
                for (int i = lane; i < a.args_post_len; i += 32) d[o + sl + i] = a.blob[a.args_post_off + i];
                o += (unsigned long long)sl + a.args_post_len;
            }
            if (lane == 0) d[o] = '\n';
        }
        if (ok && room) dst += head;
    }
    // bodies: rounds of as many lanes as fit the stage; slot of a lane = [off, off + mis + body), congruent to its destination mod 16
    const int mis = (int)(reinterpret_cast<unsigned long long>(dst) & 15ull);
    const unsigned need = body ? ((body + (unsigned)mis + 15u) & ~15u) + 16u : 0u;
    bool pending = body > 0;
    for (;;) {
        const unsigned want = __ballot_sync(0xffffffffu, pending);
        if (!want) break;
        // inclusive prefix sum of the slot sizes of the pending lanes
        unsigned x = pending ? need : 0u;
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        const bool fits = pending && x <= (unsigned)kInfoStage;
        const unsigned now = __ballot_sync(0xffffffffu, fits);
        if (!now) {
            // the first pending file alone exceeds the stage: its owner writes it directly (byte stores)
            const int l = __ffs(want) - 1;
            if (lane == l) {
                MemSink ms{dst};
                if (note) put_lit(ms, "Failed to produce valid pruned tree within max allowed attempts.\n");
                else if (pr) put_info_file<true>(ms, a, inf, gid, long_head ? 2 : 0); else put_info_file<false>(ms, a, inf, gid, long_head ? 2 : 0);
                pending = false;
            }
            continue;
        }
        const unsigned off = x - need;
        if (fits) {
            SmemSink m{(uint32_t)__cvta_generic_to_shared(stage) + off + (uint32_t)mis};
            if (note) put_lit(m, "Failed to produce valid pruned tree within max allowed attempts.\n");
            else if (pr) put_info_file<true>(m, a, inf, gid, long_head ? 2 : 0); else put_info_file<false>(m, a, inf, gid, long_head ? 2 : 0);
        }
        __syncwarp();
        unsigned todo = now;
        while (todo) {
            const int l = __ffs(todo) - 1; todo &= todo - 1;
            char* d = reinterpret_cast<char*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(dst), l));
            const unsigned len = __shfl_sync(0xffffffffu, body, l);
            const unsigned so = __shfl_sync(0xffffffffu, off, l);
            const int m = __shfl_sync(0xffffffffu, mis, l);
            flush_stage(stage + so, m, d, len);
        }
        __syncwarp();
        if (fits) pending = false;
    }
}


// ---- record-driven write kernels ---------------------------------------------------------------------------------------------------
// Bulk asynchronous copy shared -> global (TMA, cp.async.bulk): both addresses 16-byte aligned, size a multiple of 16. The
// issuing thread owns the bulk group; "wait_group.read N" returns when all but the N most recent groups have finished READING
// shared memory, i.e. when their stage buffers may be overwritten.
__device__ __forceinline__ void bulk_store(char* dst, uint32_t src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Flushes `total` staged bytes (stage[mis .. mis+total), mis = dst & 15): the 16-byte aligned body as one bulk copy issued by lane
// 0, the unaligned head and tail (< 16 bytes each) as byte stores. Every lane that wrote the stage calls this (proxy fence).
template <bool BULK> __device__ __forceinline__ void flush_stage_async(const char* stage, int mis, char* dst, unsigned total) {
    if (!BULK) { flush_stage(stage, mis, dst, total); return; }
    const int lane = threadIdx.x & 31;
    fence_async_smem();
    __syncwarp();
    unsigned head = (16u - (unsigned)mis) & 15u; if (head > total) head = total;
    const unsigned body = (total - head) & ~15u;
    if (lane == 0 && body) bulk_store(dst + head, (uint32_t)__cvta_generic_to_shared(stage + mis + head), body);
    if ((unsigned)lane < head) dst[lane] = stage[mis + lane];
    const unsigned done = head + body;
    if (done + lane < total) dst[done + lane] = stage[mis + done + lane];
}

// Pieces [0, n) of one stream of one replicate, from their records. `piece(sink, rec, i)` prints piece i, `len_of(rec)` is its
// cached length. Returns the bytes written. `trip` counts the stage buffers this warp has used (double buffering across calls).
template <bool BULK, class Rec, class F, class L>
__device__ __forceinline__ unsigned long long emit_records(int n, const Rec* recs, char* base, char* stage2, unsigned stage_bytes, unsigned& trip, F&& piece, L&& len_of) {
    const int lane = threadIdx.x & 31;
    unsigned long long running = 0;
    Rec cur; { const uint4 z = make_uint4(0, 0, 0, 0); uint4* c4 = reinterpret_cast<uint4*>(&cur); c4[0] = z; c4[1] = z; }
    if (lane < n) { const uint4* s4 = reinterpret_cast<const uint4*>(recs + lane); uint4* c4 = reinterpret_cast<uint4*>(&cur); c4[0] = s4[0]; c4[1] = s4[1]; }
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        Rec nxt = cur;
        if (i + 32 < n) { const uint4* s4 = reinterpret_cast<const uint4*>(recs + i + 32); uint4* c4 = reinterpret_cast<uint4*>(&nxt); c4[0] = s4[0]; c4[1] = s4[1]; }
        unsigned len = 0;
        if (i < n) {
            len = len_of(cur);
            if (len == 0xffffu) { CountSink c; piece(c, cur, i); len = (unsigned)c.n; }       // saturated cache: a huge name or prefix
        }
        unsigned total; const unsigned off = warp_excl_scan_u32(len, &total);
        if (total) {
            char* dst = base + running;
            const int mis = (int)(reinterpret_cast<unsigned long long>(dst) & 15ull);
            if (total + 16 <= stage_bytes) {
                char* stage = stage2 + (trip & 1u) * stage_bytes;
                if (BULK && trip >= 2) { if (lane == 0) bulk_wait_read<1>(); __syncwarp(); }       // the copy that last read this buffer is done
                if (len) { SmemSink m{(uint32_t)__cvta_generic_to_shared(stage) + (uint32_t)mis + off}; piece(m, cur, i); }
                if (!BULK) __syncwarp();
                flush_stage_async<BULK>(stage, mis, dst, total);
                if (!BULK) __syncwarp();
                ++trip;
            } else if (len) { MemSink m{dst + off}; piece(m, cur, i); }
        }
        running += total;
        cur = nxt;
    }
    return running;
}

// Finishes the raw records the label kernels left (digits of the floating-point field, length of the piece) and sums the sizes of
// the tree and row streams of every replicate: warp per replicate, records read and written back in place, coalesced.
__device__ __forceinline__ long long warp_sum_ll(long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void load_rec32(void* dst, const void* src) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src); uint4* d4 = reinterpret_cast<uint4*>(dst);
    d4[0] = s4[0]; d4[1] = s4[1];
}
template <int INSTANCE> __global__ void __launch_bounds__(256) k_finish_records(EmitArgs a) {      // (template: the header is included by two translation units)
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= a.n) return;
    const RepInfo& inf = a.info[r];
    const int status = inf.status, n_u = inf.n_u, n_p = inf.n_p, p_root_ = inf.p_root;       // (independent loads first, see k_write_tree)
    const long long off = a.node_off[r];
    if (status != REP_OK) return;
    const unsigned mask = a.stream_mask;
    const int up = n_u, nu = p_root_ >= 0 ? n_p : 0;
    unsigned long long* sz = a.sizes + r;
#pragma unroll 1
    for (int view = 0; view < 2; ++view) {          // (not unrolled: the body is ~2 000 instructions and the instruction cache 32 KB)
        const int n = view ? nu : up;
        if (mask & (1u << (view ? ST_PRUNED_TREE : ST_UNPRUNED_TREE))) {
            TreeRec* recs = a.trec + (size_t)view * a.aux_stride + off;
            const TreeRecX* recx = a.trecx ? a.trecx + (size_t)view * a.aux_stride + off : nullptr;
            const TreeRec* twin = (view && (mask & (1u << ST_UNPRUNED_TREE))) ? a.trec + off : nullptr;       // finished in the previous round of this loop
            long long sum = 0;
            for (int k = lane; k < n; k += 32) {
                TreeRec t; load_rec32(&t, recs + k);
                sum += view ? finish_tree_rec<true>(a, t, recx + k, twin) : finish_tree_rec<false>(a, t, recx + k, nullptr);
                load_rec32(recs + k, &t);
            }
            sum = warp_sum_ll(sum);
            // an empty pruned tree still prints NewickTree.toString of a null root
            if (view && inf.p_root < 0) sum = a.exclude_meta ? 2 : (a.app == 0 ? 18 : 19);
            if (lane == 0) sz[(size_t)(view ? ST_PRUNED_TREE : ST_UNPRUNED_TREE) * a.stride] = (unsigned long long)sum;
        }
        const bool wg = mask & (1u << (view ? ST_PRUNED_G2H : ST_UNPRUNED_G2H)), wl = mask & (1u << (view ? ST_PRUNED_LEAFMAP : ST_UNPRUNED_LEAFMAP));
        if ((wg || wl) && a.rows) {
            RowRec* recs = a.rows + (size_t)view * a.aux_stride + off;
            const bool twin_rows = view && (mask & ((1u << ST_UNPRUNED_G2H) | (1u << ST_UNPRUNED_LEAFMAP)));
            const RowRec* twin = twin_rows ? a.rows + off : nullptr;
            long long sg = 0, sl = 0;
            for (int k = lane; k < n; k += 32) {
                RowRec t; load_rec32(&t, recs + k);
                long long lg, ll; finish_row_rec(a, t, wg, wl, twin, (mask & (1u << ST_UNPRUNED_G2H)) != 0, &lg, &ll); sg += lg; sl += ll;
                load_rec32(recs + k, &t);
            }
            sg = warp_sum_ll(sg); sl = warp_sum_ll(sl);
            if (lane == 0) {
                // DomainTreeGen prints the gene tree of the ROOT vertex as "Host tree" (getSigma :1279)
                if (wg) sz[(size_t)(view ? ST_PRUNED_G2H : ST_UNPRUNED_G2H) * a.stride] = (unsigned long long)(sg + (a.app == 2 ? a.g2h_gene_off[2 * (view ? inf.root_gtree_p : inf.root_gtree_u) + 1] : a.g2h_head_len));
                if (wl) sz[(size_t)(view ? ST_PRUNED_LEAFMAP : ST_UNPRUNED_LEAFMAP) * a.stride] = (unsigned long long)sl;
            }
        }
        __syncwarp();          // the pruned view reads what other lanes finished in the unpruned view
    }
}

// One Newick stream (VIEW 0: unpruned, 1: pruned): warp per replicate. Everything the warp needs to find its records (status and
// vertex counts, stream offsets, first record) is loaded up front in ONE round trip: as a chain of dependent loads behind early
// returns it cost four round trips per replicate, as long as formatting the replicate's seven trips.
template <int VIEW, bool BULK>
__global__ void __launch_bounds__(kEmitWarps * 32, 8) k_write_tree(EmitArgs a) {
    extern __shared__ __align__(128) char s_stage_dyn[];          // two stage buffers of a.stage_bytes per warp (sized by the launcher from the batch's mean piece length)
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= a.n) return;
    constexpr int st = VIEW ? ST_PRUNED_TREE : ST_UNPRUNED_TREE;
    const RepInfo& inf = a.info[r];
    const int status = inf.status, n_rec = VIEW ? inf.n_p : inf.n_u, p_root = inf.p_root;
    const unsigned long long o0 = a.offsets[(size_t)st * a.stride + r], o1 = a.offsets[(size_t)st * a.stride + r + 1];
    const long long off = a.node_off[r];
    if (status != REP_OK) return;
    if (o1 > a.cap[st]) { if (lane == 0) *a.overflow = 1; return; }
    char* base = a.out[st] + o0;
    char* stage2 = s_stage_dyn + (size_t)(threadIdx.x >> 5) * 2 * a.stage_bytes;
    if (VIEW == 1 && p_root < 0) {
        // empty pruned tree: NewickTree.toString of a null root
        if (lane == 0) { MemSink m{base}; if (a.exclude_meta) put_lit(m, ";\n"); else if (a.app == 0) put_lit(m, "[NAME=PrunedTree];"); else put_lit(m, "[NAME=PrunedTree];\n"); }
        return;
    }
    const TreeRec* recs = a.trec + (size_t)VIEW * a.aux_stride + off;
    const TreeRecX* recx = a.trecx ? a.trecx + (size_t)VIEW * a.aux_stride + off : nullptr;
    unsigned trip = 0;
    emit_records<BULK>(n_rec, recs, base, stage2, (unsigned)a.stage_bytes, trip,
                       [&](auto& s, const TreeRec& t, int i) { put_tree_piece<VIEW == 1>(s, a, t, recx + i); },
                       [](const TreeRec& t) { return (unsigned)t.len; });
    if (BULK && lane == 0) bulk_wait_read<0>();         // shared memory must outlive the copies that read it
}

// One row stream (KIND 0: .guest2host, 1: .leafmap; VIEW 0: unpruned, 1: pruned): warp per replicate.
template <int KIND, int VIEW, bool BULK>
__global__ void __launch_bounds__(kEmitWarps * 32, 8) k_write_rows(EmitArgs a) {
    extern __shared__ __align__(128) char s_stage_dyn[];
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= a.n) return;
    constexpr int st = KIND == 0 ? (VIEW ? ST_PRUNED_G2H : ST_UNPRUNED_G2H) : (VIEW ? ST_PRUNED_LEAFMAP : ST_UNPRUNED_LEAFMAP);
    const RepInfo& inf = a.info[r];
    // (independent loads first, see k_write_tree)
    const int status = inf.status, n_view = VIEW ? inf.n_p : inf.n_u, p_root = inf.p_root;
    const unsigned long long o0 = a.offsets[(size_t)st * a.stride + r], o1 = a.offsets[(size_t)st * a.stride + r + 1];
    const long long off = a.node_off[r];
    if (status != REP_OK) return;
    if (o1 > a.cap[st]) { if (lane == 0) *a.overflow = 1; return; }
    char* base = a.out[st] + o0;
    char* stage2 = s_stage_dyn + (size_t)(threadIdx.x >> 5) * 2 * a.stage_bytes;
    const RowRec* recs = a.rows + (size_t)VIEW * a.aux_stride + off;
    const int np = VIEW ? (p_root >= 0 ? n_view : 0) : n_view;
    unsigned trip = 0;
    if (KIND == 0) {
        int head_len = a.g2h_head_len;
        if (a.app == 2) {
            // DomainTreeGen prints the gene tree of the ROOT vertex as "Host tree" (getSigma :1279)
            const int rg = VIEW ? inf.root_gtree_p : inf.root_gtree_u;
            head_len = a.g2h_gene_off[2 * rg + 1];
            for (int i = lane; i < head_len; i += 32) base[i] = a.blob[a.g2h_gene_off[2 * rg] + i];
        } else {
            const int mis = (int)(reinterpret_cast<unsigned long long>(base) & 15ull);
            flush_stage(a.blob + a.g2h_head_off16[mis] - mis, mis, base, (unsigned)a.g2h_head_len);      // source congruent to destination mod 16
        }
        emit_records<BULK>(np, recs, base + head_len, stage2, (unsigned)a.stage_bytes, trip,
                           [&](auto& s, const RowRec& t, int) { put_g2h_row(s, a, t); }, [](const RowRec& t) { return (unsigned)t.len_g2h; });
    } else {
        emit_records<BULK>(np, recs, base, stage2, (unsigned)a.stage_bytes, trip,
                           [&](auto& s, const RowRec& t, int) { if (t.flags & RR_LEAFMAP) put_leafmap_row(s, a, t); }, [](const RowRec& t) { return (unsigned)t.len_lm; });
    }
    if (BULK && lane == 0) bulk_wait_read<0>();
}

}  // namespace sgp
