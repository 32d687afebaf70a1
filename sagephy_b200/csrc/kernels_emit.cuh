// K4: text emission. One warp per replicate; every output file ("stream") of a replicate is a concatenation of
// independent pieces whose byte offsets are an exclusive prefix sum of their lengths:
//   *.tree       one piece per vertex in post-order:  ['(' x run | ')'] name ':' length [tag] [','] [tree tag ';' '\n']
//   *.info       one piece per line
//   *.guest2host constant header + one row per vertex in breadth-first order
//   *.leafmap    one row per leaf in breadth-first order
// Size pass  (k_emit<false>): piece lengths through a CountSink; lengths (u16) and the shortest decimal digits of every
//            branch length (Schubfach result f x 10^e) are cached per piece so that nothing is formatted twice.
// Write pass (k_emit<true>):  32 pieces at a time: cached lengths -> warp exclusive scan -> each lane formats its piece
//            once into a shared-memory staging buffer laid out congruent (mod 16) to the destination -> the warp flushes
//            the buffer with aligned 16-byte coalesced stores.
//
// Formats follow NewickVertex.toString (io/NewickVertex.java:412-435), NewickTree.toString (io/NewickTree.java:457-465),
// GuestVertex.setMeta (apps/SaGePhy/GuestVertex.java:183-425), HostTreeGen.getInfo (apps/SaGePhy/HostTreeGen.java:99-110,124-174),
// GuestTreeInHostTreeCreator.getInfo/getLeafMap/getSigma (:431-559) and the file writers in GuestTreeGen.java:66-101.
#pragma once
#include "kernels_args.cuh"

namespace sgp {



template <class Sink> __device__ __forceinline__ void put_blob(Sink& s, const char* b, int off, int len) { for (int i = 0; i < len; ++i) s.put(b[off + i]); }
__device__ __forceinline__ void put_blob(CountSink& s, const char*, int, int len) { s.n += len; }

template <class Sink> __device__ void put_name(Sink& s, const EmitArgs& a, const Node& x) {
    for (int i = 0; i < a.prefix_len; ++i) s.put(a.prefix[i]);
    put_i64(s, x.number);
    if (a.append_sigma) { s.put('_'); put_i64(s, x.sigma); }
}

template <class Sink> __device__ void put_discpt(Sink& s, const EmitArgs& a, const Node& x) {
    const double* t = a.app == 2 ? a.gene_ep_times + 3 * (size_t)(a.gene_ep_base[x.gtree] + x.epoch) : a.ep_times + 3 * (size_t)x.epoch;
    int j = 0; while (j < 3) { if (t[j] >= x.abstime) break; ++j; }
    put_lit(s, "DISCPT=("); put_i64(s, x.epoch); s.put(','); put_i64(s, j); s.put(')');
}

template <class Sink> __device__ void put_meta(Sink& s, const EmitArgs& a, const Node& x) {
    put_lit(s, "[ID="); put_i64(s, x.number);
    if (!a.is_species) {
        put_lit(s, " HOST="); put_i64(s, x.sigma); put_lit(s, " GUEST=");
        put_blob(s, a.blob, a.gname_off[2 * x.gtree], a.gname_off[2 * x.gtree + 1]);
    }
    switch (x.event) {
        case EV_DUPLICATION: if (a.is_species) put_lit(s, " VERTEXTYPE=Speciation "); else put_lit(s, " VERTEXTYPE=Duplication "); put_discpt(s, a, x); break;
        case EV_LOSS: put_lit(s, " VERTEXTYPE=Loss"); break;
        case EV_REPLACING_LOSS: put_lit(s, " VERTEXTYPE=Replacing Loss"); break;
        case EV_SPECIATION: put_lit(s, " VERTEXTYPE=Speciation "); put_discpt(s, a, x); break;
        case EV_CODIVERGENCE: put_lit(s, " VERTEXTYPE=Co-divergence "); put_discpt(s, a, x); break;
        case EV_LEAF: put_lit(s, " VERTEXTYPE=Leaf"); break;
        case EV_UNSAMPLED_LEAF: put_lit(s, " VERTEXTYPE=UnsampledLeaf"); break;
        default: {
            const bool additive = x.event == EV_ADDITIVE_TRANSFER || (x.event >= EV_AT_GG_SS && x.event <= EV_AT_DG_DS);
            if (additive) put_lit(s, " VERTEXTYPE=Additive Transfer"); else put_lit(s, " VERTEXTYPE=Replacing Transfer");
            if (x.event >= EV_AT_GG_SS) {
                const int k = (x.event - EV_AT_GG_SS) & 3;
                if (k & 2) put_lit(s, " Intergene"); else put_lit(s, " Intragene"); if (k & 1) put_lit(s, " Interspecies"); else put_lit(s, " Intraspecies");
            }
            put_lit(s, " FROMTOLINEAGE=("); put_i64(s, x.from_arc); s.put(','); put_i64(s, x.to_arc);
            if (x.event >= EV_AT_GG_SS) {
                s.put(';');
                if (x.from_g >= 0) put_blob(s, a.blob, a.gname_off[2 * x.from_g], a.gname_off[2 * x.from_g + 1]); else put_lit(s, "null");
                s.put(',');
                if (x.to_g >= 0) put_blob(s, a.blob, a.gname_off[2 * x.to_g], a.gname_off[2 * x.to_g + 1]); else put_lit(s, "null");
            }
            put_lit(s, ") "); put_discpt(s, a, x);
        }
    }
    s.put(']');
}

// piece k (post-order rank) of a tree stream; (f, e) = cached shortest decimal of the branch length (e == kNoDigits: none)
constexpr int kNoDigits = 0x7fff;
template <bool PRUNED, class Sink> __device__ void put_tree_piece(Sink& s, const EmitArgs& a, const Node* N, const int32_t* ord, int k, int root,
                                                                  unsigned long long f, int e) {
    const int vi = ord[k];
    const Node& x = N[vi];
    const bool leaf = PRUNED ? x.p_left < 0 : x.left < 0;
    if (leaf) { const int c = PRUNED ? x.chain_p : x.chain_u; for (int i = 0; i < c; ++i) s.put('('); }
    else s.put(')');
    put_name(s, a, x);
    s.put(':');
    const double bl = PRUNED ? x.p_bl : x.bl;
    if (e != kNoDigits) put_double_digits(s, bl, f, e); else put_double(s, a.T, bl);
    if (!a.exclude_meta) put_meta(s, a, x);
    if (x.flags & (PRUNED ? NF_LEFT_P : NF_LEFT_U)) s.put(',');
    if (vi == root) {
        if (!a.exclude_meta) { if (PRUNED) put_lit(s, "[NAME=PrunedTree]"); else put_lit(s, "[NAME=UnprunedTree]"); }
        s.put(';'); s.put('\n');
    }
}

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

__device__ __forceinline__ const char* event_enum_name(int e) {
    switch (e) {
        case EV_SPECIATION: return "SPECIATION"; case EV_CODIVERGENCE: return "CODIVERGENCE"; case EV_LEAF: return "LEAF";
        case EV_UNSAMPLED_LEAF: return "UNSAMPLED_LEAF"; case EV_DUPLICATION: return "DUPLICATION"; case EV_LOSS: return "LOSS";
        case EV_REPLACING_LOSS: return "REPLACING_LOSS"; case EV_ADDITIVE_TRANSFER: return "ADDITIVE_TRANSFER";
        case EV_REPLACING_TRANSFER: return "REPLACING_TRANSFER";
        case EV_AT_GG_SS: return "ADDITIVE_TRANSFER_INTRAGENE_INTRASPECIES"; case EV_AT_GG_DS: return "ADDITIVE_TRANSFER_INTRAGENE_INTERSPECIES";
        case EV_AT_DG_SS: return "ADDITIVE_TRANSFER_INTERGENE_INTRASPECIES"; case EV_AT_DG_DS: return "ADDITIVE_TRANSFER_INTERGENE_INTERSPECIES";
        case EV_RT_GG_SS: return "REPLACING_TRANSFER_INTRAGENE_INTRASPECIES"; case EV_RT_GG_DS: return "REPLACING_TRANSFER_INTRAGENE_INTERSPECIES";
        case EV_RT_DG_SS: return "REPLACING_TRANSFER_INTERGENE_INTRASPECIES"; default: return "REPLACING_TRANSFER_INTERGENE_INTERSPECIES";
    }
}

// guest2host row k (BFS rank)
template <class Sink> __device__ void put_g2h_row(Sink& s, const EmitArgs& a, const Node* N, const int32_t* bfs, int k) {
    const Node& x = N[bfs[k]];
    put_name(s, a, x); s.put('\t'); put_i64(s, x.number); s.put('\t'); put_str(s, event_enum_name(x.event)); s.put('\t');
    put_double(s, a.T, x.abstime); s.put('\t'); put_i64(s, x.sigma); s.put('\t'); put_i64(s, x.epoch); s.put('\n');
}
// leafmap row k (BFS rank); empty for interior vertices
template <bool PRUNED, class Sink> __device__ void put_leafmap_row(Sink& s, const EmitArgs& a, const Node* N, const int32_t* bfs, int k) {
    const Node& x = N[bfs[k]];
    const bool leaf = PRUNED ? x.p_left < 0 : x.left < 0;
    if (!leaf || !(x.event == EV_LEAF || x.event == EV_UNSAMPLED_LEAF)) return;
    put_name(s, a, x); s.put('\t');
    const int li = (a.app == 2 ? a.leaf_base[x.gtree] : 0) + x.sigma;
    put_blob(s, a.blob, a.leafname_off[2 * li], a.leafname_off[2 * li + 1]);
    if (a.app == 2) { s.put('\t'); put_blob(s, a.blob, a.gname_off[2 * x.gtree], a.gname_off[2 * x.gtree + 1]); }    // DomainTreeInHostTreeCreator.getLeafMap :1266
    s.put('\n');
}

__device__ __forceinline__ unsigned warp_excl_scan_u32(unsigned v, unsigned* total) {
    unsigned x = v; const int lane = threadIdx.x & 31;
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

constexpr int kEmitWarps = 4;               // warps per block
constexpr int kStageBytes = 5120;           // shared staging per warp (32 pieces; longer groups fall back to direct stores)

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

// Pieces 0..n_pieces-1 of one stream of one replicate.
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

template <bool WRITE, int STREAM>
__global__ void __launch_bounds__(kEmitWarps * 32, WRITE ? 8 : 10) k_emit(EmitArgs a) {
    __shared__ __align__(16) char s_stage[kEmitWarps][kStageBytes];
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= a.n) return;
    char* stage = s_stage[threadIdx.x >> 5];
    const RepInfo& inf = a.info[r];
    const long long gid = a.rep_base + r;
    const bool ok = inf.status == REP_OK;
    const long long off = ok ? a.node_off[r] : 0;
    const Node* N = a.nodes + off;
    const int32_t* ordU = a.aux + off; const int32_t* ordP = a.aux + a.aux_stride + off;
    const int32_t* bfsU = a.aux + 2 * a.aux_stride + off; const int32_t* bfsP = a.aux + 3 * a.aux_stride + off;
    auto no_load = [](int) { return 0u; };
    auto no_store = [](int, unsigned) {};
    {
        constexpr int st = STREAM;
        if (WRITE && a.offsets[(size_t)st * a.stride + r + 1] > a.cap[st]) { if (lane == 0) *a.overflow = 1; return; }
        char* base = WRITE ? a.out[st] + a.offsets[(size_t)st * a.stride + r] : nullptr;
        uint16_t* plen = a.plen + (size_t)st * a.aux_stride + off;       // cached piece lengths of this stream
        auto ld = [&](int i) { return (unsigned)plen[i]; };
        auto sd = [&](int i, unsigned v) { plen[i] = (uint16_t)(v > 0xffffu ? 0xffffu : v); };
        unsigned long long bytes = 0;
        if (!ok) {
            // failed replicate: the reference writes only the failure note (to <prefix>.info / <prefix>.pruned.info)
            if (st == ST_PRUNED_INFO && inf.status == REP_MAX_ATTEMPTS) {
                auto pc = [&](auto& s, int) { put_lit(s, "Failed to produce valid pruned tree within max allowed attempts.\n"); };
                if (WRITE) bytes = emit_pieces<true>(1, base, stage, pc, [](int) { return 65u; }, no_store);
                else bytes = emit_pieces<false>(1, base, stage, pc, no_load, no_store);
            }
        } else {
            if constexpr (st == ST_UNPRUNED_TREE || st == ST_PRUNED_TREE) {
                constexpr bool pr = st == ST_PRUNED_TREE;
                if (pr && inf.p_root < 0) {
                    auto pc = [&](auto& s, int) { if (a.exclude_meta) put_lit(s, ";\n"); else put_str(s, a.app == 0 ? "[NAME=PrunedTree];" : "[NAME=PrunedTree];\n"); };
                    const unsigned l = a.exclude_meta ? 2u : (a.app == 0 ? 18u : 19u);
                    if (WRITE) bytes = emit_pieces<true>(1, base, stage, pc, [l](int) { return l; }, no_store);
                    else bytes = emit_pieces<false>(1, base, stage, pc, no_load, no_store);
                } else {
                unsigned long long* df = a.dig_f + (size_t)(pr ? 1 : 0) * a.aux_stride + off;
                int16_t* de = a.dig_e + (size_t)(pr ? 1 : 0) * a.aux_stride + off;
                const int32_t* ord = pr ? ordP : ordU; const int np = pr ? inf.n_p : inf.n_u; const int root = pr ? inf.p_root : inf.root;
                if (WRITE) {
                    auto pc = [&](auto& s, int k) { if (pr) put_tree_piece<true>(s, a, N, ord, k, root, df[k], (int)de[k]); else put_tree_piece<false>(s, a, N, ord, k, root, df[k], (int)de[k]); };
                    bytes = emit_pieces<true>(np, base, stage, pc, ld, no_store);
                } else {
                    auto pc = [&](auto& s, int k) {
                        const Node& x = N[ord[k]]; const double bl = pr ? x.p_bl : x.bl;
                        uint64_t f = 0; int e = kNoDigits;
                        if (bl > 0.0 && bl < dbl_inf()) d2s_decimal(a.T, bl, &f, &e);
                        df[k] = (unsigned long long)f; de[k] = (int16_t)e;
                        if (pr) put_tree_piece<true>(s, a, N, ord, k, root, f, e); else put_tree_piece<false>(s, a, N, ord, k, root, f, e);
                    };
                    bytes = emit_pieces<false>(np, base, stage, pc, no_load, sd);
                }
                }
            }
            if constexpr (st == ST_UNPRUNED_INFO || st == ST_PRUNED_INFO) {
                constexpr bool pr = st == ST_PRUNED_INFO;
                // lines 0-1 ("# APP" and "Arguments: [...]", which may embed a whole Newick host tree): warp-cooperative copy
                const char* h0 = a.app == 0 ? "# HOSTTREEGEN\nArguments:\t" : a.app == 1 ? "# GUESTTREEGEN\nArguments:\t" : "# DOMAINTREEGEN\nArguments:\t";
                const int h0len = a.app == 0 ? 25 : a.app == 1 ? 26 : 27;
                const int seedlen = a.seed_mode ? seed_dec_len(a.sseed, a.blob, (unsigned long long)gid) : 0;
                unsigned long long o = 0;
                if (WRITE) {
                    for (int i = lane; i < h0len; i += 32) base[i] = h0[i];
                    if (a.args_pre_off16[0] >= 0) {
                        // long argument text (a host tree given as a Newick string): aligned 16-byte copies from the copy of the
                        // text whose blob offset is congruent to the destination mod 16
                        char* dst = base + h0len;
                        const int mis = (int)(reinterpret_cast<unsigned long long>(dst) & 15ull);
                        flush_stage(a.blob + a.args_pre_off16[mis] - mis, mis, dst, (unsigned)a.args_pre_len);
                    } else for (int i = lane; i < a.args_pre_len; i += 32) base[h0len + i] = a.blob[a.args_pre_off + i];
                }
                o = (unsigned long long)h0len + a.args_pre_len;
                if (a.seed_mode) {
                    if (WRITE) {
                        if (lane == 0) { MemSink m{base + o}; put_seed(m, a.sseed, a.blob, (unsigned long long)gid); }
                        for (int i = lane; i < a.args_post_len; i += 32) base[o + seedlen + i] = a.blob[a.args_post_off + i];
                    }
                    o += (unsigned long long)seedlen + a.args_post_len;
                }
                if (WRITE && lane == 0) base[o] = '\n';
                o += 1;
                auto pc = [&](auto& s, int ln) { if (pr) put_info_line<true>(s, a, inf, gid, ln + 2); else put_info_line<false>(s, a, inf, gid, ln + 2); };
                // the remaining lines are cheap: their lengths are recomputed in the write pass instead of being cached
                auto ld_info = [&](int ln) { CountSink c; if (pr) put_info_line<true>(c, a, inf, gid, ln + 2); else put_info_line<false>(c, a, inf, gid, ln + 2); return (unsigned)c.n; };
                if (WRITE) bytes = o + emit_pieces<true>(info_lines(a, pr) - 2, base + o, stage, pc, ld_info, no_store);
                else bytes = o + emit_pieces<false>(info_lines(a, pr) - 2, base, stage, pc, no_load, no_store);
            }
            if constexpr (st == ST_UNPRUNED_G2H || st == ST_PRUNED_G2H) {
                constexpr bool pr = st == ST_PRUNED_G2H;
                int head_len = a.g2h_head_len;
                if (a.app == 2) {
                    // DomainTreeGen prints the gene tree of the ROOT vertex as "Host tree" (getSigma :1279)
                    const int rg = N[pr ? (inf.p_root >= 0 ? inf.p_root : inf.root) : inf.root].gtree;
                    head_len = a.g2h_gene_off[2 * rg + 1];
                    if (WRITE) for (int i = lane; i < head_len; i += 32) base[i] = a.blob[a.g2h_gene_off[2 * rg] + i];
                } else if (WRITE) {
                    const int mis = (int)(reinterpret_cast<unsigned long long>(base) & 15ull);
                    flush_stage(a.blob + a.g2h_head_off16[mis] - mis, mis, base, (unsigned)a.g2h_head_len);      // source congruent to destination mod 16
                }
                bytes = (unsigned long long)head_len;
                const int np = pr ? (inf.p_root >= 0 ? inf.n_p : 0) : inf.n_u;
                const int32_t* bfs = pr ? bfsP : bfsU;
                auto pc = [&](auto& s, int k) { put_g2h_row(s, a, N, bfs, k); };
                if (WRITE) bytes += emit_pieces<true>(np, base + head_len, stage, pc, ld, no_store);
                else bytes += emit_pieces<false>(np, nullptr, stage, pc, no_load, sd);
            }
            if constexpr (st == ST_UNPRUNED_LEAFMAP || st == ST_PRUNED_LEAFMAP) {
                constexpr bool pr = st == ST_PRUNED_LEAFMAP;
                const int np = pr ? (inf.p_root >= 0 ? inf.n_p : 0) : inf.n_u;
                const int32_t* bfs = pr ? bfsP : bfsU;
                auto pc = [&](auto& s, int k) { if (pr) put_leafmap_row<true>(s, a, N, bfs, k); else put_leafmap_row<false>(s, a, N, bfs, k); };
                if (WRITE) bytes = emit_pieces<true>(np, base, stage, pc, ld, no_store);
                else bytes = emit_pieces<false>(np, nullptr, stage, pc, no_load, sd);
            }
        }
        if (!WRITE && lane == 0) a.sizes[(size_t)st * a.stride + r] = bytes;
    }
}

}  // namespace sgp
