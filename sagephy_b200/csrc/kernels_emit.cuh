// K4: text emission. One warp per replicate; every output file ("stream") of a replicate is a concatenation of
// independent pieces whose byte offsets are an exclusive prefix sum of their lengths:
//   *.tree       one piece per vertex in post-order:  ['(' x run | ')'] name ':' length [tag] [','] [tree tag ';' '\n']
//   *.info       a single piece
//   *.guest2host constant header + one row per vertex in breadth-first order
//   *.leafmap    one row per leaf in breadth-first order
// The same formatter runs with a CountSink (size pass) and a MemSink (write pass).
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
    const double* t = a.ep_times + 3 * (size_t)x.epoch;
    int j = 0; while (j < 3) { if (t[j] >= x.abstime) break; ++j; }
    put_str(s, "DISCPT=("); put_i64(s, x.epoch); s.put(','); put_i64(s, j); s.put(')');
}

template <class Sink> __device__ void put_meta(Sink& s, const EmitArgs& a, const Node& x) {
    put_str(s, "[ID="); put_i64(s, x.number);
    if (!a.is_species) {
        put_str(s, " HOST="); put_i64(s, x.sigma); put_str(s, " GUEST=");
        put_blob(s, a.blob, a.gname_off[2 * x.gtree], a.gname_off[2 * x.gtree + 1]);
    }
    switch (x.event) {
        case EV_DUPLICATION: put_str(s, a.is_species ? " VERTEXTYPE=Speciation " : " VERTEXTYPE=Duplication "); put_discpt(s, a, x); break;
        case EV_LOSS: put_str(s, " VERTEXTYPE=Loss"); break;
        case EV_REPLACING_LOSS: put_str(s, " VERTEXTYPE=Replacing Loss"); break;
        case EV_SPECIATION: put_str(s, " VERTEXTYPE=Speciation "); put_discpt(s, a, x); break;
        case EV_CODIVERGENCE: put_str(s, " VERTEXTYPE=Co-divergence "); put_discpt(s, a, x); break;
        case EV_LEAF: put_str(s, " VERTEXTYPE=Leaf"); break;
        case EV_UNSAMPLED_LEAF: put_str(s, " VERTEXTYPE=UnsampledLeaf"); break;
        default: {
            const bool additive = x.event == EV_ADDITIVE_TRANSFER || (x.event >= EV_AT_GG_SS && x.event <= EV_AT_DG_DS);
            put_str(s, additive ? " VERTEXTYPE=Additive Transfer" : " VERTEXTYPE=Replacing Transfer");
            if (x.event >= EV_AT_GG_SS) {
                const int k = (x.event - EV_AT_GG_SS) & 3;
                put_str(s, (k & 2) ? " Intergene" : " Intragene"); put_str(s, (k & 1) ? " Interspecies" : " Intraspecies");
            }
            put_str(s, " FROMTOLINEAGE=("); put_i64(s, x.from_arc); s.put(','); put_i64(s, x.to_arc);
            if (x.event >= EV_AT_GG_SS) {
                s.put(';');
                if (x.from_g >= 0) put_blob(s, a.blob, a.gname_off[2 * x.from_g], a.gname_off[2 * x.from_g + 1]); else put_str(s, "null");
                s.put(',');
                if (x.to_g >= 0) put_blob(s, a.blob, a.gname_off[2 * x.to_g], a.gname_off[2 * x.to_g + 1]); else put_str(s, "null");
            }
            put_str(s, ") "); put_discpt(s, a, x);
        }
    }
    s.put(']');
}

// piece k (post-order rank) of a tree stream
template <bool PRUNED, class Sink> __device__ void put_tree_piece(Sink& s, const EmitArgs& a, const Node* N, const int32_t* ord, int k, int root) {
    const int vi = ord[k];
    const Node& x = N[vi];
    const bool leaf = PRUNED ? x.p_left < 0 : x.left < 0;
    if (leaf) { const int c = PRUNED ? x.chain_p : x.chain_u; for (int i = 0; i < c; ++i) s.put('('); }
    else s.put(')');
    put_name(s, a, x);
    s.put(':'); put_double(s, a.T, PRUNED ? x.p_bl : x.bl);
    if (!a.exclude_meta) put_meta(s, a, x);
    if (x.flags & (PRUNED ? NF_LEFT_P : NF_LEFT_U)) s.put(',');
    if (vi == root) {
        if (!a.exclude_meta) put_str(s, PRUNED ? "[NAME=PrunedTree]" : "[NAME=UnprunedTree]");
        s.put(';'); s.put('\n');
    }
}

template <class Sink> __device__ void put_kv_int(Sink& s, const char* k, long long v) { put_str(s, k); put_i64(s, v); s.put('\n'); }
template <class Sink> __device__ void put_kv_dbl(Sink& s, const EmitArgs& a, const char* k, double v) { put_str(s, k); put_double(s, a.T, v); s.put('\n'); }

template <bool PRUNED, class Sink> __device__ void put_info(Sink& s, const EmitArgs& a, const RepInfo& inf, long long gid) {
    put_str(s, a.app == 0 ? "# HOSTTREEGEN\n" : a.app == 1 ? "# GUESTTREEGEN\n" : "# DOMAINTREEGEN\n");
    put_str(s, "Arguments:\t");
    put_blob(s, a.blob, a.args_pre_off, a.args_pre_len);
    if (a.seed_mode) { put_i64(s, a.seed_base + gid); put_blob(s, a.blob, a.args_post_off, a.args_post_len); }
    s.put('\n');
    put_kv_int(s, "Attempts:\t", inf.attempts);
    const int32_t* c = PRUNED ? inf.cnt_p : inf.cnt_u;
    const int nv = PRUNED ? inf.n_p : inf.n_u;
    const double total = round_sig(a.T, PRUNED ? inf.total_p : inf.total_u, 8);
    if (a.app == 0) {
        const int births = c[EV_DUPLICATION] + c[EV_SPECIATION], deaths = c[EV_LOSS];
        put_kv_int(s, "No. of vertices:\t", nv);
        put_kv_int(s, "No. of extant leaves:\t", c[EV_LEAF] + c[EV_UNSAMPLED_LEAF]);
        put_kv_int(s, "No. of births:\t", births);
        put_kv_int(s, "No. of deaths:\t", deaths);
        put_kv_dbl(s, a, "Total branch time:\t", total);
        if (!PRUNED) {
            put_kv_dbl(s, a, "Birth ML estimate:\t", round_sig(a.T, XDIV((double)births, total), 8));
            put_kv_dbl(s, a, "Death ML estimate:\t", round_sig(a.T, XDIV((double)deaths, total), 8));
        }
    } else {
        const double ben = round_sig(a.T, PRUNED ? inf.beneath_p : inf.beneath_u, 8);
        put_kv_int(s, "No. of vertices:\t", nv);
        put_kv_int(s, "No. of extant leaves:\t", c[EV_LEAF] + c[EV_UNSAMPLED_LEAF]);
        put_kv_int(s, "No. of speciations:\t", c[EV_SPECIATION]);
        put_kv_int(s, "No. of duplications:\t", c[EV_DUPLICATION]);
        put_kv_int(s, "No. of losses:\t", c[EV_LOSS]);
        put_kv_int(s, "No. of replacing losses:\t", c[EV_REPLACING_LOSS]);
        put_kv_int(s, "No. of additive transfers:\t", c[EV_ADDITIVE_TRANSFER]);
        put_kv_int(s, "No. of replacing transfers:\t", c[EV_REPLACING_TRANSFER]);
        put_kv_dbl(s, a, "Total branch time:\t", total);
        put_kv_dbl(s, a, "Total branch time beneath host stem:\t", ben);
        put_kv_dbl(s, a, "Duplication ML estimate:\t", round_sig(a.T, XDIV((double)c[EV_DUPLICATION], total), 8));
        put_kv_dbl(s, a, "Loss ML estimate:\t", round_sig(a.T, XDIV((double)c[EV_LOSS], total), 8));
        put_kv_dbl(s, a, "Replacing Loss ML estimate:\t", round_sig(a.T, XDIV((double)c[EV_REPLACING_LOSS], total), 8));
        put_kv_dbl(s, a, "Additive Transfer ML estimate:\t", round_sig(a.T, XDIV((double)c[EV_ADDITIVE_TRANSFER], total), 8));
        put_kv_dbl(s, a, "Replacing Transfer ML estimate:\t", round_sig(a.T, XDIV((double)c[EV_REPLACING_TRANSFER], total), 8));
    }
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
    put_blob(s, a.blob, a.leafname_off[2 * x.sigma], a.leafname_off[2 * x.sigma + 1]);
    s.put('\n');
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned warp_excl_scan_u32(unsigned v, unsigned* total) {
    unsigned x = v; const int lane = threadIdx.x & 31;
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

// Generic driver: pieces 0..n_pieces-1, piece lengths via CountSink, bytes via MemSink.
template <bool WRITE, class F> __device__ unsigned long long emit_pieces(int n_pieces, char* base, F&& piece) {
    const int lane = threadIdx.x & 31;
    unsigned long long running = 0;
    for (int i0 = 0; i0 < n_pieces; i0 += 32) {
        const int i = i0 + lane;
        unsigned len = 0;
        if (i < n_pieces) { CountSink c; piece(c, i); len = (unsigned)c.n; }
        unsigned total; unsigned off = warp_excl_scan_u32(len, &total);
        if (WRITE && i < n_pieces && len) { MemSink m{base + running + off}; piece(m, i); }
        running += total;
    }
    return running;
}

template <bool WRITE>
__global__ void __launch_bounds__(128) k_emit(EmitArgs a) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= a.n) return;
    const RepInfo& inf = a.info[r];
    const long long gid = a.rep_base + r;
    const bool ok = inf.status == REP_OK;
    const long long off = ok ? a.node_off[r] : 0;
    const Node* N = a.nodes + off;
    const int32_t* ordU = a.aux + off; const int32_t* ordP = a.aux + a.aux_stride + off;
    const int32_t* bfsU = a.aux + 2 * a.aux_stride + off; const int32_t* bfsP = a.aux + 3 * a.aux_stride + off;
    for (int st = 0; st < ST_COUNT; ++st) {
        if (!(a.stream_mask & (1u << st))) continue;
        char* base = WRITE ? a.out[st] + a.offsets[(size_t)st * (a.n + 1) + r] : nullptr;
        unsigned long long bytes = 0;
        if (!ok) {
            // failed replicate: the reference writes only the failure note (to <prefix>.info / <prefix>.pruned.info)
            if (st == ST_PRUNED_INFO && inf.status == REP_MAX_ATTEMPTS) {
                bytes = emit_pieces<WRITE>(1, base, [&](auto& s, int) { put_str(s, "Failed to produce valid pruned tree within max allowed attempts.\n"); });
            }
        } else switch (st) {
            case ST_UNPRUNED_TREE:
                bytes = emit_pieces<WRITE>(inf.n_u, base, [&](auto& s, int k) { put_tree_piece<false>(s, a, N, ordU, k, inf.root); });
                break;
            case ST_PRUNED_TREE:
                if (inf.p_root >= 0) bytes = emit_pieces<WRITE>(inf.n_p, base, [&](auto& s, int k) { put_tree_piece<true>(s, a, N, ordP, k, inf.p_root); });
                else bytes = emit_pieces<WRITE>(1, base, [&](auto& s, int) {
                    if (a.exclude_meta) put_str(s, ";\n"); else put_str(s, a.app == 0 ? "[NAME=PrunedTree];" : "[NAME=PrunedTree];\n"); });
                break;
            case ST_UNPRUNED_INFO:
                bytes = emit_pieces<WRITE>(1, base, [&](auto& s, int) { put_info<false>(s, a, inf, gid); });
                break;
            case ST_PRUNED_INFO:
                bytes = emit_pieces<WRITE>(1, base, [&](auto& s, int) { put_info<true>(s, a, inf, gid); });
                break;
            case ST_UNPRUNED_G2H: case ST_PRUNED_G2H: {
                const bool pr = st == ST_PRUNED_G2H;
                if (WRITE) for (int i = lane; i < a.g2h_head_len; i += 32) base[i] = a.blob[a.g2h_head_off + i];
                bytes = (unsigned long long)a.g2h_head_len;
                const int np = pr ? (inf.p_root >= 0 ? inf.n_p : 0) : inf.n_u;
                bytes += emit_pieces<WRITE>(np, WRITE ? base + a.g2h_head_len : nullptr, [&](auto& s, int k) { put_g2h_row(s, a, N, pr ? bfsP : bfsU, k); });
                break;
            }
            case ST_UNPRUNED_LEAFMAP:
                bytes = emit_pieces<WRITE>(inf.n_u, base, [&](auto& s, int k) { put_leafmap_row<false>(s, a, N, bfsU, k); });
                break;
            case ST_PRUNED_LEAFMAP:
                bytes = emit_pieces<WRITE>(inf.p_root >= 0 ? inf.n_p : 0, base, [&](auto& s, int k) { put_leafmap_row<true>(s, a, N, bfsP, k); });
                break;
        }
        if (!WRITE && lane == 0) a.sizes[(size_t)st * a.n + r] = bytes;
    }
}

}  // namespace sgp
