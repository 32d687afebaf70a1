// Text of one output piece, from its emission record (device_types.cuh: TreeRec / TreeRecX / RowRec), and the construction of
// those records from a vertex. The same templates run with a CountSink (length of the piece, stored in the record by the label
// kernel) and with a byte sink (the write kernels), so the two can never disagree.
//
// Formats follow NewickVertex.toString (io/NewickVertex.java:412-435), NewickTree.toString (io/NewickTree.java:457-465),
// GuestVertex.setMeta (apps/SaGePhy/GuestVertex.java:183-425) and GuestTreeInHostTreeCreator.getLeafMap/getSigma (:514-559),
// DomainTreeInHostTreeCreator.getLeafMap/getSigma (:1255-1300).
#pragma once
#include "kernels_args.cuh"

namespace sgp {

template <class Sink> __device__ __forceinline__ void put_blob(Sink& s, const char* b, int off, int len) { for (int i = 0; i < len; ++i) s.put(b[off + i]); }
__device__ __forceinline__ void put_blob(CountSink& s, const char*, int, int len) { s.n += len; }

// Decimal lengths kept in a finished record. A piece prints its vertex number twice (name, ID=), its host vertex twice (name,
// HOST= or the row's column) and its epoch once; measuring each number at every use was 15 % of the instructions of the write
// kernels and 24 % of the finish kernel. The finish kernel now measures each once and keeps the result in the record's `pad`
// word: bits 0-4 digits of f, then 3 bits each for number, sigma and epoch (0 = not recorded: negative, or more than 7 digits).
__device__ __forceinline__ int len_hint(int v) { return (v >= 0 && v < 10000000) ? dec_len_u32_compact((uint32_t)v) : 0; }
__device__ __forceinline__ int pack_pad(int nd, int number, int sigma, int epoch) { return nd | (len_hint(number) << 5) | (len_hint(sigma) << 8) | (len_hint(epoch) << 11); }
__device__ __forceinline__ int pad_nd(int pad) { return pad & 31; }
__device__ __forceinline__ int pad_num(int pad) { return (pad >> 5) & 7; }
__device__ __forceinline__ int pad_sig(int pad) { return (pad >> 8) & 7; }
__device__ __forceinline__ int pad_ep(int pad) { return (pad >> 11) & 7; }
template <class Sink> __device__ __forceinline__ void put_i32_len(Sink& s, int v, int len) {
    if (len == 0) { put_i32(s, v); return; }
    if (Sink::kCounts) s.skip(len); else put_u32_digits(s, (uint32_t)v, len);
}

template <class Sink> __device__ __forceinline__ void put_name(Sink& s, const EmitArgs& a, int number, int sigma, int pad) {
    for (int i = 0; i < a.prefix_len; ++i) s.put(a.prefix[i]);
    put_i32_len(s, number, pad_num(pad));
    if (a.append_sigma) { s.put('_'); put_i32_len(s, sigma, pad_sig(pad)); }
}

// f x 10^e as java.lang.Double.toString lays it out; e == kNoDigits: f holds the bits of a value without digits (0.0, NaN, ...)
// nd: number of digits of f (kept in the record: the finish kernel counts them once)
// values without digits (0.0 at the root, NaN, infinities, negatives): the general java.lang.Double.toString, out of line - it is
// rare, and inlined it put a second copy of the whole shortest-decimal conversion into every instantiation of every piece
template <class Sink> static __device__ __noinline__ void put_double_out_of_line(Sink& s, const unsigned long long* g, const double* p10, double v) {
    MathTables T; T.g = g; T.p10 = p10;
    put_double(s, T, v);
}
template <class Sink> __device__ __forceinline__ void put_decimal(Sink& s, const MathTables& T, uint64_t f, int e, int nd) {
    if (e == kNoDigits) { put_double_out_of_line(s, T.g, T.p10, __longlong_as_double((long long)f)); return; }
    // the layout depends on 1e-3 <= v < 1e7 only; v = 0.DIGITS x 10^pe, so that is -2 <= pe <= 7
    const int pe = nd + e;
    put_double_digits(s, (pe >= -2 && pe <= 7) ? 1.0 : 1e-9, f, e, nd);
}
__device__ __forceinline__ void decimal_of(const MathTables& T, double v, uint64_t* f, int* e) {
    if (v > 0.0 && v < dbl_inf()) d2s_decimal(T, v, f, e);
    else { *f = (uint64_t)__double_as_longlong(v); *e = kNoDigits; }
}
// The same as ONE out-of-line copy (the finish kernel converts at four places; inlined four times the conversion alone was
// 100 KB of code against an instruction cache of 32 KB, and `no_instruction` was that kernel's largest stall). Returns
// f; e and the digit count of f come back packed as (e & 0xffff) | nd << 16.
static __device__ __noinline__ uint64_t decimal_of_packed(const unsigned long long* g, const double* p10, double v, int* e_nd) {
    MathTables T; T.g = g; T.p10 = p10;
    uint64_t f; int e; decimal_of(T, v, &f, &e);
    *e_nd = (e & 0xffff) | ((e == kNoDigits ? 0 : dec_len_u64_compact(f)) << 16);
    return f;
}

template <class Sink> __device__ __forceinline__ void put_discpt(Sink& s, int epoch, int j) {
    put_lit(s, "DISCPT=("); put_i32(s, epoch); s.put(','); put_i32(s, j); s.put(')');
}
// second DISCPT index: first j with times[j] >= abstime among [lo, mid, up] of the vertex's epoch (GuestVertex.java:196-204)
__device__ __forceinline__ int discpt_index(const EmitArgs& a, int gtree, int epoch, double abstime) {
    const double* t = a.app == 2 ? a.gene_ep_times + 3 * (size_t)(a.gene_ep_base[gtree] + epoch) : a.ep_times + 3 * (size_t)epoch;
    int j = 0; while (j < 3) { if (t[j] >= abstime) break; ++j; }
    return j;
}

__device__ __forceinline__ bool is_transfer_event(int ev) { return ev >= EV_ADDITIVE_TRANSFER; }

// Lanes print different vertex types in the same trip, so everything that is common to several types (the key, the DISCPT tag)
// is outside the switch: a branch of a switch runs once per type present in the warp.
template <class Sink> __device__ void put_meta(Sink& s, const EmitArgs& a, const TreeRec& t, const TreeRecX* xp) {
    put_lit(s, "[ID="); put_i32_len(s, t.number, pad_num(t.pad));
    if (!a.is_species) {
        put_lit(s, " HOST="); put_i32_len(s, t.sigma, pad_sig(t.pad)); put_lit(s, " GUEST=");
        put_blob(s, a.blob, a.gname_off[2 * t.gtree], a.gname_off[2 * t.gtree + 1]);
    }
    put_lit(s, " VERTEXTYPE=");
    bool discpt = false;
    switch (t.event) {
        case EV_DUPLICATION: if (a.is_species) put_lit(s, "Speciation"); else put_lit(s, "Duplication"); discpt = true; break;
        case EV_LOSS: put_lit(s, "Loss"); break;
        case EV_REPLACING_LOSS: put_lit(s, "Replacing Loss"); break;
        case EV_SPECIATION: put_lit(s, "Speciation"); discpt = true; break;
        case EV_CODIVERGENCE: put_lit(s, "Co-divergence"); discpt = true; break;
        case EV_LEAF: put_lit(s, "Leaf"); break;
        case EV_UNSAMPLED_LEAF: put_lit(s, "UnsampledLeaf"); break;
        default: {
            const TreeRecX x = *xp;
            const bool additive = t.event == EV_ADDITIVE_TRANSFER || (t.event >= EV_AT_GG_SS && t.event <= EV_AT_DG_DS);
            if (additive) put_lit(s, "Additive Transfer"); else put_lit(s, "Replacing Transfer");
            if (t.event >= EV_AT_GG_SS) {
                const int k = (t.event - EV_AT_GG_SS) & 3;
                if (k & 2) put_lit(s, " Intergene"); else put_lit(s, " Intragene"); if (k & 1) put_lit(s, " Interspecies"); else put_lit(s, " Intraspecies");
            }
            put_lit(s, " FROMTOLINEAGE=("); put_i32(s, x.from_arc); s.put(','); put_i32(s, x.to_arc);
            if (t.event >= EV_AT_GG_SS) {
                s.put(';');
                if (x.from_g >= 0) put_blob(s, a.blob, a.gname_off[2 * x.from_g], a.gname_off[2 * x.from_g + 1]); else put_lit(s, "null");
                s.put(',');
                if (x.to_g >= 0) put_blob(s, a.blob, a.gname_off[2 * x.to_g], a.gname_off[2 * x.to_g + 1]); else put_lit(s, "null");
            }
            s.put(')'); discpt = true;
        }
    }
    if (discpt) { put_lit(s, " DISCPT=("); put_i32_len(s, t.epoch, pad_ep(t.pad)); s.put(','); s.put((char)('0' + ((t.flags >> TR_J_SHIFT) & 3))); s.put(')'); }
    s.put(']');
}


// one Newick piece
template <bool PRUNED, class Sink> __device__ void put_tree_piece(Sink& s, const EmitArgs& a, const TreeRec& t, const TreeRecX* xp) {
    {   // '(' x chain before a leaf, ')' before an interior vertex: one loop for both (no branch for the compiler to duplicate the rest into)
        const bool leaf = (t.flags & TR_LEAF) != 0;
        const int c = leaf ? t.chain : 1; const char ch = leaf ? '(' : ')';
        for (int i = 0; i < c; ++i) s.put(ch);
    }
    put_name(s, a, t.number, t.sigma, t.pad);
    s.put(':');
    put_decimal(s, a.T, t.f, t.e, pad_nd(t.pad));
    if (!a.exclude_meta) put_meta(s, a, t, xp);
    if (t.flags & TR_COMMA) s.put(',');
    if (t.flags & TR_ROOT) {
        if (!a.exclude_meta) { if (PRUNED) put_lit(s, "[NAME=PrunedTree]"); else put_lit(s, "[NAME=UnprunedTree]"); }
        s.put(';'); s.put('\n');
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
__device__ __forceinline__ int event_enum_len(int e) {
    switch (e) {
        case EV_SPECIATION: return 10; case EV_CODIVERGENCE: return 12; case EV_LEAF: return 4; case EV_UNSAMPLED_LEAF: return 14;
        case EV_DUPLICATION: return 11; case EV_LOSS: return 4; case EV_REPLACING_LOSS: return 14; case EV_ADDITIVE_TRANSFER: return 17;
        case EV_REPLACING_TRANSFER: return 18;
        default: return e <= EV_AT_DG_DS ? 40 : 41;
    }
}
template <class Sink> __device__ __forceinline__ void put_event_enum(Sink& s, int e) { put_str(s, event_enum_name(e)); }
__device__ __forceinline__ void put_event_enum(CountSink& s, int e) { s.n += event_enum_len(e); }

// .guest2host row: name, id, type, time (unrounded), host vertex / arc, host epoch
template <class Sink> __device__ void put_g2h_row(Sink& s, const EmitArgs& a, const RowRec& r) {
    put_name(s, a, r.number, r.sigma, r.pad); s.put('\t'); put_i32_len(s, r.number, pad_num(r.pad)); s.put('\t'); put_event_enum(s, r.event); s.put('\t');
    put_decimal(s, a.T, r.f, r.e, pad_nd(r.pad)); s.put('\t'); put_i32_len(s, r.sigma, pad_sig(r.pad)); s.put('\t'); put_i32_len(s, r.epoch, pad_ep(r.pad)); s.put('\n');
}
// .leafmap row of a leaf of the view (RR_LEAFMAP): name, host leaf name [, gene tree name]
template <class Sink> __device__ void put_leafmap_row(Sink& s, const EmitArgs& a, const RowRec& r) {
    put_name(s, a, r.number, r.sigma, r.pad); s.put('\t');
    const int li = (a.app == 2 ? a.leaf_base[r.gtree] : 0) + r.sigma;
    put_blob(s, a.blob, a.leafname_off[2 * li], a.leafname_off[2 * li + 1]);
    if (a.app == 2) { s.put('\t'); put_blob(s, a.blob, a.gname_off[2 * r.gtree], a.gname_off[2 * r.gtree + 1]); }    // DomainTreeInHostTreeCreator.getLeafMap :1266
    s.put('\n');
}

__device__ __forceinline__ uint16_t clamp_len(long long n) { return (uint16_t)(n > 0xffff ? 0xffff : n); }

// ---- records from a vertex ---------------------------------------------------------------------------------------------------------
// Raw record (label kernels): every field the piece prints, the floating-point value as its bits with e = kRawValue.
constexpr int kRawValue = 0x7ffe;
template <bool PRUNED> __device__ __forceinline__ TreeRec raw_tree_rec(const EmitArgs& a, const Node& x, bool is_root, TreeRecX* xout, int link) {
    TreeRec t;
    t.f = (uint64_t)__double_as_longlong(PRUNED ? x.p_bl : x.bl); t.e = (int16_t)kRawValue;
    t.link = (PRUNED && __double_as_longlong(x.p_bl) == __double_as_longlong(x.bl)) ? link : -1;
    t.number = x.number; t.sigma = x.sigma; t.epoch = x.epoch; t.event = x.event; t.gtree = x.gtree;
    const bool leaf = PRUNED ? x.p_left < 0 : x.left < 0;
    t.chain = leaf ? (PRUNED ? x.chain_p : x.chain_u) : (int16_t)0;
    uint8_t fl = leaf ? TR_LEAF : 0;
    if (x.flags & (PRUNED ? NF_LEFT_P : NF_LEFT_U)) fl |= TR_COMMA;
    if (is_root) fl |= TR_ROOT;
    const bool prints_discpt = x.event == EV_DUPLICATION || x.event == EV_SPECIATION || x.event == EV_CODIVERGENCE || is_transfer_event(x.event);
    if (!a.exclude_meta && prints_discpt) fl |= (uint8_t)(discpt_index(a, x.gtree, x.epoch, x.abstime) << TR_J_SHIFT);
    t.flags = fl;
    if (is_transfer_event(x.event) && xout) { TreeRecX xr; xr.from_arc = x.from_arc; xr.to_arc = x.to_arc; xr.from_g = x.from_g; xr.to_g = x.to_g; xr.pad = 0; *xout = xr; }
    return t;
}
template <bool PRUNED> __device__ __forceinline__ RowRec raw_row_rec(const Node& x, int link) {
    RowRec r;
    r.f = (uint64_t)__double_as_longlong(x.abstime); r.e = (int16_t)kRawValue; r.len_g2h = 0; r.link = PRUNED ? link : -1;
    r.number = x.number; r.sigma = x.sigma; r.epoch = x.epoch; r.gtree = x.gtree; r.event = x.event;
    const bool leaf = PRUNED ? x.p_left < 0 : x.left < 0;
    r.flags = (leaf && (x.event == EV_LEAF || x.event == EV_UNSAMPLED_LEAF)) ? RR_LEAFMAP : 0;
    return r;
}
// Finished record (k_finish_records): digits of the value, length of the piece. Returns the unclamped length.
// `twin`: finished unpruned-view records of the replicate; a linked pruned-view record takes its digits from there.
template <bool PRUNED> __device__ __forceinline__ long long finish_tree_rec(const EmitArgs& a, TreeRec& t, const TreeRecX* xp, const TreeRec* twin) {
    const int link = t.link;
    int nd;
    if (PRUNED && twin && link >= 0) { t.f = twin[link].f; t.e = twin[link].e; nd = pad_nd(twin[link].pad); }
    else { int e_nd; t.f = decimal_of_packed(a.T.g, a.T.p10, __longlong_as_double((long long)t.f), &e_nd); t.e = (int16_t)(e_nd & 0xffff); nd = e_nd >> 16; }
    t.pad = (int16_t)pack_pad(nd, t.number, t.sigma, t.epoch);            // (pad shares its word with link: from here on the record is finished)
    CountSink c; put_tree_piece<PRUNED>(c, a, t, xp);
    t.len = clamp_len(c.n);
    return c.n;
}
// A linked pruned-view row is the unpruned-view row of the same vertex (same text); only the leaf-map flag is its own.
__device__ __forceinline__ void finish_row_rec(const EmitArgs& a, RowRec& r, bool need_g2h, bool need_lm, const RowRec* twin, bool twin_has_g2h, long long* len_g2h, long long* len_lm) {
    long long lg = 0, ll = 0;
    const int link = r.link;
    int pad = 0;
    if (need_g2h) {
        // a linked row prints the same vertex (same number, host vertex, epoch and time) as its twin: the whole pad word carries over
        if (twin && twin_has_g2h && link >= 0) { r.f = twin[link].f; r.e = twin[link].e; pad = twin[link].pad; r.pad = (uint16_t)pad; lg = twin[link].len_g2h; if (lg == 0xffff) { CountSink c; put_g2h_row(c, a, r); lg = c.n; } }
        else {
            int e_nd; r.f = decimal_of_packed(a.T.g, a.T.p10, __longlong_as_double((long long)r.f), &e_nd); r.e = (int16_t)(e_nd & 0xffff);
            pad = pack_pad(e_nd >> 16, r.number, r.sigma, r.epoch);
            r.pad = (uint16_t)pad; CountSink c; put_g2h_row(c, a, r); lg = c.n;
        }
    } else { pad = pack_pad(0, r.number, r.sigma, r.epoch); r.pad = (uint16_t)pad; }
    if (need_lm && (r.flags & RR_LEAFMAP)) { CountSink c; put_leafmap_row(c, a, r); ll = c.n; }
    r.len_g2h = clamp_len(lg); r.len_lm = clamp_len(ll); r.pad = (uint16_t)pad; *len_g2h = lg; *len_lm = ll;
}

}  // namespace sgp
