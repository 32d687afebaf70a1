// K3: labelling, pruning and per-tree statistics. No recursion: parent links make every traversal stackless.
//
// Reference behaviour (apps/SaGePhy/PruningHelper.java):
//   :25-65  labelUnprunableVertices — prunability lattice; survivors numbered 0.. in post-order
//   :76-90  labelPrunableVertices   — remaining vertices numbered after them, again in post-order
//   :98-145 prune                   — pruned copy; a vertex with one surviving child collapses into that child, whose length
//                                      becomes round8(child + parent)
// plus the breadth-first reductions of getInfo (GuestTreeInHostTreeCreator.java:431-511, HostTreeGen.java:124-174).
//
// Two kernels share the traversal code:
//   k_label_prune_warp — one warp per tree, the tree's node records staged in shared memory (coalesced 16-byte loads and
//                        stores, each record crosses HBM exactly once in each direction), traversals as level-synchronous
//                        sweeps with all lanes busy.
//   k_label_prune      — one thread per tree straight on global memory with the reference's serial traversals, for trees that
//                        do not fit the shared-memory stage.
#pragma once
#include "emit_format.cuh"

namespace sgp {

// Post-order walk: prunability, numbering, pruned structure. ordU / ordP receive the post-orders, pr is scratch (pruned
// representative of each subtree). Vertices that do not survive get number = -2 - (rank among them), fixed up afterwards.
template <class I>
__device__ __forceinline__ int label_postorder(Node* N, I* ordU, I* ordP, I* pr, int root, const MathTables& T, int& nu_out, int& up_out) {
    int nu = 0, np = 0, up = 0;
    int v = root;
    for (;;) {
        while (N[v].left >= 0) v = N[v].left;
        bool done = false;
        for (;;) {
            Node& x = N[v];
            const int xl = x.left;
            uint8_t prun;
            if (xl < 0) {
                const bool leaf = x.event == EV_LEAF;
                prun = leaf ? PR_UNPRUNABLE : PR_PRUNABLE;
                if (leaf) { pr[v] = (I)v; x.p_bl = x.bl; x.p_left = -1; x.p_right = -1; } else pr[v] = (I)-1;
            } else {
                const int xr = x.right;
                const uint8_t pl = N[xl].prun, pq = N[xr].prun;
                if (pl == PR_PRUNABLE && pq == PR_PRUNABLE) { prun = PR_PRUNABLE; pr[v] = (I)-1; }
                else if (pl == PR_PRUNABLE || pq == PR_PRUNABLE) {
                    prun = PR_COLLAPSABLE;
                    const int c = pr[pl == PR_PRUNABLE ? xr : xl];
                    N[c].p_bl = round_sig(T, XADD(N[c].p_bl, x.bl), 8);
                    pr[v] = (I)c;
                } else {
                    prun = PR_UNPRUNABLE; x.p_left = pr[xl]; x.p_right = pr[xr]; x.p_bl = x.bl; pr[v] = (I)v;
                }
            }
            x.prun = prun;
            if (prun == PR_UNPRUNABLE) { x.number = nu; ordP[nu] = (I)v; ++nu; } else { x.number = -2 - np; ++np; }
            ordU[up++] = (I)v;
            const int p = x.parent;
            if (p < 0) { done = true; break; }
            if (N[p].left == v) { v = N[p].right; break; }
            v = p;
        }
        if (done) break;
    }
    nu_out = nu; up_out = up;
    return pr[root];
}

// Breadth-first pass over the unpruned (PRUNED = false) or pruned view: BFS order into q, getInfo sums in BFS order, event
// counts, left-child marks and '(' run lengths (a parent is always visited before its children). The pruned pass writes its
// marks into Node::pad0 so that both passes can run concurrently; the caller folds them into Node::flags afterwards.
template <class I>
__device__ __forceinline__ void label_bfs(Node* N, I* q, int start, bool pruned, double host_root_time, int32_t* cnt, double& tot_out, double& ben_out) {
    double tot = 0.0, ben = 0.0;
    if (start >= 0) {
        int h = 0, t = 0; q[t++] = (I)start;
        while (h < t) {
            Node& x = N[q[h++]];
            const int l = pruned ? x.p_left : x.left, r = pruned ? x.p_right : x.right;
            const double len = pruned ? x.p_bl : x.bl;
            if (pruned) x.pad0 |= NF_IN_PRUNED;
            if (l >= 0) {
                q[t++] = (I)l; q[t++] = (I)r;
                if (pruned) { Node& L = N[l]; L.pad0 |= NF_LEFT_P; L.chain_p = (int16_t)(x.chain_p + 1); N[r].chain_p = 0; }
                else { Node& L = N[l]; L.flags |= NF_LEFT_U; L.chain_u = (int16_t)(x.chain_u + 1); N[r].chain_u = 0; }
            }
            tot = XADD(tot, len);
            ben = XADD(ben, fmax(fmin(XSUB(host_root_time, x.abstime), len), 0.0));
            cnt[x.event]++;
        }
    }
    tot_out = tot; ben_out = ben;
}

// Shared-memory bytes per vertex of the warp kernel: the record + nine int16 work arrays.
constexpr int kLabelBytesPerNode = (int)sizeof(Node) + 9 * (int)sizeof(int16_t);
constexpr int kLabelExtraBytes = 16;
__host__ __device__ inline size_t label_smem_bytes(int cap) { return (size_t)cap * kLabelBytesPerNode + kLabelExtraBytes; }

// ---- bulk asynchronous copy global -> shared (TMA) with an mbarrier that counts the bytes as they land ----------------------------
__device__ __forceinline__ void mbar_init(uint32_t mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* src, uint32_t bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory"); }

__device__ __forceinline__ void store_rec32(void* dst, const void* src) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src); uint4* d4 = reinterpret_cast<uint4*>(dst);
    d4[0] = s4[0]; d4[1] = s4[1];
}

// Raw emission records of one labelled tree (fields only: the floating-point value as its bits, no digits, no length);
// executed by a whole warp. N = vertex records (shared or global memory), ordU / ordP / bfsU / bfsP = the four orders (int16 in
// shared memory, int32 in global memory). k_finish_records (kernels_emit.cuh) turns the values into digits, measures every
// piece and sums the stream sizes at full occupancy - inside the label kernel, which holds one warp per tree and few trees per SM,
// that arithmetic took 60 ms per 10^6 C2 trees against 9 ms on its own.
// rankU / rankB (optional, shared-memory kernel): per-vertex scratch that receives the vertex's position in the unpruned post-order /
// breadth-first order, so that a pruned-view record can point at its unpruned twin (same digits, computed once).
template <class I>
__device__ __forceinline__ void make_records(const EmitArgs& E, long long goff, const Node* N, const I* ordU, const I* ordP, const I* bfsU, const I* bfsP,
                                             int up, int nu, int root, int p_root, I* rankU, I* rankB) {
    const int lane = threadIdx.x & 31;
    const unsigned mask = E.stream_mask;
    const bool tu = mask & (1u << ST_UNPRUNED_TREE), tp = (mask & (1u << ST_PRUNED_TREE)) && p_root >= 0;
    const bool ru = E.rows && (mask & ((1u << ST_UNPRUNED_G2H) | (1u << ST_UNPRUNED_LEAFMAP))), rp = E.rows && (mask & ((1u << ST_PRUNED_G2H) | (1u << ST_PRUNED_LEAFMAP))) && p_root >= 0;
    if (tu) {
        TreeRec* out = E.trec + goff; TreeRecX* outx = E.trecx ? E.trecx + goff : nullptr;
        for (int k = lane; k < up; k += 32) {
            const int v = ordU[k]; const TreeRec t = raw_tree_rec<false>(E, N[v], v == root, outx ? outx + k : nullptr, -1); store_rec32(out + k, &t);
            if (rankU && tp) rankU[v] = (I)k;
        }
    }
    if (ru) {
        RowRec* out = E.rows + goff;
        for (int k = lane; k < up; k += 32) { const int v = bfsU[k]; const RowRec rr = raw_row_rec<false>(N[v], -1); store_rec32(out + k, &rr); if (rankB && rp) rankB[v] = (I)k; }
    }
    __syncwarp();
    if (tp) {
        TreeRec* out = E.trec + E.aux_stride + goff; TreeRecX* outx = E.trecx ? E.trecx + E.aux_stride + goff : nullptr;
        for (int k = lane; k < nu; k += 32) {
            const int v = ordP[k]; const TreeRec t = raw_tree_rec<true>(E, N[v], v == p_root, outx ? outx + k : nullptr, (rankU && tu) ? (int)rankU[v] : -1); store_rec32(out + k, &t);
        }
    }
    if (rp) {
        RowRec* out = E.rows + E.aux_stride + goff;
        for (int k = lane; k < nu; k += 32) { const int v = bfsP[k]; const RowRec rr = raw_row_rec<true>(N[v], (rankB && ru) ? (int)rankB[v] : -1); store_rec32(out + k, &rr); }
    }
}

// One warp per tree, all lanes busy. The serial recursions of the reference are restated as level-synchronous sweeps over the
// breadth-first order (a tree of ~270 vertices has ~30 levels, so a sweep is ~35 warp steps instead of ~270 dependent ones):
//   1. breadth-first order of the unpruned tree, level boundaries, left-child marks / '(' runs, event counts
//   2. bottom-up: subtree size, survivors per subtree, prunability, pruned representative, pruned children
//   2b. collapse chains (one lane per chain, chains are independent): len = round8(len + parent) going up, PruningHelper.java:131-139
//   3. top-down: position of every vertex in the post-order (= vertices before its subtree + subtree size - 1), the same among
//      survivors only; vertex numbers (survivors first, PruningHelper.java:25-90) and both post-order arrays follow directly
//   4. breadth-first order of the pruned view with its marks and counts
//   5. the getInfo sums, which are order-dependent floating-point sums: lane 0 adds the unpruned, lane 1 the pruned tree
// Labels, prunes and numbers ONE tree with a whole warp (steps 1-5 above) and leaves its emission records, getInfo terms and
// RepInfo fields. sN = the tree's vertex records (shared memory, or in place in the arena for trees that do not fit), W = nine work
// arrays of `cap` (+2) entries of index type I, s_cnt = two zero-initialisable event histograms of the warp.
template <class I>
__device__ __forceinline__ void label_tree_by_warp(const LabelArgs& a, RepInfo& inf, long long goff, int n_pool, Node* sN, I* W, int cap,
                                                   int32_t (*s_cnt)[EV_COUNT], bool in_place) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    I* bfsU = W;
    I *ordU = bfsU + cap, *ordP = ordU + cap, *bfsP = ordP + cap, *pr = bfsP /* dead before bfsP is written */;
    I *sz = bfsP + cap, *ns = sz + cap, *off = ns + cap, *offS = off + cap, *lvl = offS + cap;     // lvl: cap + 2 entries
    const int root = inf.root;
    for (int i = lane; i < 2 * EV_COUNT; i += 32) (&s_cnt[0][0])[i] = 0;
    if (lane == 0) { bfsU[0] = (I)root; lvl[0] = 0; off[root] = 0; offS[root] = 0; }
    __syncwarp();
    {
        // right children the birth-death build kernel left unlinked (NF_RIGHT_CHILD)
        for (int v = lane; v < n_pool; v += 32) {
            Node& x = sN[v]; if (x.flags & NF_RIGHT_CHILD) { sN[x.parent].right = v; x.flags &= (uint8_t)~NF_RIGHT_CHILD; }
            x.p_left = -1; x.p_right = -1; x.flags &= (uint8_t)~NF_DETACHED;          // (the samplers use p_left as a per-arc list link)
        }
        __syncwarp();
        // ---- 1. breadth-first order, levels
        int D = 0, up;
        {
            int lo = 0, hi = 1;
            while (lo < hi) {
                ++D; if (lane == 0) lvl[D] = (I)hi;
                int t = hi;
                for (int b0 = lo; b0 < hi; b0 += 32) {
                    const int k = b0 + lane; int l = -1, rr = -1;
                    if (k < hi) {
                        Node& x = sN[bfsU[k]]; l = x.left; rr = x.right;
                        if (l >= 0) { Node& L = sN[l]; L.flags |= NF_LEFT_U; L.chain_u = (I)(x.chain_u + 1); sN[rr].chain_u = 0; }
                        atomicAdd(&s_cnt[0][x.event], 1);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, l >= 0);
                    if (l >= 0) { const int pos = t + 2 * __popc(m & lt_mask); bfsU[pos] = (I)l; bfsU[pos + 1] = (I)rr; }
                    t += 2 * __popc(m);
                }
                __syncwarp();
                lo = hi; hi = t;
            }
            up = hi;
        }
        // ---- 2. bottom-up
        for (int d = D - 1; d >= 0; --d) {
            const int lo = lvl[d], hi = lvl[d + 1];
            for (int k = lo + lane; k < hi; k += 32) {
                const int v = bfsU[k]; Node& x = sN[v];
                const int l = x.left;
                uint8_t prun; int size = 1, surv = 0;
                if (l < 0) {
                    const bool leaf = x.event == EV_LEAF;
                    prun = leaf ? PR_UNPRUNABLE : PR_PRUNABLE; surv = leaf ? 1 : 0;
                    if (leaf) { pr[v] = (I)v; x.p_bl = x.bl; x.p_left = -1; x.p_right = -1; } else pr[v] = -1;
                } else {
                    const int rr = x.right;
                    const uint8_t pl = sN[l].prun, pq = sN[rr].prun;
                    size += sz[l] + sz[rr]; surv = ns[l] + ns[rr];
                    if (pl == PR_PRUNABLE && pq == PR_PRUNABLE) { prun = PR_PRUNABLE; pr[v] = -1; }
                    else if (pl == PR_PRUNABLE || pq == PR_PRUNABLE) { prun = PR_COLLAPSABLE; pr[v] = pr[pl == PR_PRUNABLE ? rr : l]; }
                    else { prun = PR_UNPRUNABLE; ++surv; x.p_left = pr[l]; x.p_right = pr[rr]; x.p_bl = x.bl; pr[v] = (I)v; }
                }
                x.prun = prun; sz[v] = (I)size; ns[v] = (I)surv;
            }
            __syncwarp();
        }
        const int nu = ns[root], p_root = pr[root];
        // ---- 2b. collapse chains: gather the survivors whose parent collapses (list in ordU, which step 3 rewrites), then fold
        {
            int n_chain = 0;
            for (int b0 = 0; b0 < up; b0 += 32) {
                const int k = b0 + lane; bool start = false; int v = -1;
                if (k < up) { v = bfsU[k]; const Node& x = sN[v]; start = x.prun == PR_UNPRUNABLE && x.parent >= 0 && sN[x.parent].prun == PR_COLLAPSABLE; }
                const unsigned m = __ballot_sync(0xffffffffu, start);
                if (start) ordU[n_chain + __popc(m & lt_mask)] = (I)v;        // ordU is free until step 3
                n_chain += __popc(m);
            }
            __syncwarp();
            for (int k = lane; k < n_chain; k += 32) {
                Node& c = sN[ordU[k]];
                double len = c.p_bl; int p = c.parent;
                while (p >= 0 && sN[p].prun == PR_COLLAPSABLE) { len = round_sig(a.T, XADD(len, sN[p].bl), 8); p = sN[p].parent; }
                c.p_bl = len;
            }
            __syncwarp();
        }
        // ---- 3. top-down: post-order positions and vertex numbers
        for (int d = 0; d < D; ++d) {
            const int lo = lvl[d], hi = lvl[d + 1];
            for (int k = lo + lane; k < hi; k += 32) {
                const int v = bfsU[k]; Node& x = sN[v];
                const int o = off[v], os = offS[v], size = sz[v], surv = ns[v];
                const int post = o + size - 1;
                ordU[post] = (I)v;
                if (x.prun == PR_UNPRUNABLE) { const int ps = os + surv - 1; x.number = ps; ordP[ps] = (I)v; }
                else x.number = nu + post - (os + surv);
                const int l = x.left;
                if (l >= 0) { const int rr = x.right; off[l] = (I)o; offS[l] = (I)os; off[rr] = (I)(o + sz[l]); offS[rr] = (I)(os + ns[l]); }
            }
            __syncwarp();
        }
        if (lane == 0 && a.has_stem_override) { sN[root].bl = a.stem_override; if (p_root >= 0) sN[p_root].p_bl = a.stem_override; }
        // ---- 4. breadth-first order of the pruned view (pr is dead from here on)
        if (p_root >= 0) {
            if (lane == 0) bfsP[0] = (I)p_root;
            __syncwarp();
            int lo = 0, hi = 1;
            while (lo < hi) {
                int t = hi;
                for (int b0 = lo; b0 < hi; b0 += 32) {
                    const int k = b0 + lane; int l = -1, rr = -1;
                    if (k < hi) {
                        Node& x = sN[bfsP[k]]; l = x.p_left; rr = x.p_right;
                        x.flags |= NF_IN_PRUNED;
                        if (l >= 0) { Node& L = sN[l]; L.flags |= NF_LEFT_P; L.chain_p = (I)(x.chain_p + 1); sN[rr].chain_p = 0; }
                        atomicAdd(&s_cnt[1][x.event], 1);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, l >= 0);
                    if (l >= 0) { const int pos = t + 2 * __popc(m & lt_mask); bfsP[pos] = (I)l; bfsP[pos + 1] = (I)rr; }
                    t += 2 * __popc(m);
                }
                __syncwarp();
                lo = hi; hi = t;
            }
        }
        __syncwarp();
        // ---- 5. the terms of the getInfo sums in breadth-first order (GuestTreeInHostTreeCreator.java:431-511). The sums themselves are
        // order-dependent floating-point additions, i.e. serial: k_info_sums adds them up with a thread per sum and full occupancy
        // (here they kept one or two lanes of the warp busy for 44 % of the kernel's instructions).
        {
            double* tU = a.terms + goff; double* tP = a.terms + a.aux_stride + goff;
            double* bU = a.terms + 2 * a.aux_stride + goff; double* bP = a.terms + 3 * a.aux_stride + goff;
            for (int k = lane; k < up; k += 32) {
                const Node& x = sN[bfsU[k]]; tU[k] = x.bl;
                if (a.want_beneath) bU[k] = fmax(fmin(XSUB(a.host_root_time, x.abstime), x.bl), 0.0);
            }
            for (int k = lane; k < nu; k += 32) {
                const Node& x = sN[bfsP[k]]; tP[k] = x.p_bl;
                if (a.want_beneath) bP[k] = fmax(fmin(XSUB(a.host_root_time, x.abstime), x.p_bl), 0.0);
            }
        }
        {
            if (a.write_nodes && !in_place) {
                uint4* dst = reinterpret_cast<uint4*>(a.nodes + goff); const uint4* src = reinterpret_cast<const uint4*>(sN);
                const int nq = n_pool * (int)(sizeof(Node) / 16);
                for (int i = lane; i < nq; i += 32) dst[i] = src[i];
            }
            if (a.write_nodes && a.aux) {
                int32_t* gOrdU = a.aux + goff; int32_t* gOrdP = a.aux + a.aux_stride + goff;
                int32_t* gBfsU = a.aux + 2 * a.aux_stride + goff; int32_t* gBfsP = a.aux + 3 * a.aux_stride + goff;
                for (int k = lane; k < up; k += 32) { gOrdU[k] = ordU[k]; gBfsU[k] = bfsU[k]; }
                for (int k = lane; k < nu; k += 32) { gOrdP[k] = ordP[k]; gBfsP[k] = bfsP[k]; }
            }
            make_records<I>(a.E, goff, sN, ordU, ordP, bfsU, bfsP, up, nu, root, p_root, off, offS);      // (off / offS: dead work arrays)
            if (lane < EV_COUNT) { inf.cnt_u[lane] = s_cnt[0][lane]; inf.cnt_p[lane] = s_cnt[1][lane]; }
            if (lane == 0) {
                inf.p_root = p_root; inf.n_u = up; inf.n_p = nu; inf.p_root_time = p_root >= 0 ? sN[p_root].abstime : 0.0;
                inf.root_gtree_u = sN[root].gtree; inf.root_gtree_p = sN[p_root >= 0 ? p_root : root].gtree;
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32) k_label_prune_warp(LabelArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int32_t s_cnt[2][EV_COUNT];
    __shared__ __align__(8) unsigned long long s_mbar;
    const int cap = a.cap;
    Node* sN = reinterpret_cast<Node*>(smem);
    int16_t* work = reinterpret_cast<int16_t*>(smem + (size_t)cap * sizeof(Node));
    const int lane = threadIdx.x;
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar), sN_addr = (uint32_t)__cvta_generic_to_shared(sN);
    if (lane == 0) mbar_init(mbar, 1);
    __syncwarp();
    uint32_t phase = 0;
    // every trip looks at 32 consecutive replicates and works through those of this launch's size class
    for (long long base = (long long)blockIdx.x * 32; base < a.n; base += (long long)gridDim.x * 32) {
      int np_lane = 0; long long goff_lane = 0;
      if (base + lane < a.n) { np_lane = (int)a.pool[base + lane]; goff_lane = a.node_off[base + lane]; }         // 0 for failed replicates
      unsigned todo = __ballot_sync(0xffffffffu, np_lane > a.min_pool && np_lane <= a.cap);
      while (todo) {
        const int b = __ffs(todo) - 1; todo &= todo - 1;
        const long long r = base + b;
        const int n_pool = __shfl_sync(0xffffffffu, np_lane, b);
        RepInfo& inf = a.info[r];
        const long long goff = __shfl_sync(0xffffffffu, goff_lane, b);
        {
            // The tree's records arrive as ONE bulk copy (TMA) instead of a loop of 16-byte loads per lane; the next tree of this
            // group is pulled into L2 meanwhile. Shared memory was last touched through the generic proxy: fence, then copy.
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)n_pool * (uint32_t)sizeof(Node);
                mbar_expect_tx(mbar, bytes);
                bulk_load(sN_addr, a.nodes + goff, bytes, mbar);
            }
            if (todo) {
                const int nb = __ffs(todo) - 1;
                const long long ngoff = __shfl_sync(0xffffffffu, goff_lane, nb); const int nnp = __shfl_sync(0xffffffffu, np_lane, nb);
                if (lane == 0) prefetch_l2_bulk(a.nodes + ngoff, (uint32_t)nnp * (uint32_t)sizeof(Node));
            }
        }
        mbar_wait(mbar, phase); phase ^= 1u;
        __syncwarp();
        label_tree_by_warp<int16_t>(a, inf, goff, n_pool, sN, work, cap, s_cnt, false);
        __syncwarp();
      }
    }
}

// Trees that do not fit the shared-memory stage (> 2 340 vertices; a guest tree on a 1000-leaf host has ~3 400): the same warp
// algorithm on the records in place (the arena, L2-resident while a warp works on one tree) with 32-bit work arrays in a global scratch
// of the block. The thread-per-tree kernel below, which these trees used to go to, took 64 ms per 8 192 such trees.
__global__ void __launch_bounds__(32) k_label_prune_warp_global(LabelArgs a) {
    __shared__ int32_t s_cnt[2][EV_COUNT];
    int32_t* work = a.work + (size_t)blockIdx.x * a.work_stride;
    for (long long r = blockIdx.x; r < a.n; r += gridDim.x) {
        const int n_pool = (int)a.pool[r];
        if (n_pool <= a.min_pool || n_pool > a.cap) continue;
        label_tree_by_warp<int32_t>(a, a.info[r], a.node_off[r], n_pool, a.nodes + a.node_off[r], work, a.cap, s_cnt, true);
        __syncwarp();
    }
}

// Fallback for trees with more than `min_pool` vertices (those the warp kernel skipped): one thread per tree on global memory.
__global__ void __launch_bounds__(128) k_label_prune(LabelArgs a) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    RepInfo& inf = a.info[r];
    if (inf.status != REP_OK || inf.n_pool <= a.cap) return;
    const long long off = a.node_off[r];
    Node* N = a.nodes + off;
    int32_t* ordU = a.aux + off; int32_t* ordP = a.aux + a.aux_stride + off;
    int32_t* bfsU = a.aux + 2 * a.aux_stride + off; int32_t* bfsP = a.aux + 3 * a.aux_stride + off;
    const int root = inf.root;
    for (int v = 0; v < inf.n_pool; ++v) { Node& x = N[v]; if (x.flags & NF_RIGHT_CHILD) { N[x.parent].right = v; x.flags &= (uint8_t)~NF_RIGHT_CHILD; } x.p_left = -1; x.p_right = -1; x.flags &= (uint8_t)~NF_DETACHED; }
    int nu = 0, up = 0;
    const int p_root = label_postorder<int32_t>(N, ordU, ordP, bfsP, root, a.T, nu, up);
    for (int k = 0; k < up; ++k) { Node& x = N[ordU[k]]; if (x.number < -1) x.number = nu + (-2 - x.number); }
    if (a.has_stem_override) { N[root].bl = a.stem_override; if (p_root >= 0) N[p_root].p_bl = a.stem_override; }
    inf.p_root = p_root; inf.n_u = up; inf.n_p = nu; inf.p_root_time = p_root >= 0 ? N[p_root].abstime : 0.0;
    inf.root_gtree_u = N[root].gtree; inf.root_gtree_p = N[p_root >= 0 ? p_root : root].gtree;
    for (int e = 0; e < EV_COUNT; ++e) { inf.cnt_u[e] = 0; inf.cnt_p[e] = 0; }
    double tot, ben;
    label_bfs<int32_t>(N, bfsU, root, false, a.host_root_time, inf.cnt_u, tot, ben);
    inf.total_u = tot; inf.beneath_u = ben;
    label_bfs<int32_t>(N, bfsP, p_root, true, a.host_root_time, inf.cnt_p, tot, ben);
    inf.total_p = tot; inf.beneath_p = ben;
    for (int k = 0; k < up; ++k) { Node& x = N[ordU[k]]; x.flags |= x.pad0; x.pad0 = 0; }
}

// The getInfo sums of the trees the warp kernel labelled: one thread per (replicate, sum), terms read in breadth-first order.
__global__ void __launch_bounds__(128) k_info_sums(LabelArgs a) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t >> 2; const int which = (int)(t & 3);
    if (r >= a.n) return;
    RepInfo& inf = a.info[r];
    if (inf.status != REP_OK || inf.n_pool > a.cap) return;          // (larger trees: k_label_prune has summed them itself)
    if (which >= 2 && !a.want_beneath) { if (which == 2) inf.beneath_u = 0.0; else inf.beneath_p = 0.0; return; }
    const double* term = a.terms + (size_t)which * a.aux_stride + a.node_off[r];
    const int cnt = (which & 1) ? (inf.p_root >= 0 ? inf.n_p : 0) : inf.n_u;
    double sum = 0.0;
    for (int k = 0; k < cnt; ++k) sum = XADD(sum, term[k]);
    if (which == 0) inf.total_u = sum; else if (which == 1) inf.total_p = sum; else if (which == 2) inf.beneath_u = sum; else inf.beneath_p = sum;
}

// Emission records of the trees the thread kernel labelled (warp per tree, vertex records and orders in global memory).
__global__ void __launch_bounds__(128) k_records_from_nodes(LabelArgs a) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= a.n) return;
    const RepInfo& inf = a.info[r];
    if (inf.status != REP_OK || inf.n_pool <= a.cap) return;
    const long long off = a.node_off[r];
    make_records<int32_t>(a.E, off, a.nodes + off, a.aux + off, a.aux + a.aux_stride + off, a.aux + 2 * a.aux_stride + off, a.aux + 3 * a.aux_stride + off,
                          inf.n_u, inf.n_p, inf.root, inf.p_root, nullptr, nullptr);
}

}  // namespace sgp
