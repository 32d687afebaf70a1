// Guest-in-host sampler (device side).
//
//  general engine  — one thread simulates one attempt exactly as the reference does, vertex by vertex in the reference's
//                    queue order, so that with MtStream the random stream is consumed in the reference's order
//                    (apps/SaGePhy/GuestTreeInHostTreeCreator.java:110-321 createUnprunedTree, :372-423 createGuestVertex,
//                    topology/Epoch.java:249-329 sampleArc/setDistance, topology/RBTreeEpochDiscretiser.java:662-683 sampleRoot).
//  BD fast path    — transfer-free processes on a 1- or 2-leaf host (HostTreeGen) in Philox mode: depth-first expansion with
//                    a small stack, vertex v of attempt a draws Philox block (v, a, replicate), so the count-only pass and
//                    the build pass regenerate identical trees.
#pragma once
#include "kernels_args.cuh"
#include "rng.cuh"
#include "jmath_ext.cuh"

namespace sgp {

struct VertexDraw { double abstime, bl; int32_t epoch; uint8_t event; };

// createGuestVertex (GuestTreeInHostTreeCreator.java:372-423). Exact Java arithmetic (jlog, true division, round8).
template <class RNG>
__device__ __forceinline__ VertexDraw draw_vertex(RNG& rng, const SimParams& P, const DevHost& H, const MathTables& T, int X, double start) {
    VertexDraw d;
    const bool isRoot = H.parent[X] < 0;
    double sum = isRoot ? XADD(P.lambda, P.mu) : XADD(XADD(P.lambda, P.mu), P.tau);
    if (sum == 0.0) sum = 1e-48;
    const double low = H.vtime[X];
    double u1 = rng.next_double();
    double bt = XDIV(-jlog(XSUB(1.0, u1)), sum);
    double et = XSUB(start, bt);
    if (et <= low) {
        et = low;
        bt = round_sig(T, XSUB(start, et), 8);
        if (H.left[X] < 0) d.event = (rng.next_double() < P.rho) ? EV_LEAF : EV_UNSAMPLED_LEAF;
        else d.event = EV_SPECIATION;
        d.epoch = H.v2e[X];
    } else {
        double rnd = rng.next_double();
        if (isRoot) d.event = (rnd < XDIV(P.lambda, sum)) ? EV_DUPLICATION : EV_LOSS;
        else if (rnd >= XDIV(XADD(P.mu, P.tau), sum)) d.event = EV_DUPLICATION;
        else if (rnd < XDIV(P.mu, sum)) d.event = EV_LOSS;
        else { double trn = rng.next_double(); d.event = (trn < P.theta) ? EV_REPLACING_TRANSFER : EV_ADDITIVE_TRANSFER; }
        int e = H.v2e[X];
        while (H.ep_up[e] < et) ++e;
        d.epoch = e;
    }
    d.abstime = et; d.bl = bt;
    return d;
}

// double-double helpers for the cumulative-weight selection (stands in for the reference's exact BigDecimal sums)
struct DD { double hi, lo; };
__device__ __forceinline__ DD dd_add_d(DD a, double b) {
    double s = a.hi + b; double bb = s - a.hi; double err = (a.hi - (s - bb)) + (b - bb);
    double lo = a.lo + err; double hi = s + lo; lo = lo - (hi - s);
    return DD{hi, lo};
}
__device__ __forceinline__ DD dd_mul_d(DD a, double b) {
    double p = a.hi * b; double e = fma(a.hi, b, -p); e = fma(a.lo, b, e);
    double hi = p + e; double lo = e - (hi - p);
    return DD{hi, lo};
}
__device__ __forceinline__ bool dd_gt(DD a, DD b) { return a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo); }

// Math.pow(Math.E, y) through the fdlibm pow restatement (bit-identical to the oracle's)
static __device__ __noinline__ double pow_e(double y) { return jpow(2.718281828459045, y); }

// Epoch.sampleArc (include: recipients must carry the same / a different HOST tag as the donor). marks[x] == mark_gen flags an "empty" arc. Returns -5 if no candidate.
template <class RNG>
__device__ int sample_arc(RNG& rng, const SimParams& P, const DevHost& H, int epoch, int donor, double t, const int32_t* marks, int mark_gen, bool include = true) {
    const int a0 = H.ep_off[epoch], a1 = H.ep_off[epoch + 1];
    const int tag = H.host_tag[donor];
    if (P.bias == 0) {
        int k = 0;
        for (int j = a0; j < a1; ++j) { int x = H.ep_arcs[j]; if (x != donor && !(marks && marks[x] == mark_gen) && ((H.host_tag[x] == tag) == include)) ++k; }
        if (k < 1) return -5;
        int pick = rng.next_int(k);
        for (int j = a0; j < a1; ++j) { int x = H.ep_arcs[j]; if (x != donor && !(marks && marks[x] == mark_gen) && ((H.host_tag[x] == tag) == include)) { if (pick == 0) return x; --pick; } }
        return -5;
    }
    // biased: weights in epoch arc order; key = running total; pick the first key > total * U, where an arc with
    // weight 0 shares (and therefore overwrites) the previous key (TreeMap.put semantics, Epoch.java:323-324).
    DD total{0.0, 0.0}; int k = 0;
    for (int j = a0; j < a1; ++j) {
        int x = H.ep_arcs[j];
        if (x == donor || (marks && marks[x] == mark_gen) || ((H.host_tag[x] == tag) != include)) continue;
        double dist = XMUL(2.0, XSUB(H.lca_time[(size_t)donor * H.nV + x], t));
        double w = (P.bias == 2) ? XMUL(P.bias_rate, pow_e(XMUL(-P.bias_rate, dist))) : XDIV(1.0, dist);
        if (w == dbl_inf()) w = 1.7976931348623157e308;
        total = dd_add_d(total, w); ++k;
    }
    if (k < 1) return -5;
    DD rnd = dd_mul_d(total, rng.next_double());
    DD cum{0.0, 0.0}; int chosen = -5; bool found = false;
    for (int j = a0; j < a1; ++j) {
        int x = H.ep_arcs[j];
        if (x == donor || (marks && marks[x] == mark_gen) || ((H.host_tag[x] == tag) != include)) continue;
        double dist = XMUL(2.0, XSUB(H.lca_time[(size_t)donor * H.nV + x], t));
        double w = (P.bias == 2) ? XMUL(P.bias_rate, pow_e(XMUL(-P.bias_rate, dist))) : XDIV(1.0, dist);
        if (w == dbl_inf()) w = 1.7976931348623157e308;
        if (found) { if (w == 0.0) { chosen = x; continue; } break; }
        cum = dd_add_d(cum, w);
        if (dd_gt(cum, rnd)) { chosen = x; found = true; }
    }
    return chosen;
}

// RBTreeEpochDiscretiser.sampleRoot
template <class RNG>
__device__ int sample_root(RNG& rng, const SimParams& P, const DevHost& H) {
    const double tip = H.ep_up[H.nE - 1];
    DD total{0.0, 0.0};
    for (int i = 0; i < H.nV; ++i) total = dd_add_d(total, XMUL(P.gene_birth, pow_e(XMUL(-P.gene_birth, XSUB(tip, H.vtime[i])))));
    DD rnd = dd_mul_d(total, rng.next_double());
    DD cum{0.0, 0.0}; int chosen = H.nV - 1; bool found = false;
    for (int i = 0; i < H.nV; ++i) {
        double w = XMUL(P.gene_birth, pow_e(XMUL(-P.gene_birth, XSUB(tip, H.vtime[i]))));
        if (found) { if (w == 0.0) { chosen = i; continue; } break; }
        cum = dd_add_d(cum, w);
        if (dd_gt(cum, rnd)) { chosen = i; found = true; }
    }
    return chosen;
}

// Workspace of one simulating thread.
struct SimWork {
    Node* nodes; int cap;         // node pool
    int32_t* queue;               // FIFO of lineages ("alive"), cap entries
    int32_t* pend;                // parked replacing transfers ("sortedlist"), cap entries
    int32_t* marks;               // per host vertex: generation stamp of "empty arc" marks (nV entries) or nullptr
    int32_t* leaf_cnt;            // per host leaf counters (nLeaves entries) or nullptr when the per-leaf window is vacuous
    int mark_gen;
};

__device__ __forceinline__ int new_node(SimWork& W, int& n, const VertexDraw& d, int sigma) {
    if (n >= W.cap) return -1;
    Node& v = W.nodes[n];
    v.abstime = d.abstime; v.bl = d.bl; v.sigma = sigma; v.left = -1; v.right = -1; v.parent = -1; v.epoch = d.epoch;
    v.from_arc = -1; v.to_arc = -1; v.number = -1; v.p_bl = 0.0; v.p_left = -1; v.p_right = -1;
    v.event = d.event; v.prun = PR_UNKNOWN; v.flags = 0; v.pad0 = 0; v.gtree = 0; v.from_g = -1; v.to_g = -1; v.chain_u = 0; v.chain_p = 0; v.pad1 = 0;
    return n++;
}

// findVertex (GuestTreeInHostTreeCreator.java:332-346): first vertex in left-before-right pre-order on arc `to`
// whose branch strictly spans time t.
static __device__ int find_victim(const Node* N, int root, int to, double t) {
    int v = root;
    while (v >= 0) {
        const Node& x = N[v];
        if (x.sigma == to && x.abstime < t && XADD(x.abstime, x.bl) > t) return v;
        if (x.left >= 0) { v = x.left; continue; }
        for (;;) {
            int p = N[v].parent;
            if (p < 0) return -1;
            if (N[p].left == v) { v = N[p].right; break; }
            v = p;
        }
    }
    return -1;
}

// One attempt of createUnprunedTree. Returns 0, or REP_NODE_OVERFLOW. n_out = vertices created, root_out = root index.
//
// Written as a state machine with ONE vertex-creation site: every loop trip either fetches the next lineage from the queue
// (stage 0) or creates exactly one vertex (stages 1-3) through the single draw_vertex call below. Threads of a warp simulate
// different replicates, so they can only re-converge on code that is textually shared — the expensive Java-exact vertex draw
// (log, divide, round8) is therefore executed by all creating lanes together instead of at eight separate call sites.
// The order in which random numbers are consumed is exactly the reference's (coin -> staying child -> sampleArc(s) ->
// moving child; left child before right child).
template <class RNG>
__device__ int simulate_attempt(RNG& rng, const SimParams& P, const DevHost& H, const MathTables& T, SimWork& W, int& n_out, int& root_out) {
    int n = 0, qh = 0, qt = 0, np = 0;
    int stage = 3;                 // 3: create the root, 0: fetch a lineage, 1: create first vertex of the pair, 2: create second
    int li = -1, c1 = -1; uint8_t ev = 0; bool stay_left = true;
    int X; double t;               // arc and start time of the vertex to create next
    if (P.ishost || !P.do_gene_birth) { X = H.root; t = H.ep_up[H.nE - 1]; }
    else { X = sample_root(rng, P, H); t = H.ep_up[H.v2e[X]]; }
    int root = -1, sigma = 0, ep = 0, first_to = -1, victim = -1;
    for (;;) {
        if (stage == 0) {
            if (qh == qt) {
                if (np == 0) break;
                int best = 0;      // release the oldest parked replacing transfer (largest abstime)
                for (int i = 1; i < np; ++i) if (W.nodes[W.pend[i]].abstime > W.nodes[W.pend[best]].abstime) best = i;
                W.queue[qt++] = W.pend[best];
                W.pend[best] = W.pend[--np];
            }
            li = W.queue[qh++];
            const Node& ln = W.nodes[li];
            ev = ln.event;
            if (ev == EV_LOSS || ev == EV_REPLACING_LOSS || ev == EV_LEAF || ev == EV_UNSAMPLED_LEAF) continue;     // lineage ends
            sigma = ln.sigma; t = ln.abstime; ep = ln.epoch;
            if (ev == EV_SPECIATION) X = H.left[sigma];
            else { X = sigma; if (ev != EV_DUPLICATION) stay_left = rng.next_double() < 0.5; }
            stage = 1;
            continue;
        }
        if (stage == 2) {
            // arc of the second vertex
            if (ev == EV_SPECIATION) X = H.right[sigma];
            else if (ev == EV_DUPLICATION) X = sigma;
            else if (ev == EV_ADDITIVE_TRANSFER) { X = sample_arc(rng, P, H, ep, sigma, t, nullptr, 0); first_to = X; victim = -1; }
            else {   // EV_REPLACING_TRANSFER: look for a lineage to replace, arc by arc
                first_to = sample_arc(rng, P, H, ep, sigma, t, nullptr, 0);
                int to = first_to;
                victim = find_victim(W.nodes, root, to, t);
                const int n_arcs = H.ep_off[ep + 1] - H.ep_off[ep];
                int n_empty = 0; const int gen = ++W.mark_gen;
                while (victim < 0 && n_empty != n_arcs - 1) {
                    if (W.marks[to] != gen) { W.marks[to] = gen; ++n_empty; }
                    if (n_empty != n_arcs - 1) { to = sample_arc(rng, P, H, ep, sigma, t, W.marks, gen); victim = find_victim(W.nodes, root, to, t); }
                }
                X = victim >= 0 ? to : first_to;
            }
        }
        // ---- the single vertex-creation site
        const VertexDraw d = draw_vertex(rng, P, H, T, X, t);
        const int idx = new_node(W, n, d, X);
        if (idx < 0) { n_out = n; return REP_NODE_OVERFLOW; }
        if (stage == 3) { root = idx; W.queue[qt++] = root; stage = 0; continue; }
        if (stage == 1) { c1 = idx; stage = 2; continue; }
        // ---- stage 2 done: attach the pair
        int lc, rc;
        if (ev == EV_SPECIATION || ev == EV_DUPLICATION) { lc = c1; rc = idx; }
        else {
            Node& ln = W.nodes[li];
            ln.from_arc = sigma; ln.to_arc = X;
            if (ev == EV_REPLACING_TRANSFER) {
                if (victim >= 0) {
                    Node& vn = W.nodes[victim];
                    vn.event = EV_REPLACING_LOSS; vn.abstime = t;
                    int w = 0;          // drop parked replacing transfers that sit in the victim's subtree (victim included)
                    for (int i = 0; i < np; ++i) {
                        int x = W.pend[i]; bool under = false;
                        for (int y = x; y >= 0; y = W.nodes[y].parent) if (y == victim) { under = true; break; }
                        if (!under) W.pend[w++] = x;
                    }
                    np = w;
                    vn.left = -1; vn.right = -1;
                } else ln.event = EV_ADDITIVE_TRANSFER;      // nothing to replace anywhere: degrade (GuestTreeInHostTreeCreator.java:228-233)
            }
            lc = stay_left ? c1 : idx; rc = stay_left ? idx : c1;
        }
        W.nodes[li].left = lc; W.nodes[li].right = rc;
        W.nodes[lc].parent = li; W.nodes[rc].parent = li;
        for (int s = 0; s < 2; ++s) {
            const int c = s == 0 ? lc : rc;
            if (W.nodes[c].event != EV_REPLACING_TRANSFER) W.queue[qt++] = c;
            else {
                bool replaced = false;     // TreeMap.put on an equal key overwrites the value
                for (int i = 0; i < np; ++i) if (W.nodes[W.pend[i]].abstime == W.nodes[c].abstime) { W.pend[i] = c; replaced = true; break; }
                if (!replaced) W.pend[np++] = c;
            }
        }
        stage = 0;
    }
    if (W.nodes[root].bl <= 1.0e-32) W.nodes[root].bl = 0.0;
    n_out = n; root_out = root;
    return REP_OK;
}

// GuestTreeMachina.unprunedIsOK (:124-169) on the attached tree. Returns sampled-leaf count through *sampled.
static __device__ bool attempt_ok(const SimParams& P, const DevHost& H, const SimWork& W, int root, int exact, int* sampled_out) {
    int sampled = 0;
    const bool per_leaf = W.leaf_cnt != nullptr;
    if (per_leaf) for (int i = 0; i < H.nLeaves; ++i) W.leaf_cnt[i] = 0;
    int v = root;
    while (v >= 0) {
        const Node& x = W.nodes[v];
        if (x.event == EV_LEAF) { ++sampled; if (per_leaf) W.leaf_cnt[H.leaf_rank[x.sigma]]++; }
        if (x.left >= 0) { v = x.left; continue; }
        for (;;) {
            int p = W.nodes[v].parent;
            if (p < 0) { v = -1; break; }
            if (W.nodes[p].left == v) { v = W.nodes[p].right; break; }
            v = p;
        }
    }
    *sampled_out = sampled;
    if (exact != -1 && sampled != exact) return false;
    if (sampled < P.min_leaves || sampled > P.max_leaves) return false;
    if (per_leaf) for (int i = 0; i < H.nLeaves; ++i) { int c = W.leaf_cnt[i]; if (c < P.minper || c > P.maxper) return false; }
    return true;
}

// ---- BD fast path ----------------------------------------------------------------------------------------------------------
// Throughput-mode vertex step for transfer-free processes on a tiny host. Not bound to Java's bit patterns (Philox
// mode is validated distributionally), but the find pass and the build pass call this same function, so they regenerate
// identical trees. One Philox4x32-10 block per vertex: words (x,y) -> waiting time, (z,w) -> event / leaf-sampling draw.
constexpr int kBdStack = 96;

struct BdVertex { double et, bt; uint8_t event; bool boundary; };

struct BdConsts {
    double inv_sum, thr_root, thr_nonroot, rho;     // 1/(lambda+mu); P(dup) on the root arc; mu/sum on other arcs
};
__device__ __forceinline__ BdConsts bd_consts(const SimParams& P) {
    double sum = P.lambda + P.mu; if (sum == 0.0) sum = 1e-48;
    BdConsts c; c.inv_sum = 1.0 / sum; c.thr_root = P.lambda / sum; c.thr_nonroot = P.mu / sum; c.rho = P.rho;
    return c;
}
// 52 random bits -> x in [1,2); 2 - x = 1 - u lies in (0,1] (never 0, so the logarithm is finite)
__device__ __forceinline__ double one_minus_u(uint32_t a, uint32_t b) {
    return 2.0 - __hiloint2double((int)(0x3ff00000u | (a >> 12)), (int)b);
}
__device__ __forceinline__ double u01_52(uint32_t a, uint32_t b) {
    return __hiloint2double((int)(0x3ff00000u | (a >> 12)), (int)b) - 1.0;
}
// Natural logarithm for 0 < x <= 1 with x a normal double (one_minus_u never produces 0, subnormals, infinities or NaNs), so
// none of the special-case handling of the library routine is needed. Table-driven (Tang): the top 7 mantissa bits select
// c_j ~ m, r = m / c_j - 1 with |r| < 2^-8, log x = (e + K_j) ln 2 + log c_j + log1p(r) with a degree-6 polynomial — 10 FP64
// instructions instead of the 27 of the reciprocal + degree-14 version. Buckets above sqrt(2) are treated as m / 2 (K_j = 1) so
// that arguments just below 1 keep a small table term. Absolute error < 3e-16 (about 1 ulp for x < 0.9), which is what an
// exponential waiting time needs. Throughput mode only; the find and build passes both call it, so they agree bit for bit.
// The table (generated by tools/gen_tables.py) is copied to shared memory by each block: lanes index it with different j.
constexpr int kLogTabSplit = 53;
__device__ static const double kLogTab[256] = {
#include "log_tables.inc"
};
__constant__ static double kLogPolyC[4] = {1.0 / 3.0, 0.2, -1.0 / 6.0, 6.93147180559945286227e-01};
__device__ __forceinline__ void load_log_table(double2* s_tab) {
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_tab[i] = make_double2(kLogTab[2 * i], kLogTab[2 * i + 1]);
    __syncthreads();
}
__device__ __forceinline__ double log_unit(double x, const double2* s_tab) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int j = (hi >> 13) & 0x7f;
    const int e = (hi >> 20) - 1023 + (j >= kLogTabSplit ? 1 : 0);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 c = s_tab[j];
    const double r = fma(m, c.x, -1.0);
    double p = fma(r, kLogPolyC[2], kLogPolyC[1]);       // log1p(r) = r + r^2 (-1/2 + r (1/3 + r (-1/4 + r (1/5 - r/6))))
    p = fma(r, p, -0.25);
    p = fma(r, p, kLogPolyC[0]);
    p = fma(r, p, -0.5);
    p = fma(__dmul_rn(r, r), p, r);
    return __dadd_rn(fma((double)e, kLogPolyC[3], c.y), p);
}

__device__ __forceinline__ BdVertex bd_vertex(const BdConsts& C, const TinyHost& H, int X, double start, U4 w, const double2* s_tab) {
    BdVertex r;
    const bool isRoot = H.parent[X] < 0, hostLeaf = H.left[X] < 0;
    const double low = H.vtime[X];
    // explicit (non-contractable) multiply / subtract: the find and build kernels must round identically
    const double bt = __dmul_rn(-log_unit(one_minus_u(w.x, w.y), s_tab), C.inv_sum);
    const double et = __dsub_rn(start, bt);
    const double u2 = u01_52(w.z, w.w);
    r.boundary = et <= low;
    r.et = r.boundary ? low : et;
    r.bt = bt;
    const uint8_t ev_boundary = hostLeaf ? (u2 < C.rho ? EV_LEAF : EV_UNSAMPLED_LEAF) : EV_SPECIATION;
    const uint8_t ev_inner = isRoot ? (u2 < C.thr_root ? EV_DUPLICATION : EV_LOSS) : (u2 >= C.thr_nonroot ? EV_DUPLICATION : EV_LOSS);
    r.event = r.boundary ? ev_boundary : ev_inner;
    return r;
}


// Single-arc host ("H0:T;"): every boundary vertex is a host leaf, every interior event is on the root arc.
__device__ __forceinline__ BdVertex bd_vertex_single(const BdConsts& C, double start, U4 w, const double2* s_tab) {
    BdVertex r;
    const double bt = __dmul_rn(-log_unit(one_minus_u(w.x, w.y), s_tab), C.inv_sum);
    const double et = __dsub_rn(start, bt);
    const double u2 = u01_52(w.z, w.w);
    r.boundary = et <= 0.0;
    r.et = r.boundary ? 0.0 : et;
    r.bt = bt;
    r.event = r.boundary ? (u2 < C.rho ? EV_LEAF : EV_UNSAMPLED_LEAF) : (u2 < C.thr_root ? EV_DUPLICATION : EV_LOSS);
    return r;
}

}  // namespace sgp
