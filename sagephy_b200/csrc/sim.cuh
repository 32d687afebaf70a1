// Guest-in-host sampler (device side).
//
//  general engine  — one lane simulates one attempt exactly as the reference does, vertex by vertex in the reference's
//                    queue order, so that with MtStream the random stream is consumed in the reference's order
//                    (apps/SaGePhy/GuestTreeInHostTreeCreator.java:110-321 createUnprunedTree, :372-423 createGuestVertex,
//                    topology/Epoch.java:249-329 sampleArc/setDistance, topology/RBTreeEpochDiscretiser.java:662-683 sampleRoot).
//                    The 32 lanes of a warp advance in lock step (one vertex per lane per trip, sim_step) and serve each other's
//                    recipient draws and victim searches (coop_*), see kernels_sample.cuh.
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

// RBTreeEpochDiscretiser.sampleRoot
template <class RNG>
__device__ __noinline__ int sample_root(RNG& rng, const SimParams& P, const DevHost& H) {
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
    // assembled in registers, written as five 16-byte stores (field-wise stores touch 32 different sectors per warp each)
    Node v;
    v.abstime = d.abstime; v.bl = d.bl; v.sigma = sigma; v.left = -1; v.right = -1; v.parent = -1; v.epoch = d.epoch;
    v.from_arc = -1; v.to_arc = -1; v.number = -1; v.p_bl = 0.0; v.p_left = -1; v.p_right = -1;
    v.event = d.event; v.prun = PR_UNKNOWN; v.flags = 0; v.pad0 = 0; v.gtree = 0; v.from_g = -1; v.to_g = -1; v.chain_u = 0; v.chain_p = 0; v.pad1 = 0;
    const uint4* src = reinterpret_cast<const uint4*>(&v); uint4* dst = reinterpret_cast<uint4*>(&W.nodes[n]);
#pragma unroll
    for (int q = 0; q < (int)(sizeof(Node) / 16); ++q) dst[q] = src[q];
    return n++;
}

// ---- warp-cooperative pieces ---------------------------------------------------------------------------------------------------
// A warp simulates 32 replicates, one per lane, and all lanes take one step of their attempts per loop trip (sim_step). The two
// operations that are linear in the host epoch or in the guest tree — the recipient draw of a transfer (Epoch.sampleArc) and the
// victim search of a replacing transfer (findVertex) — are not run by the owning lane alone (31 lanes would wait for it): the
// lanes that need one are served one after the other by the whole warp. `leader` is the owning lane; all scalar arguments are
// warp-uniform (broadcast by the caller); only the leader's random stream is advanced.
__device__ __forceinline__ DD dd_add_dd(DD a, DD b) {
    double s = a.hi + b.hi; double bb = s - a.hi; double err = (a.hi - (s - bb)) + (b.hi - bb);
    double lo = (a.lo + b.lo) + err; double hi = s + lo; lo = lo - (hi - s);
    return DD{hi, lo};
}
__device__ __forceinline__ DD dd_shfl(DD v, int src) { return DD{__shfl_sync(0xffffffffu, v.hi, src), __shfl_sync(0xffffffffu, v.lo, src)}; }

__device__ __forceinline__ bool arc_eligible(const DevHost& H, int x, int donor, int tag, const int32_t* marks, int mark_gen, bool include) {
    return x != donor && !(marks && marks[x] == mark_gen) && ((H.host_tag[x] == tag) == include);
}
__device__ __forceinline__ double arc_weight(const SimParams& P, const DevHost& H, int x, int donor, double t) {
    const double dist = XMUL(2.0, XSUB(H.lca_time[(size_t)donor * H.nV + x], t));
    double w = (P.bias == 2) ? XMUL(P.bias_rate, pow_e(XMUL(-P.bias_rate, dist))) : XDIV(1.0, dist);
    if (w == dbl_inf()) w = 1.7976931348623157e308;
    return w;
}

// Epoch.sampleArc (topology/Epoch.java:249-329) by the whole warp. include: recipients must carry the same / a different HOST tag
// as the donor; marks[x] == mark_gen flags an "empty" arc; returns -5 if there is no candidate. Uniform mode: nextInt(k) over the
// eligible arcs in epoch order. Biased modes: weights in epoch arc order, key = running total, the first key > total * U wins and
// an arc of weight 0 shares (and therefore overwrites) the previous key (TreeMap.put semantics, Epoch.java:323-324); the double-double sums are formed by a warp scan instead of a serial loop, which moves them by
// ~1e-32 relative — the same order as the double-double stand-in for BigDecimal itself (numerics).
template <class RNG>
__device__ __noinline__ int coop_sample_arc(int leader, RNG& rng, const SimParams& P, const DevHost& H, int epoch, int donor, double t, const int32_t* marks, int mark_gen, bool include = true) {
    const int lane = threadIdx.x & 31;
    const int a0 = H.ep_off[epoch], a1 = H.ep_off[epoch + 1];
    const int tag = H.host_tag[donor];
    if (P.bias == 0) {
        int k = 0;
        for (int base = a0; base < a1; base += 32) {
            const int j = base + lane; const bool el = j < a1 && arc_eligible(H, H.ep_arcs[j], donor, tag, marks, mark_gen, include);
            k += __popc(__ballot_sync(0xffffffffu, el));
        }
        if (k < 1) return -5;
        int pick = 0; if (lane == leader) pick = rng.next_int(k);
        pick = __shfl_sync(0xffffffffu, pick, leader);
        for (int base = a0; base < a1; base += 32) {
            const int j = base + lane; const int x = j < a1 ? H.ep_arcs[j] : -1;
            const bool el = j < a1 && arc_eligible(H, x, donor, tag, marks, mark_gen, include);
            const unsigned m = __ballot_sync(0xffffffffu, el); const int c = __popc(m);
            if (pick < c) return __shfl_sync(0xffffffffu, x, __fns(m, 0, pick + 1));
            pick -= c;
        }
        return -5;
    }
    DD mine{0.0, 0.0}; int k = 0;
    for (int base = a0; base < a1; base += 32) {
        const int j = base + lane; const int x = j < a1 ? H.ep_arcs[j] : -1;
        const bool el = j < a1 && arc_eligible(H, x, donor, tag, marks, mark_gen, include);
        if (el) mine = dd_add_d(mine, arc_weight(P, H, x, donor, t));
        k += __popc(__ballot_sync(0xffffffffu, el));
    }
    if (k < 1) return -5;
    for (int o = 16; o > 0; o >>= 1) { const DD other{__shfl_down_sync(0xffffffffu, mine.hi, o), __shfl_down_sync(0xffffffffu, mine.lo, o)}; if (lane < o) mine = dd_add_dd(mine, other); }
    const DD total = dd_shfl(mine, 0);
    DD rnd{0.0, 0.0}; if (lane == leader) rnd = dd_mul_d(total, rng.next_double());
    rnd = dd_shfl(rnd, leader);
    DD carry{0.0, 0.0};
    for (int base = a0; base < a1; base += 32) {
        const int j = base + lane; const int x = j < a1 ? H.ep_arcs[j] : -1;
        const bool el = j < a1 && arc_eligible(H, x, donor, tag, marks, mark_gen, include);
        DD v{el ? arc_weight(P, H, x, donor, t) : 0.0, 0.0};
        for (int o = 1; o < 32; o <<= 1) { const DD other{__shfl_up_sync(0xffffffffu, v.hi, o), __shfl_up_sync(0xffffffffu, v.lo, o)}; if (lane >= o) v = dd_add_dd(other, v); }
        const DD cum = dd_add_dd(carry, v);
        const unsigned hit = __ballot_sync(0xffffffffu, el && dd_gt(cum, rnd));
        if (hit) {
            int chosen = __shfl_sync(0xffffffffu, x, __ffs(hit) - 1);
            for (int jj = base + __ffs(hit); jj < a1; ++jj) {        // arcs of weight 0 that follow share the key and overwrite it
                const int x2 = H.ep_arcs[jj];
                if (!arc_eligible(H, x2, donor, tag, marks, mark_gen, include)) continue;
                if (arc_weight(P, H, x2, donor, t) == 0.0) chosen = x2; else break;
            }
            return chosen;
        }
        carry = dd_shfl(cum, 31);
    }
    return -5;
}

// is `v` still attached to the tree under `root`? (replaced subtrees and not-yet-attached children stay in the pool)
__device__ __forceinline__ bool node_attached(const Node* N, int v, int root) {
    while (v != root) {
        const int p = N[v].parent;
        if (p < 0 || (N[p].left != v && N[p].right != v)) return false;
        v = p;
    }
    return true;
}
// does a come before b in left-before-right pre-order? (both attached, a != b)
__device__ __forceinline__ bool preorder_before(const Node* N, int a, int b) {
    int da = 0, db = 0;
    for (int v = a; N[v].parent >= 0; v = N[v].parent) ++da;
    for (int v = b; N[v].parent >= 0; v = N[v].parent) ++db;
    int x = a, y = b;
    while (da > db) { x = N[x].parent; --da; }
    while (db > da) { y = N[y].parent; --db; }
    if (x == y) return x == a;                  // one is an ancestor of the other: the ancestor comes first
    while (N[x].parent != N[y].parent) { x = N[x].parent; y = N[y].parent; }
    return N[N[x].parent].left == x;
}
// findVertex (GuestTreeInHostTreeCreator.java:332-346: first vertex in left-before-right pre-order on arc `to` whose branch
// strictly spans time t) by the whole warp: every lane tests a slice of the node pool, the pre-order-first attached match wins.
static __device__ __noinline__ int coop_find_victim(const Node* N, int n, int root, int to, double t) {
    const int lane = threadIdx.x & 31;
    int best = -1;
    for (int base = 0; base < n; base += 32) {
        const int v = base + lane;
        bool match = false;
        if (v < n) { const Node& x = N[v]; match = x.sigma == to && x.abstime < t && XADD(x.abstime, x.bl) > t; }
        unsigned m = __ballot_sync(0xffffffffu, match);
        while (m) {
            const int cand = base + __ffs(m) - 1; m &= m - 1;
            if (!node_attached(N, cand, root)) continue;
            if (best < 0 || preorder_before(N, cand, best)) best = cand;
        }
    }
    return best;
}

// State of one attempt of createUnprunedTree between steps (a state machine with ONE vertex-creation site: every step either
// fetches the next lineage from the queue (stage 0) or creates exactly one vertex (stages 1-3), so the expensive Java-exact
// vertex draw — log, divide, round8 — is executed by all creating lanes together).
// The order in which random numbers are consumed is exactly the reference's (coin -> staying child -> sampleArc(s) ->
// moving child; left child before right child).
struct SimState {
    int n, qh, qt, np;
    int stage;                     // -1: no attempt in progress; 3: create the root, 0: fetch a lineage, 1: first vertex of the pair, 2: second
    int li, c1, X, root, sigma, ep, first_to, victim;
    uint8_t ev; bool stay_left, xfer_pending;
    double t;                      // start time of the vertex to create next
};

template <class RNG>
__device__ __forceinline__ void sim_begin(SimState& S, RNG& rng, const SimParams& P, const DevHost& H) {
    S.n = 0; S.qh = 0; S.qt = 0; S.np = 0; S.stage = 3; S.li = -1; S.c1 = -1; S.ev = 0; S.stay_left = true; S.xfer_pending = false;
    if (P.ishost || !P.do_gene_birth) { S.X = H.root; S.t = H.ep_up[H.nE - 1]; }
    else { S.X = sample_root(rng, P, H); S.t = H.ep_up[H.v2e[S.X]]; }
    S.root = -1; S.sigma = 0; S.ep = 0; S.first_to = -1; S.victim = -1;
}

// Transfers waiting for their recipient (stage 2 of a transfer vertex): served lane by lane by the whole warp.
// GuestTreeInHostTreeCreator.java:171-189 (additive) and :191-279 (replacing: look for a lineage to replace, arc by arc).
template <class RNG>
__device__ void coop_resolve_transfers(SimState& S, RNG& rng, const SimParams& P, const DevHost& H, SimWork& W) {
    unsigned req = __ballot_sync(0xffffffffu, S.stage == 2 && S.xfer_pending);
    const int lane = threadIdx.x & 31;
    while (req) {
        const int leader = __ffs(req) - 1; req &= req - 1;
        const int ep = __shfl_sync(0xffffffffu, S.ep, leader), donor = __shfl_sync(0xffffffffu, S.sigma, leader), root = __shfl_sync(0xffffffffu, S.root, leader);
        const int n = __shfl_sync(0xffffffffu, S.n, leader), ev = __shfl_sync(0xffffffffu, (int)S.ev, leader);
        const double t = __shfl_sync(0xffffffffu, S.t, leader);
        const Node* N = reinterpret_cast<const Node*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(W.nodes), leader));
        int32_t* marks = reinterpret_cast<int32_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(W.marks), leader));
        int X, first_to, victim = -1;
        if (ev == EV_ADDITIVE_TRANSFER) { X = coop_sample_arc(leader, rng, P, H, ep, donor, t, nullptr, 0); first_to = X; }
        else {
            first_to = coop_sample_arc(leader, rng, P, H, ep, donor, t, nullptr, 0);
            int to = first_to;
            victim = coop_find_victim(N, n, root, to, t);
            const int n_arcs = H.ep_off[ep + 1] - H.ep_off[ep];
            int n_empty = 0;
            int gen = 0; if (lane == leader) gen = ++W.mark_gen;
            gen = __shfl_sync(0xffffffffu, gen, leader);
            while (victim < 0 && n_empty != n_arcs - 1) {
                if (marks[to] != gen) { __syncwarp(); if (lane == leader) marks[to] = gen; __syncwarp(); ++n_empty; }
                if (n_empty != n_arcs - 1) { to = coop_sample_arc(leader, rng, P, H, ep, donor, t, marks, gen); victim = coop_find_victim(N, n, root, to, t); }
            }
            X = victim >= 0 ? to : first_to;
        }
        if (lane == leader) { S.X = X; S.first_to = first_to; S.victim = victim; S.xfer_pending = false; }
    }
}

// One step of the attempt. Returns -1 while the attempt goes on, REP_OK when the tree is complete, REP_NODE_OVERFLOW when the
// pool is full. A transfer's second vertex needs its recipient arc first: the step then only raises xfer_pending and the next
// coop_resolve_transfers call fills in X / first_to / victim.
template <class RNG>
__device__ int sim_step(SimState& S, RNG& rng, const SimParams& P, const DevHost& H, const MathTables& T, SimWork& W) {
    if (S.stage == 0) {
        // fetch lineages until one needs children (ended lineages are skipped in the same step), then create its first child
        for (;;) {
            if (S.qh == S.qt) {
                if (S.np == 0) {
                    if (W.nodes[S.root].bl <= 1.0e-32) W.nodes[S.root].bl = 0.0;
                    return REP_OK;
                }
                int best = 0;      // release the oldest parked replacing transfer (largest abstime)
                for (int i = 1; i < S.np; ++i) if (W.nodes[W.pend[i]].abstime > W.nodes[W.pend[best]].abstime) best = i;
                W.queue[S.qt++] = W.pend[best];
                W.pend[best] = W.pend[--S.np];
            }
            S.li = W.queue[S.qh++];
            const Node& ln = W.nodes[S.li];
            S.ev = ln.event;
            if (S.ev == EV_LOSS || S.ev == EV_REPLACING_LOSS || S.ev == EV_LEAF || S.ev == EV_UNSAMPLED_LEAF) continue;     // lineage ends
            S.sigma = ln.sigma; S.t = ln.abstime; S.ep = ln.epoch;
            if (S.ev == EV_SPECIATION) S.X = H.left[S.sigma];
            else { S.X = S.sigma; if (S.ev != EV_DUPLICATION) S.stay_left = rng.next_double() < 0.5; }
            S.stage = 1;
            break;
        }
    }
    if (S.stage == 2 && S.xfer_pending) return -1;          // waits for coop_resolve_transfers
    // ---- the single vertex-creation site
    const VertexDraw d = draw_vertex(rng, P, H, T, S.X, S.t);
    const int idx = new_node(W, S.n, d, S.X);
    if (idx < 0) return REP_NODE_OVERFLOW;
    if (S.stage == 3) { S.root = idx; W.queue[S.qt++] = idx; S.stage = 0; return -1; }
    if (S.stage == 1) {
        S.c1 = idx; S.stage = 2;
        // arc of the second vertex
        if (S.ev == EV_SPECIATION) S.X = H.right[S.sigma];
        else if (S.ev == EV_DUPLICATION) S.X = S.sigma;
        else S.xfer_pending = true;
        return -1;
    }
    // ---- stage 2 done: attach the pair
    const int li = S.li, c1 = S.c1;
    int lc, rc;
    if (S.ev == EV_SPECIATION || S.ev == EV_DUPLICATION) { lc = c1; rc = idx; }
    else {
        Node& ln = W.nodes[li];
        ln.from_arc = S.sigma; ln.to_arc = S.X;
        if (S.ev == EV_REPLACING_TRANSFER) {
            if (S.victim >= 0) {
                Node& vn = W.nodes[S.victim];
                vn.event = EV_REPLACING_LOSS; vn.abstime = S.t;
                int w = 0;          // drop parked replacing transfers that sit in the victim's subtree (victim included)
                for (int i = 0; i < S.np; ++i) {
                    int x = W.pend[i]; bool under = false;
                    for (int y = x; y >= 0; y = W.nodes[y].parent) if (y == S.victim) { under = true; break; }
                    if (!under) W.pend[w++] = x;
                }
                S.np = w;
                vn.left = -1; vn.right = -1;
            } else ln.event = EV_ADDITIVE_TRANSFER;      // nothing to replace anywhere: degrade (GuestTreeInHostTreeCreator.java:228-233)
        }
        lc = S.stay_left ? c1 : idx; rc = S.stay_left ? idx : c1;
    }
    W.nodes[li].left = lc; W.nodes[li].right = rc;
    W.nodes[lc].parent = li; W.nodes[rc].parent = li;
    for (int s = 0; s < 2; ++s) {
        const int c = s == 0 ? lc : rc;
        if (W.nodes[c].event != EV_REPLACING_TRANSFER) W.queue[S.qt++] = c;
        else {
            bool replaced = false;     // TreeMap.put on an equal key overwrites the value
            for (int i = 0; i < S.np; ++i) if (W.nodes[W.pend[i]].abstime == W.nodes[c].abstime) { W.pend[i] = c; replaced = true; break; }
            if (!replaced) W.pend[S.np++] = c;
        }
    }
    S.stage = 0;
    return -1;
}

// GuestTreeMachina.unprunedIsOK (:124-169) on the attached tree. Returns sampled-leaf count through *sampled.
static __device__ __noinline__ bool attempt_ok(const SimParams& P, const DevHost& H, const SimWork& W, int root, int exact, int* sampled_out) {
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
