// DomainTreeGen sampler (device side): a domain tree evolving inside a forest of gene trees.
//
// Mirrors apps/SaGePhy/DomainTreeInHostTreeCreator.java: createUnprunedTree :158-937 (sixteen spelled-out transfer branches =
// {additive, replacing} x {intra-, inter-gene} x {intra-, inter-species} x coin; folded into one parameterised path here),
// findVertex :948-964, swap :984-1002, sampleInterGeneArc :1004-1033, createGuestVertex :1042-1120, and the `-all` /
// per-leaf acceptance of GuestTreeMachina.unprunedIsOK (apps/SaGePhy/GuestTreeMachina.java:124-169).
// Same single-creation-site state machine and warp-cooperative scheme as sim_step (sim.cuh). Draw order per transfer (differs from
// GuestTreeGen): coin -> first recipient -> staying child -> further recipients (replacing only) -> moving child.
#pragma once
#include "sim.cuh"

namespace sgp {

struct DomainEnv {
    const DevHost* genes; int n_genes; double inter_gene, inter_species;
    int32_t* used;           // vertices created per gene tree in the current attempt
    // inter-gene candidate index: the vertices of gene g whose first epoch is e, ids ascending (engine.cu builds it)
    const int32_t* ev_base; const int32_t* ev_off; const int32_t* ev_list;
};

template <class RNG>
__device__ __forceinline__ VertexDraw draw_vertex_domain(RNG& rng, const SimParams& P, const DomainEnv& E, const DevHost& G, const MathTables& T, int X, double start, const double2* tab) {
    VertexDraw d;
    const bool isRoot = G.parent[X] < 0;
    const int ks = isRoot ? 0 : 1;
    const double sum = P.sum[ks];
    const double low = G.vtime[X];
    const double u1 = rng.next_double();
    double bt;          // see draw_vertex (sim.cuh)
    if constexpr (std::is_same<RNG, PhiloxStream>::value) bt = __dmul_rn(-log_unit(XSUB(1.0, u1), tab), P.inv_sum[ks]);
    else bt = XDIV(-jlog(XSUB(1.0, u1)), sum);
    double et = XSUB(start, bt);
    if (et <= low) {
        et = low;
        bt = round_sig(T, XSUB(start, et), 8);
        if (G.left[X] < 0) d.event = (rng.next_double() < P.rho) ? EV_LEAF : EV_UNSAMPLED_LEAF; else d.event = EV_CODIVERGENCE;
        d.epoch = G.v2e[X];
    } else {
        const double rnd = rng.next_double();
        if (isRoot) d.event = (rnd < P.thr_dup[0]) ? EV_DUPLICATION : EV_LOSS;
        else if (rnd >= P.thr_dup[1]) d.event = EV_DUPLICATION;
        else if (rnd < P.thr_loss[1]) d.event = EV_LOSS;
        else {
            const double trn = rng.next_double(), trn_gen = rng.next_double(), trn_spe = rng.next_double();
            d.event = (uint8_t)((trn < P.theta ? EV_RT_GG_SS : EV_AT_GG_SS) + (trn_gen < E.inter_gene ? 2 : 0) + (trn_spe < E.inter_species ? 1 : 0));
        }
        int e = G.v2e[X];
        while (G.ep_up[e] < et) ++e;
        d.epoch = e;
    }
    d.abstime = et; d.bl = bt;
    return d;
}

// ---- warp-cooperative engine (same scheme as sim.cuh: one step per lane per trip, linear-time pieces served by the warp) ----------
// sampleInterGeneArc by the whole warp: lane l looks after genes l, l + 32, ... (epoch lookup + its short candidate list), a warp
// scan of the per-lane counts turns the leader's draw into (lane, rank within the lane's candidates).
template <class RNG>
__device__ __noinline__ int coop_sample_inter_gene(int leader, RNG& rng, const DomainEnv& E, int g_from, int sigma, double t, bool intra_species, const int32_t* marks, int gen, int* g_to) {
    const int lane = threadIdx.x & 31;
    const int tag = E.genes[g_from].host_tag[sigma];
    int pick = 0, my_first = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int k = 0;          // pass 0: candidates in this lane's genes; pass 1: running rank
        for (int g0 = 0; g0 < E.n_genes; g0 += 32) {
            // gene-major order over ALL lanes: within a block of 32 genes, lane order == gene order; blocks are handled one after
            // the other, so ranks are (block, lane, position) lexicographic == the reference's (gene, vertex id) order
            const int g = g0 + lane;
            int cnt = 0, j0 = 0, j1 = 0; const DevHost* G = nullptr;
            if (g < E.n_genes && g != g_from) {
                G = &E.genes[g];
                int lo = 0, hi = G->nE;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (G->ep_up[mid] > t) hi = mid; else lo = mid + 1; }
                if (lo < G->nE && G->ep_lo[lo] < t) { j0 = E.ev_off[E.ev_base[g] + lo]; j1 = E.ev_off[E.ev_base[g] + lo + 1]; }
                for (int j = j0; j < j1; ++j) {
                    const int i = E.ev_list[j];
                    if (i == sigma || (marks && marks[i] == gen)) continue;
                    if ((G->host_tag[i] == tag) != intra_species) continue;
                    ++cnt;
                }
            }
            int incl = cnt;
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const int block_total = __shfl_sync(0xffffffffu, incl, 31);
            if (pass == 1) {
                const int before = k + incl - cnt;          // candidates ranked before this lane's first one
                const bool mine = cnt > 0 && pick >= before && pick < before + cnt;
                const unsigned who = __ballot_sync(0xffffffffu, mine);
                if (who) {
                    const int owner = __ffs(who) - 1;
                    int res = -5;
                    if (lane == owner) {
                        int left = pick - before;
                        for (int j = j0; j < j1; ++j) {
                            const int i = E.ev_list[j];
                            if (i == sigma || (marks && marks[i] == gen)) continue;
                            if ((G->host_tag[i] == tag) != intra_species) continue;
                            if (left == 0) { res = i; break; }
                            --left;
                        }
                    }
                    *g_to = g0 + owner;
                    return __shfl_sync(0xffffffffu, res, owner);
                }
            }
            k += block_total;
        }
        if (pass == 0) {
            if (k < 1) { *g_to = -1; return -5; }
            if (lane == leader) pick = rng.next_int(k);
            pick = __shfl_sync(0xffffffffu, pick, leader);
        }
    }
    (void)my_first;
    *g_to = -1; return -5;
}

// findVertex (:948-964) by the whole warp, see coop_find_victim
static __device__ __noinline__ int coop_find_victim_domain(const Node* N, int n, int root, int to, int g_to, double t) {
    const int lane = threadIdx.x & 31;
    int best = -1;
    for (int base = 0; base < n; base += 32) {
        const int v = base + lane;
        bool match = false;
        if (v < n) { const Node& x = N[v]; match = x.sigma == to && x.gtree == g_to && x.abstime < t && XADD(x.abstime, x.bl) > t; }
        unsigned m = __ballot_sync(0xffffffffu, match);
        while (m) {
            const int cand = base + __ffs(m) - 1; m &= m - 1;
            if (!node_attached(N, cand, root)) continue;
            if (best < 0 || preorder_before(N, cand, best)) best = cand;
        }
    }
    return best;
}

struct DomainState {
    int n, qh, qt, np;
    int stage;                     // -1 none; 3 root, 0 fetch, 1 first of pair, 2 second of pair, 4 swap replacement
    int li, c1, g, X, root, sigma, ep, lg, to, g_to, first_to, orig_g, victim;
    uint8_t ev; bool stay_left, transfer, replacing, inter, include;
    bool need_first, need_victim;  // recipient draws waiting for the warp
    double t;
};

template <class RNG>
__device__ __forceinline__ void domain_begin(DomainState& S, RNG& rng, const SimParams& P, const DomainEnv& E) {
    S.n = 0; S.qh = 0; S.qt = 0; S.np = 0;
    for (int g = 0; g < E.n_genes; ++g) E.used[g] = 0;
    S.stage = 3; S.li = -1; S.c1 = -1; S.ev = 0; S.stay_left = true; S.transfer = false; S.replacing = false; S.inter = false; S.include = true;
    S.need_first = false; S.need_victim = false;
    S.g = rng.next_int(E.n_genes);          // first_guest: first draw of every attempt
    if (!P.do_gene_birth) { S.X = E.genes[S.g].root; S.t = E.genes[S.g].ep_up[E.genes[S.g].nE - 1]; }
    else { S.X = sample_root(rng, P, E.genes[S.g]); S.t = E.genes[S.g].ep_up[E.genes[S.g].v2e[S.X]]; }
    S.root = -1; S.sigma = 0; S.ep = 0; S.lg = 0; S.to = -5; S.g_to = -1; S.first_to = -5; S.orig_g = -1; S.victim = -1;
}

template <class RNG>
__device__ void coop_resolve_domain(DomainState& S, RNG& rng, const SimParams& P, const DomainEnv& E, SimWork& W) {
    const int lane = threadIdx.x & 31;
    // (A) first recipient of a transfer, drawn right after the coin (createUnprunedTree :231-242)
    unsigned req = __ballot_sync(0xffffffffu, S.stage >= 0 && S.need_first);
    while (req) {
        const int leader = __ffs(req) - 1; req &= req - 1;
        const int lg = __shfl_sync(0xffffffffu, S.lg, leader), sigma = __shfl_sync(0xffffffffu, S.sigma, leader), ep = __shfl_sync(0xffffffffu, S.ep, leader);
        const bool inter = __shfl_sync(0xffffffffu, (int)S.inter, leader) != 0, include = __shfl_sync(0xffffffffu, (int)S.include, leader) != 0;
        const double t = __shfl_sync(0xffffffffu, S.t, leader);
        int to, g_to = lg;
        if (!inter) to = coop_sample_arc(leader, rng, P, E.genes[lg], ep, sigma, t, nullptr, 0, include);
        else to = coop_sample_inter_gene(leader, rng, E, lg, sigma, t, include, nullptr, 0, &g_to);
        if (lane == leader) {
            S.to = to; S.g_to = g_to; S.first_to = to; S.orig_g = g_to; S.need_first = false;
            if (to == -5) S.stage = 4;       // nothing to transfer to: the transfer vertex is swapped for a fresh vertex
        }
    }
    // (B) replacing transfer: look for a lineage to replace, recipient by recipient
    req = __ballot_sync(0xffffffffu, S.stage == 2 && S.need_victim);
    while (req) {
        const int leader = __ffs(req) - 1; req &= req - 1;
        const int lg = __shfl_sync(0xffffffffu, S.lg, leader), sigma = __shfl_sync(0xffffffffu, S.sigma, leader), ep = __shfl_sync(0xffffffffu, S.ep, leader);
        const int root = __shfl_sync(0xffffffffu, S.root, leader), n = __shfl_sync(0xffffffffu, S.n, leader);
        int to = __shfl_sync(0xffffffffu, S.to, leader), g_to = __shfl_sync(0xffffffffu, S.g_to, leader);
        const bool inter = __shfl_sync(0xffffffffu, (int)S.inter, leader) != 0, include = __shfl_sync(0xffffffffu, (int)S.include, leader) != 0;
        const double t = __shfl_sync(0xffffffffu, S.t, leader);
        const Node* N = reinterpret_cast<const Node*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(W.nodes), leader));
        int32_t* marks = reinterpret_cast<int32_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(W.marks), leader));
        const DevHost& G = E.genes[lg];
        int victim = coop_find_victim_domain(N, n, root, to, g_to, t);
        const int n_arcs = G.ep_off[ep + 1] - G.ep_off[ep];
        int n_empty = 0;
        int gen = 0; if (lane == leader) gen = ++W.mark_gen;
        gen = __shfl_sync(0xffffffffu, gen, leader);
        while (victim < 0 && n_empty != n_arcs - 1 && to != -5) {
            if (marks[to] != gen) { __syncwarp(); if (lane == leader) marks[to] = gen; __syncwarp(); ++n_empty; }
            if (n_empty != n_arcs - 1) {
                if (!inter) to = coop_sample_arc(leader, rng, P, G, ep, sigma, t, marks, gen, include);
                else to = coop_sample_inter_gene(leader, rng, E, lg, sigma, t, include, marks, gen, &g_to);
                victim = to == -5 ? -1 : coop_find_victim_domain(N, n, root, to, g_to, t);
            }
        }
        if (lane == leader) {
            S.victim = victim; S.need_victim = false;
            if (victim >= 0) { S.X = to; S.g = g_to; } else { S.X = S.first_to; S.g = S.orig_g; }
        }
    }
}

template <class RNG>
__device__ int domain_step(DomainState& S, RNG& rng, const SimParams& P, const DomainEnv& E, const MathTables& T, SimWork& W) {
    if (S.stage == 0) {
        for (;;) {
            if (S.qh == S.qt) {
                if (S.np == 0) {
                    if (W.nodes[S.root].bl <= 1.0e-32) W.nodes[S.root].bl = 0.0;
                    return REP_OK;
                }
                int best = 0;
                for (int i = 1; i < S.np; ++i) if (W.nodes[W.pend[i]].abstime > W.nodes[W.pend[best]].abstime) best = i;
                W.queue[S.qt++] = W.pend[best];
                W.pend[best] = W.pend[--S.np];
            }
            S.li = W.queue[S.qh++];
            const Node& ln = W.nodes[S.li];
            S.ev = ln.event;
            if (S.ev == EV_LOSS || S.ev == EV_REPLACING_LOSS || S.ev == EV_LEAF || S.ev == EV_UNSAMPLED_LEAF) continue;
            S.sigma = ln.sigma; S.t = ln.abstime; S.ep = ln.epoch; S.lg = ln.gtree; S.g = S.lg;
            S.transfer = S.ev >= EV_AT_GG_SS;
            S.stage = 1;
            if (S.ev == EV_CODIVERGENCE) { S.X = E.genes[S.lg].left[S.sigma]; break; }
            S.X = S.sigma;
            if (S.transfer) {
                const int code = (S.ev - EV_AT_GG_SS) & 3;
                S.replacing = S.ev >= EV_RT_GG_SS; S.inter = (code & 2) != 0; S.include = (code & 1) == 0;
                S.stay_left = rng.next_double() < 0.5;
                S.need_first = true;             // the first recipient is drawn by the warp before the staying child is created
                return -1;
            }
            break;
        }
    }
    if (S.need_first || (S.stage == 2 && S.need_victim)) return -1;
    // ---- the single vertex-creation site
    const VertexDraw d = draw_vertex_domain(rng, P, E, E.genes[S.g], T, S.X, S.t, W.logtab);
    const int idx = new_node(W, S.n, d, S.X);
    if (idx < 0) return REP_NODE_OVERFLOW;
    W.nodes[idx].gtree = (int16_t)S.g;
    E.used[S.g]++;
    if (S.stage == 3) { S.root = idx; W.queue[S.qt++] = idx; S.stage = 0; return -1; }
    if (S.stage == 1) {
        S.c1 = idx; S.stage = 2;
        if (S.ev == EV_CODIVERGENCE) { S.X = E.genes[S.lg].right[S.sigma]; S.g = S.lg; }
        else if (S.ev == EV_DUPLICATION) { S.X = S.sigma; S.g = S.lg; }
        else if (!S.replacing) { S.X = S.to; S.g = S.g_to; S.victim = -1; }
        else S.need_victim = true;
        return -1;
    }
    const int li = S.li;
    if (S.stage == 4) {
        // swap (:984-1002): idx takes the place of the transfer vertex li
        const int par = W.nodes[li].parent;
        if (li != S.root) { if (W.nodes[par].left == li) W.nodes[par].left = idx; else W.nodes[par].right = idx; W.nodes[idx].parent = par; }
        else S.root = idx;
        W.nodes[li].parent = -1;
        if (W.nodes[idx].event < EV_RT_GG_SS) W.queue[S.qt++] = idx;
        else {
            bool rep = false;
            for (int i = 0; i < S.np; ++i) if (W.nodes[W.pend[i]].abstime == W.nodes[idx].abstime) { W.pend[i] = idx; rep = true; break; }
            if (!rep) W.pend[S.np++] = idx;
        }
        S.stage = 0; return -1;
    }
    // ---- stage 2 done: attach the pair
    const int c1 = S.c1;
    int lc, rc;
    if (!S.transfer) { lc = c1; rc = idx; }
    else {
        Node& ln = W.nodes[li];
        ln.from_arc = S.sigma; ln.to_arc = S.X; ln.from_g = (int16_t)S.lg; ln.to_g = (int16_t)S.g;
        if (S.replacing) {
            if (S.victim >= 0) {
                Node& vn = W.nodes[S.victim];
                vn.event = EV_REPLACING_LOSS; vn.abstime = S.t;
                int w = 0;
                for (int i = 0; i < S.np; ++i) {
                    int x = W.pend[i]; bool under = false;
                    for (int y = x; y >= 0; y = W.nodes[y].parent) if (y == S.victim) { under = true; break; }
                    if (!under) W.pend[w++] = x;
                }
                S.np = w;
                vn.left = -1; vn.right = -1;
            } else ln.event = (uint8_t)(S.ev - 4);        // degrade to the additive kind with the same gene / species scope
        }
        lc = S.stay_left ? c1 : idx; rc = S.stay_left ? idx : c1;
    }
    W.nodes[li].left = lc; W.nodes[li].right = rc;
    W.nodes[lc].parent = li; W.nodes[rc].parent = li;
    for (int s = 0; s < 2; ++s) {
        const int c = s == 0 ? lc : rc;
        if (W.nodes[c].event < EV_RT_GG_SS) W.queue[S.qt++] = c;
        else {
            bool rep = false;
            for (int i = 0; i < S.np; ++i) if (W.nodes[W.pend[i]].abstime == W.nodes[c].abstime) { W.pend[i] = c; rep = true; break; }
            if (!rep) W.pend[S.np++] = c;
        }
    }
    S.stage = 0;
    return -1;
}

// unprunedIsOK for DomainTreeGen: the per-host-leaf window compares SPECIES leaf ids with gene-tree vertex ids (as the
// reference does, GuestTreeMachina.java:133-157 with DomainTreeInHostTreeCreator.getHostLeaves :1123-1125); `-all`.
static __device__ bool attempt_ok_domain(const SimParams& P, const DevHost& S, const DomainEnv& E, const SimWork& W, int root, int exact, int cnt_size,
                                         bool per_leaf, bool all_genes, int* sampled_out) {
    int sampled = 0;
    if (per_leaf) for (int i = 0; i < cnt_size; ++i) W.leaf_cnt[i] = 0;
    int v = root;
    while (v >= 0) {
        const Node& x = W.nodes[v];
        if (x.event == EV_LEAF) { ++sampled; if (per_leaf && x.sigma < cnt_size) W.leaf_cnt[x.sigma]++; }
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
    if (per_leaf) for (int l = 0; l < S.nV; ++l) { if (S.left[l] >= 0) continue; const int c = l < cnt_size ? W.leaf_cnt[l] : 0; if (c < P.minper || c > P.maxper) return false; }
    if (all_genes) for (int g = 0; g < E.n_genes; ++g) if (E.used[g] == 0) return false;
    return true;
}

}  // namespace sgp
