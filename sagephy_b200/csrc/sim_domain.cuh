// DomainTreeGen sampler (device side): a domain tree evolving inside a forest of gene trees.
//
// Mirrors apps/SaGePhy/DomainTreeInHostTreeCreator.java: createUnprunedTree :158-937 (sixteen spelled-out transfer branches =
// {additive, replacing} x {intra-, inter-gene} x {intra-, inter-species} x coin; folded into one parameterised path here),
// findVertex :948-964, swap :984-1002, sampleInterGeneArc :1004-1033, createGuestVertex :1042-1120, and the `-all` /
// per-leaf acceptance of GuestTreeMachina.unprunedIsOK (apps/SaGePhy/GuestTreeMachina.java:124-169).
// Same single-creation-site state machine as simulate_attempt (sim.cuh). Draw order per transfer (differs from
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
__device__ __forceinline__ VertexDraw draw_vertex_domain(RNG& rng, const SimParams& P, const DomainEnv& E, const DevHost& G, const MathTables& T, int X, double start) {
    VertexDraw d;
    const bool isRoot = G.parent[X] < 0;
    double sum = isRoot ? XADD(P.lambda, P.mu) : XADD(XADD(P.lambda, P.mu), P.tau);
    if (sum == 0.0) sum = 1e-48;
    const double low = G.vtime[X];
    double bt = XDIV(-jlog(XSUB(1.0, rng.next_double())), sum);
    double et = XSUB(start, bt);
    if (et <= low) {
        et = low;
        bt = round_sig(T, XSUB(start, et), 8);
        if (G.left[X] < 0) d.event = (rng.next_double() < P.rho) ? EV_LEAF : EV_UNSAMPLED_LEAF; else d.event = EV_CODIVERGENCE;
        d.epoch = G.v2e[X];
    } else {
        const double rnd = rng.next_double();
        if (isRoot) d.event = (rnd < XDIV(P.lambda, sum)) ? EV_DUPLICATION : EV_LOSS;
        else if (rnd >= XDIV(XADD(P.mu, P.tau), sum)) d.event = EV_DUPLICATION;
        else if (rnd < XDIV(P.mu, sum)) d.event = EV_LOSS;
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

// sampleInterGeneArc: uniform over the arcs (vertex ids) of the OTHER gene trees that are "alive" at time t in the reference's
// sense (vertex time < t < upper time of the epoch directly above the vertex) with the requested species relation; ids are
// compared / marked across trees exactly as the reference does. Returns -5 when there is no candidate.
template <class RNG>
__device__ int sample_inter_gene(RNG& rng, const DomainEnv& E, int g_from, int sigma, double t, bool intra_species, const int32_t* marks, int gen, int* g_to) {
    const int tag = E.genes[g_from].host_tag[sigma];
    int k = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int pick = 0;
        if (pass == 1) { if (k < 1) { *g_to = -1; return -5; } pick = rng.next_int(k); }
        for (int g = 0; g < E.n_genes; ++g) {
            if (g == g_from) continue;
            const DevHost& G = E.genes[g];
            // the epoch of gene g that contains t: first epoch whose upper time exceeds t (epochs ascend and are contiguous)
            int lo = 0, hi = G.nE;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (G.ep_up[mid] > t) hi = mid; else lo = mid + 1; }
            if (lo >= G.nE || !(G.ep_lo[lo] < t)) continue;
            const int j0 = E.ev_off[E.ev_base[g] + lo], j1 = E.ev_off[E.ev_base[g] + lo + 1];
            for (int j = j0; j < j1; ++j) {
                const int i = E.ev_list[j];
                if (i == sigma || (marks && marks[i] == gen)) continue;
                if ((G.host_tag[i] == tag) != intra_species) continue;
                if (pass == 0) ++k;
                else { if (pick == 0) { *g_to = g; return i; } --pick; }
            }
        }
    }
    *g_to = -1; return -5;
}

static __device__ int find_victim_domain(const Node* N, int root, int to, int g_to, double t) {
    int v = root;
    while (v >= 0) {
        const Node& x = N[v];
        if (x.sigma == to && x.gtree == g_to && x.abstime < t && XADD(x.abstime, x.bl) > t) return v;
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

template <class RNG>
__device__ int simulate_attempt_domain(RNG& rng, const SimParams& P, const DomainEnv& E, const MathTables& T, SimWork& W, int& n_out, int& root_out) {
    int n = 0, qh = 0, qt = 0, np = 0;
    for (int g = 0; g < E.n_genes; ++g) E.used[g] = 0;
    int stage = 3;          // 3 root, 0 fetch, 1 first of pair, 2 second of pair, 4 swap replacement
    int li = -1, c1 = -1; uint8_t ev = 0; bool stay_left = true, transfer = false, replacing = false, inter = false, include = true;
    int g = rng.next_int(E.n_genes);          // first_guest: first draw of every attempt
    int X; double t;
    if (!P.do_gene_birth) { X = E.genes[g].root; t = E.genes[g].ep_up[E.genes[g].nE - 1]; }
    else { X = sample_root(rng, P, E.genes[g]); t = E.genes[g].ep_up[E.genes[g].v2e[X]]; }
    int root = -1, sigma = 0, ep = 0, lg = 0, to = -5, g_to = -1, first_to = -5, orig_g = -1, victim = -1;
    for (;;) {
        if (stage == 0) {
            if (qh == qt) {
                if (np == 0) break;
                int best = 0;
                for (int i = 1; i < np; ++i) if (W.nodes[W.pend[i]].abstime > W.nodes[W.pend[best]].abstime) best = i;
                W.queue[qt++] = W.pend[best];
                W.pend[best] = W.pend[--np];
            }
            li = W.queue[qh++];
            const Node& ln = W.nodes[li];
            ev = ln.event;
            if (ev == EV_LOSS || ev == EV_REPLACING_LOSS || ev == EV_LEAF || ev == EV_UNSAMPLED_LEAF) continue;
            sigma = ln.sigma; t = ln.abstime; ep = ln.epoch; lg = ln.gtree; g = lg;
            transfer = ev >= EV_AT_GG_SS;
            if (ev == EV_CODIVERGENCE) { X = E.genes[lg].left[sigma]; stage = 1; continue; }
            X = sigma; stage = 1;
            if (transfer) {
                const int code = (ev - EV_AT_GG_SS) & 3;
                replacing = ev >= EV_RT_GG_SS; inter = (code & 2) != 0; include = (code & 1) == 0;
                stay_left = rng.next_double() < 0.5;
                if (!inter) { to = sample_arc(rng, P, E.genes[lg], ep, sigma, t, nullptr, 0, include); g_to = lg; }
                else to = sample_inter_gene(rng, E, lg, sigma, t, include, nullptr, 0, &g_to);
                first_to = to; orig_g = g_to;
                if (to == -5) stage = 4;       // nothing to transfer to: the transfer vertex is swapped for a fresh vertex
            }
            continue;
        }
        if (stage == 2) {
            if (ev == EV_CODIVERGENCE) { X = E.genes[lg].right[sigma]; g = lg; }
            else if (ev == EV_DUPLICATION) { X = sigma; g = lg; }
            else if (!replacing) { X = to; g = g_to; victim = -1; }
            else {
                const DevHost& G = E.genes[lg];
                victim = find_victim_domain(W.nodes, root, to, g_to, t);
                const int n_arcs = G.ep_off[ep + 1] - G.ep_off[ep];
                int n_empty = 0; const int gen = ++W.mark_gen;
                while (victim < 0 && n_empty != n_arcs - 1 && to != -5) {
                    if (W.marks[to] != gen) { W.marks[to] = gen; ++n_empty; }
                    if (n_empty != n_arcs - 1) {
                        if (!inter) to = sample_arc(rng, P, G, ep, sigma, t, W.marks, gen, include);
                        else to = sample_inter_gene(rng, E, lg, sigma, t, include, W.marks, gen, &g_to);
                        victim = to == -5 ? -1 : find_victim_domain(W.nodes, root, to, g_to, t);
                    }
                }
                if (victim >= 0) { X = to; g = g_to; } else { X = first_to; g = orig_g; }
            }
        }
        // ---- the single vertex-creation site
        const VertexDraw d = draw_vertex_domain(rng, P, E, E.genes[g], T, X, t);
        const int idx = new_node(W, n, d, X);
        if (idx < 0) { n_out = n; return REP_NODE_OVERFLOW; }
        W.nodes[idx].gtree = (int16_t)g;
        E.used[g]++;
        if (stage == 3) { root = idx; W.queue[qt++] = root; stage = 0; continue; }
        if (stage == 1) { c1 = idx; stage = 2; continue; }
        if (stage == 4) {
            // swap (:984-1002): idx takes the place of the transfer vertex li
            const int par = W.nodes[li].parent;
            if (li != root) { if (W.nodes[par].left == li) W.nodes[par].left = idx; else W.nodes[par].right = idx; W.nodes[idx].parent = par; }
            else root = idx;
            W.nodes[li].parent = -1;
            if (W.nodes[idx].event < EV_RT_GG_SS) W.queue[qt++] = idx;
            else {
                bool rep = false;
                for (int i = 0; i < np; ++i) if (W.nodes[W.pend[i]].abstime == W.nodes[idx].abstime) { W.pend[i] = idx; rep = true; break; }
                if (!rep) W.pend[np++] = idx;
            }
            stage = 0; continue;
        }
        // ---- stage 2 done: attach the pair
        int lc, rc;
        if (!transfer) { lc = c1; rc = idx; }
        else {
            Node& ln = W.nodes[li];
            ln.from_arc = sigma; ln.to_arc = X; ln.from_g = (int16_t)lg; ln.to_g = (int16_t)g;
            if (replacing) {
                if (victim >= 0) {
                    Node& vn = W.nodes[victim];
                    vn.event = EV_REPLACING_LOSS; vn.abstime = t;
                    int w = 0;
                    for (int i = 0; i < np; ++i) {
                        int x = W.pend[i]; bool under = false;
                        for (int y = x; y >= 0; y = W.nodes[y].parent) if (y == victim) { under = true; break; }
                        if (!under) W.pend[w++] = x;
                    }
                    np = w;
                    vn.left = -1; vn.right = -1;
                } else ln.event = (uint8_t)(ev - 4);        // degrade to the additive kind with the same gene / species scope
            }
            lc = stay_left ? c1 : idx; rc = stay_left ? idx : c1;
        }
        W.nodes[li].left = lc; W.nodes[li].right = rc;
        W.nodes[lc].parent = li; W.nodes[rc].parent = li;
        for (int s = 0; s < 2; ++s) {
            const int c = s == 0 ? lc : rc;
            if (W.nodes[c].event < EV_RT_GG_SS) W.queue[qt++] = c;
            else {
                bool rep = false;
                for (int i = 0; i < np; ++i) if (W.nodes[W.pend[i]].abstime == W.nodes[c].abstime) { W.pend[i] = c; rep = true; break; }
                if (!rep) W.pend[np++] = c;
            }
        }
        stage = 0;
    }
    if (W.nodes[root].bl <= 1.0e-32) W.nodes[root].bl = 0.0;
    n_out = n; root_out = root;
    return REP_OK;
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
