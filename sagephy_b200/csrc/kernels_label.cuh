// K3: labelling, pruning and per-tree statistics — one thread per accepted replicate, no recursion (parent links make
// every traversal stackless).
//
// Reference behaviour (apps/SaGePhy/PruningHelper.java):
//   :25-65  labelUnprunableVertices — prunability lattice; survivors numbered 0.. in post-order
//   :76-90  labelPrunableVertices   — remaining vertices numbered after them, again in post-order
//   :98-145 prune                   — pruned copy; a vertex with one surviving child collapses into that child, whose length
//                                      becomes round8(child + parent)
// plus the breadth-first reductions of getInfo (GuestTreeInHostTreeCreator.java:431-511, HostTreeGen.java:124-174).
#pragma once
#include "kernels_args.cuh"

namespace sgp {


__global__ void __launch_bounds__(128) k_label_prune(LabelArgs a) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    RepInfo& inf = a.info[r];
    if (inf.status != REP_OK) return;
    const long long off = a.node_off[r];
    Node* N = a.nodes + off;
    int32_t* ordU = a.aux + off; int32_t* ordP = a.aux + a.aux_stride + off;
    int32_t* bfsU = a.aux + 2 * a.aux_stride + off; int32_t* bfsP = a.aux + 3 * a.aux_stride + off;
    int32_t* pr = bfsP;     // temporary: pruned representative of each subtree
    const int root = inf.root;
    int nu = 0, np = 0, up = 0;
    // ---- post-order walk: prunability, numbering, pruned structure
    int v = root;
    for (;;) {
        while (N[v].left >= 0) v = N[v].left;
        bool done = false;
        for (;;) {
            Node& x = N[v];
            if (x.left < 0) {
                x.prun = (x.event == EV_LEAF) ? PR_UNPRUNABLE : PR_PRUNABLE;
                if (x.event == EV_LEAF) { pr[v] = v; x.p_bl = x.bl; x.p_left = -1; x.p_right = -1; } else pr[v] = -1;
            } else {
                const uint8_t pl = N[x.left].prun, pq = N[x.right].prun;
                if (pl == PR_PRUNABLE && pq == PR_PRUNABLE) { x.prun = PR_PRUNABLE; pr[v] = -1; }
                else if (pl == PR_PRUNABLE || pq == PR_PRUNABLE) {
                    x.prun = PR_COLLAPSABLE;
                    const int c = pr[pl == PR_PRUNABLE ? x.right : x.left];
                    N[c].p_bl = round_sig(a.T, XADD(N[c].p_bl, x.bl), 8);
                    pr[v] = c;
                } else {
                    x.prun = PR_UNPRUNABLE; x.p_left = pr[x.left]; x.p_right = pr[x.right]; x.p_bl = x.bl; pr[v] = v;
                }
            }
            if (x.prun == PR_UNPRUNABLE) { x.number = nu; ordP[nu] = v; ++nu; } else { x.number = -2 - np; ++np; }
            ordU[up++] = v;
            const int p = x.parent;
            if (p < 0) { done = true; break; }
            if (N[p].left == v) { v = N[p].right; break; }
            v = p;
        }
        if (done) break;
    }
    const int p_root = pr[root];
    for (int k = 0; k < up; ++k) { Node& x = N[ordU[k]]; if (x.number < -1) x.number = nu + (-2 - x.number); }
    // ---- top-down: left-child flags and '(' run lengths (reverse post-order visits parents before children)
    for (int k = up - 1; k >= 0; --k) {
        Node& x = N[ordU[k]];
        if (x.left >= 0) { Node& L = N[x.left]; L.flags |= NF_LEFT_U; L.chain_u = (int16_t)(x.chain_u + 1); N[x.right].chain_u = 0; }
    }
    for (int k = nu - 1; k >= 0; --k) {
        Node& x = N[ordP[k]];
        x.flags |= NF_IN_PRUNED;
        if (x.p_left >= 0) { Node& L = N[x.p_left]; L.flags |= NF_LEFT_P; L.chain_p = (int16_t)(x.chain_p + 1); N[x.p_right].chain_p = 0; }
    }
    if (a.has_stem_override) { N[root].bl = a.stem_override; if (p_root >= 0) N[p_root].p_bl = a.stem_override; }
    // ---- breadth-first orders and the getInfo reductions (sum order = BFS order)
    inf.p_root = p_root; inf.n_u = up; inf.n_p = nu; inf.p_root_time = p_root >= 0 ? N[p_root].abstime : 0.0;
    for (int e = 0; e < EV_COUNT; ++e) { inf.cnt_u[e] = 0; inf.cnt_p[e] = 0; }
    double tot = 0.0, ben = 0.0;
    {
        int h = 0, t = 0; bfsU[t++] = root;
        while (h < t) {
            const Node& x = N[bfsU[h++]];
            if (x.left >= 0) { bfsU[t++] = x.left; bfsU[t++] = x.right; }
            tot = XADD(tot, x.bl);
            ben = XADD(ben, fmax(fmin(XSUB(a.host_root_time, x.abstime), x.bl), 0.0));
            inf.cnt_u[x.event]++;
        }
    }
    inf.total_u = tot; inf.beneath_u = ben;
    tot = 0.0; ben = 0.0;
    if (p_root >= 0) {
        int h = 0, t = 0; bfsP[t++] = p_root;
        while (h < t) {
            const Node& x = N[bfsP[h++]];
            if (x.p_left >= 0) { bfsP[t++] = x.p_left; bfsP[t++] = x.p_right; }
            tot = XADD(tot, x.p_bl);
            ben = XADD(ben, fmax(fmin(XSUB(a.host_root_time, x.abstime), x.p_bl), 0.0));
            inf.cnt_p[x.event]++;
        }
    }
    inf.total_p = tot; inf.beneath_p = ben;
}

}  // namespace sgp
