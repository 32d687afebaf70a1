// BD fast path kernels (transfer-free processes on a 1- or 2-leaf host, Philox mode) — see sim.cuh.
#pragma once
#include "sim.cuh"

namespace sgp {

// ---- BD fast path ------------------------------------------------------------------------------------------------------------

// Count-only pass: every loop iteration expands one vertex of the current attempt of the current replicate of this thread,
// so the 32 lanes of a warp stay converged while they sit in different attempts / replicates.
__global__ void __launch_bounds__(256) k_bd_find(BdArgs a) {
    double st_time[kBdStack]; int8_t st_arc[kBdStack];
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    long long r = -1; uint32_t rep_lo = 0, rep_hi = 0;
    int sp = 0; uint32_t vidx = 0; int attempts = 0; int sampled = 0, c0 = 0, c1 = 0; bool active = false, rejected = false; int exact = -1;
    uint64_t verts = 0; int status = REP_PENDING;
    const int leaf0 = a.H.nV == 1 ? 0 : a.H.left[a.H.root];
    for (;;) {
        if (sp == 0) {
            bool need_new_rep = !active;
            if (active) {
                // attempt finished
                verts += vidx;
                ++attempts;
                bool ok = !rejected && sampled >= a.P.min_leaves && sampled <= a.P.max_leaves && (exact == -1 || sampled == exact);
                if (ok && a.H.nV > 1) ok = c0 >= a.P.minper && c0 <= a.P.maxper && c1 >= a.P.minper && c1 <= a.P.maxper;
                if (ok) { status = REP_OK; }
                else if (attempts > a.P.max_attempts) { status = REP_MAX_ATTEMPTS; }
                if (status != REP_PENDING) {
                    RepInfo inf; memset(&inf, 0, sizeof(inf));
                    inf.status = status; inf.exact = exact; inf.root = -1; inf.p_root = -1; inf.vertices_sampled = verts;
                    if (status == REP_OK) { inf.acc_attempt = (uint32_t)(attempts - 1); inf.n_pool = (int)vidx; inf.sampled_leaves = sampled; inf.attempts = attempts + 1; }
                    else inf.attempts = attempts;
                    a.info[r] = inf;
                    need_new_rep = true;
                }
            }
            if (need_new_rep) {
                long long slot = (long long)atomicAdd(a.work_counter, 1ULL);
                if (slot >= a.n) break;
                r = slot; const unsigned long long gid = (unsigned long long)(a.rep_base + r);
                rep_lo = (uint32_t)gid; rep_hi = (uint32_t)(gid >> 32);
                attempts = 0; verts = 0; status = REP_PENDING; active = true; exact = -1;
                if (a.P.n_sizes > 0) { PhiloxStream ph; ph.init(a.seed, gid, 0xFFFFFFFFu); exact = a.sizes[ph.next_int(a.P.n_sizes)]; }
            }
            // start attempt `attempts`
            st_time[0] = a.H.tip; st_arc[0] = (int8_t)a.H.root; sp = 1; vidx = 0; sampled = 0; c0 = 0; c1 = 0; rejected = false;
        }
        --sp;
        const double start = st_time[sp]; const int X = st_arc[sp];
        U4 w = philox4x32_10(vidx, (uint32_t)attempts, rep_lo, rep_hi, k0, k1);
        ++vidx;
        BdVertex v = bd_vertex(a.P, a.H, X, start, u01_from_words(w.x, w.y), u01_from_words(w.z, w.w), false, a.T);
        if (v.event == EV_LEAF) {
            ++sampled; if (X == leaf0) ++c0; else ++c1;
            if (sampled > a.P.max_leaves) { rejected = true; sp = 0; }       // early abort: the attempt can no longer be accepted
        } else if (v.event == EV_DUPLICATION) {
            if (sp + 2 > kBdStack) { rejected = true; sp = 0; status = REP_STACK_OVERFLOW; }
            else { st_time[sp] = v.et; st_arc[sp] = (int8_t)X; st_time[sp + 1] = v.et; st_arc[sp + 1] = (int8_t)X; sp += 2; }
        } else if (v.event == EV_SPECIATION) {
            if (sp + 2 > kBdStack) { rejected = true; sp = 0; status = REP_STACK_OVERFLOW; }
            else { st_time[sp] = v.et; st_arc[sp] = (int8_t)a.H.right[X]; st_time[sp + 1] = v.et; st_arc[sp + 1] = (int8_t)a.H.left[X]; sp += 2; }
        }
        if (status == REP_STACK_OVERFLOW && sp == 0) {
            // record and move on (the host re-runs such replicates through the general engine)
            RepInfo inf; memset(&inf, 0, sizeof(inf)); inf.status = REP_STACK_OVERFLOW; inf.exact = exact; inf.root = -1; inf.p_root = -1; inf.attempts = attempts;
            a.info[r] = inf; active = false;
        }
    }
}

// Build pass: regenerates the accepted attempt depth-first straight into the arena (no scratch).
__global__ void __launch_bounds__(256) k_bd_build(BdArgs a) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    RepInfo& inf = a.info[r];
    if (inf.status != REP_OK) return;
    Node* N = a.nodes + a.node_off[r];
    const unsigned long long gid = (unsigned long long)(a.rep_base + r);
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32), rep_lo = (uint32_t)gid, rep_hi = (uint32_t)(gid >> 32);
    double st_time[kBdStack]; int8_t st_arc[kBdStack]; int32_t st_par[kBdStack];   // parent index; bit 31 set = right child
    int sp = 0; uint32_t vidx = 0;
    st_time[0] = a.H.tip; st_arc[0] = (int8_t)a.H.root; st_par[0] = -1; sp = 1;
    while (sp > 0) {
        --sp;
        const double start = st_time[sp]; const int X = st_arc[sp]; const int32_t pw = st_par[sp];
        U4 w = philox4x32_10(vidx, inf.acc_attempt, rep_lo, rep_hi, k0, k1);
        BdVertex v = bd_vertex(a.P, a.H, X, start, u01_from_words(w.x, w.y), u01_from_words(w.z, w.w), true, a.T);
        const int me = (int)vidx++;
        Node& nd = N[me];
        nd.abstime = v.et; nd.bl = v.bl; nd.sigma = X; nd.left = -1; nd.right = -1; nd.from_arc = -1; nd.to_arc = -1; nd.number = -1;
        nd.p_bl = 0.0; nd.p_left = -1; nd.p_right = -1; nd.event = v.event; nd.prun = PR_UNKNOWN; nd.flags = 0; nd.pad0 = 0; nd.gtree = 0;
        nd.from_g = -1; nd.to_g = -1; nd.chain_u = 0; nd.chain_p = 0; nd.pad1 = 0;
        int ep = a.H.v2e[X];
        if (v.event == EV_DUPLICATION || v.event == EV_LOSS) { while (a.H.ep_up[ep] < v.et) ++ep; }
        nd.epoch = ep;
        if (pw == -1) nd.parent = -1;
        else { int p = pw & 0x7fffffff; nd.parent = p; if (pw < 0) N[p].right = me; else N[p].left = me; }
        if (v.event == EV_DUPLICATION) {
            st_time[sp] = v.et; st_arc[sp] = (int8_t)X; st_par[sp] = (int32_t)(0x80000000u | (uint32_t)me);
            st_time[sp + 1] = v.et; st_arc[sp + 1] = (int8_t)X; st_par[sp + 1] = me; sp += 2;
        } else if (v.event == EV_SPECIATION) {
            st_time[sp] = v.et; st_arc[sp] = (int8_t)a.H.right[X]; st_par[sp] = (int32_t)(0x80000000u | (uint32_t)me);
            st_time[sp + 1] = v.et; st_arc[sp + 1] = (int8_t)a.H.left[X]; st_par[sp + 1] = me; sp += 2;
        }
    }
    if (N[0].bl <= 1.0e-32) N[0].bl = 0.0;
    inf.root = 0;
    if ((int)vidx != inf.n_pool) inf.status = REP_MODEL_ERROR;
}

}  // namespace sgp
