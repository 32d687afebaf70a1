// BD fast path kernels (transfer-free processes on a 1- or 2-leaf host, Philox mode) — see sim.cuh.
//
// k_bd_find is the hot kernel of HostTreeGen batches: acceptance of a 100-leaf window is ~0.36 %, so ~276 attempts are
// thrown away per accepted tree. It therefore only COUNTS (no vertex is stored) and it is warp-cooperative: the 32 lanes of a
// warp work on the same replicate, each on its own attempt index claimed from a per-warp shared-memory counter, so
//   * every loop iteration expands exactly one vertex in every lane (no divergence between "phases" of the algorithm),
//   * the geometric tail of the attempt count is divided by 32 (no idle SMs at the end of a batch),
//   * the accepted attempt is the smallest accepted index (ballot-free: shared-memory atomicMin), i.e. exactly what the
//     sequential retry loop of GuestTreeMachina.sampleGuestTree (apps/SaGePhy/GuestTreeMachina.java:81-98) would accept.
// k_bd_build regenerates the accepted attempt of each replicate into the exactly-sized node arena.
#pragma once
#include "sim.cuh"

namespace sgp {

constexpr int kBdFindThreads = 256;

__global__ void __launch_bounds__(kBdFindThreads) k_bd_find(BdArgs a) {
    __shared__ double2 s_logtab[128];
    load_log_table(s_logtab);
    __shared__ int s_next[kBdFindThreads / 32];     // next unclaimed attempt index of the warp's replicate
    __shared__ int s_best[kBdFindThreads / 32];     // smallest accepted attempt index so far
    double st_time[kBdStack]; int8_t st_arc[kBdStack];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const BdConsts C = bd_consts(a.P);
    const int leaf0 = a.H.nV == 1 ? 0 : a.H.left[a.H.root];
    const bool two_leaves = a.H.nV > 1;
    volatile int* v_best = &s_best[wib];
    for (;;) {
        long long r = 0;
        if (lane == 0) r = (long long)atomicAdd(a.work_counter, 1ULL);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= a.n) break;
        const unsigned long long gid = (unsigned long long)(a.rep_base + r);
        const uint32_t rep_lo = (uint32_t)gid, rep_hi = (uint32_t)(gid >> 32);
        int exact = -1;
        if (a.P.n_sizes > 0) { PhiloxStream ph; ph.init(a.seed, gid, 0xFFFFFFFFu); exact = a.sizes[ph.next_int(a.P.n_sizes)]; }
        if (lane == 0) { s_next[wib] = 0; s_best[wib] = 0x7fffffff; }
        __syncwarp();
        bool idle = false; int my_a = -1, sp = 0; uint32_t vidx = 0; int sampled = 0, c0 = 0, c1 = 0; bool rejected = false, overflow = false;
        int acc_a = 0x7fffffff, acc_n = 0, acc_sampled = 0;
        unsigned long long v_completed = 0, v_abandoned = 0; int n_completed = 0;
        U4 w_next{0, 0, 0, 0};
        for (;;) {
            if (sp == 0 && !idle) {
                if (my_a >= 0) {          // attempt my_a ran to completion (or was cut short by the max-leaves test)
                    v_completed += vidx; ++n_completed;
                    bool ok = !rejected && sampled >= a.P.min_leaves && sampled <= a.P.max_leaves && (exact == -1 || sampled == exact);
                    if (ok && two_leaves) ok = c0 >= a.P.minper && c0 <= a.P.maxper && c1 >= a.P.minper && c1 <= a.P.maxper;
                    if (ok) { atomicMin(&s_best[wib], my_a); acc_a = my_a; acc_n = (int)vidx; acc_sampled = sampled; }
                }
                const int na = atomicAdd(&s_next[wib], 1);
                if (na > *v_best || na > a.P.max_attempts) { my_a = -1; idle = true; }
                else {
                    my_a = na; st_time[0] = a.H.tip; st_arc[0] = (int8_t)a.H.root; sp = 1; vidx = 0; sampled = 0; c0 = 0; c1 = 0; rejected = false;
                    w_next = philox4x32_10(0u, (uint32_t)my_a, rep_lo, rep_hi, k0, k1);
                }
            }
            if (__all_sync(0xffffffffu, idle)) break;
            if (my_a >= 0) {
                if (my_a > *v_best) { v_abandoned += vidx; sp = 0; my_a = -1; idle = true; continue; }     // a smaller index was accepted meanwhile
                --sp;
                const double start = st_time[sp]; const int X = st_arc[sp];
                const U4 w = w_next;
                ++vidx;
                w_next = philox4x32_10(vidx, (uint32_t)my_a, rep_lo, rep_hi, k0, k1);     // next vertex's block: independent of this vertex's math
                const BdVertex v = bd_vertex(C, a.H, X, start, w, s_logtab);
                if (v.event == EV_LEAF) {
                    ++sampled; if (X == leaf0) ++c0; else ++c1;
                    if (sampled > a.P.max_leaves) { rejected = true; sp = 0; }          // the attempt can no longer be accepted
                } else if (v.event == EV_DUPLICATION || v.event == EV_SPECIATION) {
                    if (sp + 2 > kBdStack) { rejected = true; overflow = true; sp = 0; }
                    else {
                        const bool dup = v.event == EV_DUPLICATION;
                        st_time[sp] = v.et; st_arc[sp] = (int8_t)(dup ? X : a.H.right[X]);
                        st_time[sp + 1] = v.et; st_arc[sp + 1] = (int8_t)(dup ? X : a.H.left[X]); sp += 2;
                    }
                }
            }
        }
        // resolve: the lane that owns the smallest accepted index reports
        const int best = *v_best;
        for (int o = 16; o > 0; o >>= 1) {
            v_completed += __shfl_xor_sync(0xffffffffu, v_completed, o); v_abandoned += __shfl_xor_sync(0xffffffffu, v_abandoned, o);
            n_completed += __shfl_xor_sync(0xffffffffu, n_completed, o);
        }
        const bool any_overflow = __any_sync(0xffffffffu, overflow);
        if (best != 0x7fffffff) {
            if (acc_a == best) {
                RepInfo inf; memset(&inf, 0, sizeof(inf));
                inf.status = REP_OK; inf.exact = exact; inf.root = -1; inf.p_root = -1; inf.acc_attempt = (uint32_t)best; inf.n_pool = acc_n;
                inf.sampled_leaves = acc_sampled; inf.attempts = best + 2;       // trees generated (best + 1) + 1, GuestTreeMachina.java:97,107
                inf.vertices_sampled = v_completed; inf.vertices_abandoned = v_abandoned; inf.attempts_completed = n_completed;
                a.info[r] = inf;
            }
        } else if (lane == 0) {
            RepInfo inf; memset(&inf, 0, sizeof(inf));
            inf.status = any_overflow ? REP_STACK_OVERFLOW : REP_MAX_ATTEMPTS; inf.exact = exact; inf.root = -1; inf.p_root = -1;
            inf.attempts = a.P.max_attempts + 1; inf.vertices_sampled = v_completed; inf.vertices_abandoned = v_abandoned; inf.attempts_completed = n_completed;
            a.info[r] = inf;
        }
        __syncwarp();
    }
}

// Single-arc specialisation (HostTreeGen without -bi; requires vtime[0] == 0): no arc bookkeeping, and the vertex step is
// branch-free. The lineage stack keeps its top element in a register: slot d is the start time of parked lineage d (1-based,
// slot 0 is a dummy so that the refill load never needs a guard), `tos` mirrors slot `depth`. A duplication stores the sibling and
// continues with one child; every other event continues with `tos` and issues the refill load one whole vertex step before its
// value can be needed. Lanes of a warp take different events every trip, so selects instead of branches halve the issue count.
//
// The first kBdSmemStack slots of every lane's stack live in SHARED memory, laid out [slot][thread]: lanes sit at different
// depths, so a per-thread local-memory array made every stack access touch up to 32 different L1 lines (32 wavefronts per
// instruction), while [slot][thread] hits 32 different banks whatever the depths are (2 wavefronts for 8-byte words). The C2
// workload is below depth 16 in 99.997 % of its vertex steps; deeper lineages spill to a local-memory tail.
constexpr int kBdSmemStack = 16;
constexpr int kBdLogCopies = 4;         // replicated log table (see log_unit)
__device__ __forceinline__ int lds_volatile(uint32_t addr) { int v; asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ int atoms_add(uint32_t addr, int v) { int o; asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(o) : "r"(addr), "r"(v) : "memory"); return o; }
__device__ __forceinline__ void atoms_min(uint32_t addr, int v) { asm volatile("red.shared.min.s32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
// Variants (template parameters; the engine picks one, SGP_FIND_VARIANT overrides it for A/B measurements):
//   SMEM_STACK  first kBdSmemStack stack slots in shared memory [slot][thread] instead of a per-thread local array
//   LOGCOPIES   replication factor of the log table
//   PTXCTL      control words through ld.volatile.shared / atom.shared with addresses computed once
template <bool SMEM_STACK, int LOGCOPIES, bool PTXCTL, int MINBLOCKS>
__global__ void __launch_bounds__(kBdFindThreads, MINBLOCKS) k_bd_find_single(BdArgs a) {
    __shared__ double2 s_logtab[128 * LOGCOPIES];
    load_log_table_rep<LOGCOPIES>(s_logtab);
    __shared__ int s_next[kBdFindThreads / 32];
    __shared__ int s_best[kBdFindThreads / 32];
    __shared__ double s_st[SMEM_STACK ? kBdSmemStack + 1 : 1][SMEM_STACK ? kBdFindThreads : 1];
    double st_hi[SMEM_STACK ? kBdStack - kBdSmemStack : kBdStack + 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const BdConsts C = bd_consts(a.P);
    const double tip = a.H.tip;
    const int min_leaves = a.P.min_leaves, max_leaves = a.P.max_leaves, max_attempts = a.P.max_attempts;
    // 32-bit shared-window addresses computed once (generic pointers make the compiler rebuild the window base at every use)
    const uint32_t a_next = (uint32_t)__cvta_generic_to_shared(&s_next[wib]), a_best = (uint32_t)__cvta_generic_to_shared(&s_best[wib]);
    volatile int* v_best = &s_best[wib];
    auto best_now = [&]() { return PTXCTL ? lds_volatile(a_best) : *v_best; };
    double* const my_st = &s_st[0][SMEM_STACK ? threadIdx.x : 0];          // slot d at my_st[d * kBdFindThreads]
    if (SMEM_STACK) my_st[0] = 0.0; else st_hi[0] = 0.0;
    for (;;) {
        long long r = 0;
        if (lane == 0) r = (long long)atomicAdd(a.work_counter, 1ULL);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= a.n) break;
        const unsigned long long gid = (unsigned long long)(a.rep_base + r);
        const uint32_t rep_lo = (uint32_t)gid, rep_hi = (uint32_t)(gid >> 32);
        int exact = -1;
        if (a.P.n_sizes > 0) { PhiloxStream ph; ph.init(a.seed, gid, 0xFFFFFFFFu); exact = a.sizes[ph.next_int(a.P.n_sizes)]; }
        if (lane == 0) { s_next[wib] = 0; s_best[wib] = 0x7fffffff; }
        __syncwarp();
        // lane state: 0 = needs an attempt, 1 = running attempt my_a, 2 = idle (nothing left to claim)
        int state = 0, my_a = -1, depth = 0; uint32_t vidx = 0; int sampled = 0; bool cut_short = false;
        int ovf_a = 0x7fffffff;          // smallest attempt index that outgrew the stack (it counts as rejected: the replicate is refused if it would have won)
        double cur = 0.0, tos = 0.0;
        int acc_a = 0x7fffffff, acc_n = 0, acc_sampled = 0;
        unsigned long long v_completed = 0, v_abandoned = 0; int n_completed = 0;
        U4 w_next{0, 0, 0, 0};
        // Per-lane loop: a lane leaves it when there is nothing left to claim and waits for the others at the reconvergence
        // point below (no warp vote per vertex step).
        for (;;) {
            if (state == 0) {
                const int na = PTXCTL ? atoms_add(a_next, 1) : atomicAdd(&s_next[wib], 1);
                if (na > best_now() || na > max_attempts) break;
                my_a = na; state = 1; cur = tip; tos = 0.0; depth = 0; vidx = 0; sampled = 0;
                w_next = philox4x32_10_rk(0u, (uint32_t)my_a, rep_lo, rep_hi, a.keys);
            }
            if (my_a > best_now()) { cut_short = true; break; }      // a smaller index was accepted meanwhile
            const U4 w = w_next;
            ++vidx;
            w_next = philox4x32_10_rk(vidx, (uint32_t)my_a, rep_lo, rep_hi, a.keys);   // next vertex's block: independent of this vertex's math
            const BdVertex v = bd_vertex_single<LOGCOPIES>(C, cur, w, s_logtab, lane & (LOGCOPIES - 1));
            const bool dup = v.event == EV_DUPLICATION;
            sampled += v.event == EV_LEAF ? 1 : 0;
            const bool ovf = dup && depth >= kBdStack;
            const int nd = dup ? depth + 1 : depth - 1;
            double refill;
            if constexpr (SMEM_STACK) {
                const bool deep = nd > kBdSmemStack;                 // rare: the slot lives in the local-memory tail
                if (dup && !ovf && !deep) my_st[nd * kBdFindThreads] = v.et;
                refill = my_st[(deep || nd < 0 ? 0 : nd) * kBdFindThreads];
                if (deep) {
                    const int k = (nd <= kBdStack ? nd : kBdStack) - kBdSmemStack - 1;
                    if (dup && !ovf) st_hi[k] = v.et;
                    refill = st_hi[k];
                }
            } else {
                if (dup && !ovf) st_hi[depth + 1] = v.et;
                refill = st_hi[nd > 0 ? (nd <= kBdStack ? nd : 0) : 0];
            }
            cur = dup ? v.et : tos;
            tos = dup ? v.et : refill;
            depth = nd;
            if (nd < 0 || sampled > max_leaves || ovf) {       // ran to completion, or can no longer be accepted
                state = 0;
                if (ovf && my_a < ovf_a) ovf_a = my_a;
                v_completed += vidx; ++n_completed;
                const bool ok = !ovf && sampled >= min_leaves && sampled <= max_leaves && (exact == -1 || sampled == exact);
                if (ok) { if (PTXCTL) atoms_min(a_best, my_a); else atomicMin(&s_best[wib], my_a); acc_a = my_a; acc_n = (int)vidx; acc_sampled = sampled; }
            }
        }
        if (cut_short) v_abandoned += vidx;
        __syncwarp();
        const int best = best_now();
        for (int o = 16; o > 0; o >>= 1) {
            v_completed += __shfl_xor_sync(0xffffffffu, v_completed, o); v_abandoned += __shfl_xor_sync(0xffffffffu, v_abandoned, o);
            n_completed += __shfl_xor_sync(0xffffffffu, n_completed, o);
        }
        // An attempt deeper than the stack is treated as rejected. If its index is below the accepted one, the sequential retry
        // loop of the reference might have accepted IT: refuse the replicate instead of returning a biased tree.
        const int first_ovf = __reduce_min_sync(0xffffffffu, ovf_a);
        if (best != 0x7fffffff && first_ovf > best) {
            if (acc_a == best) {
                RepInfo inf; memset(&inf, 0, sizeof(inf));
                inf.status = REP_OK; inf.exact = exact; inf.root = -1; inf.p_root = -1; inf.acc_attempt = (uint32_t)best; inf.n_pool = acc_n;
                inf.sampled_leaves = acc_sampled; inf.attempts = best + 2;
                inf.vertices_sampled = v_completed; inf.vertices_abandoned = v_abandoned; inf.attempts_completed = n_completed;
                a.info[r] = inf;
            }
        } else if (lane == 0) {
            RepInfo inf; memset(&inf, 0, sizeof(inf));
            inf.status = first_ovf != 0x7fffffff ? REP_STACK_OVERFLOW : REP_MAX_ATTEMPTS; inf.exact = exact; inf.root = -1; inf.p_root = -1;
            inf.attempts = max_attempts + 1; inf.vertices_sampled = v_completed; inf.vertices_abandoned = v_abandoned; inf.attempts_completed = n_completed;
            a.info[r] = inf;
        }
        __syncwarp();
    }
}

// Build pass: regenerates the accepted attempt depth-first straight into the arena (no scratch). Thread per replicate; the 32
// records a warp produces per step pass through shared memory so that they leave as 16-byte stores in which five consecutive
// lanes cover one record (whole sectors) - a lane storing its own 80-byte record made every store instruction touch 32 sectors
// (the kernel sat on lg_throttle). A right child is not linked into its parent here: that 4-byte write to a record stored
// hundreds of steps earlier came back from DRAM as a read-modify-write (29 KB written + 8 KB read per 16 KB tree); the record
// carries NF_RIGHT_CHILD instead and the label kernels, which hold the tree in shared memory, enter the link.
constexpr int kBdBuildThreads = 256;
__global__ void __launch_bounds__(kBdBuildThreads) k_bd_build(BdArgs a) {
    __shared__ double2 s_logtab[128];
    __shared__ __align__(16) Node s_node[kBdBuildThreads / 32][32];
    __shared__ Node* s_dst[kBdBuildThreads / 32][32];
    load_log_table(s_logtab);
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    bool active = r < a.n && a.info[r < a.n ? r : 0].status == REP_OK;
    RepInfo* inf = active ? &a.info[r] : nullptr;
    Node* N = active ? a.nodes + a.node_off[r] : nullptr;
    const unsigned long long gid = (unsigned long long)(a.rep_base + r);
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32), rep_lo = (uint32_t)gid, rep_hi = (uint32_t)(gid >> 32);
    const uint32_t acc = active ? inf->acc_attempt : 0u;
    const BdConsts C = bd_consts(a.P);
    double st_time[kBdStack]; int8_t st_arc[kBdStack]; int32_t st_par[kBdStack];   // parent index; bit 31 set = right child
    int sp = 0; uint32_t vidx = 0;
    if (active) { st_time[0] = a.H.tip; st_arc[0] = (int8_t)a.H.root; st_par[0] = -1; sp = 1; }
    while (__any_sync(0xffffffffu, sp > 0)) {
        Node* dst = nullptr;
        if (sp > 0) {
            --sp;
            const double start = st_time[sp]; const int X = st_arc[sp]; const int32_t pw = st_par[sp];
            const U4 w = philox4x32_10(vidx, acc, rep_lo, rep_hi, k0, k1);
            const BdVertex v = a.H.nV == 1 ? bd_vertex_single(C, start, w, s_logtab) : bd_vertex(C, a.H, X, start, w, s_logtab);
            const int me = (int)vidx++;
            const bool internal = v.event == EV_DUPLICATION || v.event == EV_SPECIATION;
            Node nd;
            nd.abstime = v.et;
            nd.bl = v.boundary ? round_sig(a.T, start - v.et, 8) : v.bt;       // GuestTreeInHostTreeCreator.java:385 vs :378
            if (me == 0 && nd.bl <= 1.0e-32) nd.bl = 0.0;
            nd.sigma = X; nd.left = internal ? me + 1 : -1; nd.right = -1; nd.from_arc = -1; nd.to_arc = -1; nd.number = -1;       // depth-first: the left child is the next vertex
            nd.p_bl = 0.0; nd.p_left = -1; nd.p_right = -1; nd.event = v.event; nd.prun = PR_UNKNOWN; nd.flags = (pw != -1 && pw < 0) ? NF_RIGHT_CHILD : 0; nd.pad0 = 0; nd.gtree = 0;
            nd.from_g = -1; nd.to_g = -1; nd.chain_u = 0; nd.chain_p = 0; nd.pad1 = 0;
            int ep = a.H.v2e[X];
            if (!v.boundary) { while (a.H.ep_up[ep] < v.et) ++ep; }
            nd.epoch = ep;
            nd.parent = pw == -1 ? -1 : (pw & 0x7fffffff);
            {
                const uint4* src = reinterpret_cast<const uint4*>(&nd); uint4* stg = reinterpret_cast<uint4*>(&s_node[wib][lane]);
#pragma unroll
                for (int q = 0; q < (int)(sizeof(Node) / 16); ++q) stg[q] = src[q];
            }
            dst = &N[me];
            if (internal) {
                const bool dup = v.event == EV_DUPLICATION;
                st_time[sp] = v.et; st_arc[sp] = (int8_t)(dup ? X : a.H.right[X]); st_par[sp] = (int32_t)(0x80000000u | (uint32_t)me);
                st_time[sp + 1] = v.et; st_arc[sp + 1] = (int8_t)(dup ? X : a.H.left[X]); st_par[sp + 1] = me; sp += 2;
            }
        }
        s_dst[wib][lane] = dst;
        __syncwarp();
        // 32 records x 5 quarters = 160 16-byte items, item i = quarter i % 5 of the record of lane i / 5
#pragma unroll
        for (int j = 0; j < (int)(sizeof(Node) / 16); ++j) {
            const int item = lane + 32 * j, slot = item / 5, q = item - slot * 5;
            Node* d = s_dst[wib][slot];
            if (d) reinterpret_cast<uint4*>(d)[q] = reinterpret_cast<const uint4*>(&s_node[wib][slot])[q];
        }
        __syncwarp();
    }
    if (active) {
        inf->root = 0;
        if ((int)vidx != inf->n_pool) inf->status = REP_MODEL_ERROR;
    }
}

}  // namespace sgp
