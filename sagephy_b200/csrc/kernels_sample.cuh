// Sampling kernels (K1) of the general engine: "find" decides which attempt of each replicate is accepted, "build" regenerates
// that attempt into the exactly-sized node arena.
//
// A warp simulates 32 replicates, one per lane, in ONE flat loop: every trip, lanes without work claim a replicate (find) and
// every lane with an attempt in progress takes one step of it (sim_step / domain_step: exactly one vertex is created per step),
// so the lanes meet again at the top of each trip and share the expensive Java-exact vertex draw. The pieces that are linear in
// the host epoch or the guest tree (recipient draws, victim searches) are served by the whole warp (coop_resolve_*). All 32
// lanes stay in the loop until the warp has nothing left to do, because those pieces need them.
#pragma once
#include <type_traits>
#include "sim_domain.cuh"

namespace sgp {

template <bool DOMAIN> using EngineState = typename std::conditional<DOMAIN, DomainState, SimState>::type;

template <bool DOMAIN, class RNG, class ARGS>
__device__ __forceinline__ void engine_begin(EngineState<DOMAIN>& S, RNG& rng, const ARGS& a, const DomainEnv& E) {
    if constexpr (DOMAIN) domain_begin(S, rng, a.P, E); else sim_begin(S, rng, a.P, a.H);
}
template <bool DOMAIN, class RNG, class ARGS>
__device__ __forceinline__ void engine_resolve(EngineState<DOMAIN>& S, RNG& rng, const ARGS& a, const DomainEnv& E, SimWork& W) {
    if constexpr (DOMAIN) coop_resolve_domain(S, rng, a.P, E, W); else coop_resolve_transfers(S, rng, a.P, a.H, W);
}
template <bool DOMAIN, bool COUNT, class RNG, class ARGS>
__device__ __forceinline__ int engine_step(EngineState<DOMAIN>& S, RNG& rng, const ARGS& a, const DomainEnv& E, SimWork& W) {
    if constexpr (DOMAIN) return domain_step(S, rng, a.P, E, a.T, W); else return sim_step<COUNT>(S, rng, a.P, a.H, a.T, W);
}

template <bool REPLAY, bool DOMAIN>
__global__ void __launch_bounds__(128) k_find_general(FindArgs a) {
    __shared__ double2 s_logtab[128];
    load_log_table(s_logtab);
    const int worker = blockIdx.x * blockDim.x + threadIdx.x;
    SimWork W; W.logtab = s_logtab;
    W.nodes = a.scratch_nodes + (size_t)worker * a.cap; W.cap = a.cap;
    W.queue = a.scratch_queue + (size_t)worker * a.cap; W.pend = a.scratch_pend + (size_t)worker * a.cap;
    const int mark_n = DOMAIN ? a.max_gene_nv : a.H.nV, cnt_n = DOMAIN ? a.max_gene_nv : a.H.nLeaves;
    W.marks = a.scratch_marks + (size_t)worker * mark_n;
    W.leaf_cnt = a.per_leaf_check ? a.scratch_leafcnt + (size_t)worker * cnt_n : nullptr;
    W.mark_gen = 0;
    W.arc_head = (!DOMAIN && a.scratch_arc_head) ? a.scratch_arc_head + (size_t)worker * a.H.nV : nullptr; W.arc_gen = 0;
    for (int i = 0; i < mark_n; ++i) W.marks[i] = 0;
    DomainEnv E{a.genes, a.n_genes, a.inter_gene, a.inter_species, DOMAIN ? a.scratch_used + (size_t)worker * a.n_genes : nullptr, a.ev_base, a.ev_off, a.ev_list};
    MtStream mt; PhiloxStream ph;
    if (REPLAY) { mt.mt = a.mt_state + (size_t)(worker >> 5) * 624 * 32 + (worker & 31); mt.stride = 32; }
    EngineState<DOMAIN> S; S.stage = -1;
    bool have = false, exhausted = false, again = false, big = false;
    long long r = 0; RepInfo inf; int attempts = 0; uint64_t verts = 0, pos = 0;
    bool copy_pending = false; long long copy_r = 0; int copy_n = 0;
    for (;;) {
        if (a.arena) {
            // One-pass mode: the trees accepted in the previous trip leave their lanes' scratch pools for the arena before the lanes
            // start anything new. The whole warp copies (16 bytes per lane, contiguous on both sides) - regenerating the tree in a
            // second kernel cost 40 % of the sampling time, a per-lane copy as much as that.
            unsigned cp = __ballot_sync(0xffffffffu, copy_pending);
            while (cp) {
                const int l = __ffs(cp) - 1; cp &= cp - 1;
                const int nn = __shfl_sync(0xffffffffu, copy_n, l);
                const long long rr = __shfl_sync(0xffffffffu, copy_r, l);
                const uint4* s4 = reinterpret_cast<const uint4*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(W.nodes), l));
                unsigned long long off = 0;
                if ((threadIdx.x & 31) == 0) off = atomicAdd(a.arena_cursor, (unsigned long long)nn);
                off = __shfl_sync(0xffffffffu, off, 0);
                if (off + (unsigned long long)nn <= a.arena_cap) {
                    uint4* d4 = reinterpret_cast<uint4*>(a.arena + off);
                    const int nq = nn * (int)(sizeof(Node) / 16);
                    // eight loads in flight per lane (the source is cold: one load per trip made the copy latency-bound)
                    int i = threadIdx.x & 31;
                    for (; i + 7 * 32 < nq; i += 8 * 32) {
                        uint4 v[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = s4[i + q * 32];
#pragma unroll
                        for (int q = 0; q < 8; ++q) d4[i + q * 32] = v[q];
                    }
                    for (; i < nq; i += 32) d4[i] = s4[i];
                    if ((threadIdx.x & 31) == 0) a.node_off_out[rr] = (long long)off;
                }       // else: the cursor ends beyond the capacity; the engine sees that and lets the build kernel regenerate every tree
            }
            copy_pending = false;
            __syncwarp();
        }
        if (!have && !exhausted) {
            long long slot = 0;
            if (!again) slot = (long long)atomicAdd(a.work_counter, 1ULL);
            if (!again && slot >= a.n) exhausted = true;
            else {
                if (!again) r = a.rep_ids ? a.rep_ids[slot] : slot;
                again = false;
                const long long gid = a.rep_base + r;
                memset(&inf, 0, sizeof(inf));
                inf.status = REP_PENDING; inf.exact = -1; inf.root = -1; inf.p_root = -1;
                if (REPLAY) { uint32_t key[4]; seed_words(a.sseed, (unsigned long long)gid, key); mt.seed(key); }
                else ph.init(a.seed, (uint64_t)gid, 0xFFFFFFFFu);
                if (a.P.n_sizes > 0) inf.exact = a.sizes[REPLAY ? mt.next_int(a.P.n_sizes) : ph.next_int(a.P.n_sizes)];
                attempts = 0; verts = 0; have = true; S.stage = -1;
            }
        }
        if constexpr (!DOMAIN) {
            // the per-host-leaf tallies of an attempt start at zero: cleared by the whole warp for every lane that starts one
            if (a.per_leaf_check) {
                unsigned starting = __ballot_sync(0xffffffffu, have && S.stage < 0 && attempts <= a.P.max_attempts);
                while (starting) {
                    const int l = __ffs(starting) - 1; starting &= starting - 1;
                    int32_t* cnt = reinterpret_cast<int32_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(W.leaf_cnt), l));
                    for (int i = threadIdx.x & 31; i < a.H.nLeaves; i += 32) cnt[i] = 0;
                }
                __syncwarp();
            }
        }
        if (have && S.stage < 0) {          // start the next attempt of this replicate (GuestTreeMachina.java:81-98)
            if (attempts > a.P.max_attempts) { inf.status = REP_MAX_ATTEMPTS; inf.attempts = attempts; inf.vertices_sampled = verts; a.info[r] = inf; have = false; }
            else if (REPLAY) { pos = mt.position(); engine_begin<DOMAIN>(S, mt, a, E); }
            else { ph.begin_attempt((uint32_t)attempts); engine_begin<DOMAIN>(S, ph, a, E); }
        }
        if (!__ballot_sync(0xffffffffu, have || !exhausted)) break;
        if (REPLAY) engine_resolve<DOMAIN>(S, mt, a, E, W); else engine_resolve<DOMAIN>(S, ph, a, E, W);
        if (have && S.stage >= 0) {
            const int st = REPLAY ? engine_step<DOMAIN, true>(S, mt, a, E, W) : engine_step<DOMAIN, true>(S, ph, a, E, W);
            if (st >= 0) {
                verts += (uint64_t)S.n;
                const int n = S.n, root = S.root;
                S.stage = -1;
                if (st != REP_OK) {
                    int spare = -1;
                    if (st == REP_NODE_OVERFLOW && !big && a.big_count > 0) { spare = atomicAdd(a.big_next, 1); if (spare >= a.big_count) spare = -1; }
                    if (spare >= 0) {
                        // the tree outgrew the pool: move to a spare big pool and run the replicate again from its first attempt
                        W.nodes = a.big_nodes + (size_t)spare * a.big_cap; W.queue = a.big_queue + (size_t)spare * a.big_cap; W.pend = a.big_pend + (size_t)spare * a.big_cap;
                        W.cap = a.big_cap; big = true; again = true; have = false;
                    } else { inf.status = st; inf.attempts = attempts; inf.vertices_sampled = verts; a.info[r] = inf; have = false; }
                } else {
                    ++attempts;
                    int sampled = 0; bool ok;
                    if constexpr (DOMAIN) ok = attempt_ok_domain(a.P, a.H, E, W, root, inf.exact, a.max_gene_nv, a.per_leaf_check != 0, a.all_genes != 0, &sampled);
                    else { ok = attempt_ok_counted(a.P, S, W, inf.exact); sampled = S.sampled; }
                    if (ok) {
                        inf.status = REP_OK; inf.n_pool = n; inf.sampled_leaves = sampled; inf.root = root;
                        copy_pending = true; copy_r = r; copy_n = n;
                        if (REPLAY) { inf.acc_attempt = (uint32_t)pos; inf.acc_hi = (uint32_t)(pos >> 32); } else { inf.acc_attempt = (uint32_t)(attempts - 1); inf.acc_hi = 0; }
                        ++attempts;     // GuestTreeMachina.java:107
                        inf.attempts = attempts; inf.vertices_sampled = verts; a.info[r] = inf; have = false;
                    }
                }
            }
        }
    }
}


template <bool REPLAY, bool DOMAIN>
__global__ void __launch_bounds__(128) k_build_general(BuildArgs a) {
    __shared__ double2 s_logtab[128];
    load_log_table(s_logtab);
    const int worker = blockIdx.x * blockDim.x + threadIdx.x;
    SimWork W; W.logtab = s_logtab; W.nodes = nullptr; W.cap = 0; W.queue = nullptr; W.pend = nullptr;
    const int mark_n = DOMAIN ? a.max_gene_nv : a.H.nV;
    W.marks = a.scratch_marks + (size_t)worker * mark_n; W.leaf_cnt = nullptr; W.mark_gen = 0;
    W.arc_head = (!DOMAIN && a.scratch_arc_head) ? a.scratch_arc_head + (size_t)worker * a.H.nV : nullptr; W.arc_gen = 0;
    for (int i = 0; i < mark_n; ++i) W.marks[i] = 0;
    DomainEnv E{a.genes, a.n_genes, a.inter_gene, a.inter_species, DOMAIN ? a.scratch_used + (size_t)worker * a.n_genes : nullptr, a.ev_base, a.ev_off, a.ev_list};
    MtStream mt; PhiloxStream ph;
    if (REPLAY) { mt.mt = a.mt_state + (size_t)(worker >> 5) * 624 * 32 + (worker & 31); mt.stride = 32; }
    EngineState<DOMAIN> S; S.stage = -1;
    bool have = false, exhausted = false;
    long long r = 0;
    // same flat loop as the find kernel: a lane that has finished its tree takes the next one (trees differ a lot in size; one
    // tree per thread would leave most lanes of a warp waiting for its largest tree)
    for (;;) {
        if (!have && !exhausted) {
            const long long t = (long long)atomicAdd(a.work_counter, 1ULL);
            if (t >= a.n) exhausted = true;
            else {
                if (!a.order) r = t;
                else {
                    // large trees are started early but one per warp: 32 of them in one warp would queue up behind each other's
                    // cooperative recipient draws for their whole (long) life
                    long long nb = *a.n_big; if (nb * 32 > a.n) nb = 0;
                    const long long w = t >> 5;
                    if ((t & 31) == 0 && w < nb) r = a.order[w];
                    else { const long long before = (t + 31) >> 5 < nb ? (t + 31) >> 5 : nb; r = a.order[nb + (t - before)]; }
                    if (nb == 0) r = a.order[t];
                }
                const RepInfo& inf = a.info[r];
                if (inf.status == REP_OK) {
                    const long long off = a.node_off[r];
                    W.nodes = a.nodes + off; W.cap = inf.n_pool; W.queue = a.aux + off; W.pend = a.aux + a.aux_stride + off;
                    const long long gid = a.rep_base + r;
                    if (REPLAY) {
                        uint32_t key[4]; seed_words(a.sseed, (unsigned long long)gid, key); mt.seed(key); mt.skip(((uint64_t)inf.acc_hi << 32) | inf.acc_attempt);
                        engine_begin<DOMAIN>(S, mt, a, E);
                    } else { ph.init(a.seed, (uint64_t)gid, inf.acc_attempt); engine_begin<DOMAIN>(S, ph, a, E); }
                    have = true;
                }
            }
        }
        if (!__ballot_sync(0xffffffffu, have || !exhausted)) break;
        if (REPLAY) engine_resolve<DOMAIN>(S, mt, a, E, W); else engine_resolve<DOMAIN>(S, ph, a, E, W);
        if (have) {
            const int rc = REPLAY ? engine_step<DOMAIN, false>(S, mt, a, E, W) : engine_step<DOMAIN, false>(S, ph, a, E, W);
            if (rc >= 0) {
                RepInfo& inf = a.info[r];
                if (rc != REP_OK || S.n != inf.n_pool) inf.status = REP_MODEL_ERROR; else inf.root = S.root;
                S.stage = -1; have = false;
            }
        }
    }
}

}  // namespace sgp
