// K5: BranchRelaxer — per-branch rate sampling (IID models), retry until every rate lies in [min, max], length x rate, and
// emission of the relaxed Newick tree + .info report.
//
// Reference behaviour: apps/SaGePhy/BranchRelaxer.java:94-101 (retry loop), :148-158 (isOK), :166-173 (lengths x rates),
// :186-197 + io/NewickTreeWriter.java:100-113 (tree text), :118-135 (.info text); rate models in apps/SaGePhy/IID*RateModel.java
// on top of math/{LogNormal,Normal,Gamma,Uniform,Exponential}Distribution.java.
//
//   k_relax_rates<REPLAY> one thread per replicate draws the rate vector in vertex-id order from the replicate's stream
//                         (MT19937 replay: exactly the reference's draw order; Philox: block (vertex, attempt, replicate)).
//   k_relax_apply         Philox fast path: one thread per (replicate, vertex) evaluates attempt 0 in parallel — 8 B read +
//                         16 B written per branch, coalesced; replicates with a rate outside [min, max] are flagged and
//                         finished by k_relax_rates starting from attempt 1.
//   k_relax_emit<WRITE>   same piece machinery as the tree generators (kernels_emit.cuh).
#pragma once
#include "kernels_emit.cuh"
#include "jmath_ext.cuh"
#include "rng.cuh"

namespace sgp {

// java.util.Random.nextGaussian on an MT stream (polar method; the second deviate is cached)
struct GaussCache { bool have; double next; };
__device__ inline double mt_next_gaussian(MtStream& mt, GaussCache& gc) {
    if (gc.have) { gc.have = false; return gc.next; }
    double v1, v2, s;
    do {
        v1 = XSUB(XMUL(2.0, mt.next_double()), 1.0); v2 = XSUB(XMUL(2.0, mt.next_double()), 1.0);
        s = XADD(XMUL(v1, v1), XMUL(v2, v2));
    } while (s >= 1 || s == 0);
    const double mult = jsqrt(XDIV(XMUL(-2.0, jlog(s)), s));
    gc.next = XMUL(v2, mult); gc.have = true;
    return XMUL(v1, mult);
}

// rate from the reference's quantile transforms; *err set when the reference would throw (gamma outside [2e-6, 0.999998])
__device__ inline double rate_from_uniform(const RelaxArgs& a, double u, int* err) {
    switch (a.model) {
        case RM_EXPONENTIAL: return XDIV(-jlog(XSUB(1.0, u)), a.pa);
        case RM_GAMMA: return XDIV(XMUL(jchi2_quantile_pre(u, XMUL(2.0, a.pa), a.gamma_lng, a.ln2, err), a.pb), 2.0);
        case RM_LOGNORMAL: return jexp(XADD(XMUL(jnormal_quantile(u), a.sigma), a.pa));
        default: return a.pa;
    }
}

template <bool REPLAY>
__device__ inline double draw_rate(const RelaxArgs& a, MtStream& mt, GaussCache& gc, unsigned long long gid, uint32_t attempt, int x, int* err) {
    if (a.model == RM_CONSTANT) return a.pa;
    if (REPLAY) {
        switch (a.model) {
            case RM_NORMAL: return XADD(XMUL(mt_next_gaussian(mt, gc), a.sigma), a.pa);
            case RM_UNIFORM: { const double d = mt.next_double(); const bool up = mt.next(1) != 0; const double w = XSUB(a.pb, a.pa);
                               return up ? XADD(XMUL(w, d), a.pa) : XADD(XMUL(w, XSUB(1.0, d)), a.pa); }
            case RM_SAMPLES: return a.samples[mt.next_int(a.n_samples)];
            default: return rate_from_uniform(a, mt.next_double(), err);
        }
    }
    // Philox: vertex x of attempt `attempt` owns the counter blocks (x, attempt | trial << 24, replicate)
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    switch (a.model) {
        case RM_NORMAL: {
            for (uint32_t trial = 0;; ++trial) {
                const U4 w = philox4x32_10((uint32_t)x, attempt ^ (trial << 20) ^ 0x80000000u, (uint32_t)gid, (uint32_t)(gid >> 32), k0, k1);
                const double v1 = 2.0 * u01_from_words(w.x, w.y) - 1.0, v2 = 2.0 * u01_from_words(w.z, w.w) - 1.0, s = v1 * v1 + v2 * v2;
                if (s < 1 && s != 0) return v1 * sqrt(-2.0 * log(s) / s) * a.sigma + a.pa;
            }
        }
        case RM_UNIFORM: { const U4 w = philox4x32_10((uint32_t)x, attempt, (uint32_t)gid, (uint32_t)(gid >> 32), k0, k1);
                           const double d = u01_from_words(w.x, w.y); return (w.z & 1u) ? (a.pb - a.pa) * d + a.pa : (a.pb - a.pa) * (1.0 - d) + a.pa; }
        case RM_SAMPLES: { PhiloxStream ph; ph.init(a.seed, gid, attempt); ph.block = (uint32_t)x << 8; return a.samples[ph.next_int(a.n_samples)]; }
        default: { const U4 w = philox4x32_10((uint32_t)x, attempt, (uint32_t)gid, (uint32_t)(gid >> 32), k0, k1);
                   return rate_from_uniform(a, u01_from_words(w.x, w.y), err); }
    }
}

// ---- autocorrelated models (rates depend on the parent's rate; vertices in RTree.getTopologicalOrdering() order) ---------------
// One uniform / Gaussian per call: MT replay takes the next value of the replicate's stream (the reference's draw order);
// Philox addresses draw `k` of vertex x of attempt `attempt`.
template <bool REPLAY>
__device__ inline double ac_uniform(const RelaxArgs& a, MtStream& mt, unsigned long long gid, uint32_t attempt, int x, uint32_t k) {
    if (REPLAY) return mt.next_double();
    const U4 w = philox4x32_10((uint32_t)x * 256u + k, attempt ^ 0x40000000u, (uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    return u01_from_words(w.x, w.y);
}
template <bool REPLAY>
__device__ inline double ac_gaussian(const RelaxArgs& a, MtStream& mt, GaussCache& gc, unsigned long long gid, uint32_t attempt, int x, uint32_t k) {
    if (REPLAY) return mt_next_gaussian(mt, gc);
    for (uint32_t trial = 0;; ++trial) {
        const U4 w = philox4x32_10((uint32_t)x * 256u + k, attempt ^ (trial << 20) ^ 0xC0000000u, (uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        const double v1 = XSUB(XMUL(2.0, u01_from_words(w.x, w.y)), 1.0), v2 = XSUB(XMUL(2.0, u01_from_words(w.z, w.w)), 1.0), s = XADD(XMUL(v1, v1), XMUL(v2, v2));
        if (s < 1 && s != 0) return XMUL(v1, jsqrt(XDIV(XMUL(-2.0, jlog(s)), s)));
    }
}
// new LogNormalDistribution(mu, sigma2).sampleValue(prng), math/LogNormalDistribution.java:39-48,196-205; *err when the constructor throws
__device__ inline double ac_lognormal(double mu, double sigma2, double u, int* err) {
    if (sigma2 <= 0.0) { *err = 1; return 0.0; }
    return jexp(XADD(XMUL(jnormal_quantile(u), jsqrt(sigma2)), mu));
}
// Fills rates[base ..] for one attempt; vr (the `relaxed` slots, overwritten later) holds the vertex rates of ACRY07 / ACLBPL07.
template <bool REPLAY>
__device__ inline void ac_rates(const RelaxArgs& a, const RelaxTreeView& tv, long long base, MtStream& mt, GaussCache& gc, unsigned long long gid, uint32_t attempt, int* err) {
    const int32_t* topo = a.topo + tv.v0; const int32_t* par = a.parent + tv.v0; const double* len = a.orig + tv.v0;
    double* rates = a.rates + base; double* vr = a.relaxed + base;
    const int n = tv.nV;
    if (a.model == RM_ACTK98) {                 // apps/SaGePhy/ACThorneKishino98RateModel.java:58-74
        rates[topo[0]] = a.pa;
        for (int i = 1; i < n && !*err; ++i) {
            const int x = topo[i], xp = par[x];
            const double deltat = XDIV(XADD(len[x], len[xp]), 2.0);
            const double sigma2 = XMUL(a.pb, deltat);
            rates[x] = ac_lognormal(XSUB(jlog(rates[xp]), XDIV(sigma2, 2.0)), sigma2, ac_uniform<REPLAY>(a, mt, gid, attempt, x, 0), err);
        }
    } else if (a.model == RM_ACRY07) {          // apps/SaGePhy/ACRannalaYang07RateModel.java:56-88
        const int root = topo[0]; double dt = len[root];
        if (dt == 0.0) { rates[root] = a.pa; vr[root] = a.pa; }
        else {
            const double s2 = XMUL(dt, a.pb);
            double r = ac_lognormal(XSUB(jlog(a.pa), XDIV(s2, 2.0)), s2, ac_uniform<REPLAY>(a, mt, gid, attempt, root, 0), err); rates[root] = r;
            if (!*err) r = ac_lognormal(XSUB(jlog(r), XDIV(s2, 2.0)), s2, ac_uniform<REPLAY>(a, mt, gid, attempt, root, 1), err);
            vr[root] = r;
        }
        for (int i = 1; i < n && !*err; ++i) {
            const int x = topo[i], xp = par[x];
            dt = XDIV(len[x], 2.0);
            const double s2 = XMUL(dt, a.pb);
            double r = ac_lognormal(XSUB(jlog(vr[xp]), XDIV(s2, 2.0)), s2, ac_uniform<REPLAY>(a, mt, gid, attempt, x, 0), err); rates[x] = r;
            if (!*err) r = ac_lognormal(XSUB(jlog(r), XDIV(s2, 2.0)), s2, ac_uniform<REPLAY>(a, mt, gid, attempt, x, 1), err);
            vr[x] = r;
        }
    } else if (a.model == RM_ACABY02) {         // apps/SaGePhy/ACArisBrosouYang02RateModel.java:49-64
        rates[topo[0]] = a.pa;
        for (int i = 1; i < n && !*err; ++i) {
            const int x = topo[i], xp = par[x];
            const double lambda = XDIV(1.0, rates[xp]);
            if (lambda <= 0) { *err = 1; break; }           // ExponentialDistribution constructor throws
            rates[x] = XDIV(-jlog(XSUB(1.0, ac_uniform<REPLAY>(a, mt, gid, attempt, x, 0))), lambda);
        }
    } else {                                    // apps/SaGePhy/ACLepageBryantPhillipeLartillot07RateModel.java:75-115: CIR, 200 Euler steps per branch
        const int STEPS = 200;
        for (int j = 0; j < n && !*err; ++j) {
            const int x = topo[j];
            double ir = j == 0 ? a.pa : vr[par[x]];
            double r = 0.0;
            const double dt = XDIV(len[x], (double)STEPS);
            if (j == 0 && dt == 0.0) { vr[x] = a.pa; rates[x] = a.pa; continue; }
            if (!(dt > 0.0) && !(dt != dt)) { *err = 1; break; }      // NormalDistribution(0, dt) throws for dt <= 0 (NaN passes the check)
            const double sd = jsqrt(dt);
            for (int i = 0; i < STEPS; ++i) {
                const double root_ir = jsqrt(ir);
                const double g = XADD(XMUL(ac_gaussian<REPLAY>(a, mt, gc, gid, attempt, x, (uint32_t)i), sd), 0.0);
                ir = XADD(XADD(ir, XMUL(XMUL(a.pc, XSUB(a.pb, ir)), dt)), XMUL(XMUL(a.pd, root_ir), g));
                r = XADD(r, ir);
            }
            vr[x] = ir; rates[x] = XDIV(r, (double)STEPS);
        }
    }
}

// IIDRK11 (apps/SaGePhy/IIDRasmussenKellis11RateModel.java:110-157): per guest vertex one Gamma(k, theta) draw per host arc the
// branch passes over (lowest arc first), combined with the precomputed weights and the scale factor (a.pa).
template <bool REPLAY>
__device__ inline void rk11_rates(const RelaxArgs& a, const RelaxTreeView& tv, long long base, MtStream& mt, unsigned long long gid, uint32_t attempt, int* err) {
    double* rates = a.rates + base;
    for (int x = 0; x < tv.nV && !*err; ++x) {
        double r = 0.0;
        for (int s = a.rk_off[x]; s < a.rk_off[x + 1]; ++s) {
            const double k = a.rk_k[s], theta = a.rk_theta[s];
            if (!(k > 0.0) || !(theta > 0.0)) { *err = 1; break; }       // GammaDistribution constructor throws / PARAMS missing (NullPointerException)
            const double u = ac_uniform<REPLAY>(a, mt, gid, attempt, x, (uint32_t)(s - a.rk_off[x]));
            const double draw = XDIV(XMUL(jchi2_quantile(u, XMUL(2.0, k), err), theta), 2.0);
            if (*err) break;
            r = XADD(r, XMUL(a.rk_w[s], draw));
        }
        rates[x] = XMUL(r, a.pa);
    }
}

__device__ __forceinline__ long long relax_slot(const RelaxArgs& a, long long r, const RelaxTreeView** tv) {
    const long long t = r % a.n_trees; *tv = &a.trees[t];
    return (r / a.n_trees) * a.sum_v + a.trees[t].v0;
}

// status[r] on entry: 0 = run from attempt `first_attempt`; anything else = skip. REP_* codes on exit.
template <bool REPLAY>
__global__ void __launch_bounds__(128) k_relax_rates(RelaxArgs a, int first_attempt, int only_flagged) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    if (only_flagged && a.status[r] != REP_PENDING) return;
    const RelaxTreeView* tv; const long long base = relax_slot(a, r, &tv);
    const int nV = tv->nV; const unsigned long long gid = (unsigned long long)(a.rep_base + r);
    if (nV == 0) { a.status[r] = REP_MODEL_ERROR; a.attempts[r] = 0; return; }       // chained stage: the source replicate has no tree
    MtStream mt; GaussCache gc{false, 0.0};
    if (REPLAY) { const int worker = blockIdx.x * blockDim.x + threadIdx.x; mt.mt = a.mt_state + (size_t)(worker >> 5) * 624 * 32 + (worker & 31); mt.stride = 32;
                  uint32_t key[4]; seed_words(a.sseed, (unsigned long long)gid, key); mt.seed(key); }
    int attempts = first_attempt; int status = REP_PENDING;
    for (;;) {
        if (attempts > a.max_attempts) { status = REP_MAX_ATTEMPTS; break; }
        bool ok = true; int err = 0;
        if (a.model >= RM_ACTK98) {
            if (a.model == RM_RK11) rk11_rates<REPLAY>(a, *tv, base, mt, gid, (uint32_t)attempts, &err);
            else ac_rates<REPLAY>(a, *tv, base, mt, gc, gid, (uint32_t)attempts, &err);
            if (!err) for (int x = 0; x < nV; ++x) { const double rate = a.rates[base + x]; if (x != tv->ignore && (rate < a.min_rate || rate > a.max_rate)) ok = false; }
        } else
        for (int x = 0; x < nV; ++x) {
            const double rate = draw_rate<REPLAY>(a, mt, gc, gid, (uint32_t)attempts, x, &err);
            a.rates[base + x] = rate;
            if (x != tv->ignore && (rate < a.min_rate || rate > a.max_rate)) ok = false;      // isOK: rates are NOT short-circuited in the draw
        }
        ++attempts;
        if (err) { status = REP_MODEL_ERROR; break; }       // the reference aborts with an exception (ChiSquared.java:24)
        if (ok) { status = REP_OK; break; }
    }
    a.attempts[r] = attempts; a.status[r] = status;
    if (status == REP_OK) for (int x = 0; x < nV; ++x) a.relaxed[base + x] = XMUL(a.orig[tv->v0 + x], a.rates[base + x]);
}

// Philox fast path, attempt 0, one thread per (replicate, vertex). A draw on which the reference would throw ends the run whatever
// the other rates are (BranchRelaxer.java:137-140), so its mark must win over "rate outside [min, max]": kApplyError > REP_PENDING
// under atomicMax; k_relax_finish_apply turns it into REP_MODEL_ERROR.
constexpr int kApplyError = 7;
__global__ void __launch_bounds__(256) k_relax_apply(RelaxArgs a, long long total_slots) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_slots) return;
    // slot -> (replicate, vertex): slots are grouped by rounds of n_trees replicates
    const long long round = g / a.sum_v; const long long in_round = g - round * a.sum_v;
    int lo = 0, hi = a.n_trees - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (a.trees[mid].v0 <= in_round) lo = mid; else hi = mid - 1; }
    const long long r = round * a.n_trees + lo;
    if (r >= a.n) return;
    const RelaxTreeView& tv = a.trees[lo];
    const int x = (int)(in_round - tv.v0);
    MtStream mt; GaussCache gc{false, 0.0}; int err = 0;
    const double rate = draw_rate<false>(a, mt, gc, (unsigned long long)(a.rep_base + r), 0u, x, &err);
    a.rates[g] = rate;
    a.relaxed[g] = XMUL(a.orig[tv.v0 + x], rate);
    if (err) atomicMax(&a.status[r], kApplyError);
    else if (x != tv.ignore && (rate < a.min_rate || rate > a.max_rate)) atomicMax(&a.status[r], (int)REP_PENDING);
}
// Same work with one WARP per replicate (grid-stride over replicates): used when every replicate has its own tree (chained
// stage), where the slot -> tree search of k_relax_apply would be a 20-step binary search per vertex. Lanes take the vertices of the
// tree 32 at a time: coalesced 8-byte reads of the original lengths, coalesced writes of rates and relaxed lengths.
__global__ void __launch_bounds__(256) k_relax_apply_per_rep(RelaxArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp0; r < a.n; r += n_warps) {
        const RelaxTreeView* tvp; const long long base = relax_slot(a, r, &tvp);
        const RelaxTreeView tv = *tvp;
        const unsigned long long gid = (unsigned long long)(a.rep_base + r);
        int flag = 0;
        for (int x = lane; x < tv.nV; x += 32) {
            MtStream mt; GaussCache gc{false, 0.0}; int err = 0;
            const double rate = draw_rate<false>(a, mt, gc, gid, 0u, x, &err);
            a.rates[base + x] = rate;
            a.relaxed[base + x] = XMUL(a.orig[tv.v0 + x], rate);
            if (err) flag = kApplyError;
            else if (x != tv.ignore && (rate < a.min_rate || rate > a.max_rate) && flag < (int)REP_PENDING) flag = (int)REP_PENDING;
        }
        flag = __reduce_max_sync(0xffffffffu, flag);
        if (lane == 0 && flag) atomicMax(&a.status[r], flag);
    }
}
__global__ void k_relax_finish_apply(RelaxArgs a) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    if (a.trees[r % a.n_trees].nV == 0) { a.status[r] = REP_MODEL_ERROR; a.attempts[r] = 0; return; }
    if (a.status[r] == 0) { a.status[r] = REP_OK; a.attempts[r] = 1; }       // status was zero-initialised == "no violation seen"
    else if (a.status[r] == kApplyError) { a.status[r] = REP_MODEL_ERROR; a.attempts[r] = 1; }
}

// ---- chained stage: generator trees -> templates ---------------------------------------------------------------------------
// What io/NewickTreeReader.java:158-296 + BranchRelaxer.java:79-93 would extract from the generator's own Newick text: vertices
// numbered in post-order (= the order the text lists them), lengths (Double.toString round-trips, so the parsed value is the
// stored double), leaf / sibling structure, names.
__global__ void k_ingest_count(IngestArgs g) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > g.n_src) return;
    long long nv = 0;
    if (r < g.n_src && g.info[r].status == REP_OK) nv = g.pruned ? g.info[r].n_p : g.info[r].n_u;
    g.nv[r] = nv;
}
__global__ void __launch_bounds__(256) k_ingest_fill(IngestArgs g) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= g.n_src) return;
    const long long v0 = g.v0[r]; const int nV = (int)(g.v0[r + 1] - v0);
    if (lane == 0) { RelaxTreeView tv; tv.nV = nV; tv.root = nV - 1; tv.ignore = -1; tv.v0 = v0; g.views[r] = tv; }
    if (nV == 0) return;
    const Node* N = g.nodes + g.node_off[r];
    const int32_t* ord = g.aux + (g.pruned ? g.aux_stride : 0) + g.node_off[r];        // post-order of the chosen view
    if (g.topo) {
        // ids of the chosen view: position in its post-order; then BFS order and parents in those ids
        int32_t* pid = g.post_id + g.node_off[r];
        const int32_t* bfs = g.aux + (g.pruned ? 3 : 2) * g.aux_stride + g.node_off[r];
        for (int k = lane; k < nV; k += 32) pid[ord[k]] = k;
        __syncwarp();
        if (lane == 0) g.parent[v0 + nV - 1] = -1;          // the root is last in post-order
        for (int k = lane; k < nV; k += 32) {
            g.topo[v0 + k] = pid[bfs[k]];
            const Node& x = N[ord[k]];
            const int l = g.pruned ? x.p_left : x.left, rr = g.pruned ? x.p_right : x.right;
            if (l >= 0) { g.parent[v0 + pid[l]] = k; g.parent[v0 + pid[rr]] = k; }
        }
    }
    for (int k = lane; k < nV; k += 32) {
        const Node& x = N[ord[k]];
        const bool leaf = g.pruned ? x.p_left < 0 : x.left < 0;
        uint8_t f = leaf ? 1 : 0;
        if (x.flags & (g.pruned ? NF_LEFT_P : NF_LEFT_U)) f |= 2;      // a first child: a ',' follows it
        if (leaf || g.keep_interior) f |= 4;
        g.orig[v0 + k] = g.pruned ? x.p_bl : x.bl;
        g.chain[v0 + k] = g.pruned ? x.chain_p : x.chain_u;
        g.vflags[v0 + k] = f; g.number[v0 + k] = x.number; g.sigma[v0 + k] = x.sigma;
    }
}

// ---- emission -------------------------------------------------------------------------------------------------------------------
// (f, e): shortest-decimal digits of the relaxed length, computed once by the size pass and cached (rdig_f / rdig_e)
template <class Sink> __device__ void put_relax_vertex(Sink& s, const RelaxArgs& a, const RelaxTreeView& tv, long long base, int x, unsigned long long dig_f, int dig_e) {
    const long long t = tv.v0 + x; const uint8_t f = a.vflags[t];
    if (f & 1) { for (int i = 0; i < a.chain[t]; ++i) s.put('('); } else s.put(')');
    if (f & 4) {
        if (a.syn_names) {
            for (int i = 0; i < a.syn_prefix_len; ++i) s.put(a.syn_prefix[i]);
            put_i64(s, a.syn_number[t]);
            if (a.syn_append_sigma) { s.put('_'); put_i64(s, a.syn_sigma[t]); }
        } else put_blob(s, a.blob, a.name_off[2 * t], a.name_off[2 * t + 1]);
    }
    const double len = a.relaxed[base + x];
    if (len == len) { s.put(':'); put_decimal(s, a.T, dig_f, dig_e, dig_e == kNoDigits ? 0 : dec_len_u64(dig_f)); }
    if (a.do_meta) {
        put_lit(s, "[ID="); put_i64(s, x); put_lit(s, " ORIGBL="); put_double(s, a.T, a.orig[t]); put_lit(s, " RATE="); put_double(s, a.T, a.rates[base + x]); s.put(']');
    }
    if (f & 2) s.put(',');
    if (x == tv.root) { s.put(';'); s.put('\n'); }
}
// info pieces: 0 header, 1 args, 2 attempts, then 3 arrays x nV elements, then the constant model tail
template <class Sink> __device__ void put_relax_info(Sink& s, const RelaxArgs& a, const RelaxTreeView& tv, long long base, long long gid, int attempts, int k) {
    const int nV = tv.nV;
    if (k == 0) { put_lit(s, "# BRANCHLENGTHRELAXER\n"); return; }
    if (k == 1) {
        put_lit(s, "Arguments:\t"); put_blob(s, a.blob, a.args_pre_off, a.args_pre_len);
        if (a.seed_mode) { put_seed(s, a.sseed, a.blob, (unsigned long long)gid); put_blob(s, a.blob, a.args_post_off, a.args_post_len); }
        s.put('\n'); return;
    }
    if (k == 2) { put_lit(s, "Attempts:\t"); put_i64(s, attempts); s.put('\n'); return; }
    k -= 3;
    if (k < 3 * nV) {
        const int arr = k / nV, x = k - arr * nV;
        if (x == 0) put_str(s, arr == 0 ? "Original lengths:\t[" : arr == 1 ? "Rates:\t[" : "Relaxed lengths:\t[");
        else { s.put(','); s.put(' '); }
        put_double(s, a.T, arr == 0 ? a.orig[tv.v0 + x] : arr == 1 ? a.rates[base + x] : a.relaxed[base + x]);
        if (x == nV - 1) { s.put(']'); s.put('\n'); }
        return;
    }
    put_blob(s, a.blob, a.model_tail_off, a.model_tail_len);
}

template <bool WRITE>
__global__ void __launch_bounds__(kEmitWarps * 32, 4) k_relax_emit(RelaxArgs a) {
    __shared__ __align__(16) char s_stage[kEmitWarps][kStageBytes];
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= a.n) return;
    char* stage = s_stage[threadIdx.x >> 5];
    const RelaxTreeView* tvp; const long long base = relax_slot(a, r, &tvp); const RelaxTreeView tv = *tvp;
    const long long gid = a.rep_base + r;
    const bool ok = a.status[r] == REP_OK;
    auto no_store = [](int, unsigned) {}; auto no_load = [](int) { return 0u; };
    // stream 0: relaxed tree
    {
        char* out = WRITE ? a.out[0] + a.offsets[r] : nullptr;
        uint16_t* pl = a.plen_tree + base;
        unsigned long long bytes = 0;
        if (ok) {
            unsigned long long* df = a.rdig_f + base; int16_t* de = a.rdig_e + base;
            if (WRITE) {
                auto pc = [&](auto& s, int x) { put_relax_vertex(s, a, tv, base, x, df[x], (int)de[x]); };
                bytes = emit_pieces<true>(tv.nV, out, stage, pc, [&](int i) { return (unsigned)pl[i]; }, no_store);
            } else {
                auto pc = [&](auto& s, int x) {
                    uint64_t f = 0; int e = kNoDigits; const double len = a.relaxed[base + x];
                    if (len == len) decimal_of(a.T, len, &f, &e);
                    df[x] = (unsigned long long)f; de[x] = (int16_t)e;
                    put_relax_vertex(s, a, tv, base, x, (unsigned long long)f, e);
                };
                bytes = emit_pieces<false>(tv.nV, out, stage, pc, no_load, [&](int i, unsigned v) { pl[i] = (uint16_t)(v > 0xffffu ? 0xffffu : v); });
            }
        }
        if (!WRITE && lane == 0) a.sizes[r] = bytes;
    }
    // stream 1: .info (only with -o)
    if (a.has_outfile) {
        char* out = WRITE ? a.out[1] + a.offsets[(size_t)(a.n + 1) + r] : nullptr;
        uint16_t* pl = a.plen_info + 3 * base + 8 * r;
        unsigned long long bytes = 0;
        if (ok) {
            // lines 0-1 ("# BRANCHLENGTHRELAXER" and "Arguments: [...]", which may embed whole Newick trees and exceed the 16-bit
            // piece-length cache): warp-cooperative copy, length known in closed form
            const char* h0 = "# BRANCHLENGTHRELAXER\nArguments:\t"; const int h0len = 33;
            const int seedlen = a.seed_mode ? seed_dec_len(a.sseed, a.blob, (unsigned long long)gid) : 0;
            if (WRITE) {
                for (int i = lane; i < h0len; i += 32) out[i] = h0[i];
                for (int i = lane; i < a.args_pre_len; i += 32) out[h0len + i] = a.blob[a.args_pre_off + i];
            }
            unsigned long long o = (unsigned long long)h0len + a.args_pre_len;
            if (a.seed_mode) {
                if (WRITE) {
                    if (lane == 0) { MemSink m{out + o}; put_seed(m, a.sseed, a.blob, (unsigned long long)gid); }
                    for (int i = lane; i < a.args_post_len; i += 32) out[o + seedlen + i] = a.blob[a.args_post_off + i];
                }
                o += (unsigned long long)seedlen + a.args_post_len;
            }
            if (WRITE && lane == 0) out[o] = '\n';
            o += 1;
            const int np = 1 + 3 * tv.nV + 1; const int att = a.attempts[r];
            auto pc = [&](auto& s, int k) { put_relax_info(s, a, tv, base, gid, att, k + 2); };
            if (WRITE) bytes = o + emit_pieces<true>(np, out + o, stage, pc, [&](int i) { return (unsigned)pl[i]; }, no_store);
            else bytes = o + emit_pieces<false>(np, out, stage, pc, no_load, [&](int i, unsigned v) { pl[i] = (uint16_t)(v > 0xffffu ? 0xffffu : v); });
        }
        if (!WRITE && lane == 0) a.sizes[(size_t)a.n + r] = bytes;
    }
}

}  // namespace sgp
