// Data layout shared by the sampler, label/prune and emit kernels (see DESIGN.md "Data layout in HBM").
#pragma once
#include <cstdint>
#include "jmath.cuh"

namespace sgp {

// GuestVertex.Event (apps/SaGePhy/GuestVertex.java:11-33), same ordinal order.
enum Ev : uint8_t {
    EV_SPECIATION = 0, EV_CODIVERGENCE = 1, EV_LEAF = 2, EV_UNSAMPLED_LEAF = 3, EV_DUPLICATION = 4, EV_LOSS = 5, EV_REPLACING_LOSS = 6,
    EV_ADDITIVE_TRANSFER = 7, EV_REPLACING_TRANSFER = 8,
    EV_AT_GG_SS = 9, EV_AT_GG_DS = 10, EV_AT_DG_SS = 11, EV_AT_DG_DS = 12,      // additive: intra/inter gene x intra/inter species
    EV_RT_GG_SS = 13, EV_RT_GG_DS = 14, EV_RT_DG_SS = 15, EV_RT_DG_DS = 16,     // replacing
    EV_COUNT = 17
};
enum Prn : uint8_t { PR_UNPRUNABLE = 0, PR_PRUNABLE = 1, PR_COLLAPSABLE = 2, PR_UNKNOWN = 3 };

enum NodeFlags : uint8_t { NF_LEFT_U = 1, NF_LEFT_P = 2, NF_IN_PRUNED = 4,
                           NF_RIGHT_CHILD = 8 /* set by the birth-death build kernel: the label kernels enter the vertex as its parent's right child */,
                           NF_DETACHED = 16 /* general engine, find pass: the vertex hangs below a replaced lineage (not part of the final tree) */ };

// One guest-tree vertex. 80 bytes, 16-byte aligned (five 16-byte vector accesses).
struct __align__(16) Node {
    double abstime;        // absolute time of the vertex
    double bl;             // branch length in the unpruned tree
    int32_t sigma;         // host vertex / arc
    int32_t left, right;   // children in the unpruned tree (-1: none)
    int32_t parent;
    int32_t epoch;         // host epoch number
    int32_t from_arc, to_arc;
    int32_t number;        // vertex number (name = prefix + number)
    double p_bl;           // branch length in the pruned tree
    int32_t p_left, p_right;
    uint8_t event, prun, flags, pad0;
    int16_t gtree;         // host tree index the vertex lives in (DomainTreeGen: gene tree)
    int16_t from_g, to_g;  // transfer endpoints' host tree index (DomainTreeGen)
    int16_t chain_u;       // '(' run written before this leaf, unpruned tree
    int16_t chain_p;       //   "    pruned tree
    int16_t pad1;
};
static_assert(sizeof(Node) == 80, "Node layout");

// ---- emission records -------------------------------------------------------------------------------------------------------------
// The label kernel, which holds a whole tree in shared memory, leaves one compact record per output PIECE in the order the piece is
// written, so that the write kernels read their input sequentially (coalesced, each byte once) instead of gathering 80-byte vertex
// records through an order array once per stream and pass. A record carries everything the piece prints, the shortest-decimal
// digits of its floating-point field (computed once) and the piece's length in bytes.
constexpr int kNoDigits = 0x7fff;      // TreeRec::e / RowRec::e: f holds the raw bits of the double (0.0, NaN, infinities, negatives)
enum TreeRecFlags : uint8_t { TR_LEAF = 1, TR_COMMA = 2, TR_ROOT = 4, TR_J_SHIFT = 3 /* bits 3-4: second DISCPT index */ };
// One Newick piece ['(' x chain | ')'] name ':' length [tag] [','] [tree tag ';' '\n'], post-order of its view.
struct __align__(16) TreeRec {
    uint64_t f;            // branch length = f x 10^e (shortest decimal), or raw bits if e == kNoDigits
    int32_t number, sigma, epoch;
    int16_t e, chain;
    uint8_t event, flags;
    int16_t gtree;
    union {
        struct { uint16_t len; int16_t pad; };     // finished: bytes of the piece (0xffff: longer, measured again by the writer)
        int32_t link;      // raw pruned-view record: index of the same vertex's record in the unpruned view if both print the same length (-1: none)
    };
};
static_assert(sizeof(TreeRec) == 32, "TreeRec layout");
// Transfer vertices only (GuestTreeGen / DomainTreeGen): same index as the TreeRec, touched for transfer events alone.
struct __align__(16) TreeRecX { int32_t from_arc, to_arc; int16_t from_g, to_g; int32_t pad; };
static_assert(sizeof(TreeRecX) == 16, "TreeRecX layout");
enum RowRecFlags : uint8_t { RR_LEAFMAP = 1 };
// One .guest2host row (and the .leafmap row of the same vertex, if it has one), breadth-first order of its view.
struct __align__(16) RowRec {
    uint64_t f;            // vertex time = f x 10^e, or raw bits if e == kNoDigits
    int32_t number, sigma, epoch;
    int16_t e, gtree;
    uint8_t event, flags;
    uint16_t len_g2h;
    union {
        struct { uint16_t len_lm, pad; };
        int32_t link;      // raw pruned-view record: index of the same vertex's row in the unpruned view (-1: none)
    };
};
static_assert(sizeof(RowRec) == 32, "RowRec layout");

// Per-replicate record produced by find/build/label and consumed by emit + the summary histograms.
struct RepInfo {
    int32_t status;          // sgp_rep_status
    int32_t attempts;        // value printed as "Attempts:" (trees generated + 1)
    uint32_t acc_attempt;    // index of the accepted attempt (Philox key) / MT words drawn before it (replay)
    uint32_t acc_hi;         // high half of the MT word position
    int32_t n_pool;          // vertices created in the accepted attempt (size of the node pool)
    int32_t root, p_root;
    int32_t n_u, n_p;        // vertices in the unpruned / pruned tree
    int32_t sampled_leaves;
    int32_t exact;           // -sizes draw (-1: none)
    int16_t root_gtree_u, root_gtree_p;   // host tree index of the root vertex of the unpruned / pruned view (.guest2host header, DomainTreeGen)
    double total_u, beneath_u, total_p, beneath_p;
    double p_root_time;      // time of the pruned tree's root (0 if the pruned tree is empty)
    int32_t cnt_u[EV_COUNT];
    int32_t cnt_p[EV_COUNT];
    uint64_t vertices_sampled;   // vertices created in attempts that ran to completion (work counter)
    uint64_t vertices_abandoned; // vertices of attempts that were cut off because a smaller attempt index had been accepted
    int32_t attempts_completed; int32_t pad2;
};

enum RepStatus : int32_t { REP_OK = 0, REP_MAX_ATTEMPTS = 1, REP_NODE_OVERFLOW = 2, REP_STACK_OVERFLOW = 3, REP_MODEL_ERROR = 4, REP_PENDING = 5 };

// Read-only host-tree tables in device memory.
struct DevHost {
    int32_t nV, nE, root, nLeaves, max_arcs, has_lca;
    int32_t ep_sorted;       // the upper epoch times increase with the epoch number (always, unless the host tree's times are inconsistent)
    const int32_t *parent, *left, *right, *host_tag, *v2e, *ep_off, *ep_arcs, *leaf_rank;
    const double *vtime, *ep_lo, *ep_mid, *ep_up, *lca_time;
};

struct SimParams {
    double lambda, mu, tau, theta, rho, bias_rate, gene_birth;
    int32_t bias;            // 0 none, 1 simple (any unrecognised string), 2 exponential
    int32_t do_gene_birth, ishost;
    int32_t min_leaves, max_leaves, minper, maxper, max_attempts;
    int32_t n_sizes;         // > 0: -sizes list present
    // createGuestVertex's rate sums and event thresholds for the root arc ([0]) and every other arc ([1]): sum, 1/sum,
    // P(dup) boundary and P(loss) boundary. IEEE division is correctly rounded on host and device alike, so taking them from
    // here instead of dividing per vertex does not change a bit (GuestTreeInHostTreeCreator.java:372-412).
    double sum[2], inv_sum[2], thr_dup[2], thr_loss[2];
};

}  // namespace sgp
