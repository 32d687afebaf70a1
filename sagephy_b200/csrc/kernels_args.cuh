// Kernel argument blocks (plain structs passed by value), shared between the kernel translation units and the host engine.
#pragma once
#include "device_types.cuh"
#include "rng.cuh"

namespace sgp {

// Tiny host tree (HostTreeGen: "H0:T;" or "(H0:T,H1:T):0.0;") passed by value to the BD fast path.
struct TinyHost {
    int32_t nV, root;
    int32_t left[3], right[3], parent[3], v2e[3];
    double vtime[3];
    double ep_lo[2], ep_mid[2], ep_up[2];
    double tip;
};

struct FindArgs {
    SimParams P;
    DevHost H;
    MathTables T;
    long long n;                 // replicates in this launch
    const long long* rep_ids;    // optional indirection (re-runs of overflowed replicates); nullptr = identity
    long long rep_base;          // global id of local replicate 0 (multi-GPU shard offset)
    unsigned long long seed;     // Philox key
    SeedSpec sseed;              // MT base seed (replicate i == reference run with -s seed+i)
    RepInfo* info;               // [n_total]
    unsigned long long* work_counter;
    // scratch (general engine): per worker thread
    Node* scratch_nodes; int32_t* scratch_queue; int32_t* scratch_pend; int cap;
    int32_t* scratch_marks; int32_t* scratch_leafcnt;      // nV / nLeaves ints per worker
    unsigned long long* scratch_arc_head;                  // nV entries per worker, zeroed (SimWork::arc_head); nullptr: none
    uint32_t* mt_state;          // 624 words per worker, warp-interleaved (replay mode)
    const int32_t* sizes;        // -sizes list
    int per_leaf_check;
    // DomainTreeGen: forest of gene trees (H above is then the species tree)
    const DevHost* genes; int n_genes; int max_gene_nv; int all_genes; double inter_gene, inter_species;
    int32_t* scratch_used;       // n_genes ints per worker
    // spare scratch pools of big_cap vertices: a lane whose tree outgrows its own pool takes one (if any is left) and restarts the
    // replicate at once instead of leaving it to a later, latency-bound launch
    Node* big_nodes; int32_t* big_queue; int32_t* big_pend; int big_cap, big_count; int* big_next;
    const int32_t* ev_base; const int32_t* ev_off; const int32_t* ev_list;    // inter-gene candidate index: vertices of gene g whose first epoch is e = ev_list[ev_off[ev_base[g] + e] .. ev_off[ev_base[g] + e + 1])
    // one-pass mode: an accepted tree is copied from the lane's scratch pool into the arena by the whole warp (bump allocation;
    // node_off_out[r] = its first record). arena == nullptr: the build kernel regenerates the accepted attempts instead.
    Node* arena; unsigned long long arena_cap; unsigned long long* arena_cursor; long long* node_off_out;
};

struct BuildArgs {
    SimParams P;
    DevHost H;
    MathTables T;
    long long n; long long rep_base; unsigned long long seed; SeedSpec sseed;
    RepInfo* info;
    const long long* node_off;   // exclusive scan of n_pool
    unsigned long long* work_counter;   // build slots are claimed dynamically, like replicates in the find kernel
    const long long* n_big;      // order[0 .. *n_big) are the replicates with unusually large trees
    const long long* order;      // slot t builds replicate order[t] (largest trees first: their serial latency is the kernel's tail); nullptr = identity
    Node* nodes;                 // arena
    int32_t* aux; long long aux_stride;    // 4 int arrays of length total_nodes: aux + k*aux_stride
    int32_t* scratch_marks; int32_t* scratch_leafcnt;
    unsigned long long* scratch_arc_head;
    uint32_t* mt_state;
    const DevHost* genes; int n_genes; int max_gene_nv; int all_genes; double inter_gene, inter_species;
    int32_t* scratch_used;
    const int32_t* ev_base; const int32_t* ev_off; const int32_t* ev_list;
};

struct BdArgs {
    SimParams P;
    TinyHost H;
    MathTables T;
    long long n; long long rep_base; unsigned long long seed;
    PhiloxKeys keys;                  // round keys of `seed`
    RepInfo* info;
    unsigned long long* work_counter;
    const int32_t* sizes;
    // build only
    const long long* node_off; Node* nodes;
};

enum Stream : int { ST_UNPRUNED_TREE = 0, ST_PRUNED_TREE = 1, ST_UNPRUNED_INFO = 2, ST_PRUNED_INFO = 3,
                    ST_UNPRUNED_G2H = 4, ST_PRUNED_G2H = 5, ST_UNPRUNED_LEAFMAP = 6, ST_PRUNED_LEAFMAP = 7, ST_COUNT = 8 };

struct EmitArgs {
    long long n; long long rep_base;
    const RepInfo* info;
    const long long* node_off;
    const Node* nodes;
    const int32_t* aux; long long aux_stride;
    MathTables T;
    int app;                       // 0 HostTreeGen, 1 GuestTreeGen, 2 DomainTreeGen
    int exclude_meta, append_sigma, is_species, do_ml_pruned;
    unsigned stream_mask;
    char prefix[24]; int prefix_len;
    const char* blob;              // constant strings, see below
    int args_pre_off, args_pre_len, args_post_off, args_post_len, seed_mode; SeedSpec sseed;
    int g2h_head_off, g2h_head_len;
    int args_pre_off16[16];        // same 16 shifted copies for the "Arguments:" text when it is long (it may embed a whole host tree); [0] < 0: not built
    int g2h_head_off16[16];        // 16 copies of the header, copy k starting at a blob offset == k (mod 16): aligned 16-byte copies for any destination
    const int32_t* gname_off;      // per host tree: offset/len pairs into blob (GUEST= value / file name)
    const int32_t* leafname_off;   // per host vertex: offset/len pairs into blob (leafmap second column)
    const int32_t* leaf_base;      // DomainTreeGen: first entry of gene tree g in leafname_off; g2h header of gene g = g2h_gene_off[2g], [2g+1]
    const int32_t* g2h_gene_off;
    const double* ep_times;        // 3 doubles per epoch [lo, mid, up] (DISCPT index)
    const double* gene_ep_times; const int32_t* gene_ep_base;     // DomainTreeGen: epochs of gene tree g start at gene_ep_base[g]
    // emission records written by the label kernel (device_types.cuh), [2 views][aux_stride], replicate r at node_off[r]
    TreeRec* trec;                 // Newick pieces, post-order of the view
    TreeRecX* trecx;               // transfer endpoints, same index (GuestTreeGen / DomainTreeGen; nullptr for HostTreeGen)
    RowRec* rows;                  // .guest2host / .leafmap rows, breadth-first order of the view (nullptr if neither stream is wanted)
    // sizes / offsets are laid out [ST_COUNT][stride] over ALL replicates of the run; a launch covers the `n` replicates that
    // start at the position these pointers already point to (sub-batches of a pipelined run)
    long long stride;
    unsigned long long* sizes;     // size pass output: sizes[st * stride + r]
    const unsigned long long* offsets;   // write pass input: offsets[st * stride + r], end of the replicate at [.. + r + 1]
    char* out[ST_COUNT];
    unsigned long long cap[ST_COUNT];    // bytes allocated per stream: a replicate that would end beyond it is not written ...
    int* overflow;                       // ... and raises this flag (the engine then grows the buffers and writes again)
    int stage_bytes;                     // write kernels: bytes of one shared-memory stage buffer (two per warp), a multiple of 128
};

struct LabelArgs {
    long long n;
    RepInfo* info;
    const long long* node_off;
    Node* nodes;
    int32_t* aux; long long aux_stride;     // ordU, ordP, bfsU, bfsP (nullptr: not kept)
    MathTables T;
    double host_root_time;                  // hostTree.getVertexTime(root)
    int has_stem_override; double stem_override;    // HostTreeGen -stem (HostTreeGen.java:74-79)
    const long long* pool;                  // vertices per replicate (0: failed)
    int min_pool, cap;                      // warp kernel: size class (min_pool, cap]; thread kernel: trees above cap
    double* terms; int want_beneath;        // [4][aux_stride] terms of the getInfo sums in breadth-first order: total U, total P, beneath-stem U, P (the last two if want_beneath)
    int32_t* work; long long work_stride; int work_blocks;   // k_label_prune_warp_global: work_blocks blocks with work_stride >= 9 * (max_pool + 2) ints of scratch each
    int write_nodes;                        // write the labelled vertex records and the order arrays back (chained stages, oversized trees)
    EmitArgs E;                             // formats and the record arrays the kernel fills in
};

}  // namespace sgp

namespace sgp {

// ---- BranchRelaxer ------------------------------------------------------------------------------------------------------------
enum RelaxModel : int { RM_CONSTANT = 0, RM_EXPONENTIAL = 1, RM_GAMMA = 2, RM_LOGNORMAL = 3, RM_NORMAL = 4, RM_UNIFORM = 5, RM_SAMPLES = 6,
                        RM_ACTK98 = 7, RM_ACRY07 = 8, RM_ACABY02 = 9, RM_ACLBPL07 = 10,      // autocorrelated: parent -> child along the tree
                        RM_RK11 = 11 };                                                     // host-specific gamma rates (single input tree)

// One input tree ("template"): vertex ids are the reference's post-order numbers, so id order == Newick emission order.
struct RelaxTreeView {
    int32_t nV, root, ignore;          // ignore: root vertex whose missing length was replaced by 0 (-1 otherwise)
    long long v0;                      // offset of this tree's vertices in the flat per-vertex template arrays
};

struct RelaxArgs {
    long long n; long long rep_base; unsigned long long seed;
    int model; double pa, pb, sigma;   // model parameters (a, b, sqrt(variance))
    double pc, pd;                     // ACLBPL07: theta, sigma (pa = start rate, pb = mu)
    double gamma_lng, ln2;             // IIDGamma: jln_gamma(k) and jlog(2.0), constants of every quantile evaluation of the run
    const int32_t* topo;               // per template vertex slot: vertex ids in RTree.getTopologicalOrdering() order (BFS from the root)
    const int32_t* parent;             // per template vertex: parent id (-1: root)
    double min_rate, max_rate; int max_attempts;
    const double* samples; int n_samples;
    const int32_t* rk_off; const double* rk_k; const double* rk_theta; const double* rk_w;    // IIDRK11 segments per vertex of the (single) template
    int n_trees; long long sum_v;      // replicate r relaxes template r % n_trees; its vertex slot base is (r / n_trees) * sum_v + v0
    const RelaxTreeView* trees;
    const double* orig;                // per template vertex: original length (NaN: absent)
    const int16_t* chain;              // '(' run before a leaf
    const uint8_t* vflags;             // bit0: leaf, bit1: a ',' follows (not the last child), bit2: print the name
    const int32_t* name_off;           // offset/len pairs into blob
    const char* blob;
    // chained stage (trees ingested from a tree-generator batch): names are <prefix><number>[_<sigma>] as the generator prints them
    int syn_names; const int32_t* syn_number; const int32_t* syn_sigma; char syn_prefix[24]; int syn_prefix_len, syn_append_sigma;
    MathTables T;
    double* rates; double* relaxed;    // [total vertex slots]
    int32_t* attempts; int32_t* status;// [n]
    uint32_t* mt_state;
    // emission
    int do_meta; int has_outfile;
    int args_pre_off, args_pre_len, args_post_off, args_post_len, seed_mode; SeedSpec sseed;
    int model_tail_off, model_tail_len;            // "Model:\t...\nModel parameters:\n..." (constant per run)
    uint16_t* plen_tree; uint16_t* plen_info;      // cached piece lengths: [vertex slots], [3 * vertex slots + 8 * n]
    unsigned long long* rdig_f; int16_t* rdig_e;   // cached shortest-decimal digits of the relaxed lengths: [vertex slots]
    unsigned long long* sizes;                     // [2][n]
    const unsigned long long* offsets;             // [2][n+1]
    char* out[2];
};

// Device-side ingest of generator trees into BranchRelaxer templates (chained stage).
struct IngestArgs {
    long long n_src; const RepInfo* info; const long long* node_off; const Node* nodes; const int32_t* aux; long long aux_stride;
    int pruned, keep_interior;
    long long* nv;                     // [n_src + 1] vertices per tree (0: unusable source replicate), then scanned into v0
    const long long* v0;
    RelaxTreeView* views; double* orig; int16_t* chain; uint8_t* vflags; int32_t* number; int32_t* sigma;
    int32_t* topo; int32_t* parent;    // autocorrelated models
    int32_t* post_id;                  // scratch, one entry per source vertex record: its id in the chosen view's post-order
};

}  // namespace sgp
