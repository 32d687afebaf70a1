/*
 * sagephy_b200 — C ABI of the B200-native batched phylogeny simulator.
 *
 * This is the drop-in boundary for SaGePhy's data-parallel hot path. The reference (Java, no FFI of its own) exposes the
 * path through two in-JVM seams and the process-level CLI; each entry point below names the seam it replaces
 * (paths relative to src/main/java/uc/sgp/sagephy/ of the reference):
 *
 *   sgp_app_run          <- the CLI contract  `java -jar sagephy.jar <App> [options] <positionals>`:
 *                           apps/SaGePhyStarter.java:75-103 (first token = app), apps/SaGePhy/HostTreeGen.java:25-116,
 *                           GuestTreeGen.java:21-107, DomainTreeGen.java:21-107, BranchRelaxer.java:48-141.
 *                           Same options, defaults (from the code) and positional counts, plus a replicate count.
 *     tree generators    <- in-JVM seam 1: UnprunedGuestTreeCreator (apps/SaGePhy/UnprunedGuestTreeCreator.java:17-59) driven
 *                           by GuestTreeMachina.sampleGuestTree (apps/SaGePhy/GuestTreeMachina.java:69-115), followed by
 *                           PruningHelper / GuestVertex.setMeta / NewickTreeWriter / getInfo / getSigma / getLeafMap — all
 *                           behind sgp_app_run("HostTreeGen" | "GuestTreeGen" | "DomainTreeGen", ...).
 *     BranchRelaxer      <- in-JVM seam 2: RateModel.getRates (apps/SaGePhy/RateModel.java:13-29) + BranchRelaxer.main's retry
 *                           loop and rescaling (apps/SaGePhy/BranchRelaxer.java:94-101,148-197) — sgp_app_run("BranchRelaxer").
 *   sgp_relax_run_on_batch <- the file hand-over between two CLI runs (generator writes <prefix>.pruned.tree, BranchRelaxer
 *                           reads it: apps/SaGePhy/BranchRelaxerParameters.java:167-179), kept on the device.
 *   sgp_batch_write_files[_layout] <- the apps' own file fan-out (apps/SaGePhy/GuestTreeGen.java:66-101, HostTreeGen.java:82-110).
 *
 * All pointers are plain host pointers; no CUDA or torch types cross the boundary. Calls block until the outputs are
 * resident (in HBM for the *_device accessors, in pinned host memory for the host accessors). There is no CPU fallback:
 * every entry point fails with SGP_ERR_NO_DEVICE if no CUDA device is usable.
 */
#ifndef SAGEPHY_B200_H
#define SAGEPHY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgp_context sgp_context;
typedef struct sgp_batch sgp_batch;

typedef enum {
    SGP_OK = 0,
    SGP_ERR_NO_DEVICE = 1,      /* no usable CUDA device / CUDA runtime failure */
    SGP_ERR_INVALID_ARGUMENT = 2,
    SGP_ERR_PARSE = 3,          /* malformed Newick / number */
    SGP_ERR_UNSUPPORTED = 4,    /* an application name the reference does not have */
    SGP_ERR_INTERNAL = 5,
    SGP_USAGE = 6               /* wrong positional count or -h: usage text is in the batch's stdout */
} sgp_status;

/* per-replicate status (GuestTreeMachina retry budget: apps/SaGePhy/MaxAttemptsException.java:8-20) */
typedef enum { SGP_REP_OK = 0, SGP_REP_MAX_ATTEMPTS = 1, SGP_REP_NODE_OVERFLOW = 2, SGP_REP_STACK_OVERFLOW = 3, SGP_REP_MODEL_ERROR = 4 } sgp_rep_status;

typedef enum { SGP_RNG_PHILOX = 0, SGP_RNG_MT_REPLAY = 1 } sgp_rng_mode;

/* output streams of a tree-generator batch (file suffixes of the reference in the comments) */
typedef enum {
    SGP_STREAM_UNPRUNED_TREE = 0,     /* .unpruned.tree        */
    SGP_STREAM_PRUNED_TREE = 1,       /* .pruned.tree          */
    SGP_STREAM_UNPRUNED_INFO = 2,     /* .unpruned.info        */
    SGP_STREAM_PRUNED_INFO = 3,       /* .pruned.info          */
    SGP_STREAM_UNPRUNED_G2H = 4,      /* .unpruned.guest2host  */
    SGP_STREAM_PRUNED_G2H = 5,        /* .pruned.guest2host    */
    SGP_STREAM_UNPRUNED_LEAFMAP = 6,  /* .unpruned.leafmap     */
    SGP_STREAM_PRUNED_LEAFMAP = 7,    /* .pruned.leafmap       */
    SGP_STREAM_COUNT = 8
} sgp_stream;
/* BranchRelaxer batches use stream 0 for the relaxed trees (file = the -o value, or stdout) and stream 1 for <-o value>.info */

/* Batch extension of the reference CLI (everything else in argv is the reference's own grammar). */
typedef struct {
    int64_t n_replicates;       /* number of independent replicates (reference: 1 per JVM launch) */
    int64_t replicate_offset;   /* global index of this call's first replicate (multi-GPU sharding by index range) */
    int32_t rng_mode;           /* sgp_rng_mode; MT replay: replicate i == reference run with -s (seed + i) */
    uint32_t stream_mask;       /* bit s set = emit stream s; 0 = whatever the reference would write for these options */
    int32_t keep_on_device;     /* 1: leave stream bytes in HBM only (host accessors then copy lazily) */
    int32_t flags;              /* SGP_FLAG_* */
} sgp_batch_opts;

/* BranchRelaxer only: the <tree> positional (file or string) holds one Newick tree per line; replicate i relaxes tree i % T. */
#define SGP_FLAG_MULTI_TREE_INPUT 1
/* Tree generators only: keep the vertex records of the accepted trees on the device so that the batch can feed
 * sgp_relax_run_on_batch without a host round trip (costs ~100 bytes of HBM per vertex until the batch is freed). */
#define SGP_FLAG_KEEP_TREES 2
/* Tree generators only: one sub-batch, all phases in order on one stream. This is the default since the pipelined mode measured
 * slower on a B200 (see SGP_FLAG_PIPELINE); the flag is kept so that callers which set it keep working. */
#define SGP_FLAG_SERIAL 4
/* Tree generators only (large HostTreeGen batches in Philox mode): run the sampler of sub-batch k+1 on a second stream next to
 * build / label / emit of sub-batch k. Same bytes as the serial run; the per-phase times of sgp_timings then overlap. Opt-in:
 * on a B200 the persistent sampler and the latency-bound kernels slow each other down more than the overlap gains
 * (10^6 C2 replicates: 481 ms pipelined against 401 ms serial). */
#define SGP_FLAG_PIPELINE 8

/* summary counters of a batch (merged across GPUs with an allreduce(sum) by the caller) */
#define SGP_HIST_LEAF_BINS 1025
typedef struct {
    int64_t n_ok, n_failed;
    int64_t attempts_total;           /* sum of "Attempts:" over replicates */
    int64_t vertices_sampled;         /* vertices created in attempts that ran to completion (work of the sampler) */
    int64_t vertices_abandoned;       /* vertices of attempts cut off because a smaller attempt index was accepted (parallel search overhead) */
    int64_t attempts_completed;       /* attempts that ran to completion (>= the attempts the sequential retry loop needs) */
    int64_t event_counts[17];         /* unpruned trees, indexed by GuestVertex.Event ordinal */
    int64_t leaf_hist[SGP_HIST_LEAF_BINS];   /* sampled leaves per accepted tree (last bin = overflow) */
    int64_t out_bytes[SGP_STREAM_COUNT];
} sgp_summary;

typedef struct {
    double find_ms, build_ms, label_ms, emit_size_ms, emit_write_ms, d2h_ms, total_ms;   /* CUDA-event times of the last run */
    int64_t kernel_launches;
    int64_t h2d_bytes;          /* host->device bytes of the run: host-tree tables, constant strings, tree templates */
    int64_t sub_batches;        /* > 1: pipelined run, phase times overlap (find on one stream, the rest on another) */
} sgp_timings;

sgp_status sgp_context_create(int device, sgp_context** out);
void sgp_context_destroy(sgp_context* ctx);
const char* sgp_last_error(const sgp_context* ctx);

/* Runs `<app> argv...` for opts->n_replicates replicates. app = "HostTreeGen" | "GuestTreeGen" | "DomainTreeGen" |
 * "BranchRelaxer". Returns SGP_USAGE (with a batch holding the usage text) exactly when the reference would print usage. */
sgp_status sgp_app_run(sgp_context* ctx, const char* app, int argc, const char* const* argv, const sgp_batch_opts* opts, sgp_batch** out);

/* Multi-GPU front end of sgp_app_run: the replicates [replicate_offset, replicate_offset + n_replicates) are cut into
 * contiguous index ranges, one per device, and run concurrently (one host thread and one context per device). Replicate ids are
 * global (Philox counters / seed + i), so the shards together are bit-identical to a single-device run; no data crosses
 * between GPUs. `out` receives one batch per device in index order (entries may be NULL after an error); `merged` (may be NULL)
 * receives the element-wise sum of the shard summaries. The same device may be listed more than once. */
typedef struct sgp_multi sgp_multi;
sgp_status sgp_multi_create(const int* devices, int n_devices, sgp_multi** out);
int sgp_multi_size(const sgp_multi* m);
sgp_context* sgp_multi_context(sgp_multi* m, int i);
void sgp_multi_destroy(sgp_multi* m);
sgp_status sgp_app_run_multi(sgp_multi* m, const char* app, int argc, const char* const* argv, const sgp_batch_opts* opts,
                             sgp_batch** out /* sgp_multi_size(m) entries */, sgp_summary* merged);
/* sgp_batch_write_files_layout over the shards of one sgp_app_run_multi call: the files are what a single batch holding all the
 * replicates would have written (packed: streams concatenated in index order, one .idx with run-wide offsets). */
sgp_status sgp_batches_write_files_layout(sgp_batch* const* shards, int n_shards, int layout, int64_t first_index);

/* Chained stage: BranchRelaxer over the trees of a tree-generator batch that is still resident on the
 * device — the in-memory equivalent of writing <prefix>.pruned.tree / .unpruned.tree per replicate and launching
 * `java ... BranchRelaxer <that file> <model> <args>` on each (apps/SaGePhy/BranchRelaxerParameters.java:80-117 reads the file,
 * io/NewickTreeReader.java:158-296 numbers its vertices in post-order). `src` must come from sgp_app_run with
 * SGP_FLAG_KEEP_TREES and live on the same context; which_tree = SGP_STREAM_PRUNED_TREE or SGP_STREAM_UNPRUNED_TREE.
 * argv is BranchRelaxer's own argv; its <tree> positional is only echoed in the "Arguments:" line of the .info stream.
 * Replicate i relaxes tree (i mod n_src) with the seed / Philox coordinates of replicate i, so n_replicates = k * n_src gives
 * k independent draws per tree. Source replicates that failed (or whose pruned tree is empty) yield SGP_REP_MODEL_ERROR. */
sgp_status sgp_relax_run_on_batch(sgp_context* ctx, sgp_batch* src, int which_tree, int argc, const char* const* argv,
                                  const sgp_batch_opts* opts, sgp_batch** out);

/* Chained stage: DomainTreeGen over a forest of gene trees that are the pruned trees [first_tree,
 * first_tree + n_trees) of a GuestTreeGen batch still resident on the device - the in-memory equivalent of writing those
 * <prefix>_<i>.pruned.tree files into a directory and launching `java ... DomainTreeGen <species tree> <that directory> ...`
 * (apps/SaGePhy/DomainTreeGenParameters.java:136-151 reads every file of the directory, io/NewickTreeReader.java:158-296 numbers the
 * vertices in post-order and captures HOST=, io/PrIMENewickTreeVerifier.java:237-270 turns lengths into vertex times, one
 * RBTreeEpochDiscretiser per gene tree: apps/SaGePhy/DomainTreeInHostTreeCreator.java:122-130). The vertex ids, lengths, names and
 * HOST tags are read from the kept vertex records by the ingest kernels (~25 bytes per vertex cross PCIe; no Newick text is written
 * or parsed); the epoch tables of the forest are then built like those of any host tree. Gene tree i is named
 * "<out prefix of src>_<name_first_index + i>.pruned.tree" (the file SGP_LAYOUT_PER_REPLICATE writes for it) and the forest is
 * ordered by name, as a directory listing is. argv is DomainTreeGen's own argv; its <guest tree directory> positional is only
 * echoed in the "Arguments:" line. `src` must come from sgp_app_run("GuestTreeGen" | "HostTreeGen", ...) with SGP_FLAG_KEEP_TREES. */
sgp_status sgp_domain_run_on_batch(sgp_context* ctx, sgp_batch* src, int64_t first_tree, int64_t n_trees, int64_t name_first_index,
                                   int argc, const char* const* argv, const sgp_batch_opts* opts, sgp_batch** out);

/* Batch accessors. Stream s is the concatenation over replicates, replicate i occupying bytes [offsets[i], offsets[i+1]). */
int64_t sgp_batch_n(const sgp_batch* b);
const char* sgp_batch_stream_data(sgp_batch* b, int stream);             /* host pointer (pinned), copies from HBM on first use */
const void* sgp_batch_stream_device_data(const sgp_batch* b, int stream);/* device pointer */
const uint64_t* sgp_batch_stream_offsets(sgp_batch* b, int stream);      /* n+1 entries, host */
/* Copies stream bytes / the n+1 offsets from HBM into a caller-owned host buffer (pinned memory gives full PCIe rate). */
sgp_status sgp_batch_copy_stream(sgp_batch* b, int stream, void* dst, uint64_t dst_capacity);
sgp_status sgp_batch_copy_offsets(sgp_batch* b, int stream, uint64_t* dst /* n+1 */);
/* Asynchronous variants: enqueue the copy on the context's copy stream so that it overlaps the kernels of the next
 * sgp_app_run on the same context (dst must be pinned). The batch must stay alive until sgp_context_wait_copies returns. */
sgp_status sgp_batch_copy_stream_async(sgp_batch* b, int stream, void* dst, uint64_t dst_capacity);
sgp_status sgp_batch_copy_offsets_async(sgp_batch* b, int stream, uint64_t* dst /* n+1 */);
sgp_status sgp_context_wait_copies(sgp_context* ctx);
uint32_t sgp_batch_stream_mask(const sgp_batch* b);                      /* bit s set = stream s was emitted */
/* BranchRelaxer: flat per-branch arrays (replicate-major, vertex-id order): which = 0 rates, 1 relaxed lengths. */
const double* sgp_batch_values(sgp_batch* b, int which, int64_t* count);
uint64_t sgp_batch_stream_bytes(const sgp_batch* b, int stream);
const char* sgp_batch_stream_suffix(const sgp_batch* b, int stream);     /* file suffix the reference uses, "" if unused */
const int32_t* sgp_batch_status(sgp_batch* b);                           /* n entries, sgp_rep_status */
const int32_t* sgp_batch_attempts(sgp_batch* b);                         /* n entries, value of "Attempts:" */
const int32_t* sgp_batch_sampled_leaves(sgp_batch* b);                   /* n entries */
const double* sgp_batch_root_times(sgp_batch* b);                        /* n entries: time of the pruned tree's root */
const double* sgp_batch_total_times(sgp_batch* b);                       /* n entries: total unpruned branch time (unrounded) */
const int32_t* sgp_batch_event_counts(sgp_batch* b);                     /* n x 17, unpruned trees */
const char* sgp_batch_stdout(const sgp_batch* b);                        /* what the reference prints to stdout (usage, -q trees) */
const char* sgp_batch_stderr(const sgp_batch* b);
const char* sgp_batch_out_prefix(const sgp_batch* b);                    /* the <out prefix> positional ("" with -q) */
void sgp_batch_summary(sgp_batch* b, sgp_summary* out);
void sgp_batch_timings(const sgp_batch* b, sgp_timings* out);
/* Writes the files the reference would write (n == 1: identical names; n > 1: one file per stream, replicates concatenated). */
sgp_status sgp_batch_write_files(sgp_batch* b);
/* Output layouts for n > 1 (downstream tools such as the reference's partial.py read one tree / leaf map per
 * file, partial.py:6-21,67-103):
 *   SGP_LAYOUT_PACKED        one file per stream, replicates concatenated, plus "<file>.idx": a "#sagephy-idx app=<App> n=<n> first=<k>" line,
 *                            then one line "offset<TAB>length<TAB>status" per replicate (tools/unpack_batch.py splits it again)
 *   SGP_LAYOUT_PER_REPLICATE replicate i written exactly as the reference run with output prefix "<prefix>_<i>" would write it
 *                            (first_index + i is used in the name; failed replicates write only their failure note). */
typedef enum { SGP_LAYOUT_PACKED = 0, SGP_LAYOUT_PER_REPLICATE = 1 } sgp_layout;
sgp_status sgp_batch_write_files_layout(sgp_batch* b, int layout, int64_t first_index);
void sgp_batch_free(sgp_batch* b);

/* Probes used by the parity tests: run the device implementations of the Java-exact primitives on `n` inputs. */
sgp_status sgp_probe_double_to_string(sgp_context* ctx, const double* v, int64_t n, char* out /* n*32 bytes, NUL padded */);
sgp_status sgp_probe_math(sgp_context* ctx, int fn /* 0 log, 1 log10, 2 exp, 3 round8, 4 normal quantile, 5 pow(x,0.37), 6 chi2 quantile k=4, 7 chi2 quantile k=0.25, 8 pow(e,x), 9 throughput-mode log on (0,1] */, const double* v, int64_t n, double* out);
/* MT19937 words of replicate `replicate_index` of a run seeded with the decimal integer `seed_decimal` (any size; the reference
 * parses it with java.math.BigInteger and keeps the low 16 bytes: math/PRNG.java:29-37,58-60; replicate i uses seed + i). */
sgp_status sgp_probe_mt(sgp_context* ctx, const char* seed_decimal, int64_t replicate_index, int64_t n_words, uint32_t* out_words);
/* Host-side seed logic (no GPU needed): the four big-endian key words MT19937 is seeded with for seed + replicate_index, and the
 * decimal text of seed + replicate_index as it appears in the "Arguments:" line of replay-mode .info output. */
sgp_status sgp_probe_seed(const char* seed_decimal, int64_t replicate_index, uint32_t key_out[4], char* decimal_out, int32_t decimal_cap);
/* Philox4x32-10 blocks for n (counter, key) pairs: variant 0 = key schedule computed per block, 1 = precomputed round keys (the
 * sampler's form). ctx == NULL runs the host build of the same source (no GPU needed). Pinned by the Random123 known answers. */
sgp_status sgp_probe_philox(sgp_context* ctx, int variant, const uint32_t* ctr /* n x 4 */, const uint32_t* key /* n x 2 */, int64_t n, uint32_t* out /* n x 4 */);
/* Issue-rate microbenchmarks (roofline denominators for the sampler), giga warp-instructions per second, whole chip:
 * independent DFMA chains, independent IMAD chains, and an IMAD + SHF + LOP3 mix (what Philox is made of). */
sgp_status sgp_probe_issue_peaks(sgp_context* ctx, double* dfma_ginst, double* imad_ginst, double* mixed_int_ginst);
/* Host-side argument handling of the three generators (no GPU needed): everything sgp_app_run does before the first kernel -
 * the reference's option grammar, number parsing, host / gene tree reading and the checks of its constructors
 * (apps/SaGePhy/{Host,Guest,Domain}TreeGen.java + *Parameters.java). Returns what sgp_app_run would return at that point
 * (SGP_OK, SGP_USAGE or the error) and copies the text the run would leave on stderr into `err`. With -hybrid this is the whole
 * run: the reference cannot create a guest vertex in that mode (GuestVertex.java:100), see csrc/hybrid_host.hpp. */
sgp_status sgp_probe_parse_app(const char* app, int argc, const char* const* argv, char* err, int32_t err_cap);
/* Host-side table preparation ("K0", no GPU needed): parses a Newick host tree as GuestTreeGen would (stem override as given,
 * NaN = keep the tree's own) and returns the tables the kernels consume, for CPU tests of the host logic.
 * Sizes: parent/left/right/v2e/vtime [nV], ep_lo/ep_up [nE], ep_off [nE+1], ep_arcs [ep_off[nE]], lca_time [nV*nV].
 * Any output pointer may be NULL; returns SGP_ERR_INVALID_ARGUMENT when a capacity (in elements) is too small. */
sgp_status sgp_probe_host_model(const char* newick, double stem, int32_t* n_vertices, int32_t* n_epochs,
                                int32_t* parent, int32_t* left, int32_t* right, int32_t* v2e, double* vtime, int32_t cap_v,
                                double* ep_lo, double* ep_up, int32_t* ep_off, int32_t cap_e, int32_t* ep_arcs, int32_t cap_arcs,
                                double* lca_time, int64_t cap_lca, char* error, int32_t error_cap);

#ifdef __cplusplus
}
#endif
#endif
