// Host-side engine interface (C++). The C ABI in capi.cpp is a thin shell over this.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include "host_model.hpp"

namespace sgp {

struct SimParamsHost {
    double lambda = 0, mu = 0, tau = 0, theta = 0, rho = 1, bias_rate = 1, gene_birth = 1;
    int bias = 0;             // 0 none, 1 simple, 2 exponential
    bool do_gene_birth = false, ishost = false;
    int min_leaves = 2, max_leaves = 1000, minper = 0, maxper = 10, max_attempts = 10000;
};

// A run's seed: any decimal integer (the reference parses it with java.math.BigInteger), see SeedSpec in rng.cuh.
struct SeedHost {
    unsigned long long lo = 0, hi = 0;        // low 128 bits, two's complement
    bool wide = false, negative = false;
    unsigned long long dec_lo19 = 0; std::string hs[2];
    unsigned long long philox_key = 0;
    std::string text;                          // canonical decimal text
};

struct TreeGenConfig {
    int app = 0;              // 0 HostTreeGen, 1 GuestTreeGen, 2 DomainTreeGen
    SimParamsHost P;
    HostModel host;                      // HostTreeGen: the 1- or 2-leaf start tree; GuestTreeGen: species tree; DomainTreeGen: species tree
    std::vector<HostModel> genes;        // DomainTreeGen: gene trees (sorted by file name)
    double inter_gene = 0.5, inter_species = 0.5; bool all_genes = false;
    bool quiet = false, exclude_meta = false, append_sigma = false, is_species = false;
    std::string vertex_prefix = "H", out_prefix;
    bool has_stem_override = false; double stem_override = 0;   // HostTreeGen -stem (applied to the finished trees)
    std::vector<int> sizes;
    bool has_seed = false; SeedHost seed;
    std::vector<std::string> args;      // the app's argv (for the "Arguments:" line)
    int seed_arg_index = -1;            // position of the seed value inside args
};

struct RelaxTreeTemplate {
    int nV = 0, root = -1, ignore = -1;
    std::vector<double> orig; std::vector<int16_t> chain; std::vector<uint8_t> flags; std::vector<std::string> names;
    std::vector<int32_t> topo, parent;    // RTree.getTopologicalOrdering() (BFS from the root) and parent ids, for the autocorrelated models
};

struct RelaxConfig {
    int model = 0; double pa = 0, pb = 0, pc = 0, pd = 0, sigma = 0;
    std::vector<double> samples;
    double min_rate = 1e-64, max_rate = 1e64; int max_attempts = 10000;
    bool do_meta = false, keep_interior = false, has_outfile = false; std::string outfile;
    std::vector<RelaxTreeTemplate> trees;
    bool has_seed = false; SeedHost seed;
    std::vector<std::string> args; int seed_arg_index = -1;
    std::string model_name; std::vector<std::pair<std::string, std::string>> params;   // in java.util.HashMap iteration order
    // IIDRK11 (apps/SaGePhy/IIDRasmussenKellis11RateModel.java): guest vertex x draws one Gamma(k, theta) rate per host arc its branch
    // passes over, segments rk_off[x] .. rk_off[x+1] in draw order with their (deterministic) weights; rate = pa * sum w * draw
    std::vector<int32_t> rk_off; std::vector<double> rk_k, rk_theta, rk_w;
};

struct BatchOpts {
    long long n = 1, offset = 0;
    int rng_mode = 0;         // 0 Philox, 1 MT replay
    unsigned stream_mask = 0;
    bool keep_on_device = false;
    bool multi_tree_input = false;
    bool keep_trees = false;  // tree generators: leave the vertex records on the device for a chained stage
    bool serial = false;      // never pipeline (also the fallback when an output estimate of a pipelined run was too small)
    bool pipeline = false;    // opt-in: sampler of sub-batch k+1 on a second stream next to build / label / emit of sub-batch k
};

struct Timings { double find_ms = 0, build_ms = 0, label_ms = 0, emit_size_ms = 0, emit_write_ms = 0, d2h_ms = 0, total_ms = 0; long long launches = 0, h2d_bytes = 0; int sub_batches = 1; };

struct Summary {
    long long n_ok = 0, n_failed = 0, attempts_total = 0, vertices_sampled = 0, vertices_abandoned = 0, attempts_completed = 0;
    long long event_counts[17] = {0};
    long long leaf_hist[1025] = {0};
    long long out_bytes[8] = {0};
};

// Result of a run. Stream bytes live in device memory; host copies are made on demand into pinned memory.
class Batch {
public:
    Batch();
    ~Batch();
    long long n = 0;
    int device = 0;
    void* stream = nullptr;   // cudaStream_t the device buffers were allocated on
    int app = 0;
    std::string out_text, err_text, out_prefix;
    bool quiet = false;
    const char* suffix[8] = {"", "", "", "", "", "", "", ""};
    // device
    char* d_stream[8] = {nullptr}; unsigned long long stream_bytes[8] = {0};
    unsigned long long* d_offsets = nullptr;      // [8][n+1]
    long long offsets_stride = 0;
    void* d_info = nullptr;                       // RepInfo[n]
    // kept only with BatchOpts::keep_trees: vertex records, order arrays (4 x aux_stride), first record of each replicate
    void* d_nodes = nullptr; int32_t* d_aux = nullptr; long long* d_node_off = nullptr; long long aux_stride = 0;
    std::string name_prefix; bool name_append_sigma = false, name_exclude_meta = false;
    unsigned stream_mask = 0;
    double* d_values[2] = {nullptr, nullptr}; long long n_values = 0;   // BranchRelaxer: rates, relaxed lengths (flat, replicate-major)
    std::vector<double> h_values[2];
    const double* values_host(int which);
    // host (lazy)
    char* h_stream[8] = {nullptr};
    std::vector<unsigned long long> h_offsets;    // [8][n+1]
    std::vector<int32_t> status, attempts, sampled, events; std::vector<double> root_time, total_time;
    bool have_info = false, have_stats = false;
    Summary summary; Timings timings;
    const char* stream_host(int s);
    bool copy_stream_to(int s, void* dst);
    bool copy_stream_to_async(int s, void* dst, void* copy_stream);
    bool copy_offsets_to_async(int s, unsigned long long* dst, void* copy_stream);
    bool copy_offsets_to(int s, unsigned long long* dst);
    void fetch_info();        // status, attempts, sampled leaves
    void fetch_stats();       // tree times, event counts
    void fetch_offsets();
};

class Context {
public:
    explicit Context(int device);
    ~Context();
    int device;
    std::string last_error;
    struct Impl; std::unique_ptr<Impl> impl;
    void run_treegen(const TreeGenConfig& cfg, const BatchOpts& opts, Batch& out);
    // src != nullptr: chained run over the trees of a tree-generator batch (pruned or unpruned view), cfg.trees is ignored
    void run_relax(const RelaxConfig& cfg, const BatchOpts& opts, Batch& out, const Batch* src = nullptr, bool src_pruned = true);
    // Chained stage: the trees [first, first + count) of a generator batch kept with keep_trees, as parsed Newick trees would look
    // (post-order ids, lengths, names, HOST tags) - read from the vertex records on the device, never printed or parsed.
    std::vector<FlatTree> ingest_trees(const Batch& src, long long first, long long count, bool pruned);
    void* copy_stream();          // second CUDA stream for device->host copies that overlap the next batch's kernels
    void wait_copies();
    void probe_d2s(const double* v, long long n, char* out);
    void probe_math(int fn, const double* v, long long n, double* out);
    void probe_mt(const SeedHost& seed, long long offset, long long n_words, uint32_t* out);
    void probe_issue(double* dfma, double* imad, double* ddep);
    void probe_philox(int variant, const uint32_t* ctr, const uint32_t* key, long long n, uint32_t* out);
};

// tables.cpp: the device's seed arithmetic (rng.cuh) compiled for the host: MT key words and decimal text of seed + i
void host_seed_probe(const SeedHost& h, unsigned long long i, uint32_t key[4], std::string& text);

void host_philox_probe(int variant, const uint32_t* ctr, const uint32_t* key, long long n, uint32_t* out);

// apps.cpp: reference CLI grammar -> configs
struct AppResult { int status = 0; std::string out_text, err_text; };   // status: 0 run, 6 usage, other = error code
// forest != nullptr (DomainTreeGen, chained stage): gene trees and their names come from memory, the <guest tree directory> positional is only echoed
void parse_treegen_app(const std::string& app, const std::vector<std::string>& args, TreeGenConfig& cfg, AppResult& res, const std::vector<std::pair<std::string, FlatTree>>* forest = nullptr);
void parse_relax_app(const std::vector<std::string>& args, bool multi_tree_input, RelaxConfig& cfg, AppResult& res, bool trees_from_batch = false);

}  // namespace sgp
