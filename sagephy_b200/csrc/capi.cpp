// C ABI (include/sagephy_b200.h) over the C++ engine.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <thread>
#include "../../include/sagephy_b200.h"
#include "engine.hpp"
#include "apps_common.hpp"

using namespace sgp;

struct sgp_context { std::unique_ptr<Context> ctx; std::string last_error; };
struct sgp_batch { Batch b; std::string stdout_cache; bool stdout_ready = false; sgp_context* owner = nullptr; };

namespace { thread_local std::string g_create_error; }

extern "C" {

sgp_status sgp_context_create(int device, sgp_context** out) {
    if (!out) return SGP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    try {
        auto* c = new sgp_context();
        c->ctx.reset(new Context(device));
        *out = c;
        return SGP_OK;
    } catch (std::exception& e) {
        g_create_error = e.what();
        std::fprintf(stderr, "sagephy_b200: %s\n", e.what());
        return SGP_ERR_NO_DEVICE;
    }
}
void sgp_context_destroy(sgp_context* ctx) { delete ctx; }
const char* sgp_last_error(const sgp_context* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

sgp_status sgp_app_run(sgp_context* ctx, const char* app, int argc, const char* const* argv, const sgp_batch_opts* opts, sgp_batch** out) {
    if (!ctx || !app || !out || argc < 0) return SGP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    sgp_batch_opts o{}; if (opts) o = *opts; if (o.n_replicates <= 0) o.n_replicates = 1;
    std::vector<std::string> args; for (int i = 0; i < argc; ++i) args.emplace_back(argv[i] ? argv[i] : "");
    auto* b = new sgp_batch(); b->owner = ctx;
    try {
        if (std::string(app) == "BranchRelaxer") {
            RelaxConfig rc; AppResult rr;
            parse_relax_app(args, (o.flags & SGP_FLAG_MULTI_TREE_INPUT) != 0, rc, rr);
            b->b.out_text = rr.out_text; b->b.err_text = rr.err_text;
            if (rr.status == 6) { *out = b; return SGP_USAGE; }
            if (rr.status != 0) { ctx->last_error = rr.err_text; *out = b; return (sgp_status)rr.status; }
            BatchOpts bo; bo.n = o.n_replicates; bo.offset = o.replicate_offset; bo.rng_mode = o.rng_mode; bo.stream_mask = o.stream_mask; bo.keep_on_device = o.keep_on_device != 0;
            ctx->ctx->run_relax(rc, bo, b->b);
            *out = b;
            return SGP_OK;
        }
        TreeGenConfig cfg; AppResult res;
        parse_treegen_app(app, args, cfg, res);
        b->b.out_text = res.out_text; b->b.err_text = res.err_text;
        if (res.status == 6) { *out = b; return SGP_USAGE; }
        if (res.status != 0) { ctx->last_error = res.err_text; *out = b; return (sgp_status)res.status; }
        if (o.rng_mode == SGP_RNG_MT_REPLAY && !cfg.has_seed) {
            // the reference would seed from /dev/random; replay of an unknown seed is meaningless but harmless
        }
        BatchOpts bo; bo.n = o.n_replicates; bo.offset = o.replicate_offset; bo.rng_mode = o.rng_mode; bo.stream_mask = o.stream_mask; bo.keep_on_device = o.keep_on_device != 0; bo.keep_trees = (o.flags & SGP_FLAG_KEEP_TREES) != 0; bo.serial = (o.flags & SGP_FLAG_SERIAL) != 0; bo.pipeline = (o.flags & SGP_FLAG_PIPELINE) != 0;
        ctx->ctx->run_treegen(cfg, bo, b->b);
        *out = b;
        return SGP_OK;
    } catch (SgpError& e) {
        ctx->last_error = e.what();
        b->b.err_text += std::string(e.what()) + "\n\nUse option -h or --help to show usage.\n";
        *out = b;
        const std::string w = e.what();
        if (w.find("CUDA") != std::string::npos) return SGP_ERR_NO_DEVICE;
        if (w.find("Newick") != std::string::npos || w.find("NumberFormat") != std::string::npos) return SGP_ERR_PARSE;
        return SGP_ERR_INVALID_ARGUMENT;
    } catch (std::exception& e) {
        ctx->last_error = e.what(); b->b.err_text += e.what(); *out = b;
        return SGP_ERR_INTERNAL;
    }
}

sgp_status sgp_relax_run_on_batch(sgp_context* ctx, sgp_batch* src, int which_tree, int argc, const char* const* argv, const sgp_batch_opts* opts, sgp_batch** out) {
    if (!ctx || !src || !out || argc < 0) return SGP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (which_tree != SGP_STREAM_PRUNED_TREE && which_tree != SGP_STREAM_UNPRUNED_TREE) { ctx->last_error = "which_tree must be SGP_STREAM_PRUNED_TREE or SGP_STREAM_UNPRUNED_TREE"; return SGP_ERR_INVALID_ARGUMENT; }
    if (src->owner != ctx || !src->b.d_nodes) { ctx->last_error = "source batch holds no trees on this context (run the generator with SGP_FLAG_KEEP_TREES)"; return SGP_ERR_INVALID_ARGUMENT; }
    sgp_batch_opts o{}; if (opts) o = *opts; if (o.n_replicates <= 0) o.n_replicates = src->b.n;
    std::vector<std::string> args; for (int i = 0; i < argc; ++i) args.emplace_back(argv[i] ? argv[i] : "");
    auto* b = new sgp_batch(); b->owner = ctx;
    try {
        RelaxConfig rc; AppResult rr;
        parse_relax_app(args, false, rc, rr, true);
        b->b.out_text = rr.out_text; b->b.err_text = rr.err_text;
        if (rr.status == 6) { *out = b; return SGP_USAGE; }
        if (rr.status != 0) { ctx->last_error = rr.err_text; *out = b; return (sgp_status)rr.status; }
        BatchOpts bo; bo.n = o.n_replicates; bo.offset = o.replicate_offset; bo.rng_mode = o.rng_mode; bo.stream_mask = o.stream_mask; bo.keep_on_device = o.keep_on_device != 0;
        ctx->ctx->run_relax(rc, bo, b->b, &src->b, which_tree == SGP_STREAM_PRUNED_TREE);
        *out = b;
        return SGP_OK;
    } catch (SgpError& e) {
        ctx->last_error = e.what(); b->b.err_text += std::string(e.what()) + "\n"; *out = b;
        return std::string(e.what()).find("CUDA") != std::string::npos ? SGP_ERR_NO_DEVICE : SGP_ERR_INVALID_ARGUMENT;
    } catch (std::exception& e) {
        ctx->last_error = e.what(); b->b.err_text += e.what(); *out = b;
        return SGP_ERR_INTERNAL;
    }
}

sgp_status sgp_domain_run_on_batch(sgp_context* ctx, sgp_batch* src, int64_t first_tree, int64_t n_trees, int64_t name_first_index, int argc, const char* const* argv, const sgp_batch_opts* opts, sgp_batch** out) {
    if (!ctx || !src || !out || argc < 0) return SGP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (src->owner != ctx || !src->b.d_nodes) { ctx->last_error = "source batch holds no trees on this context (run the generator with SGP_FLAG_KEEP_TREES)"; return SGP_ERR_INVALID_ARGUMENT; }
    sgp_batch_opts o{}; if (opts) o = *opts; if (o.n_replicates <= 0) o.n_replicates = 1;
    std::vector<std::string> args; for (int i = 0; i < argc; ++i) args.emplace_back(argv[i] ? argv[i] : "");
    auto* b = new sgp_batch(); b->owner = ctx;
    try {
        std::vector<FlatTree> trees = ctx->ctx->ingest_trees(src->b, first_tree, n_trees, true);
        // gene tree i carries the name SGP_LAYOUT_PER_REPLICATE gives its file (the GUEST= value and the forest order follow from it)
        std::vector<std::pair<std::string, FlatTree>> forest;
        const std::string base = cli::file_name(src->b.out_prefix);
        for (size_t i = 0; i < trees.size(); ++i) forest.emplace_back(base + "_" + std::to_string(name_first_index + (int64_t)i) + ".pruned.tree", std::move(trees[i]));
        TreeGenConfig cfg; AppResult res;
        parse_treegen_app("DomainTreeGen", args, cfg, res, &forest);
        b->b.out_text = res.out_text; b->b.err_text = res.err_text;
        if (res.status == 6) { *out = b; return SGP_USAGE; }
        if (res.status != 0) { ctx->last_error = res.err_text; *out = b; return (sgp_status)res.status; }
        BatchOpts bo; bo.n = o.n_replicates; bo.offset = o.replicate_offset; bo.rng_mode = o.rng_mode; bo.stream_mask = o.stream_mask; bo.keep_on_device = o.keep_on_device != 0; bo.keep_trees = (o.flags & SGP_FLAG_KEEP_TREES) != 0;
        ctx->ctx->run_treegen(cfg, bo, b->b);
        *out = b;
        return SGP_OK;
    } catch (SgpError& e) {
        ctx->last_error = e.what(); b->b.err_text += std::string(e.what()) + "\n"; *out = b;
        return std::string(e.what()).find("CUDA") != std::string::npos ? SGP_ERR_NO_DEVICE : SGP_ERR_INVALID_ARGUMENT;
    } catch (std::exception& e) {
        ctx->last_error = e.what(); b->b.err_text += e.what(); *out = b;
        return SGP_ERR_INTERNAL;
    }
}

// ---- multi-GPU front end ---------------------------------------------------------------------------------------------------------
struct sgp_multi { std::vector<sgp_context*> ctx; };

sgp_status sgp_multi_create(const int* devices, int n_devices, sgp_multi** out) {
    if (!devices || n_devices <= 0 || !out) return SGP_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    auto* m = new sgp_multi();
    for (int i = 0; i < n_devices; ++i) {
        sgp_context* c = nullptr;
        const sgp_status st = sgp_context_create(devices[i], &c);
        if (st != SGP_OK) { for (auto* x : m->ctx) sgp_context_destroy(x); delete m; return st; }
        m->ctx.push_back(c);
    }
    *out = m;
    return SGP_OK;
}
int sgp_multi_size(const sgp_multi* m) { return m ? (int)m->ctx.size() : 0; }
sgp_context* sgp_multi_context(sgp_multi* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[(size_t)i] : nullptr; }
void sgp_multi_destroy(sgp_multi* m) { if (!m) return; for (auto* x : m->ctx) sgp_context_destroy(x); delete m; }

sgp_status sgp_app_run_multi(sgp_multi* m, const char* app, int argc, const char* const* argv, const sgp_batch_opts* opts, sgp_batch** out, sgp_summary* merged) {
    if (!m || m->ctx.empty() || !app || !out || argc < 0) return SGP_ERR_INVALID_ARGUMENT;
    const int G = (int)m->ctx.size();
    sgp_batch_opts o{}; if (opts) o = *opts; if (o.n_replicates <= 0) o.n_replicates = 1;
    std::vector<sgp_status> st((size_t)G, SGP_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) out[g] = nullptr;
    for (int g = 0; g < G; ++g) {
        const int64_t lo = o.n_replicates * g / G, hi = o.n_replicates * (g + 1) / G;
        if (hi <= lo) continue;                      // more devices than replicates
        th.emplace_back([&, g, lo, hi]() {
            sgp_batch_opts og = o; og.n_replicates = hi - lo; og.replicate_offset = o.replicate_offset + lo;
            st[(size_t)g] = sgp_app_run(m->ctx[(size_t)g], app, argc, argv, &og, &out[g]);
        });
    }
    for (auto& t : th) t.join();
    sgp_status rc = SGP_OK;
    for (int g = 0; g < G; ++g) if (st[(size_t)g] != SGP_OK && rc == SGP_OK) rc = st[(size_t)g];
    if (merged) {
        std::memset(merged, 0, sizeof *merged);
        for (int g = 0; g < G; ++g) {
            if (!out[g] || st[(size_t)g] != SGP_OK) continue;
            sgp_summary s; sgp_batch_summary(out[g], &s);
            merged->n_ok += s.n_ok; merged->n_failed += s.n_failed; merged->attempts_total += s.attempts_total; merged->vertices_sampled += s.vertices_sampled;
            merged->vertices_abandoned += s.vertices_abandoned; merged->attempts_completed += s.attempts_completed;
            for (int i = 0; i < 17; ++i) merged->event_counts[i] += s.event_counts[i];
            for (int i = 0; i < SGP_HIST_LEAF_BINS; ++i) merged->leaf_hist[i] += s.leaf_hist[i];
            for (int i = 0; i < 8; ++i) merged->out_bytes[i] += s.out_bytes[i];
        }
    }
    return rc;
}

sgp_status sgp_batches_write_files_layout(sgp_batch* const* shards, int n_shards, int layout, int64_t first_index) {
    if (!shards || n_shards <= 0 || (layout != SGP_LAYOUT_PACKED && layout != SGP_LAYOUT_PER_REPLICATE)) return SGP_ERR_INVALID_ARGUMENT;
    std::vector<sgp_batch*> bs; for (int i = 0; i < n_shards; ++i) if (shards[i]) bs.push_back(shards[i]);
    if (bs.empty()) return SGP_ERR_INVALID_ARGUMENT;
    if (layout == SGP_LAYOUT_PER_REPLICATE) {
        int64_t first = first_index;
        for (auto* b : bs) { const sgp_status st = sgp_batch_write_files_layout(b, layout, first); if (st != SGP_OK) return st; first += b->b.n; }
        return SGP_OK;
    }
    sgp_batch* b0 = bs[0];
    if (b0->b.quiet || !b0->b.d_offsets) return SGP_OK;
    long long n_total = 0; for (auto* b : bs) { b->b.fetch_info(); b->b.fetch_offsets(); n_total += b->b.n; }
    static const char* kApps[] = {"HostTreeGen", "GuestTreeGen", "DomainTreeGen", "BranchRelaxer"};
    for (int s = 0; s < 8; ++s) {
        if (!(b0->b.stream_mask & (1u << s))) continue;
        const std::string sfx = b0->b.suffix[s];
        std::ofstream f(b0->b.out_prefix + sfx, std::ios::binary), idx(b0->b.out_prefix + sfx + ".idx", std::ios::binary);
        if (!f || !idx) return SGP_ERR_INVALID_ARGUMENT;
        idx << "#sagephy-idx app=" << kApps[b0->b.app & 3] << " n=" << n_total << " first=" << first_index << '\n';
        unsigned long long base = 0;
        for (auto* b : bs) {
            const char* data = b->b.stream_host(s);
            if (!data && b->b.stream_bytes[s]) return SGP_ERR_INTERNAL;
            f.write(data, (std::streamsize)b->b.stream_bytes[s]);
            const unsigned long long* off = b->b.h_offsets.data() + (size_t)s * (b->b.n + 1);
            for (long long i = 0; i < b->b.n; ++i) idx << (base + off[i]) << '\t' << (off[i + 1] - off[i]) << '\t' << b->b.status[i] << '\n';
            base += b->b.stream_bytes[s];
        }
    }
    return SGP_OK;
}

int64_t sgp_batch_n(const sgp_batch* b) { return b ? b->b.n : 0; }
const char* sgp_batch_stream_data(sgp_batch* b, int s) { return b ? b->b.stream_host(s) : nullptr; }
const void* sgp_batch_stream_device_data(const sgp_batch* b, int s) { return (b && s >= 0 && s < 8) ? b->b.d_stream[s] : nullptr; }
const uint64_t* sgp_batch_stream_offsets(sgp_batch* b, int s) {
    if (!b || s < 0 || s >= 8 || !b->b.d_offsets) return nullptr;
    b->b.fetch_offsets();
    return (const uint64_t*)(b->b.h_offsets.data() + (size_t)s * (b->b.n + 1));
}
sgp_status sgp_batch_copy_stream(sgp_batch* b, int s, void* dst, uint64_t cap) {
    if (!b || s < 0 || s >= 8 || !dst) return SGP_ERR_INVALID_ARGUMENT;
    if (cap < b->b.stream_bytes[s]) return SGP_ERR_INVALID_ARGUMENT;
    return b->b.copy_stream_to(s, dst) ? SGP_OK : SGP_ERR_NO_DEVICE;
}
sgp_status sgp_batch_copy_offsets(sgp_batch* b, int s, uint64_t* dst) {
    if (!b || s < 0 || s >= 8 || !dst || !b->b.d_offsets) return SGP_ERR_INVALID_ARGUMENT;
    return b->b.copy_offsets_to(s, (unsigned long long*)dst) ? SGP_OK : SGP_ERR_NO_DEVICE;
}
sgp_status sgp_batch_copy_stream_async(sgp_batch* b, int s, void* dst, uint64_t cap) {
    if (!b || !b->owner || s < 0 || s >= 8 || !dst || cap < b->b.stream_bytes[s]) return SGP_ERR_INVALID_ARGUMENT;
    return b->b.copy_stream_to_async(s, dst, b->owner->ctx->copy_stream()) ? SGP_OK : SGP_ERR_NO_DEVICE;
}
sgp_status sgp_batch_copy_offsets_async(sgp_batch* b, int s, uint64_t* dst) {
    if (!b || !b->owner || s < 0 || s >= 8 || !dst || !b->b.d_offsets) return SGP_ERR_INVALID_ARGUMENT;
    return b->b.copy_offsets_to_async(s, (unsigned long long*)dst, b->owner->ctx->copy_stream()) ? SGP_OK : SGP_ERR_NO_DEVICE;
}
sgp_status sgp_context_wait_copies(sgp_context* ctx) { if (!ctx) return SGP_ERR_INVALID_ARGUMENT; ctx->ctx->wait_copies(); return SGP_OK; }
uint32_t sgp_batch_stream_mask(const sgp_batch* b) { return b ? b->b.stream_mask : 0u; }
const double* sgp_batch_values(sgp_batch* b, int which, int64_t* count) {
    if (!b) return nullptr;
    if (count) *count = b->b.n_values;
    return b->b.values_host(which);
}
uint64_t sgp_batch_stream_bytes(const sgp_batch* b, int s) { return (b && s >= 0 && s < 8) ? b->b.stream_bytes[s] : 0; }
const char* sgp_batch_stream_suffix(const sgp_batch* b, int s) { return (b && s >= 0 && s < 8) ? b->b.suffix[s] : ""; }
const int32_t* sgp_batch_status(sgp_batch* b) { if (!b) return nullptr; b->b.fetch_info(); return b->b.status.data(); }
const int32_t* sgp_batch_attempts(sgp_batch* b) { if (!b) return nullptr; b->b.fetch_info(); return b->b.attempts.data(); }
const int32_t* sgp_batch_sampled_leaves(sgp_batch* b) { if (!b) return nullptr; b->b.fetch_info(); if (b->b.sampled.size() != (size_t)b->b.n) b->b.sampled.assign((size_t)b->b.n, 0); return b->b.sampled.data(); }
const double* sgp_batch_root_times(sgp_batch* b) { if (!b) return nullptr; b->b.fetch_stats(); return b->b.root_time.data(); }
const double* sgp_batch_total_times(sgp_batch* b) { if (!b) return nullptr; b->b.fetch_stats(); return b->b.total_time.data(); }
const int32_t* sgp_batch_event_counts(sgp_batch* b) { if (!b) return nullptr; b->b.fetch_stats(); return b->b.events.data(); }

const char* sgp_batch_stdout(const sgp_batch* cb) {
    sgp_batch* b = const_cast<sgp_batch*>(cb);
    if (!b) return "";
    if (!b->stdout_ready) {
        b->stdout_cache = b->b.out_text;
        const int qs = b->b.app == 3 ? 0 : 1;      // BranchRelaxer prints the relaxed tree, the generators the pruned tree
        if (b->b.quiet && b->b.d_stream[qs]) { const char* t = b->b.stream_host(qs); if (t) b->stdout_cache.append(t, b->b.stream_bytes[qs]); }
        b->stdout_ready = true;
    }
    return b->stdout_cache.c_str();
}
const char* sgp_batch_stderr(const sgp_batch* b) { return b ? b->b.err_text.c_str() : ""; }
const char* sgp_batch_out_prefix(const sgp_batch* b) { return b ? b->b.out_prefix.c_str() : ""; }

void sgp_batch_summary(sgp_batch* b, sgp_summary* out) {
    if (!b || !out) return;
    const Summary& s = b->b.summary;
    out->n_ok = s.n_ok; out->n_failed = s.n_failed; out->attempts_total = s.attempts_total; out->vertices_sampled = s.vertices_sampled; out->vertices_abandoned = s.vertices_abandoned; out->attempts_completed = s.attempts_completed;
    for (int i = 0; i < 17; ++i) out->event_counts[i] = s.event_counts[i];
    for (int i = 0; i < SGP_HIST_LEAF_BINS; ++i) out->leaf_hist[i] = s.leaf_hist[i];
    for (int i = 0; i < 8; ++i) out->out_bytes[i] = s.out_bytes[i];
}
void sgp_batch_timings(const sgp_batch* b, sgp_timings* out) {
    if (!b || !out) return;
    const Timings& t = b->b.timings;
    out->find_ms = t.find_ms; out->build_ms = t.build_ms; out->label_ms = t.label_ms; out->emit_size_ms = t.emit_size_ms; out->emit_write_ms = t.emit_write_ms;
    out->d2h_ms = t.d2h_ms; out->total_ms = t.total_ms; out->kernel_launches = t.launches; out->h2d_bytes = t.h2d_bytes; out->sub_batches = t.sub_batches;
}

sgp_status sgp_batch_write_files(sgp_batch* b) {
    if (!b) return SGP_ERR_INVALID_ARGUMENT;
    if (b->b.quiet || !b->b.d_offsets) return SGP_OK;
    b->b.fetch_info();
    bool all_failed = true; for (long long i = 0; i < b->b.n; ++i) if (b->b.status[i] == 0) all_failed = false;
    for (int s = 0; s < 8; ++s) {
        if (!(b->b.stream_mask & (1u << s))) continue;
        std::string sfx = b->b.suffix[s];
        if (all_failed) {
            // the reference writes only the failure note: <prefix>.info (HostTreeGen.java:65) / <prefix>.pruned.info (GuestTreeGen.java:50)
            if (s != SGP_STREAM_PRUNED_INFO) continue;
            if (b->b.app == 0) sfx = ".info";
        }
        const char* data = b->b.stream_host(s);
        if (!data) return SGP_ERR_INTERNAL;
        std::ofstream f(b->b.out_prefix + sfx, std::ios::binary);
        if (!f) return SGP_ERR_INVALID_ARGUMENT;
        f.write(data, (std::streamsize)b->b.stream_bytes[s]);
    }
    return SGP_OK;
}
sgp_status sgp_batch_write_files_layout(sgp_batch* b, int layout, int64_t first_index) {
    if (!b || (layout != SGP_LAYOUT_PACKED && layout != SGP_LAYOUT_PER_REPLICATE)) return SGP_ERR_INVALID_ARGUMENT;
    if (b->b.quiet || !b->b.d_offsets) return SGP_OK;
    b->b.fetch_info(); b->b.fetch_offsets();
    const long long n = b->b.n;
    for (int s = 0; s < 8; ++s) {
        if (!(b->b.stream_mask & (1u << s))) continue;
        const char* data = b->b.stream_host(s);
        if (!data && b->b.stream_bytes[s]) return SGP_ERR_INTERNAL;
        const unsigned long long* off = b->b.h_offsets.data() + (size_t)s * (n + 1);
        const std::string sfx = b->b.suffix[s];
        if (layout == SGP_LAYOUT_PACKED) {
            std::ofstream f(b->b.out_prefix + sfx, std::ios::binary), idx(b->b.out_prefix + sfx + ".idx", std::ios::binary);
            if (!f || !idx) return SGP_ERR_INVALID_ARGUMENT;
            f.write(data, (std::streamsize)b->b.stream_bytes[s]);
            static const char* kApps[] = {"HostTreeGen", "GuestTreeGen", "DomainTreeGen", "BranchRelaxer"};
            idx << "#sagephy-idx app=" << kApps[b->b.app & 3] << " n=" << n << " first=" << first_index << '\n';
            for (long long i = 0; i < n; ++i) idx << off[i] << '\t' << (off[i + 1] - off[i]) << '\t' << b->b.status[i] << '\n';
        } else {
            for (long long i = 0; i < n; ++i) {
                std::string name = sfx;
                if (b->b.status[i] != 0) {
                    // a failed reference run leaves only the failure note (HostTreeGen.java:65: <prefix>.info; GuestTreeGen.java:50: <prefix>.pruned.info)
                    if (s != SGP_STREAM_PRUNED_INFO || off[i + 1] == off[i]) continue;
                    if (b->b.app == 0) name = ".info";
                }
                std::ofstream f(b->b.out_prefix + "_" + std::to_string(first_index + i) + name, std::ios::binary);
                if (!f) return SGP_ERR_INVALID_ARGUMENT;
                f.write(data + off[i], (std::streamsize)(off[i + 1] - off[i]));
            }
        }
    }
    return SGP_OK;
}
void sgp_batch_free(sgp_batch* b) { delete b; }

sgp_status sgp_probe_double_to_string(sgp_context* ctx, const double* v, int64_t n, char* out) {
    if (!ctx) return SGP_ERR_INVALID_ARGUMENT;
    try { ctx->ctx->probe_d2s(v, n, out); return SGP_OK; } catch (std::exception& e) { ctx->last_error = e.what(); return SGP_ERR_NO_DEVICE; }
}
sgp_status sgp_probe_math(sgp_context* ctx, int fn, const double* v, int64_t n, double* out) {
    if (!ctx) return SGP_ERR_INVALID_ARGUMENT;
    try { ctx->ctx->probe_math(fn, v, n, out); return SGP_OK; } catch (std::exception& e) { ctx->last_error = e.what(); return SGP_ERR_NO_DEVICE; }
}
sgp_status sgp_probe_mt(sgp_context* ctx, const char* seed, int64_t rep, int64_t n_words, uint32_t* out) {
    if (!ctx || !seed) return SGP_ERR_INVALID_ARGUMENT;
    try { ctx->ctx->probe_mt(cli::parse_seed(seed), rep, n_words, out); return SGP_OK; }
    catch (SgpError& e) { ctx->last_error = e.what(); return SGP_ERR_PARSE; }
    catch (std::exception& e) { ctx->last_error = e.what(); return SGP_ERR_NO_DEVICE; }
}
sgp_status sgp_probe_seed(const char* seed, int64_t rep, uint32_t key_out[4], char* dec, int32_t cap) {
    if (!seed || rep < 0) return SGP_ERR_INVALID_ARGUMENT;
    try {
        const SeedHost h = cli::parse_seed(seed);
        std::string text;
        host_seed_probe(h, (unsigned long long)rep, key_out, text);
        if (dec) { if ((int)text.size() + 1 > cap) return SGP_ERR_INVALID_ARGUMENT; std::memcpy(dec, text.c_str(), text.size() + 1); }
        return SGP_OK;
    } catch (std::exception&) { return SGP_ERR_PARSE; }
}
sgp_status sgp_probe_philox(sgp_context* ctx, int variant, const uint32_t* ctr, const uint32_t* key, int64_t n, uint32_t* out) {
    if (!ctr || !key || !out || n < 0) return SGP_ERR_INVALID_ARGUMENT;
    if (!ctx) { host_philox_probe(variant, ctr, key, n, out); return SGP_OK; }
    try { ctx->ctx->probe_philox(variant, ctr, key, n, out); return SGP_OK; } catch (std::exception& e) { ctx->last_error = e.what(); return SGP_ERR_NO_DEVICE; }
}
sgp_status sgp_probe_issue_peaks(sgp_context* ctx, double* dfma, double* imad, double* ddep) {
    if (!ctx) return SGP_ERR_INVALID_ARGUMENT;
    try { ctx->ctx->probe_issue(dfma, imad, ddep); return SGP_OK; } catch (std::exception& e) { ctx->last_error = e.what(); return SGP_ERR_NO_DEVICE; }
}

sgp_status sgp_probe_parse_app(const char* app, int argc, const char* const* argv, char* err, int32_t err_cap) {
    if (!app || argc < 0) return SGP_ERR_INVALID_ARGUMENT;
    std::vector<std::string> args; for (int i = 0; i < argc; ++i) args.emplace_back(argv[i] ? argv[i] : "");
    std::string text; sgp_status st = SGP_OK;
    try {
        TreeGenConfig cfg; AppResult res;
        parse_treegen_app(app, args, cfg, res);
        text = res.err_text;
        st = res.status == 6 ? SGP_USAGE : (sgp_status)res.status;
    } catch (SgpError& e) {
        text = std::string(e.what()) + "\n\nUse option -h or --help to show usage.\n";
        const std::string w = e.what();
        st = (w.find("Newick") != std::string::npos || w.find("NumberFormat") != std::string::npos) ? SGP_ERR_PARSE : SGP_ERR_INVALID_ARGUMENT;
    } catch (std::exception& e) { text = e.what(); st = SGP_ERR_INVALID_ARGUMENT; }
    if (err && err_cap > 0) std::snprintf(err, (size_t)err_cap, "%s", text.c_str());
    return st;
}

sgp_status sgp_probe_host_model(const char* newick, double stem, int32_t* n_vertices, int32_t* n_epochs,
                                int32_t* parent, int32_t* left, int32_t* right, int32_t* v2e, double* vtime, int32_t cap_v,
                                double* ep_lo, double* ep_up, int32_t* ep_off, int32_t cap_e, int32_t* ep_arcs, int32_t cap_arcs,
                                double* lca_time, int64_t cap_lca, char* error, int32_t error_cap) {
    if (!newick) return SGP_ERR_INVALID_ARGUMENT;
    try {
        HostModel h;
        const bool keep = stem != stem;
        h.build(parse_newick(newick), nullptr, keep ? nullptr : &stem, lca_time != nullptr);
        if (n_vertices) *n_vertices = h.nV;
        if (n_epochs) *n_epochs = h.nE;
        auto copy = [](auto* dst, const auto& src, long long cap) { if (!dst) return true; if ((long long)src.size() > cap) return false; for (size_t i = 0; i < src.size(); ++i) dst[i] = src[i]; return true; };
        bool ok = copy(parent, h.parent, cap_v) && copy(left, h.left, cap_v) && copy(right, h.right, cap_v) && copy(v2e, h.v2e, cap_v) && copy(vtime, h.vtime, cap_v);
        ok = ok && copy(ep_lo, h.ep_lo, cap_e) && copy(ep_up, h.ep_up, cap_e) && copy(ep_off, h.ep_off, (long long)cap_e + 1) && copy(ep_arcs, h.ep_arcs, cap_arcs) && copy(lca_time, h.lca_time, cap_lca);
        return ok ? SGP_OK : SGP_ERR_INVALID_ARGUMENT;
    } catch (std::exception& e) {
        if (error && error_cap > 0) std::snprintf(error, (size_t)error_cap, "%s", e.what());
        return SGP_ERR_PARSE;
    }
}

}  // extern "C"
