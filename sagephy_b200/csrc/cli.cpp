// `sagephy <App> [batch options] [reference options] <positionals>` — native stand-in for
// `java -jar sagephy-1.0.0.jar <App> ...` (reference: apps/SaGePhyStarter.java:75-103 strips the app name and forwards
// the rest). Batch options (new surface, removed before the reference grammar sees argv):
//   -n / --replicates N         number of replicates (default 1 => the same files the reference writes)
//   -rng / --rng philox|mt-replay   (default: mt-replay when -s is given and N == 1, else philox)
//   -gpu / --gpu K              CUDA device index (default 0)
//   --gpus K                    shard the replicates by index range over devices 0..K-1 of this node (one host thread per device; the
//                               files are the same bytes a single device writes: replicate ids are global)
//   --out-layout packed|per-replicate   for N > 1: one file per stream + .idx (default), or one set of files per replicate named
//                               <prefix>_<i>.* (what N reference runs with prefixes <prefix>_0 ... would leave behind)
// Exit code is 0 in every case the reference exits 0 (including usage and max-attempts failures).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/sagephy_b200.h"

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fputs("Usage: sagephy <HostTreeGen|GuestTreeGen|DomainTreeGen|BranchRelaxer> [-n N] [-rng philox|mt-replay] [-gpu K] [--gpus K] [--out-layout packed|per-replicate] [options] <args>\n", stdout);
        return 0;
    }
    std::string app = argv[1];
    std::vector<const char*> rest;
    sgp_batch_opts o; std::memset(&o, 0, sizeof o); o.n_replicates = 1; o.rng_mode = -1;
    int gpu = 0, gpus = 1; bool has_seed = false; int layout = -1;
    for (int i = 2; i < argc; ++i) {
        std::string t = argv[i];
        if ((t == "-n" || t == "--replicates") && i + 1 < argc) { o.n_replicates = std::atoll(argv[++i]); continue; }
        if ((t == "-rng" || t == "--rng") && i + 1 < argc) { std::string m = argv[++i]; o.rng_mode = (m == "mt-replay" || m == "mt") ? SGP_RNG_MT_REPLAY : SGP_RNG_PHILOX; continue; }
        if ((t == "-gpu" || t == "--gpu") && i + 1 < argc) { gpu = std::atoi(argv[++i]); continue; }
        if (t == "--gpus" && i + 1 < argc) { gpus = std::atoi(argv[++i]); continue; }
        if (t == "--out-layout" && i + 1 < argc) { layout = std::string(argv[++i]) == "per-replicate" ? SGP_LAYOUT_PER_REPLICATE : SGP_LAYOUT_PACKED; continue; }
        if (t == "-s" || t == "--seed") has_seed = true;
        rest.push_back(argv[i]);
    }
    if (o.rng_mode < 0) o.rng_mode = (has_seed && o.n_replicates == 1) ? SGP_RNG_MT_REPLAY : SGP_RNG_PHILOX;
    if (gpus > 1 && o.n_replicates > 1) {
        std::vector<int> dev; for (int g = 0; g < gpus; ++g) dev.push_back(gpu + g);
        if (const char* e = std::getenv("SAGEPHY_DEVICES")) {      // explicit device list "0,0" (e.g. two shards on one GPU)
            dev.clear(); for (const char* p = e; *p;) { dev.push_back(std::atoi(p)); while (*p && *p != ',') ++p; if (*p == ',') ++p; }
        }
        sgp_multi* m = nullptr;
        if (sgp_multi_create(dev.data(), (int)dev.size(), &m) != SGP_OK) { std::fprintf(stderr, "sagephy: %s\n", sgp_last_error(nullptr)); return 1; }
        std::vector<sgp_batch*> shards(dev.size(), nullptr);
        sgp_summary sum;
        const sgp_status st = sgp_app_run_multi(m, app.c_str(), (int)rest.size(), rest.data(), &o, shards.data(), &sum);
        sgp_batch* first = nullptr; for (auto* b : shards) if (b) { first = b; break; }
        if (first) { std::fputs(sgp_batch_stdout(first), stdout); std::fputs(sgp_batch_stderr(first), stderr); }
        if (st == SGP_OK) {
            if (first && *sgp_batch_out_prefix(first) == 0) for (auto* b : shards) if (b && b != first) std::fputs(sgp_batch_stdout(b), stdout);      // -q: trees of every shard, in order
            if (sum.n_failed > 0) std::fputs("Failed to produce valid pruned tree within max allowed attempts.\n", stderr);
            sgp_batches_write_files_layout(shards.data(), (int)shards.size(), layout < 0 ? SGP_LAYOUT_PACKED : layout, o.replicate_offset);
        }
        for (auto* b : shards) if (b) sgp_batch_free(b);
        sgp_multi_destroy(m);
        return (st == SGP_ERR_NO_DEVICE) ? 1 : 0;
    }
    sgp_context* ctx = nullptr;
    if (sgp_context_create(gpu, &ctx) != SGP_OK) { std::fprintf(stderr, "sagephy: %s\n", sgp_last_error(nullptr)); return 1; }
    sgp_batch* b = nullptr;
    sgp_status st = sgp_app_run(ctx, app.c_str(), (int)rest.size(), rest.data(), &o, &b);
    if (b) {
        std::fputs(sgp_batch_stdout(b), stdout);
        std::fputs(sgp_batch_stderr(b), stderr);
        if (st == SGP_OK) {
            const int32_t* status = sgp_batch_status(b);
            bool any_fail = false; for (int64_t i = 0; i < sgp_batch_n(b); ++i) if (status[i] != SGP_REP_OK) any_fail = true;
            if (any_fail) std::fputs("Failed to produce valid pruned tree within max allowed attempts.\n", stderr);
            if (o.n_replicates == 1 && layout < 0) sgp_batch_write_files(b);             // the reference's own file names
            else sgp_batch_write_files_layout(b, layout < 0 ? SGP_LAYOUT_PACKED : layout, o.replicate_offset);
        }
        sgp_batch_free(b);
    }
    sgp_context_destroy(ctx);
    return (st == SGP_ERR_NO_DEVICE) ? 1 : 0;
}
