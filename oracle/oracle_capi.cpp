// ORACLE — test infrastructure only (see oracle/README.md). PARITY UNPINNED.
// C entry points for ctypes (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference) and a tiny CLI.
#include <chrono>
#include <cstring>
#include "apps.hpp"
#include "relax.hpp"

using namespace oracle;

namespace {
struct Handle { Outputs o; RepStats st; long no_recipient = 0; };

void dispatch(const std::vector<std::string>& argv, Outputs& o, RepStats* st) {
    if (argv.empty()) { o.out += "Usage: <HostTreeGen|GuestTreeGen|DomainTreeGen|BranchRelaxer> [options] <args>\n"; return; }
    std::vector<std::string> rest(argv.begin() + 1, argv.end());
    const std::string& app = argv[0];
    if (app == "HostTreeGen") run_host_tree_gen(rest, o, st);
    else if (app == "GuestTreeGen") run_guest_tree_gen(rest, o, st);
    else if (app == "BranchRelaxer") run_branch_relaxer(rest, o, st);
    else if (app == "DomainTreeGen") run_domain_tree_gen(rest, o, st);
    else { o.out += "Unknown application: " + app + "\n"; if (st) st->status = 2; }
}

// Replace the value following -s/--seed by seed+i (replicate i == reference run with "-s seed+i").
std::vector<std::string> with_seed_offset(const std::vector<std::string>& argv, long long i) {
    std::vector<std::string> a = argv;
    for (size_t k = 0; k + 1 < a.size(); ++k)
        if (a[k] == "-s" || a[k] == "--seed") { a[k + 1] = decimal_add(a[k + 1], (unsigned long long)i); break; }
    return a;
}
}  // namespace

extern "C" {

void* oracle_run(int argc, const char** argv) {
    auto* h = new Handle();
    std::vector<std::string> a(argv, argv + argc);
    no_recipient_hits() = 0;
    dispatch(a, h->o, &h->st);
    h->no_recipient = no_recipient_hits();
    return h;
}
int oracle_n_files(void* h) { return (int)((Handle*)h)->o.files.size(); }
const char* oracle_file_name(void* h, int i) { return ((Handle*)h)->o.files[i].first.c_str(); }
const char* oracle_file_data(void* h, int i) { return ((Handle*)h)->o.files[i].second.data(); }
long long oracle_file_len(void* h, int i) { return (long long)((Handle*)h)->o.files[i].second.size(); }
const char* oracle_stdout(void* h) { return ((Handle*)h)->o.out.c_str(); }
const char* oracle_stderr(void* h) { return ((Handle*)h)->o.err.c_str(); }
int oracle_attempts(void* h) { return ((Handle*)h)->st.attempts; }
int oracle_status(void* h) { return ((Handle*)h)->st.status; }
long long oracle_no_recipient(void* h) { return ((Handle*)h)->no_recipient; }
void oracle_free(void* h) { delete (Handle*)h; }

// Batch: replicate i runs with seed+i. Per-replicate stats are written as 32 int64/doubles-in-slots rows.
// stats row layout (doubles): [attempts, status, n_unpruned, n_pruned, sampled_leaves, pruned_root_time, total_time,
//                              vertices_created, out_bytes, ev[0..16]]   -> 26 doubles
// Returns elapsed seconds of the loop.
double oracle_batch(int argc, const char** argv, long long first, long long n, double* stats /* n x 26 or null */) {
    std::vector<std::string> a(argv, argv + argc);
    auto t0 = std::chrono::steady_clock::now();
    for (long long i = 0; i < n; ++i) {
        Outputs o; RepStats st;
        dispatch(with_seed_offset(a, first + i), o, &st);
        uint64_t bytes = o.out.size();
        for (auto& f : o.files) bytes += f.second.size();
        if (stats) {
            double* r = stats + i * 26;
            r[0] = st.attempts; r[1] = st.status; r[2] = st.n_unpruned; r[3] = st.n_pruned; r[4] = st.sampled_leaves; r[5] = st.pruned_root_time;
            r[6] = st.total_time; r[7] = (double)st.vertices_created; r[8] = (double)bytes;
            for (int e = 0; e < 17; ++e) r[9 + e] = st.ev[e];
        }
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// probes
void oracle_mt_words(const char* seed_decimal, int n, unsigned* out) { JavaMT m(seed_decimal); for (int i = 0; i < n; ++i) out[i] = m.next_word(); }
void oracle_mt_doubles(const char* seed_decimal, int n, double* out) { JavaMT m(seed_decimal); for (int i = 0; i < n; ++i) out[i] = m.nextDouble(); }
void oracle_mt_ints(const char* seed_decimal, int bound, int n, int* out) { JavaMT m(seed_decimal); for (int i = 0; i < n; ++i) out[i] = m.nextInt(bound); }
int oracle_double_to_string(double v, char* buf, int cap) { std::string s = java_double_to_string(v); std::snprintf(buf, cap, "%s", s.c_str()); return (int)s.size(); }
double oracle_log(double x) { return j_log(x); }
double oracle_log10(double x) { return j_log10(x); }
double oracle_exp(double x) { return j_exp(x); }
double oracle_round_sig(double x, int n) { return round_sig(x, n); }
double oracle_pow(double x, double y) { return j_pow(x, y); }
double oracle_normal_quantile(double p) { return normal_quantile(p); }
double oracle_gamma_quantile(double p, double k, double theta) { return chi2_quantile(p, 2 * k) * theta / 2; }
long long oracle_round(double x) { return j_round(x); }

}  // extern "C"

#ifdef ORACLE_MAIN
int main(int argc, char** argv) {
    std::vector<std::string> a(argv + 1, argv + argc);
    Outputs o; RepStats st;
    dispatch(a, o, &st);
    for (auto& f : o.files) { std::ofstream out(f.first, std::ios::binary); out << f.second; }
    std::fputs(o.out.c_str(), stdout);
    std::fputs(o.err.c_str(), stderr);
    return 0;
}
#endif
