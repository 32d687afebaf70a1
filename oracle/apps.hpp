// ORACLE — test infrastructure only (see oracle/README.md). PARITY UNPINNED.
//
// App front-ends: argument grammar, defaults (from the code, not the README) and file fan-out. Restates
//   apps/SaGePhy/HostTreeGen.java:25-116,  HostTreeGenParameters.java:15-123
//   apps/SaGePhy/GuestTreeGen.java:21-107, GuestTreeGenParameters.java:24-164
//   JCommander 1.18 grammar as used there: options anywhere, Boolean options have arity 0, others take the next token,
//   everything else is a positional ("main parameter"); an unknown "-x" token raises (caught by main -> stack trace).
#pragma once
#include <dirent.h>
#include <fstream>
#include <map>
#include <sstream>
#include "treegen.hpp"
#include "hybrid.hpp"

namespace oracle {

struct Outputs {
    std::vector<std::pair<std::string, std::string>> files;    // (path, content) in write order
    std::string out, err;                                      // stdout / stderr
    void file(const std::string& path, const std::string& content) { files.emplace_back(path, content); }
};

struct OptSpec { std::vector<std::string> names; int arity; };

struct ParsedArgs {
    std::map<std::string, std::vector<std::string>> opts;      // canonical (first) name -> values ("" for flags)
    std::vector<std::string> pos;
    bool has(const std::string& k) const { return opts.count(k) != 0; }
    std::string get(const std::string& k, const std::string& def) const { auto it = opts.find(k); return it == opts.end() ? def : it->second.back(); }
};

inline ParsedArgs jc_parse(const std::vector<std::string>& args, const std::vector<OptSpec>& specs) {
    ParsedArgs p;
    for (size_t i = 0; i < args.size(); ++i) {
        const std::string& a = args[i];
        const OptSpec* hit = nullptr;
        for (auto& s : specs) for (auto& n : s.names) if (n == a) hit = &s;
        if (hit) {
            if (hit->arity == 0) p.opts[hit->names[0]].push_back("");
            else {
                if (i + (size_t)hit->arity >= args.size())
                    throw std::invalid_argument("ParameterException: Expected a value after parameter " + a);
                for (int k = 0; k < hit->arity; ++k) p.opts[hit->names[0]].push_back(args[++i]);
            }
        } else if (!a.empty() && a[0] == '-' && a.size() > 1 && !(a[1] >= '0' && a[1] <= '9') && a[1] != '.') {
            throw std::invalid_argument("ParameterException: Unknown option: " + a);
        } else {
            p.pos.push_back(a);
        }
    }
    return p;
}

inline std::string arrays_to_string(const std::vector<std::string>& a) {   // java.util.Arrays.toString(Object[])
    std::string s = "[";
    for (size_t i = 0; i < a.size(); ++i) { if (i) s += ", "; s += a[i]; }
    return s + "]";
}
inline double parse_double(const std::string& s) {
    std::string t = s;   // Double.parseDouble trims whitespace
    size_t b = t.find_first_not_of(" \t\n\r"), e = t.find_last_not_of(" \t\n\r");
    if (b == std::string::npos) throw std::invalid_argument("NumberFormatException: empty String");
    t = t.substr(b, e - b + 1);
    char* end = nullptr; double v = std::strtod(t.c_str(), &end);
    if (end == t.c_str() || (*end != '\0' && !((*end == 'd' || *end == 'D' || *end == 'f' || *end == 'F') && end[1] == '\0')))
        throw std::invalid_argument("NumberFormatException: For input string: \"" + s + "\"");
    return v;
}
inline int parse_int(const std::string& s) {
    char* end = nullptr; long v = std::strtol(s.c_str(), &end, 10);
    if (end == s.c_str() || *end != '\0') throw std::invalid_argument("NumberFormatException: For input string: \"" + s + "\"");
    return (int)v;
}
inline std::string trim_java(const std::string& s) {   // String.trim(): strips chars <= ' '
    size_t b = 0, e = s.size();
    while (b < e && (unsigned char)s[b] <= ' ') ++b;
    while (e > b && (unsigned char)s[e - 1] <= ' ') --e;
    return s.substr(b, e - b);
}
inline std::vector<int> read_leaf_sizes(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("FileNotFoundException: " + path);
    std::vector<int> v; std::string line;
    while (std::getline(f, line)) v.push_back(parse_int(line));
    return v;
}
inline bool file_exists(const std::string& p) { std::ifstream f(p); return (bool)f; }
inline std::string read_file(const std::string& p) { std::ifstream f(p, std::ios::binary); std::stringstream ss; ss << f.rdbuf(); return ss.str(); }
inline std::string base_name(const std::string& p) { size_t k = p.find_last_of('/'); return k == std::string::npos ? p : p.substr(k + 1); }

static const char* kFailMsg = "Failed to produce valid pruned tree within max allowed attempts.";

// Per-replicate summary used by the statistical (KS) tests and by the CPU-baseline timing.
struct RepStats { int attempts = 0; int status = 0; int n_unpruned = 0; int n_pruned = 0; int sampled_leaves = 0; double pruned_root_time = 0; double total_time = 0; int ev[17] = {0}; uint64_t vertices_created = 0; uint64_t out_bytes = 0; };

inline void fill_stats(RepStats& st, const MachinaResult& r) {
    st.attempts = r.attempts;
    std::deque<const GVertex*> q; q.push_back(r.unpruned);
    while (!q.empty()) { const GVertex* v = q.front(); q.pop_front(); ++st.n_unpruned; st.total_time += v->bl; st.ev[(int)v->event]++; if (v->event == Event::LEAF) ++st.sampled_leaves; if (!v->is_leaf()) { q.push_back(v->lc); q.push_back(v->rc); } }
    if (r.pruned) { q.push_back(r.pruned); st.pruned_root_time = r.pruned->abstime; }
    while (!q.empty()) { const GVertex* v = q.front(); q.pop_front(); ++st.n_pruned; if (!v->is_leaf()) { q.push_back(v->lc); q.push_back(v->rc); } }
}

// ---- HostTreeGen -----------------------------------------------------------------------------------------------------------
inline const std::vector<OptSpec>& host_specs() {
    static const std::vector<OptSpec> s = {
        {{"-h", "--help"}, 0}, {{"-q", "--quiet"}, 0}, {{"-s", "--seed"}, 1}, {{"-min", "--min-leaves"}, 1}, {{"-max", "--max-leaves"}, 1},
        {{"-sizes", "--leaf-sizes-file"}, 1}, {{"-nox", "--no-auxiliary-tags"}, 0}, {{"-p", "--leaf-sampling-probability"}, 1},
        {{"-bi", "--start-with-bifurcation"}, 0}, {{"-a", "--max-attempts"}, 1}, {{"-stem", "--override-stem"}, 1}, {{"-vp", "--vertex-prefix"}, 1}};
    return s;
}

inline void run_host_tree_gen(const std::vector<std::string>& args, Outputs& o, RepStats* st = nullptr) {
    try {
        ParsedArgs p = jc_parse(args, host_specs());
        bool quiet = p.has("-q");
        size_t noargs = quiet ? 3 : 4;
        if (args.empty() || p.has("-h") || p.pos.size() != noargs) {
            o.out += "Usage:\n    java -jar sagephy-X.Y.Z.jar HostTreeGen [options] <time interval> <birth rate> <death rate> <out prefix>\n\n";
            return;
        }
        double T = parse_double(p.pos[0]);
        if (T <= 0.0) throw std::invalid_argument("IllegalArgumentException: Invalid time interval.");
        bool bi = p.has("-bi");
        std::string Ts = java_double_to_string(T);
        std::string S = bi ? "(H0:" + Ts + ",H1:" + Ts + "):0.0;" : "H0:" + Ts + ";";
        auto hostNewick = nw_read_tree(S);
        std::vector<int> sizes; bool hasSizes = p.has("-sizes");
        if (hasSizes) sizes = read_leaf_sizes(p.get("-sizes", ""));
        bool excludeMeta = p.has("-nox");
        if (!p.has("-s")) throw std::invalid_argument("oracle: a seed (-s) is required (the reference would read /dev/random)");
        Machina machine(p.get("-s", ""), parse_int(p.get("-min", "2")), parse_int(p.get("-max", "1000")), bi ? 1 : 0, INT32_MAX,
                        hasSizes ? &sizes : nullptr, parse_int(p.get("-a", "10000")), p.get("-vp", "H"), excludeMeta, false, true);
        double birth = parse_double(p.pos[1]); if (birth < 0) throw std::invalid_argument("IllegalArgumentException: Invalid rate.");
        double death = parse_double(p.pos[2]); if (death < 0) throw std::invalid_argument("IllegalArgumentException: Invalid rate.");
        double rho = parse_double(p.get("-p", "1.0")); if (rho < 0) throw std::invalid_argument("IllegalArgumentException: Invalid leaf sampling prob.");
        HostModel host(*hostNewick, std::nullopt, std::nullopt);
        GuestInHostCreator motor(host, birth, death, 0.0, 0.0, "", 1.0, false, 0.0, true, rho);
        std::string prefix = quiet ? "" : trim_java(p.pos[3]);
        MachinaResult r;
        try {
            r = machine.sample(motor, false);
        } catch (MaxAttemptsException&) {
            if (!quiet) o.file(prefix + ".info", std::string(kFailMsg) + "\n");
            o.err += std::string(kFailMsg) + "\n";
            if (st) { st->status = 1; st->attempts = machine.attempts; st->vertices_created = motor.vertices_created; }
            return;
        }
        if (p.has("-stem")) { double s = parse_double(p.get("-stem", "")); r.unpruned->bl = s; if (r.pruned) r.pruned->bl = s; }
        const char* um = excludeMeta ? nullptr : "[NAME=UnprunedTree]";
        const char* pm = excludeMeta ? nullptr : "[NAME=PrunedTree]";
        if (quiet) {
            if (excludeMeta) o.out += (r.pruned ? gtree_to_string(r.pruned, pm) : std::string(";")) + "\n";
            else o.out += (r.pruned ? gtree_to_string(r.pruned, pm) : std::string("[NAME=PrunedTree];")) + "\n";
        } else {
            o.file(prefix + ".unpruned.tree", gtree_to_string(r.unpruned, um) + "\n");
            if (excludeMeta) o.file(prefix + ".pruned.tree", r.pruned ? gtree_to_string(r.pruned, pm) + "\n" : std::string(";\n"));
            else o.file(prefix + ".pruned.tree", r.pruned ? gtree_to_string(r.pruned, pm) + "\n" : std::string("[NAME=PrunedTree];"));
            std::string head = "# HOSTTREEGEN\nArguments:\t" + arrays_to_string(args) + "\nAttempts:\t" + std::to_string(r.attempts) + "\n";
            o.file(prefix + ".unpruned.info", head + host_tree_gen_info(r.unpruned, true));
            o.file(prefix + ".pruned.info", head + host_tree_gen_info(r.pruned, false));
        }
        if (st) { fill_stats(*st, r); st->vertices_created = motor.vertices_created; }
    } catch (std::exception& e) {
        o.err += std::string(e.what()) + "\n\nUse option -h or --help to show usage.\n";
        if (st) st->status = 2;
    }
}

// getHostHybridGraphCreator (GuestTreeGenParameters.java:182-197, DomainTreeGenParameters.java:205-220) followed by
// GuestTreeMachina.sampleGuestTree -> createUnprunedTree (GuestTreeInHybridGraphCreator.java:86-96): never returns, see hybrid.hpp.
[[noreturn]] inline void run_hybrid_host(const ParsedArgs& p, size_t hostPos, size_t dupPos, size_t lossPos) {
    const std::string& str = p.pos[hostPos];
    double dup = parse_double(p.pos[dupPos]);
    double loss = parse_double(p.pos[lossPos]);
    const std::vector<std::string>& hybrid = p.opts.at("-hybrid");
    const size_t h0 = hybrid.size() - 3;                        // JCommander keeps the values of the last occurrence
    double timespan = parse_double(hybrid[h0]);
    double dupfact = parse_double(hybrid[h0 + 1]);
    double lossfact = parse_double(hybrid[h0 + 2]);
    gml::GMLGraph host = gml::first_graph(gml::readGML(file_exists(str) ? read_file(str) : str));
    double rho = parse_double(p.get("-p", "1.0"));
    if (p.has("-stem")) (void)parse_double(p.get("-stem", ""));
    hybrid_creator_and_first_vertex(host, dup, loss, dupfact, lossfact, timespan, rho);
}

// ---- GuestTreeGen -----------------------------------------------------------------------------------------------------------
inline const std::vector<OptSpec>& guest_specs() {
    static const std::vector<OptSpec> s = {
        {{"-h", "--help"}, 0}, {{"-q", "--quiet"}, 0}, {{"-s", "--seed"}, 1}, {{"-min", "--min-leaves"}, 1}, {{"-max", "--max-leaves"}, 1},
        {{"-minper", "--min-leaves-per-host-leaf"}, 1}, {{"-maxper", "--max-leaves-per-host-leaf"}, 1}, {{"-sizes", "--leaf-sizes-file"}, 1},
        {{"-nox", "--no-auxiliary-tags"}, 0}, {{"-p", "--leaf-sampling-probability"}, 1}, {{"-mpr", "--enforce-most-parsimonious-reconciliation"}, 0},
        {{"-a", "--max-attempts"}, 1}, {{"-stem", "--override-host-stem"}, 1}, {{"-vp", "--vertex-prefix"}, 1}, {{"-vph", "--vertex-prefix-host-map"}, 0},
        {{"-hybrid", "--hybrid-host-graph"}, 3}, {{"-rt", "--replacing-transfers"}, 1}, {{"-db", "--distance-bias"}, 1}, {{"-dbr", "--distance-bias-rate"}, 1},
        {{"-gb", "--gene-birth-sampling"}, 0}, {{"-gbc", "--gene-birth-coefficient"}, 1}};
    return s;
}

inline void run_guest_tree_gen(const std::vector<std::string>& args, Outputs& o, RepStats* st = nullptr) {
    try {
        ParsedArgs p = jc_parse(args, guest_specs());
        bool quiet = p.has("-q");
        size_t noargs = quiet ? 4 : 5;
        if (args.empty() || p.has("-h") || p.pos.size() != noargs) {
            o.out += "Usage:\n    java -jar sagephy-X.Y.Z.jar GuestTreeGen [options] <host tree> <dup rate> <loss rate> <trans rate> <out prefix>\n\n";
            return;
        }
        std::vector<int> sizes; bool hasSizes = p.has("-sizes");
        if (hasSizes) sizes = read_leaf_sizes(p.get("-sizes", ""));
        bool excludeMeta = p.has("-nox");
        if (!p.has("-s")) throw std::invalid_argument("oracle: a seed (-s) is required (the reference would read /dev/random)");
        Machina machina(p.get("-s", ""), parse_int(p.get("-min", "2")), parse_int(p.get("-max", "1000")), parse_int(p.get("-minper", "0")),
                        parse_int(p.get("-maxper", "10")), hasSizes ? &sizes : nullptr, parse_int(p.get("-a", "10000")), p.get("-vp", "G"),
                        excludeMeta, true, false);
        if (p.has("-hybrid")) run_hybrid_host(p, 0, 1, 2);     // GuestTreeGen.java:42
        // getHostTreeCreator (GuestTreeGenParameters.java:127-136)
        std::unique_ptr<NTree> hostNewick; std::optional<std::string> treeName;
        if (file_exists(p.pos[0])) { hostNewick = nw_read_tree(read_file(p.pos[0])); treeName = base_name(p.pos[0]); }
        else hostNewick = nw_read_tree(p.pos[0]);
        std::optional<double> stem = p.has("-stem") ? parse_double(p.get("-stem", "")) : 0.0;
        HostModel host(*hostNewick, treeName, stem);
        GuestInHostCreator motor(host, parse_double(p.pos[1]), parse_double(p.pos[2]), parse_double(p.pos[3]), parse_double(p.get("-rt", "0.5")),
                                 p.get("-db", "simple"), parse_double(p.get("-dbr", "1.0")), p.has("-gb"), parse_double(p.get("-gbc", "1.0")), false,
                                 parse_double(p.get("-p", "1.0")));
        std::string prefix = quiet ? "" : trim_java(p.pos[4]);
        MachinaResult r;
        try {
            r = machina.sample(motor, false);
        } catch (MaxAttemptsException&) {
            if (!quiet) o.file(prefix + ".pruned.info", std::string(kFailMsg) + "\n");
            o.err += std::string(kFailMsg) + "\n";
            if (st) { st->status = 1; st->attempts = machina.attempts; }
            return;
        }
        const char* um = excludeMeta ? nullptr : "[NAME=UnprunedTree]";
        const char* pm = excludeMeta ? nullptr : "[NAME=PrunedTree]";
        if (quiet) {
            if (excludeMeta) o.out += (r.pruned ? gtree_to_string(r.pruned, pm) : std::string(";")) + "\n";
            else o.out += (r.pruned ? gtree_to_string(r.pruned, pm) : std::string("[NAME=PrunedTree];")) + "\n";
        } else {
            o.file(prefix + ".unpruned.tree", gtree_to_string(r.unpruned, um) + "\n");
            if (excludeMeta) o.file(prefix + ".pruned.tree", r.pruned ? gtree_to_string(r.pruned, pm) + "\n" : std::string(";\n"));
            else o.file(prefix + ".pruned.tree", r.pruned ? gtree_to_string(r.pruned, pm) + "\n" : std::string("[NAME=PrunedTree];\n"));
            std::string head = "# GUESTTREEGEN\nArguments:\t" + arrays_to_string(args) + "\nAttempts:\t" + std::to_string(r.attempts) + "\n";
            o.file(prefix + ".unpruned.info", head + motor.info(r.unpruned, true));
            if (!r.pruned) throw std::runtime_error("NullPointerException (pruned tree is empty)");   // GuestTreeGen.java:86 dereferences null
            o.file(prefix + ".pruned.info", head + motor.info(r.pruned, true));
            if (!excludeMeta) {
                o.file(prefix + ".unpruned.guest2host", motor.sigma(r.unpruned));
                o.file(prefix + ".pruned.guest2host", motor.sigma(r.pruned));
                o.file(prefix + ".unpruned.leafmap", motor.leaf_map(r.unpruned));
                o.file(prefix + ".pruned.leafmap", motor.leaf_map(r.pruned));
            }
        }
        if (st) { fill_stats(*st, r); st->vertices_created = motor.vertices_created; }
    } catch (std::exception& e) {
        o.err += std::string(e.what()) + "\n\nUse option -h or --help to show usage.\n";
        if (st) st->status = 2;
    }
}


// ---- DomainTreeGen (apps/SaGePhy/DomainTreeGen.java:21-107, DomainTreeGenParameters.java:30-151) ---------------------------------
inline const std::vector<OptSpec>& domain_specs() {
    static const std::vector<OptSpec> s = {
        {{"-h", "--help"}, 0}, {{"-q", "--quiet"}, 0}, {{"-s", "--seed"}, 1}, {{"-min", "--min-leaves"}, 1}, {{"-max", "--max-leaves"}, 1},
        {{"-minper", "--min-leaves-per-host-leaf"}, 1}, {{"-maxper", "--max-leaves-per-host-leaf"}, 1}, {{"-sizes", "--leaf-sizes-file"}, 1},
        {{"-nox", "--no-auxiliary-tags"}, 0}, {{"-p", "--leaf-sampling-probability"}, 1}, {{"-mpr", "--enforce-most-parsimonious-reconciliation"}, 0},
        {{"-a", "--max-attempts"}, 1}, {{"-stem", "--override-host-stem"}, 1}, {{"-vp", "--vertex-prefix"}, 1}, {{"-vph", "--vertex-prefix-host-map"}, 0},
        {{"-hybrid", "--hybrid-host-graph"}, 3}, {{"-rt", "--replacing-transfers"}, 1}, {{"-db", "--distance-bias"}, 1}, {{"-dbr", "--distance-bias-rate"}, 1},
        {{"-dbs", "--domain-birth-sampling"}, 0}, {{"-dbc", "--domain-birth-coefficient"}, 1}, {{"-ig", "--inter-gene-transfers"}, 1},
        {{"-is", "--inter-species-transfers"}, 1}, {{"-all", "--all-gene-trees"}, 0}};
    return s;
}
// Gene-tree files of the directory in lexicographic file-name order (the reference uses File.listFiles(), whose order is
// OS-dependent; both the oracle and the product fix it to sorted order — SURVEY H7).
inline std::vector<std::string> list_dir_sorted(const std::string& dir) {
    std::vector<std::string> v; DIR* d = opendir(dir.c_str());
    if (!d) throw std::runtime_error("NullPointerException: cannot list gene tree directory " + dir);
    while (dirent* e = readdir(d)) { std::string n = e->d_name; if (n == "." || n == "..") continue; v.push_back(dir + "/" + n); }
    closedir(d); std::sort(v.begin(), v.end()); return v;
}

inline void run_domain_tree_gen(const std::vector<std::string>& args, Outputs& o, RepStats* st = nullptr) {
    try {
        ParsedArgs p = jc_parse(args, domain_specs());
        bool quiet = p.has("-q");
        size_t noargs = quiet ? 5 : 6;
        if (args.empty() || p.has("-h") || p.pos.size() != noargs) {
            o.out += "Usage:\n    java -jar sagephy-X.Y.Z.jar DomainTreeGen [options] <host tree file or string> <guest tree directory> <dup rate> <loss rate> <trans rate> <out prefix>\n\n";
            return;
        }
        std::vector<int> sizes; bool hasSizes = p.has("-sizes");
        if (hasSizes) sizes = read_leaf_sizes(p.get("-sizes", ""));
        bool excludeMeta = p.has("-nox");
        if (!p.has("-s")) throw std::invalid_argument("oracle: a seed (-s) is required (the reference would read /dev/random)");
        Machina machina(p.get("-s", ""), parse_int(p.get("-min", "2")), parse_int(p.get("-max", "1000")), parse_int(p.get("-minper", "0")),
                        parse_int(p.get("-maxper", "10")), hasSizes ? &sizes : nullptr, parse_int(p.get("-a", "10000")), p.get("-vp", "D"), excludeMeta, true, false);
        if (p.has("-hybrid")) run_hybrid_host(p, 0, 2, 3);     // DomainTreeGen.java:42
        std::unique_ptr<NTree> spNewick; std::optional<std::string> spName;
        if (file_exists(p.pos[0])) { spNewick = nw_read_tree(read_file(p.pos[0])); spName = base_name(p.pos[0]); } else spNewick = nw_read_tree(p.pos[0]);
        std::optional<double> stem = p.has("-stem") ? parse_double(p.get("-stem", "")) : 0.0;
        HostModel species(*spNewick, spName, stem);
        std::vector<std::unique_ptr<NTree>> geneNewicks; std::vector<std::unique_ptr<HostModel>> geneModels; std::vector<HostModel*> genePtrs;
        for (const std::string& f : list_dir_sorted(p.pos[1])) {
            geneNewicks.push_back(nw_read_tree(read_file(f)));
            geneModels.push_back(std::make_unique<HostModel>(*geneNewicks.back(), std::optional<std::string>(base_name(f)), std::nullopt));
            genePtrs.push_back(geneModels.back().get());
        }
        if (genePtrs.empty()) throw std::invalid_argument("IllegalArgumentException: n must be positive (no gene trees in directory)");
        DomainCreator motor(species, genePtrs, parse_double(p.pos[2]), parse_double(p.pos[3]), parse_double(p.pos[4]), parse_double(p.get("-rt", "0.5")),
                            p.get("-db", "none"), parse_double(p.get("-dbr", "1.0")), p.has("-dbs"), parse_double(p.get("-dbc", "1.0")),
                            parse_double(p.get("-ig", "0.5")), parse_double(p.get("-is", "0.5")), parse_double(p.get("-p", "1.0")));
        std::string prefix = quiet ? "" : trim_java(p.pos[5]);
        MachinaResult r;
        try { r = machina.sample(motor, p.has("-all")); }
        catch (MaxAttemptsException&) {
            if (!quiet) o.file(prefix + ".pruned.info", std::string(kFailMsg) + "\n");
            o.err += std::string(kFailMsg) + "\n";
            if (st) { st->status = 1; st->attempts = machina.attempts; }
            return;
        }
        const char* um = excludeMeta ? nullptr : "[NAME=UnprunedTree]"; const char* pm = excludeMeta ? nullptr : "[NAME=PrunedTree]";
        if (quiet) {
            if (excludeMeta) o.out += (r.pruned ? gtree_to_string(r.pruned, pm) : std::string(";")) + "\n";
            else o.out += (r.pruned ? gtree_to_string(r.pruned, pm) : std::string("[NAME=PrunedTree];")) + "\n";
        } else {
            o.file(prefix + ".unpruned.tree", gtree_to_string(r.unpruned, um) + "\n");
            if (excludeMeta) o.file(prefix + ".pruned.tree", r.pruned ? gtree_to_string(r.pruned, pm) + "\n" : std::string(";\n"));
            else o.file(prefix + ".pruned.tree", r.pruned ? gtree_to_string(r.pruned, pm) + "\n" : std::string("[NAME=PrunedTree];\n"));
            std::string head = "# DOMAINTREEGEN\nArguments:\t" + arrays_to_string(args) + "\nAttempts:\t" + std::to_string(r.attempts) + "\n";
            o.file(prefix + ".unpruned.info", head + motor.info(r.unpruned, true));
            if (!r.pruned) throw std::runtime_error("NullPointerException (pruned tree is empty)");
            o.file(prefix + ".pruned.info", head + motor.info(r.pruned, true));
            if (!excludeMeta) {
                o.file(prefix + ".unpruned.guest2host", motor.sigma(r.unpruned)); o.file(prefix + ".pruned.guest2host", motor.sigma(r.pruned));
                o.file(prefix + ".unpruned.leafmap", motor.leaf_map(r.unpruned)); o.file(prefix + ".pruned.leafmap", motor.leaf_map(r.pruned));
            }
        }
        if (st) { fill_stats(*st, r); st->vertices_created = motor.vertices_created; }
    } catch (std::exception& e) {
        o.err += std::string(e.what()) + "\n\nUse option -h or --help to show usage.\n";
        if (st) st->status = 2;
    }
}

}  // namespace oracle
