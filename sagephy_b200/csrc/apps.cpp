// Reference command-line grammar -> engine configuration.
//
// Mirrors (src/main/java/uc/sgp/sagephy/apps/SaGePhy/): HostTreeGenParameters.java:15-123, HostTreeGen.java:25-58,
// GuestTreeGenParameters.java:24-164, GuestTreeGen.java:21-44. Defaults are the ones in the code (e.g. -db defaults to
// "simple", -vph to true, the host stem to 0 -> 1e-64), not the README's. Options may appear anywhere; Boolean options take
// no value; a token starting with '-' that is not an option and not a number is an error, as in JCommander.
#include <dirent.h>
#include <algorithm>
#include "apps_common.hpp"
#include "hybrid_host.hpp"

namespace sgp {

using namespace cli;

namespace {

const Opt kHostOpts[] = {
    {"-h", "--help", 0}, {"-q", "--quiet", 0}, {"-s", "--seed", 1}, {"-min", "--min-leaves", 1}, {"-max", "--max-leaves", 1},
    {"-sizes", "--leaf-sizes-file", 1}, {"-nox", "--no-auxiliary-tags", 0}, {"-p", "--leaf-sampling-probability", 1},
    {"-bi", "--start-with-bifurcation", 0}, {"-a", "--max-attempts", 1}, {"-stem", "--override-stem", 1}, {"-vp", "--vertex-prefix", 1}};

const Opt kGuestOpts[] = {
    {"-h", "--help", 0}, {"-q", "--quiet", 0}, {"-s", "--seed", 1}, {"-min", "--min-leaves", 1}, {"-max", "--max-leaves", 1},
    {"-minper", "--min-leaves-per-host-leaf", 1}, {"-maxper", "--max-leaves-per-host-leaf", 1}, {"-sizes", "--leaf-sizes-file", 1},
    {"-nox", "--no-auxiliary-tags", 0}, {"-p", "--leaf-sampling-probability", 1}, {"-mpr", "--enforce-most-parsimonious-reconciliation", 0},
    {"-a", "--max-attempts", 1}, {"-stem", "--override-host-stem", 1}, {"-vp", "--vertex-prefix", 1}, {"-vph", "--vertex-prefix-host-map", 0},
    {"-hybrid", "--hybrid-host-graph", 3}, {"-rt", "--replacing-transfers", 1}, {"-db", "--distance-bias", 1}, {"-dbr", "--distance-bias-rate", 1},
    {"-gb", "--gene-birth-sampling", 0}, {"-gbc", "--gene-birth-coefficient", 1}};

const Opt kDomainOpts[] = {
    {"-h", "--help", 0}, {"-q", "--quiet", 0}, {"-s", "--seed", 1}, {"-min", "--min-leaves", 1}, {"-max", "--max-leaves", 1},
    {"-minper", "--min-leaves-per-host-leaf", 1}, {"-maxper", "--max-leaves-per-host-leaf", 1}, {"-sizes", "--leaf-sizes-file", 1},
    {"-nox", "--no-auxiliary-tags", 0}, {"-p", "--leaf-sampling-probability", 1}, {"-mpr", "--enforce-most-parsimonious-reconciliation", 0},
    {"-a", "--max-attempts", 1}, {"-stem", "--override-host-stem", 1}, {"-vp", "--vertex-prefix", 1}, {"-vph", "--vertex-prefix-host-map", 0},
    {"-hybrid", "--hybrid-host-graph", 3}, {"-rt", "--replacing-transfers", 1}, {"-db", "--distance-bias", 1}, {"-dbr", "--distance-bias-rate", 1},
    {"-dbs", "--domain-birth-sampling", 0}, {"-dbc", "--domain-birth-coefficient", 1}, {"-ig", "--inter-gene-transfers", 1},
    {"-is", "--inter-species-transfers", 1}, {"-all", "--all-gene-trees", 0}};

std::vector<std::string> list_dir(const std::string& dir) {
    std::vector<std::string> out;
    DIR* d = opendir(dir.c_str());
    if (!d) throw SgpError("cannot list gene tree directory " + dir);
    while (dirent* e = readdir(d)) { const std::string n = e->d_name; if (n != "." && n != "..") out.push_back(dir + "/" + n); }
    closedir(d);
    std::sort(out.begin(), out.end());
    return out;
}

void common_window(const Parsed& p, TreeGenConfig& cfg) {
    cfg.P.min_leaves = to_int(p.str("-min", "2")); cfg.P.max_leaves = to_int(p.str("-max", "1000")); cfg.P.max_attempts = to_int(p.str("-a", "10000"));
    if (p.on("-sizes")) cfg.sizes = load_sizes(p.str("-sizes", ""));
    cfg.exclude_meta = p.on("-nox"); cfg.quiet = p.on("-q");
    if (p.on("-s")) { cfg.has_seed = true; cfg.seed = parse_seed(p.str("-s", "0")); cfg.seed_arg_index = p.val_index.at("-s"); }
    else { cfg.has_seed = false; cfg.seed = random_seed(); cfg.seed_arg_index = -1; }
}

// getHostHybridGraphCreator (GuestTreeGenParameters.java:182-197, DomainTreeGenParameters.java:205-220): rates, the three -hybrid
// values, then the GML graph (a file if one of that name exists, else the argument itself), then -p and -stem.
[[noreturn]] void hybrid_host(const Parsed& p, int host_pos, int dup_pos, int loss_pos) {
    const double dup = to_double(p.positional[dup_pos]), loss = to_double(p.positional[loss_pos]);
    const std::vector<std::string>& h = p.val.at("-hybrid");
    const double timespan = to_double(h[h.size() - 3]), dupfact = to_double(h[h.size() - 2]), lossfact = to_double(h[h.size() - 1]);
    const std::string& arg = p.positional[host_pos];
    const std::string text = exists(arg) ? slurp(arg) : arg;
    hybrid::Graph g = hybrid::read_first_graph(text);
    const double rho = to_double(p.str("-p", "1.0"));
    if (p.on("-stem")) (void)to_double(p.str("-stem", "0"));
    hybrid::create_and_sample(g, dup, loss, dupfact, lossfact, timespan, rho);
}

}  // namespace

void parse_treegen_app(const std::string& app, const std::vector<std::string>& args, TreeGenConfig& cfg, AppResult& res, const std::vector<std::pair<std::string, FlatTree>>* forest) {
    cfg.args = args;
    if (app == "HostTreeGen") {
        Parsed p = scan(args, kHostOpts, (int)(sizeof kHostOpts / sizeof kHostOpts[0]));
        const size_t want = p.on("-q") ? 3 : 4;
        if (args.empty() || p.on("-h") || p.positional.size() != want) {
            res.status = 6;
            res.out_text = "Usage:\n    java -jar sagephy-X.Y.Z.jar HostTreeGen [options] <time interval> <birth rate> <death rate> <out prefix>\n"
                           "    (sagephy_b200: same options plus the batch options of sgp_batch_opts / `sagephy -n <replicates> -rng <philox|mt-replay>`)\n\n";
            return;
        }
        cfg.app = 0;
        common_window(p, cfg);
        const double T = to_double(p.positional[0]);
        if (T <= 0.0) throw SgpError("Invalid time interval.");
        const std::string Ts = jdouble(T);
        const bool bi = p.on("-bi");
        const std::string newick = bi ? "(H0:" + Ts + ",H1:" + Ts + "):0.0;" : "H0:" + Ts + ";";
        cfg.host.build(parse_newick(newick), nullptr, nullptr, false);
        cfg.P.lambda = to_double(p.positional[1]); if (cfg.P.lambda < 0) throw SgpError("Invalid rate.");
        cfg.P.mu = to_double(p.positional[2]); if (cfg.P.mu < 0) throw SgpError("Invalid rate.");
        cfg.P.tau = 0; cfg.P.theta = 0; cfg.P.bias = 1 /* "" is treated as "simple"; unused with tau = 0 */; cfg.P.bias_rate = 1.0;
        cfg.P.do_gene_birth = false; cfg.P.gene_birth = 0; cfg.P.ishost = true;
        cfg.P.rho = to_double(p.str("-p", "1.0"));
        if (cfg.P.rho < 0) throw SgpError("Invalid leaf sampling prob.");
        if (cfg.P.rho > 1) throw SgpError("Cannot have leaf sampling probability outside [0,1].");
        cfg.P.minper = bi ? 1 : 0; cfg.P.maxper = 2147483647;
        cfg.vertex_prefix = p.str("-vp", "H"); cfg.append_sigma = false; cfg.is_species = true;
        if (p.on("-stem")) { cfg.has_stem_override = true; cfg.stem_override = to_double(p.str("-stem", "0")); }
        cfg.out_prefix = cfg.quiet ? std::string() : java_trim(p.positional[3]);
        return;
    }
    if (app == "GuestTreeGen") {
        Parsed p = scan(args, kGuestOpts, (int)(sizeof kGuestOpts / sizeof kGuestOpts[0]));
        const size_t want = p.on("-q") ? 4 : 5;
        if (args.empty() || p.on("-h") || p.positional.size() != want) {
            res.status = 6;
            res.out_text = "Usage:\n    java -jar sagephy-X.Y.Z.jar GuestTreeGen [options] <host tree> <dup rate> <loss rate> <trans rate> <out prefix>\n"
                           "    (sagephy_b200: same options plus the batch options of sgp_batch_opts / `sagephy -n <replicates> -rng <philox|mt-replay>`)\n\n";
            return;
        }
        cfg.app = 1;
        common_window(p, cfg);
        cfg.P.minper = to_int(p.str("-minper", "0")); cfg.P.maxper = to_int(p.str("-maxper", "10"));
        if (p.on("-hybrid")) hybrid_host(p, 0, 1, 2);             // never returns: see hybrid_host.hpp
        cfg.P.lambda = to_double(p.positional[1]); cfg.P.mu = to_double(p.positional[2]); cfg.P.tau = to_double(p.positional[3]);
        if (cfg.P.lambda < 0 || cfg.P.mu < 0 || cfg.P.tau < 0) throw SgpError("Cannot have rate less than 0.");
        cfg.P.theta = to_double(p.str("-rt", "0.5"));
        const std::string db = p.str("-db", "simple");
        cfg.P.bias = db == "none" ? 0 : db == "exponential" ? 2 : 1;
        cfg.P.bias_rate = to_double(p.str("-dbr", "1.0"));
        cfg.P.do_gene_birth = p.on("-gb"); cfg.P.gene_birth = to_double(p.str("-gbc", "1.0")); cfg.P.ishost = false;
        cfg.P.rho = to_double(p.str("-p", "1.0"));
        if (cfg.P.rho < 0 || cfg.P.rho > 1) throw SgpError("Cannot have leaf sampling probability outside [0,1].");
        cfg.vertex_prefix = p.str("-vp", "G"); cfg.append_sigma = true; cfg.is_species = false;
        const double stem = p.on("-stem") ? to_double(p.str("-stem", "0")) : 0.0;
        const std::string& hostarg = p.positional[0];
        if (exists(hostarg)) { std::string nm = file_name(hostarg); cfg.host.build(parse_newick(slurp(hostarg)), &nm, &stem, cfg.P.bias != 0 && cfg.P.tau > 0); }
        else cfg.host.build(parse_newick(hostarg), nullptr, &stem, cfg.P.bias != 0 && cfg.P.tau > 0);
        cfg.out_prefix = cfg.quiet ? std::string() : java_trim(p.positional[4]);
        return;
    }
    if (app == "DomainTreeGen") {
        Parsed p = scan(args, kDomainOpts, (int)(sizeof kDomainOpts / sizeof kDomainOpts[0]));
        const size_t want = p.on("-q") ? 5 : 6;
        if (args.empty() || p.on("-h") || p.positional.size() != want) {
            res.status = 6;
            res.out_text = "Usage:\n    java -jar sagephy-X.Y.Z.jar DomainTreeGen [options] <host tree file or string> <guest tree directory> <dup rate> <loss rate> <trans rate> <out prefix>\n"
                           "    (sagephy_b200: same options plus the batch options of sgp_batch_opts / `sagephy -n <replicates> -rng <philox|mt-replay>`)\n\n";
            return;
        }
        cfg.app = 2;
        common_window(p, cfg);
        cfg.P.minper = to_int(p.str("-minper", "0")); cfg.P.maxper = to_int(p.str("-maxper", "10"));
        if (p.on("-hybrid")) hybrid_host(p, 0, 2, 3);             // never returns: see hybrid_host.hpp
        cfg.P.lambda = to_double(p.positional[2]); cfg.P.mu = to_double(p.positional[3]); cfg.P.tau = to_double(p.positional[4]);
        cfg.P.theta = to_double(p.str("-rt", "0.5"));
        const std::string db = p.str("-db", "none");                 // DomainTreeGenParameters.java: default "none"
        cfg.P.bias = db == "none" ? 0 : db == "exponential" ? 2 : 1;
        cfg.P.bias_rate = to_double(p.str("-dbr", "1.0"));
        cfg.P.do_gene_birth = p.on("-dbs"); cfg.P.gene_birth = to_double(p.str("-dbc", "1.0")); cfg.P.ishost = false;
        cfg.P.rho = to_double(p.str("-p", "1.0"));
        cfg.inter_gene = to_double(p.str("-ig", "0.5")); cfg.inter_species = to_double(p.str("-is", "0.5")); cfg.all_genes = p.on("-all");
        if (cfg.P.lambda < 0 || cfg.P.mu < 0 || cfg.P.tau < 0 || cfg.P.gene_birth < 0) throw SgpError("Cannot have rate less than 0.");
        auto bad = [](double v) { return v < 0 || v > 1; };
        if (bad(cfg.P.theta) || bad(cfg.inter_gene) || bad(cfg.inter_species) || bad(cfg.P.rho)) throw SgpError("Cannot have probability outside [0,1].");
        cfg.vertex_prefix = p.str("-vp", "D"); cfg.append_sigma = true; cfg.is_species = false;
        const double stem = p.on("-stem") ? to_double(p.str("-stem", "0")) : 0.0;
        const std::string& sp = p.positional[0];
        if (exists(sp)) { std::string nm = file_name(sp); cfg.host.build(parse_newick(slurp(sp)), &nm, &stem, false); }
        else cfg.host.build(parse_newick(sp), nullptr, &stem, false);
        // gene trees: every file of the directory, in sorted file-name order (the reference's File.listFiles() order is OS-dependent)
        if (forest) {
            // chained stage: the forest is already in memory (sorted by name like a directory listing would be)
            if (forest->empty()) throw SgpError("no gene trees in the source batch");
            std::vector<size_t> idx(forest->size()); for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
            std::sort(idx.begin(), idx.end(), [&](size_t x, size_t y) { return (*forest)[x].first < (*forest)[y].first; });
            cfg.genes.resize(forest->size());
            for (size_t i = 0; i < idx.size(); ++i) { const auto& g = (*forest)[idx[i]]; cfg.genes[i].build(g.second, &g.first, nullptr, cfg.P.bias != 0 && cfg.P.tau > 0); }
        } else {
        std::vector<std::string> files = list_dir(p.positional[1]);
        if (files.empty()) throw SgpError("no gene trees in directory " + p.positional[1]);
        cfg.genes.resize(files.size());
        for (size_t i = 0; i < files.size(); ++i) { std::string nm = file_name(files[i]); cfg.genes[i].build(parse_newick(slurp(files[i])), &nm, nullptr, cfg.P.bias != 0 && cfg.P.tau > 0); }
        }
        cfg.out_prefix = cfg.quiet ? std::string() : java_trim(p.positional[5]);
        return;
    }
    res.status = 4;
    res.err_text = "Unknown application: " + app + "\n";
}

}  // namespace sgp
