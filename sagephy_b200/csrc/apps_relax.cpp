// BranchRelaxer front-end: reference grammar -> RelaxConfig.
// Mirrors apps/SaGePhy/BranchRelaxerParameters.java:28-117 (options, model factory, tree loading) and
// BranchRelaxer.java:79-93 (RTree / names / lengths preparation, root without length -> 0 and ignored by the range check).
#include <climits>
#include <set>
#include "apps_common.hpp"

namespace sgp {
using namespace cli;

namespace {
const Opt kRelaxOpts[] = {{"-h", "--help", 0}, {"-o", "--output-file", 1}, {"-s", "--seed", 1}, {"-x", "--auxiliary-tags", 0},
    {"-innms", "--keep-interior-names", 0}, {"-min", "--min-rate", 1}, {"-max", "--max-rate", 1}, {"-a", "--max-attempts", 1}};

bool same_nocase(const std::string& a, const char* b) {
    size_t n = 0; while (b[n]) ++n;
    if (a.size() != n) return false;
    for (size_t i = 0; i < n; ++i) { char x = a[i], y = b[i]; if (x >= 'A' && x <= 'Z') x += 32; if (y >= 'A' && y <= 'Z') y += 32; if (x != y) return false; }
    return true;
}
// iteration order of a java.util.HashMap<String,String> filled by put() in the given order
std::vector<std::pair<std::string, std::string>> java_hashmap_order(const std::vector<std::pair<std::string, std::string>>& kv, int initial_capacity) {
    auto hash = [](const std::string& s) { uint32_t h = 0; for (unsigned char c : s) h = 31u * h + c; return h ^ (h >> 16); };
    int cap = 1; while (cap < initial_capacity) cap <<= 1;
    size_t count = 0;
    for (size_t i = 0; i < kv.size(); ++i) { ++count; if ((double)count > 0.75 * cap) cap <<= 1; }
    std::vector<std::pair<std::string, std::string>> out;     // bucket order, insertion order inside a bucket
    for (int b = 0; b < cap; ++b) for (auto& e : kv) if ((int)(hash(e.first) & (uint32_t)(cap - 1)) == b) out.push_back(e);
    return out;
}
// strict = PrIMENewickTreeVerifier.runStandardTestSuite (io/PrIMENewickTreeVerifier.java:37-46), which the reference runs for the
// models whose lengthsMustBeUltrametric() is true; for plain Newick input it can fail on leaf names (:79-94) and on missing
// branch lengths (:101-103,143-152). Arc/vertex-time tags are not modelled.
RelaxTreeTemplate make_template(const FlatTree& t, bool keep_interior, bool strict) {
    RelaxTreeTemplate r; r.nV = t.n; r.root = t.root;
    if (strict) {
        bool any_name = false, any_bl = false;
        for (int v = 0; v < t.n; ++v) { any_name |= (bool)t.has_name[v]; any_bl |= (bool)t.has_bl[v]; }
        if (any_name) {
            std::set<std::string> seen;
            for (int i = 0; i < t.n; ++i) {
                const int v = t.bfs[i];
                if (t.n_children[v] != 0) continue;
                if (!t.has_name[v]) throw SgpError("NewickIOException: Missing name on Newick vertex " + std::to_string(v) + ".");
                if (!seen.insert(t.name[v]).second) throw SgpError("NewickIOException: Duplicate vertex name '" + t.name[v] + "' found in vertex " + std::to_string(v) + ".");
            }
        }
        if (any_bl) for (int v = 0; v < t.n; ++v) if (!t.has_bl[v] && v != t.root) throw SgpError("NewickIOException: Missing branch length on Newick vertex " + std::to_string(v) + ".");
    }
    r.topo.assign(t.bfs.begin(), t.bfs.end()); r.parent.assign(t.parent.begin(), t.parent.end());
    for (int v = 0; v < t.n; ++v) {
        if (t.n_children[v] == 1) throw SgpError("Cannot create RTree from collapsable NewickTree.");
        if (t.n_children[v] == 0 && !t.has_name[v]) throw SgpError("Missing leaf name in vertex " + std::to_string(v) + ".");
    }
    r.orig.assign(t.n, std::nan("")); bool any = false;
    for (int v = 0; v < t.n; ++v) if (t.has_bl[v]) { r.orig[v] = t.bl[v]; any = true; }
    if (!any) throw SgpError("Cannot create time map; times missing in tree.");
    if (std::isnan(r.orig[r.root])) { r.orig[r.root] = 0.0; r.ignore = r.root; }
    r.chain.assign(t.n, 0); r.flags.assign(t.n, 0); r.names.resize(t.n);
    for (int i = 0; i < t.n; ++i) {          // BFS order: parents before children
        const int v = t.bfs[i]; const int p = t.parent[v];
        const bool first = p >= 0 && t.first_child[p] == v;
        r.chain[v] = (int16_t)(first ? r.chain[p] + 1 : 0);
        uint8_t f = 0;
        if (t.n_children[v] == 0) f |= 1;
        if (p >= 0 && t.next_sibling[v] >= 0) f |= 2;
        if (t.has_name[v] && (t.n_children[v] == 0 || keep_interior)) { f |= 4; r.names[v] = t.name[v]; }
        r.flags[v] = f;
    }
    return r;
}

// ---- IIDRK11 --------------------------------------------------------------------------------------------------------------------
// What the constructor of apps/SaGePhy/IIDRasmussenKellis11RateModel.java:60-101 derives from its inputs, plus the part of getRates
// (:110-157) that does not depend on the random draws: which host arcs the branch above every guest vertex passes over, lowest
// first, and the weight of each. Times come from branch lengths only (io/PrIMENewickTree.java:733-779 with the rule of
// io/PrIMENewickTreeVerifier.java:237-270); VT= / AT= / TT= tags are not modelled.
struct RkTree {
    int n = 0, root = -1; std::vector<int> parent, left, right; std::vector<double> vt, at; std::vector<double> k, theta; std::vector<uint8_t> has_params;
    void build(const FlatTree& t) {
        n = t.n; root = t.root; parent = t.parent; left.assign(n, -1); right.assign(n, -1);
        for (int v = 0; v < n; ++v) {
            if (t.n_children[v] != 0 && t.n_children[v] != 2) throw SgpError("Cannot create RBTree from non-bifurcating NewickTree.");
            if (t.n_children[v] == 2) { left[v] = t.first_child[v]; right[v] = t.next_sibling[left[v]]; }
        }
        at.assign(n, std::nan("")); vt.assign(n, std::nan("")); bool any = false;
        for (int v = 0; v < n; ++v) if (t.has_bl[v]) { at[v] = t.bl[v]; any = true; }
        if (!any) throw SgpError("Cannot create time map; times missing in tree.");
        for (int i = n - 1; i > 0; --i) {
            const int v = t.bfs[i];
            if (t.n_children[v] == 0) vt[v] = 0.0;
            if (!std::isnan(vt[v])) { const int a = parent[v]; if (std::isnan(vt[a])) vt[a] = vt[v] + at[v]; }
        }
        if (n == 1) vt[0] = 0.0;
        k.assign(n, std::nan("")); theta.assign(n, std::nan("")); has_params.assign(n, 0);
        for (int v = 0; v < n; ++v) {         // PARAMS="?(k,theta)"? (io/PrIMENewickTree.java:74,382-387)
            if (!t.has_meta[v]) continue;
            const std::string& m = t.meta[v]; size_t pos = m.find("PARAMS=");
            if (pos == std::string::npos) continue;
            size_t q = pos + 7; if (q < m.size() && m[q] == '"') ++q;
            if (q >= m.size() || m[q] != '(') continue;
            const size_t e = m.find(')', q); if (e == std::string::npos) continue;
            std::string body; for (size_t j = q + 1; j < e; ++j) if (m[j] != ' ') body.push_back(m[j]);
            const size_t c = body.find(','); if (c == std::string::npos) continue;
            k[v] = to_double(body.substr(0, c)); theta[v] = to_double(body.substr(c + 1)); has_params[v] = 1;
        }
    }
    int lca(int x, int y) const {            // topology/RBTree.java:243-258
        std::vector<int> seen;
        for (;;) {
            if (x != -1) { for (int s : seen) if (s == x) return x; seen.push_back(x); x = parent[x]; }
            if (y != -1) { for (int s : seen) if (s == y) return y; seen.push_back(y); y = parent[y]; }
        }
    }
};

void build_rk11(const FlatTree& gt, const FlatTree& ht, const std::string& map_path, double scale, RelaxConfig& cfg) {
    RkTree G, H; std::vector<int> sigma;
    try {
        H.build(ht); G.build(gt);
        // io/GuestHostMapReader.java:44-56 into the LinkedHashMap of topology/GuestHostMap.java:25-34
        std::ifstream f(map_path); if (!f) throw SgpError("FileNotFoundException: " + map_path + " (No such file or directory)");
        std::vector<std::pair<std::string, std::string>> map; std::string line;
        while (std::getline(f, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            const std::string ln = java_trim(line);
            if (ln.empty()) break;
            std::vector<std::string> parts; size_t i = 0;
            while (i < ln.size()) { size_t j = i; while (j < ln.size() && ln[j] != ' ' && ln[j] != '\t') ++j; parts.push_back(ln.substr(i, j - i)); i = j; while (i < ln.size() && (ln[i] == ' ' || ln[i] == '\t')) ++i; }
            if (parts.size() < 2) throw SgpError("ArrayIndexOutOfBoundsException: 1");
            bool found = false; for (auto& m : map) if (m.first == parts[0]) { m.second = parts[1]; found = true; break; }
            if (!found) map.emplace_back(parts[0], parts[1]);
        }
        // topology/MPRMap.java:60-86: leaves from the map, the rest by the LCA of the children's host vertices
        sigma.assign(G.n, 0); int n_leaves = 0;
        for (int l = 0; l < G.n; ++l) {
            if (G.left[l] != -1) continue;
            ++n_leaves;
            const std::string gname = gt.has_name[l] ? gt.name[l] : std::string("null");
            const std::string* hn = nullptr; for (auto& m : map) if (m.first == gname) hn = &m.second;
            if (!hn) throw SgpError("Cannot find guest leaf " + gname + " in guest-to-host leaf map.");
            int hv = -1; for (int v = 0; v < H.n; ++v) if (ht.has_name[v] && ht.name[v] == *hn) hv = v;
            if (hv < 0) throw SgpError("Cannot find host leaf " + *hn + " corresponding to guest leaf " + gname + ".");
            sigma[l] = hv;
        }
        if (n_leaves != (int)map.size()) throw SgpError("Mismatch in the number of guest tree leaves and guest-to-host map items: " + std::to_string(n_leaves) + " vs. " + std::to_string(map.size()) + ".");
        for (int i = G.n - 1; i >= 0; --i) { const int x = gt.bfs[i]; if (G.left[x] != -1) sigma[x] = H.lca(sigma[G.left[x]], sigma[G.right[x]]); }      // children before parents
        bool any_params = false; for (int v = 0; v < H.n; ++v) any_params |= H.has_params[v] != 0;
        if (!any_params) throw SgpError("NullPointerException");
        if (std::isnan(H.at[H.root])) H.at[H.root] = 0.0;
        if (!H.has_params[H.root]) { H.k[H.root] = 1000.0; H.theta[H.root] = 0.001; H.has_params[H.root] = 1; }
        if (std::isnan(G.at[G.root])) G.at[G.root] = 0.0;
    } catch (SgpError& e) {
        throw SgpError(std::string("Invalid input to rate model: ") + e.what());
    }
    if (std::fabs(G.vt[G.root] + G.at[G.root] - H.vt[H.root] - H.at[H.root]) > 1e-4) throw SgpError("Invalid guest and host trees: Time scales do not agree.");
    auto enclosing = [&](int x) {
        int lo = sigma[x]; const double xt = G.vt[x];
        for (;;) { if (lo == H.root) return lo; const int lop = H.parent[lo]; if (H.vt[lop] >= xt) return lo; lo = lop; }
    };
    cfg.pa = scale; cfg.rk_off.assign(1, 0);
    for (int x = 0; x < G.n; ++x) {
        const int lo = enclosing(x), up = G.parent[x] == -1 ? H.root : enclosing(G.parent[x]);
        const double l = G.at[x];
        const size_t first = cfg.rk_w.size();
        for (int h = lo;; h = H.parent[h]) {
            if (h < 0) throw SgpError("Invalid guest and host trees: a guest branch leaves the host tree (ArrayIndexOutOfBoundsException in the reference)");
            cfg.rk_k.push_back(H.has_params[h] ? H.k[h] : std::nan("")); cfg.rk_theta.push_back(H.has_params[h] ? H.theta[h] : std::nan(""));
            cfg.rk_w.push_back(H.at[h] / l);
            if (h == up) break;
        }
        const size_t cnt = cfg.rk_w.size() - first;
        if (cnt > 255) throw SgpError("IIDRK11: a guest branch spans more than 255 host arcs (not supported)");
        if (cnt == 1) cfg.rk_w[first] = 1.0;
        else {
            double dt = G.vt[x] - H.vt[lo];
            cfg.rk_w[first] = std::max(0.0, cfg.rk_w[first] - dt / l);
            if (G.parent[x] != -1) { dt = G.vt[G.parent[x]] - H.vt[up]; cfg.rk_w.back() = std::max(0.0, dt / l); }
        }
        cfg.rk_off.push_back((int32_t)cfg.rk_w.size());
    }
}
}  // namespace

void parse_relax_app(const std::vector<std::string>& args, bool multi_tree_input, RelaxConfig& cfg, AppResult& res, bool trees_from_batch) {
    cfg.args = args;
    Parsed p = scan(args, kRelaxOpts, (int)(sizeof kRelaxOpts / sizeof kRelaxOpts[0]));
    if (p.on("-h") || args.empty()) {
        res.status = 6;
        res.out_text = "Usage:\n    java -jar sagephy-X.Y.Z.jar BranchRelaxer [options] <tree> <model> <arg1> <arg2> <...>\n"
                       "Supported models (accelerated): Constant <rate>, IIDGamma <k> <theta>, IIDLogNormal <mu> <sigma2>, IIDNormal <mu> <sigma2>,\n"
                       "    IIDUniform <a> <b>, IIDExponential <lambda>, IIDSamplesFromFile <filename>,\n"
                       "    ACTK98 <start rate> <v>, ACRY07 <start rate> <sigma2>, ACABY02 <start rate>, ACLBPL07 <start rate> <mu> <theta> <sigma>,\n"
                       "    IIDRK11 <host tree> <guest-to-host map> <scale factor>\n\n";
        return;
    }
    auto arg = [&](size_t i) -> const std::string& {
        if (i >= p.positional.size()) throw SgpError("IndexOutOfBoundsException: Index: " + std::to_string(i) + ", Size: " + std::to_string(p.positional.size()));
        return p.positional[i];
    };
    if (p.on("-s")) { cfg.has_seed = true; cfg.seed = parse_seed(p.str("-s", "0")); cfg.seed_arg_index = p.val_index.at("-s"); }
    else { cfg.has_seed = false; cfg.seed = random_seed(); }
    const std::string& model = arg(1);
    std::vector<std::pair<std::string, std::string>> kv; int hcap = 2; bool strict = false;
    if (same_nocase(model, "Constant")) { cfg.model = 0; cfg.pa = to_double(arg(2)); cfg.model_name = "ConstantRates"; kv = {{"rate", jdouble(cfg.pa)}}; hcap = 1; }
    else if (same_nocase(model, "IIDExponential")) {
        cfg.model = 1; cfg.pa = to_double(arg(2)); if (cfg.pa <= 0) throw SgpError("Cannot have non-positive rate for exponential distribution.");
        cfg.model_name = "IIDExponentialRates"; kv = {{"lambda", jdouble(cfg.pa)}}; hcap = 1;
    } else if (same_nocase(model, "IIDGamma")) {
        cfg.model = 2; cfg.pa = to_double(arg(2)); cfg.pb = to_double(arg(3));
        if (cfg.pa <= 0 || cfg.pb <= 0) throw SgpError("Cannot have non-positive shape or scale for gamma distribution.");
        cfg.model_name = "IIDGammaRates"; kv = {{"k", jdouble(cfg.pa)}, {"theta", jdouble(cfg.pb)}};
    } else if (same_nocase(model, "IIDLogNormal")) {
        cfg.model = 3; cfg.pa = to_double(arg(2)); const double s2 = to_double(arg(3));
        if (s2 <= 0) throw SgpError("Cannot have non-positive sigma^2 for log-normal distribution.");
        cfg.sigma = std::sqrt(s2); cfg.model_name = "IIDLogNormalRates"; kv = {{"mu", jdouble(cfg.pa)}, {"sigma2", jdouble(cfg.sigma * cfg.sigma)}};
    } else if (same_nocase(model, "IIDNormal")) {
        cfg.model = 4; cfg.pa = to_double(arg(2)); cfg.pb = to_double(arg(3));
        if (cfg.pb <= 0) throw SgpError("Cannot have non-positive variance for normal distribution.");
        cfg.sigma = std::sqrt(cfg.pb); cfg.model_name = "IIDNormalRates"; kv = {{"mu", jdouble(cfg.pa)}, {"sigma2", jdouble(cfg.pb)}};
    } else if (same_nocase(model, "IIDUniform")) {
        cfg.model = 5; cfg.pa = to_double(arg(2)); cfg.pb = to_double(arg(3));
        if (cfg.pa >= cfg.pb) throw SgpError("Invalid range for uniform distribution.");
        cfg.model_name = "IIDUniformRates"; kv = {{"a", jdouble(cfg.pa)}, {"b", jdouble(cfg.pb)}};
    } else if (same_nocase(model, "IIDSamplesFromFile")) {
        cfg.model = 6; cfg.model_name = "IIDSamplesFromFileRates";
        std::ifstream f(arg(2)); if (!f) throw SgpError("java.io.FileNotFoundException: " + arg(2));
        std::string line; while (std::getline(f, line)) cfg.samples.push_back(to_double(line));
        if (cfg.samples.empty()) throw SgpError("samples file is empty: " + arg(2));
        char buf[PATH_MAX]; const char* rp = realpath(arg(2).c_str(), buf);
        kv = {{"filename", rp ? std::string(rp) : arg(2)}};
    } else if (same_nocase(model, "ACTK98")) {          // apps/SaGePhy/ACThorneKishino98RateModel.java
        cfg.model = 7; cfg.pa = to_double(arg(2)); cfg.pb = to_double(arg(3)); cfg.model_name = "ThorneKishino98Rates";
        kv = {{"start rate", jdouble(cfg.pa)}, {"v", jdouble(cfg.pb)}}; strict = true;
    } else if (same_nocase(model, "ACRY07")) {          // apps/SaGePhy/ACRannalaYang07RateModel.java
        cfg.model = 8; cfg.pa = to_double(arg(2)); cfg.pb = to_double(arg(3)); cfg.model_name = "RannalaYang07Rates";
        kv = {{"start rate", jdouble(cfg.pa)}, {"sigma2", jdouble(cfg.pb)}}; strict = true;
    } else if (same_nocase(model, "ACABY02")) {         // apps/SaGePhy/ACArisBrosouYang02RateModel.java
        cfg.model = 9; cfg.pa = to_double(arg(2)); cfg.model_name = "ArisBrosouYang02Rates"; kv = {{"start rate", jdouble(cfg.pa)}}; hcap = 1;
    } else if (same_nocase(model, "ACLBPL07")) {        // apps/SaGePhy/ACLepageBryantPhillipeLartillot07RateModel.java
        cfg.model = 10; cfg.pa = to_double(arg(2)); cfg.pb = to_double(arg(3)); cfg.pc = to_double(arg(4)); cfg.pd = to_double(arg(5));
        if (2 * cfg.pc * cfg.pb < cfg.pd * cfg.pd) throw SgpError("Invalid parameters for CIR process: 0 rate may be achieved.");
        cfg.model_name = "LepageBryantPhillipeLartillot07Rates";
        kv = {{"start rate", jdouble(cfg.pa)}, {"mu", jdouble(cfg.pb)}, {"theta", jdouble(cfg.pc)}, {"sigma", jdouble(cfg.pd)}};
    } else if (same_nocase(model, "IIDRK11")) {         // apps/SaGePhy/IIDRasmussenKellis11RateModel.java; tables built below from the trees
        cfg.model = 11; cfg.model_name = "IIDRasmussenKellis11Rates"; strict = true; hcap = 2;
        (void)arg(2); (void)arg(3); (void)to_double(arg(4));
        if (trees_from_batch || multi_tree_input) throw SgpError("IIDRK11 relaxes one guest tree against one host tree: chained and multi-tree input are not supported");
    } else throw SgpError("Invalid rate model identifier: " + model);
    cfg.params = java_hashmap_order(kv, hcap);
    cfg.min_rate = to_double(p.str("-min", "1e-64")); cfg.max_rate = to_double(p.str("-max", "1e64")); cfg.max_attempts = to_int(p.str("-a", "10000"));
    cfg.do_meta = p.on("-x"); cfg.keep_interior = p.on("-innms");
    cfg.has_outfile = p.on("-o"); if (cfg.has_outfile) cfg.outfile = java_trim(p.str("-o", ""));
    if (trees_from_batch) { (void)arg(0); return; }      // chained stage: the trees are already on the device
    const std::string text = exists(arg(0)) ? slurp(arg(0)) : arg(0);
    if (cfg.model == 11) {
        const FlatTree gt = parse_newick(text);
        cfg.trees.push_back(make_template(gt, cfg.keep_interior, true));
        const FlatTree ht = parse_newick(exists(arg(2)) ? slurp(arg(2)) : arg(2));
        (void)make_template(ht, false, true);          // the host tree is read with the same strict checks (BranchRelaxerParameters.java:109-110)
        build_rk11(gt, ht, arg(3), to_double(arg(4)), cfg);
    } else if (!multi_tree_input) cfg.trees.push_back(make_template(parse_newick(text), cfg.keep_interior, strict));
    else {
        size_t b = 0;
        while (b < text.size()) {
            size_t e = text.find('\n', b); if (e == std::string::npos) e = text.size();
            std::string line = text.substr(b, e - b); b = e + 1;
            if (line.find_first_not_of(" \t\r") == std::string::npos) continue;
            cfg.trees.push_back(make_template(parse_newick(line), cfg.keep_interior, strict));
        }
        if (cfg.trees.empty()) throw SgpError("no tree found in the multi-tree input");
    }
}

}  // namespace sgp
