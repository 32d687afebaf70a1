// ORACLE — test infrastructure only (see oracle/README.md). PARITY UNPINNED.
//
// BranchRelaxer and its IID rate models. Restates
//   apps/SaGePhy/BranchRelaxer.java:48-141,148-197        (main loop, isOK, getRelaxedLengths, toNewickTree)
//   apps/SaGePhy/BranchRelaxerParameters.java:80-117,167-179 (model factory, tree loading)
//   apps/SaGePhy/{Constant,IIDExponential,IIDGamma,IIDLogNormal,IIDNormal,IIDUniform,IIDSamplesFromFile}RateModel.java
//   math/LogNormalDistribution.java:39-48,196-205, math/Normal.java:17-36,48-86, math/NormalDistribution.java:44-54,217-221,
//   math/GammaDistribution.java:53-62,255-257, math/Gamma.java:16-57,83-88,127-204, math/ChiSquared.java:23-84,
//   math/UniformDistribution.java:40-49,210-212, math/RealInterval.java:177-240, math/ExponentialDistribution.java:79-81,155-159
//   topology/RTree.java:53-86, io/NewickTreeWriter.java:100-113, topology/DoubleMap.java:183-185 (Arrays.toString)
//   java.util.Random.nextGaussian / nextBoolean (JDK specification: polar method on StrictMath.log / sqrt)
#pragma once
#include "apps.hpp"
#include "jpow.hpp"
#include <set>

namespace oracle {

inline double JavaMT::nextGaussian() {
    if (have_gauss) { have_gauss = false; return next_gauss; }
    double v1, v2, s;
    do { v1 = 2 * nextDouble() - 1; v2 = 2 * nextDouble() - 1; s = v1 * v1 + v2 * v2; } while (s >= 1 || s == 0);
    double multiplier = std::sqrt(-2 * j_log(s) / s);
    next_gauss = v2 * multiplier; have_gauss = true;
    return v1 * multiplier;
}

inline double normal_cdf(double x) {     // math/Normal.java:17-36
    if (x < -39) return 0.0;
    if (x > 9) return 1.0;
    const double b1 = 0.319381530, b2 = -0.356563782, b3 = 1.781477937, b4 = -1.821255978, b5 = 1.330274429, p = 0.2316419, c = 0.39894228;
    if (x >= 0.0) { double t = 1.0 / (1.0 + p * x); return (1.0 - c * j_exp(-x * x / 2.0) * t * (t * (t * (t * (t * b5 + b4) + b3) + b2) + b1)); }
    double t = 1.0 / (1.0 - p * x);
    return (c * j_exp(-x * x / 2.0) * t * (t * (t * (t * (t * b5 + b4) + b3) + b2) + b1));
}
inline double normal_quantile(double p) {    // math/Normal.java:48-86 (Voutier 2010)
    if (p < 0.0 || p > 1.0) throw std::invalid_argument("IllegalArgumentException: Cannot compute quantile for probability not in [0,1].");
    if (0.025 <= p && p <= 0.975) {
        const double a0 = 0.151015505647689, a1 = -0.5303572634357367, a2 = 1.365020122861334, b0 = 0.132089632343748, b1 = -0.7607324991323768;
        double q = p - 0.5; double r = j_pow(q, 2);
        return (q * (a2 + (a1 * r + a0) / (r * r + b1 * r + b0)));
    }
    if (1e-50 < p && p < 1.0 - 1e-16) {
        const double c3 = -1.000182518730158122, cp0 = 16.682320830719986527, cp1 = 4.120411523939115059, cp2 = 0.029814187308200211,
                     d0 = 7.173787663925508066, d1 = 8.759693508958633869;
        if (p < 0.5) { double r = std::sqrt(j_log(1.0 / j_pow(p, 2))); return (c3 * r + cp2 + (cp1 * r + cp0) / (r * r + d1 * r + d0)); }
        double r = std::sqrt(j_log(1.0 / j_pow(1.0 - p, 2)));
        return -(c3 * r + cp2 + (cp1 * r + cp0) / (r * r + d1 * r + d0));
    }
    return (p < 0.5 ? -HUGE_VAL : HUGE_VAL);
}
inline double ln_gamma(double xx) {      // math/Gamma.java:44-57
    static const double LANCZOS[14] = {57.1562356658629235, -59.5979603554754912, 14.1360979747417471, -0.491913816097520199, .339946499848118887e-4,
        .465236289270485756e-4, -.983744753048795646e-4, .158088703224912494e-3, -.210264441724104883e-3, .217439618115212643e-3,
        -.164318106536763890e-3, .844182239838527433e-4, -.261908384015814087e-4, .368991826595316234e-5};
    const double SQRT_2_PI = std::sqrt(2 * 3.141592653589793);
    double x = xx, y = xx;
    double tmp = x + 5.24218750000000000;
    tmp = (x + 0.5) * j_log(tmp) - tmp;
    double ser = 0.999999999999997092;
    for (int j = 0; j < 14; ++j) ser += LANCZOS[j] / ++y;
    return (tmp + j_log(SQRT_2_PI * ser / x));
}
inline double inc_gamma_ratio(double x, double k) {     // math/Gamma.java:127-204 (AS 32)
    if (x < 0.0 || k <= 0.0) throw std::invalid_argument("IllegalArgumentException: Invalid parameters for incomplete gamma ratio.");
    if (x == 0.0) return 0.0;
    const double accuracy = 1.0e-8, overflow = 1.0e30, xbig = 1.0E6, alimit = 1000.0;
    double g_in = 0.0, g = ln_gamma(k);
    double factor = j_exp(k * j_log(x) - x - g);
    if (k > alimit) { double pn1 = 3 * std::sqrt(k) * (j_pow(x / k, 1.0 / 3.0) + 1.0 / (9.0 * k) - 1.0); return normal_cdf(pn1); }
    if (x > xbig) return 1.0;
    if (x > 1.0 && x >= k) {
        double a = 1.0 - k, b = a + x + 1.0, term = 0.0, pn[6];
        pn[0] = 1.0; pn[1] = x; pn[2] = x + 1.0; pn[3] = x * b;
        g_in = pn[2] / pn[3];
        double dif = 1.0, rn = 0.0;
        for (;;) {
            a += 1.0; b += 2.0; term += 1.0;
            double an = a * term;
            for (int i = 0; i < 2; i++) pn[i + 4] = b * pn[i + 2] - an * pn[i];
            if (pn[5] != 0.0) {
                rn = pn[4] / pn[5]; dif = std::fabs(g_in - rn);
                if (dif <= accuracy) { if (dif <= accuracy * rn) return 1.0 - factor * g_in; }
                g_in = rn;
            }
            for (int i = 0; i < 4; i++) pn[i] = pn[i + 2];
            if (std::fabs(pn[4]) >= overflow) for (int i = 0; i < 4; i++) pn[i] = pn[i] / overflow;
        }
    }
    g_in = 1.0; double term = 1.0, rn = k;
    do { rn += 1.0; term *= x / rn; g_in += term; } while (term > accuracy);
    return g_in * factor / k;
}
inline double chi2_quantile(double p, double k) {     // math/ChiSquared.java:23-84 (AS 91)
    if (p < 2e-6 || p > 0.999998 || k <= 0.0) throw std::invalid_argument("IllegalArgumentException: Invalid arguments for chi^2 quantile function.");
    const double E = 0.5e-6, AA = j_log(2.0);
    double XX = 0.5 * k, C = XX - 1.0, G = ln_gamma(XX), ch;
    if (k < (-1.24 * j_log(p))) {
        ch = j_pow(p * XX * j_exp(G + XX * AA), 1.0 / XX);
        if (ch < E) return ch;
    } else if (k <= 0.32) {
        ch = 0.4; double a = j_log(1.0 - p), Q;
        do {
            Q = ch;
            double P1 = 1.0 + ch * (4.67 + ch), P2 = ch * (6.73 + ch * (6.66 + ch));
            double T = -0.5 + (4.67 + 2.0 * ch) / P1 - (6.73 + ch * (13.32 + 3.0 * ch)) / P2;
            ch = ch - (1.0 - j_exp(a + G + 0.5 * ch + C * AA) * P2 / P1) / T;
        } while (std::fabs(Q / ch - 1) > 0.01);
    } else {
        double X = normal_quantile(p), P1 = 0.222222 / k;
        ch = k * j_pow((X * std::sqrt(P1) + 1.0 - P1), 3);
        if (ch > (2.2 * k + 6.0)) ch = -2.0 * (j_log(1.0 - p) - C * j_log(0.5 * ch) + G);
    }
    double Q;
    do {
        Q = ch;
        double P1 = 0.5 * ch, P2 = p - inc_gamma_ratio(P1, XX);
        double T = P2 * j_exp(XX * AA + G + P1 - C * j_log(ch));
        double B = T / ch, A = 0.5 * T - B * C;
        double S1 = (210 + A * (140 + A * (105 + A * (84 + A * (70 + 60 * A))))) / 420;
        double S2 = (420 + A * (735 + A * (966 + A * (1141 + 1278 * A)))) / 2520;
        double S3 = (210 + A * (462 + A * (707 + 932 * A))) / 2520;
        double S4 = (252 + A * (672 + 1182 * A) + C * (294 + A * (889 + 1740 * A))) / 5040;
        double S5 = (84 + 264 * A + C * (175 + 606 * A)) / 2520;
        double S6 = (120 + C * (346 + 127 * C)) / 5040;
        ch = ch + T * (1 + 0.5 * T * S1 - B * C * (S1 - B * (S2 - B * (S3 - B * (S4 - B * (S5 - B * S6))))));
    } while (std::fabs(Q / ch - 1) > E);
    return ch;
}

struct RateModelO {
    enum Kind { CONSTANT, EXPONENTIAL, GAMMA, LOGNORMAL, NORMAL, UNIFORM, SAMPLES, ACTK98, ACRY07, ACABY02, ACLBPL07, IIDRK11 } kind = CONSTANT;
    double a = 0, b = 0, c = 0, d = 0, sigma = 0; std::vector<double> samples; std::string file;
    std::string name() const {
        static const char* n[] = {"ConstantRates", "IIDExponentialRates", "IIDGammaRates", "IIDLogNormalRates", "IIDNormalRates", "IIDUniformRates", "IIDSamplesFromFileRates",
                                  "ThorneKishino98Rates", "RannalaYang07Rates", "ArisBrosouYang02Rates", "LepageBryantPhillipeLartillot07Rates", "IIDRasmussenKellis11Rates"};
        return n[kind];
    }
    bool strict_tree() const { return kind == ACTK98 || kind == ACRY07 || kind == IIDRK11; }      // lengthsMustBeUltrametric()
    // new LogNormalDistribution(mu, sigma2).sampleValue(prng)  (math/LogNormalDistribution.java:39-48,196-205)
    static double lognormal(double mu, double sigma2, JavaMT& prng) {
        if (sigma2 <= 0.0) throw std::invalid_argument("IllegalArgumentException: Cannot have non-positive sigma^2 for log-normal distribution.");
        const double sg = std::sqrt(sigma2);
        return j_exp(normal_quantile(prng.nextDouble()) * sg + mu);
    }
    // getRates of the autocorrelated models: vertices in RTree.getTopologicalOrdering() order (breadth first from the root,
    // topology/RTree.java:488-500); topo[i] / parent[] hold vertex numbers.
    void rates_autocorrelated(const std::vector<int>& topo, const std::vector<int>& parent, const std::vector<double>& len, JavaMT& prng, std::vector<double>& rates) const {
        const int n = (int)topo.size();
        if (kind == ACTK98) {                       // apps/SaGePhy/ACThorneKishino98RateModel.java:58-74
            rates[topo[0]] = a;
            for (int i = 1; i < n; ++i) {
                const int x = topo[i], xp = parent[x];
                const double deltat = (len[x] + len[xp]) / 2;
                const double sigma2 = b * deltat;
                rates[x] = lognormal(j_log(rates[xp]) - sigma2 / 2, sigma2, prng);
            }
        } else if (kind == ACRY07) {                // apps/SaGePhy/ACRannalaYang07RateModel.java:56-88
            std::vector<double> vr(n);
            const int root = topo[0]; double dt = len[root];
            if (dt == 0.0) { rates[root] = a; vr[root] = a; }
            else {
                double r = lognormal(j_log(a) - dt * b / 2, dt * b, prng); rates[root] = r;
                r = lognormal(j_log(r) - dt * b / 2, dt * b, prng); vr[root] = r;
            }
            for (int i = 1; i < n; ++i) {
                const int x = topo[i], xp = parent[x];
                dt = len[x] / 2;
                double r = lognormal(j_log(vr[xp]) - dt * b / 2, dt * b, prng); rates[x] = r;
                r = lognormal(j_log(r) - dt * b / 2, dt * b, prng); vr[x] = r;
            }
        } else if (kind == ACABY02) {               // apps/SaGePhy/ACArisBrosouYang02RateModel.java:49-64
            rates[topo[0]] = a;
            for (int i = 1; i < n; ++i) {
                const int x = topo[i], xp = parent[x];
                const double lambda = 1.0 / rates[xp];
                if (lambda <= 0) throw std::invalid_argument("IllegalArgumentException: Invalid rate for exponential distribution.");
                rates[x] = -j_log(1.0 - prng.nextDouble()) / lambda;
            }
        } else {                                    // apps/SaGePhy/ACLepageBryantPhillipeLartillot07RateModel.java:75-115 (CIR, 200 Euler steps per branch)
            const int STEPS = 200;
            std::vector<double> vr(n);
            for (int j = 0; j < n; ++j) {
                const int x = topo[j];
                double ir = j == 0 ? a : vr[parent[x]];
                double r = 0.0;
                const double dt = len[x] / STEPS;
                if (j == 0 && dt == 0.0) { vr[x] = a; rates[x] = a; continue; }
                if (dt <= 0.0) throw std::invalid_argument("IllegalArgumentException: Cannot have non-positive variance for normal distribution.");
                const double sd = std::sqrt(dt);
                for (int i = 0; i < STEPS; ++i) {
                    const double root_ir = std::sqrt(ir);
                    ir = ir + c * (b - ir) * dt + d * root_ir * (prng.nextGaussian() * sd + 0.0);
                    r += ir;
                }
                vr[x] = ir; rates[x] = r / STEPS;
            }
        }
    }
    // java.util.HashMap iteration order of the parameter names (bucket = (h ^ h>>>16) & (cap-1)); all key sets here are
    // single-character-hash friendly: {"mu","sigma2"}: hash("mu")=3496 -> bucket 0 of 2... computed in params() explicitly.
    std::vector<std::pair<std::string, std::string>> params() const;
    double sample(JavaMT& prng) const {
        switch (kind) {
            case CONSTANT: return a;
            case EXPONENTIAL: return -j_log(1.0 - prng.nextDouble()) / a;
            case GAMMA: return chi2_quantile(prng.nextDouble(), 2 * a) * b / 2;
            case LOGNORMAL: return j_exp(normal_quantile(prng.nextDouble()) * sigma + a);
            case NORMAL: return prng.nextGaussian() * sigma + a;
            case UNIFORM: { double d = prng.nextDouble(); return prng.nextBoolean() ? (b - a) * d + a : (b - a) * (1.0 - d) + a; }
            case SAMPLES: return samples[prng.nextInt((int)samples.size())];
            default: break;      // autocorrelated models: rates_autocorrelated
        }
        return 0;
    }
};

inline int java_string_hash(const std::string& s) { int32_t h = 0; for (unsigned char c : s) h = 31 * h + c; return h; }
inline std::vector<std::pair<std::string, std::string>> hashmap_order(std::vector<std::pair<std::string, std::string>> kv, int initialCapacity) {
    // HashMap(initialCapacity): table size = next power of two >= initialCapacity (min 1), doubled while size > 0.75 * cap
    int cap = 1; while (cap < initialCapacity) cap <<= 1;
    size_t inserted = 0; std::vector<std::vector<std::pair<std::string, std::string>>> tab(cap);
    auto bucket = [&](const std::string& k, int c) { uint32_t h = (uint32_t)java_string_hash(k); h ^= (h >> 16); return (int)(h & (uint32_t)(c - 1)); };
    for (auto& e : kv) {
        tab[bucket(e.first, cap)].push_back(e); ++inserted;
        if ((double)inserted > 0.75 * cap) {
            int nc = cap * 2; std::vector<std::vector<std::pair<std::string, std::string>>> nt(nc);
            for (auto& bk : tab) for (auto& x : bk) nt[bucket(x.first, nc)].push_back(x);
            tab.swap(nt); cap = nc;
        }
    }
    std::vector<std::pair<std::string, std::string>> out;
    for (auto& bk : tab) for (auto& x : bk) out.push_back(x);
    return out;
}
inline std::vector<std::pair<std::string, std::string>> RateModelO::params() const {
    auto D = [](double v) { return java_double_to_string(v); };
    switch (kind) {
        case CONSTANT: return hashmap_order({{"rate", D(a)}}, 1);
        case EXPONENTIAL: return hashmap_order({{"lambda", D(a)}}, 1);
        case GAMMA: return hashmap_order({{"k", D(a)}, {"theta", D(b)}}, 2);
        case LOGNORMAL: return hashmap_order({{"mu", D(a)}, {"sigma2", D(sigma * sigma)}}, 2);
        case NORMAL: return hashmap_order({{"mu", D(a)}, {"sigma2", D(b)}}, 2);
        case UNIFORM: return hashmap_order({{"a", D(a)}, {"b", D(b)}}, 2);
        case SAMPLES: return hashmap_order({{"filename", file}}, 2);
        case ACTK98: return hashmap_order({{"start rate", D(a)}, {"v", D(b)}}, 2);
        case ACRY07: return hashmap_order({{"start rate", D(a)}, {"sigma2", D(b)}}, 2);
        case ACABY02: return hashmap_order({{"start rate", D(a)}}, 1);
        case ACLBPL07: return hashmap_order({{"start rate", D(a)}, {"mu", D(b)}, {"theta", D(c)}, {"sigma", D(d)}}, 2);
        case IIDRK11: return {};       // getModelParameters returns an empty map (IIDRasmussenKellis11RateModel.java:104-108)
    }
    return {};
}

// ---- IIDRK11: host-specific gamma rates (apps/SaGePhy/IIDRasmussenKellis11RateModel.java:60-180) --------------------------------------
// Everything the constructor derives from the two trees and the leaf map: RBTree arrays (topology/RBTree.java:56-85), times from
// branch lengths (io/PrIMENewickTree.java:733-779 with io/PrIMENewickTreeVerifier.java:237-270: scanning the BFS list backwards, the
// first child to report sets its parent's time), PARAMS=(k,theta) per host vertex (io/PrIMENewickTree.java:74,382-387), the
// most-parsimonious reconciliation sigma (topology/MPRMap.java:60-86,147-156 on the naive LCA of topology/RBTree.java:243-258).
struct RK11Tree {
    int n = 0, root = -1;
    std::vector<int> parent, left, right; std::vector<double> vt, at; std::vector<std::optional<std::string>> names;
    std::vector<std::optional<std::pair<double, double>>> params;
    void build(const NTree& t) {
        auto bfs = t.vertices_bfs(); n = (int)bfs.size(); root = t.root->number;
        parent.assign(n, -1); left.assign(n, -1); right.assign(n, -1); names.resize(n); params.resize(n);
        for (NVertex* v : bfs) if (!v->children.empty() && v->children.size() != 2) throw TopologyException("Cannot create RBTree from non-bifurcating NewickTree.");
        bool any_bl = false;
        at.assign(n, std::nan("")); vt.assign(n, std::nan(""));
        for (NVertex* v : bfs) {
            const int i = v->number;
            if (v->parent) parent[i] = v->parent->number;
            if (!v->children.empty()) { left[i] = v->children[0]->number; right[i] = v->children[1]->number; }
            names[i] = v->name;
            if (v->bl) { at[i] = *v->bl; any_bl = true; }
            if (v->meta) {
                // PARAMS="?(\([0-9+-.eE ]+[,0-9+-.eE ]+\))"? -> "[k,theta]" -> doubles
                const std::string& m = *v->meta; size_t pos = m.find("PARAMS=");
                if (pos != std::string::npos) {
                    size_t k = pos + 7; if (k < m.size() && m[k] == '"') ++k;
                    if (k < m.size() && m[k] == '(') {
                        size_t e = m.find(')', k);
                        if (e != std::string::npos) {
                            std::string body; for (size_t q = k + 1; q < e; ++q) if (m[q] != ' ') body.push_back(m[q]);
                            size_t c = body.find(',');
                            if (c != std::string::npos) params[i] = std::make_pair(parse_double(body.substr(0, c)), parse_double(body.substr(c + 1)));
                        }
                    }
                }
            }
        }
        if (!any_bl) throw NewickIOException("Cannot create time map; times missing in tree.");
        for (int i = n - 1; i > 0; --i) {
            NVertex* v = bfs[i]; const int x = v->number;
            if (v->is_leaf()) vt[x] = 0.0;
            if (!std::isnan(vt[x])) { const int a = v->parent->number; if (std::isnan(vt[a])) vt[a] = vt[x] + at[x]; }
        }
        if (n == 1) vt[0] = 0.0;
    }
    bool has_params() const { for (auto& p : params) if (p) return true; return false; }
    int lca(int x, int y) const {
        std::vector<int> visited;
        for (;;) {
            if (x != -1) { if (std::find(visited.begin(), visited.end(), x) != visited.end()) return x; visited.push_back(x); x = parent[x]; }
            if (y != -1) { if (std::find(visited.begin(), visited.end(), y) != visited.end()) return y; visited.push_back(y); y = parent[y]; }
        }
    }
};
struct RK11Model {
    RK11Tree host, guest; std::vector<int> sigma; double scale = 1.0;
    void init(const NTree& g, const NTree& h, const std::vector<std::pair<std::string, std::string>>& gs, double scaleFact) {
        try {
            host.build(h); guest.build(g); scale = scaleFact;
            // LinkedHashMap semantics of GuestHostMap.add: a repeated guest leaf keeps its position and takes the last host leaf
            std::vector<std::pair<std::string, std::string>> map;
            for (auto& e : gs) { bool found = false; for (auto& m : map) if (m.first == e.first) { m.second = e.second; found = true; break; } if (!found) map.push_back(e); }
            sigma.assign(guest.n, 0);
            int n_leaves = 0;
            for (int l = 0; l < guest.n; ++l) {
                if (guest.left[l] != -1) continue;
                ++n_leaves;
                const std::string gname = guest.names[l] ? *guest.names[l] : std::string("null");
                const std::string* hn = nullptr; for (auto& m : map) if (m.first == gname) hn = &m.second;
                if (!hn) throw std::invalid_argument("IllegalArgumentException: Cannot find guest leaf " + gname + " in guest-to-host leaf map.");
                int hv = -1; for (int v = 0; v < host.n; ++v) if (host.names[v] && *host.names[v] == *hn) hv = v;       // HashMap.put: the last vertex with that name wins
                if (hv < 0) throw std::invalid_argument("IllegalArgumentException: Cannot find host leaf " + *hn + " corresponding to guest leaf " + gname + ".");
                sigma[l] = hv;
            }
            if (n_leaves != (int)map.size()) throw std::invalid_argument("IllegalArgumentException: Mismatch in the number of guest tree leaves and guest-to-host map items: " + std::to_string(n_leaves) + " vs. " + std::to_string(map.size()) + ".");
            compute_sigma(guest.root);
            if (!host.has_params()) throw std::runtime_error("NullPointerException");        // getVertexParamsMap returns null
            if (std::isnan(host.at[host.root])) host.at[host.root] = 0.0;
            if (!host.params[host.root]) host.params[host.root] = std::make_pair(1000.0, 0.001);
            if (std::isnan(guest.at[guest.root])) guest.at[guest.root] = 0.0;
        } catch (std::exception& e) {
            throw std::invalid_argument(std::string("IllegalArgumentException: Invalid input to rate model: ") + e.what());
        }
        const int x = guest.root, y = host.root;
        if (std::fabs(guest.vt[x] + guest.at[x] - host.vt[y] - host.at[y]) > 1e-4) throw std::invalid_argument("IllegalArgumentException: Invalid guest and host trees: Time scales do not agree.");
    }
    int compute_sigma(int x) {
        if (guest.left[x] == -1) return sigma[x];
        const int l = compute_sigma(guest.left[x]), r = compute_sigma(guest.right[x]);
        return sigma[x] = host.lca(l, r);
    }
    int enclosing_arc(int x) const {
        int lo = sigma[x]; const double xt = guest.vt[x];
        for (;;) {
            if (lo == host.root) return lo;
            const int lop = host.parent[lo];
            if (host.vt[lop] >= xt) return lo;
            lo = lop;
        }
    }
    void rates(JavaMT& prng, std::vector<double>& out) const {       // getRates :110-157
        std::vector<double> ws, rs;
        for (int x = 0; x < guest.n; ++x) {
            const int lo = enclosing_arc(x);
            const int up = guest.parent[x] == -1 ? host.root : enclosing_arc(guest.parent[x]);
            ws.clear(); rs.clear();
            const double l = guest.at[x];
            int h = lo;
            for (;;) {
                if (!host.params[h]) throw std::runtime_error("NullPointerException");
                const double k = host.params[h]->first, theta = host.params[h]->second;
                if (k <= 0.0 || theta <= 0.0) throw std::invalid_argument("IllegalArgumentException: Cannot have non-positive shape or scale for gamma distribution.");
                ws.push_back(host.at[h] / l);
                rs.push_back(chi2_quantile(prng.nextDouble(), 2 * k) * theta / 2);
                if (h == up) break;
                h = host.parent[h];
                if (h < 0) throw std::out_of_range("ArrayIndexOutOfBoundsException: -1");
            }
            if (ws.size() == 1) ws[0] = 1.0;
            else {
                double dt = guest.vt[x] - host.vt[lo];
                ws[0] = std::max(0.0, ws[0] - dt / l);
                if (guest.parent[x] != -1) { dt = guest.vt[guest.parent[x]] - host.vt[up]; ws.back() = std::max(0.0, dt / l); }
            }
            double r = 0.0;
            for (size_t i = 0; i < ws.size(); ++i) r += ws[i] * rs[i];
            r *= scale;
            out[x] = r;
        }
    }
};
// io/GuestHostMapReader.java:44-56: one "guest host" pair per line, split on blanks and tabs, reading stops at the first empty line
inline std::vector<std::pair<std::string, std::string>> read_guest_host_map(const std::string& path) {
    std::ifstream f(path); if (!f) throw std::invalid_argument("FileNotFoundException: " + path + " (No such file or directory)");
    std::vector<std::pair<std::string, std::string>> out; std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::string ln = trim_java(line);
        if (ln.empty()) break;
        std::vector<std::string> parts; size_t i = 0;
        while (i < ln.size()) { size_t j = i; while (j < ln.size() && ln[j] != ' ' && ln[j] != '\t') ++j; parts.push_back(ln.substr(i, j - i)); i = j; while (i < ln.size() && (ln[i] == ' ' || ln[i] == '\t')) ++i; }
        if (parts.size() < 2) throw std::out_of_range("ArrayIndexOutOfBoundsException: 1");
        out.push_back({parts[0], parts[1]});
    }
    return out;
}

// PrIMENewickTreeVerifier.runStandardTestSuite (io/PrIMENewickTreeVerifier.java:37-46) as far as plain Newick input can
// trip it: leaf names present and unique (:79-94), branch lengths on every vertex but the root (:101-103,143-152)
inline void strict_newick_checks(const NTree& t) {
    auto bfs = t.vertices_bfs(); const int n = (int)bfs.size();
    bool any_name = false, any_bl = false; for (NVertex* v : bfs) { any_name |= (bool)v->name; any_bl |= (bool)v->bl; }
    if (any_name) {
        std::set<std::string> seen;
        for (NVertex* v : bfs) {
            if (!v->is_leaf()) continue;
            if (!v->name) throw NewickIOException("Missing name on Newick vertex " + std::to_string(v->number) + ".");
            if (!seen.insert(*v->name).second) throw NewickIOException("Duplicate vertex name '" + *v->name + "' found in vertex " + std::to_string(v->number) + ".");
        }
    }
    if (any_bl) for (int i = 0; i < n; ++i) for (NVertex* v : bfs) if (v->number == i && !v->bl && v != t.root) throw NewickIOException("Missing branch length on Newick vertex " + std::to_string(i) + ".");
}

inline const std::vector<OptSpec>& relax_specs() {
    static const std::vector<OptSpec> s = {{{"-h", "--help"}, 0}, {{"-o", "--output-file"}, 1}, {{"-s", "--seed"}, 1}, {{"-x", "--auxiliary-tags"}, 0},
        {{"-innms", "--keep-interior-names"}, 0}, {{"-min", "--min-rate"}, 1}, {{"-max", "--max-rate"}, 1}, {{"-a", "--max-attempts"}, 1}};
    return s;
}
inline bool ieq(const std::string& a, const char* b) { if (a.size() != std::strlen(b)) return false; for (size_t i = 0; i < a.size(); ++i) if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false; return true; }

inline void relax_to_string(const NTree& t, const NVertex* v, const std::vector<double>& len, const std::vector<double>& orig, const std::vector<double>& rate,
                            bool keepInterior, bool withMeta, std::string& sb) {
    if (!v->children.empty()) {
        sb.push_back('(');
        for (size_t i = 0; i < v->children.size(); ++i) { if (i) sb.push_back(','); relax_to_string(t, v->children[i], len, orig, rate, keepInterior, withMeta, sb); }
        sb.push_back(')');
    }
    int x = v->number;
    if (v->name && (v->is_leaf() || keepInterior)) sb += *v->name;
    if (!std::isnan(len[x])) { sb.push_back(':'); sb += java_double_to_string(len[x]); }
    if (withMeta) sb += "[ID=" + std::to_string(x) + " ORIGBL=" + java_double_to_string(orig[x]) + " RATE=" + java_double_to_string(rate[x]) + "]";
}
inline std::string arr_to_string(const std::vector<double>& a) { std::string s = "["; for (size_t i = 0; i < a.size(); ++i) { if (i) s += ", "; s += java_double_to_string(a[i]); } return s + "]"; }

inline void run_branch_relaxer(const std::vector<std::string>& args, Outputs& o, RepStats* st = nullptr) {
    try {
        ParsedArgs p = jc_parse(args, relax_specs());
        if (p.has("-h") || args.size() < 1) { o.out += "Usage:\n    java -jar sagephy-X.Y.Z.jar BranchRelaxer [options] <tree> <model> <arg1> <arg2> <...>\n\n"; return; }
        if (!p.has("-s")) throw std::invalid_argument("oracle: a seed (-s) is required (the reference would read /dev/random)");
        JavaMT prng(p.get("-s", ""));
        auto arg = [&](size_t i) -> const std::string& { if (i >= p.pos.size()) throw std::out_of_range("IndexOutOfBoundsException: Index: " + std::to_string(i) + ", Size: " + std::to_string(p.pos.size())); return p.pos[i]; };
        RateModelO m; const std::string& model = arg(1);
        if (ieq(model, "Constant")) { m.kind = RateModelO::CONSTANT; m.a = parse_double(arg(2)); }
        else if (ieq(model, "IIDExponential")) { m.kind = RateModelO::EXPONENTIAL; m.a = parse_double(arg(2)); if (m.a <= 0) throw std::invalid_argument("IllegalArgumentException: Cannot have non-positive rate for exponential distribution."); }
        else if (ieq(model, "IIDGamma")) { m.kind = RateModelO::GAMMA; m.a = parse_double(arg(2)); m.b = parse_double(arg(3)); if (m.a <= 0 || m.b <= 0) throw std::invalid_argument("IllegalArgumentException: Cannot have non-positive shape or scale for gamma distribution."); }
        else if (ieq(model, "IIDLogNormal")) { m.kind = RateModelO::LOGNORMAL; m.a = parse_double(arg(2)); double s2 = parse_double(arg(3)); if (s2 <= 0) throw std::invalid_argument("IllegalArgumentException: Cannot have non-positive sigma^2 for log-normal distribution."); m.sigma = std::sqrt(s2); }
        else if (ieq(model, "IIDNormal")) { m.kind = RateModelO::NORMAL; m.a = parse_double(arg(2)); m.b = parse_double(arg(3)); if (m.b <= 0) throw std::invalid_argument("IllegalArgumentException: Cannot have non-positive variance for normal distribution."); m.sigma = std::sqrt(m.b); }
        else if (ieq(model, "IIDUniform")) { m.kind = RateModelO::UNIFORM; m.a = parse_double(arg(2)); m.b = parse_double(arg(3)); if (m.a >= m.b) throw std::invalid_argument("IllegalArgumentException: Invalid range for uniform distribution."); }
        else if (ieq(model, "IIDSamplesFromFile")) {
            m.kind = RateModelO::SAMPLES; m.file = arg(2);
            std::ifstream f(m.file); if (!f) throw std::invalid_argument("IllegalArgumentException: java.io.FileNotFoundException: " + m.file);
            std::string line; while (std::getline(f, line)) m.samples.push_back(parse_double(line));
        }
        else if (ieq(model, "ACTK98")) { m.kind = RateModelO::ACTK98; m.a = parse_double(arg(2)); m.b = parse_double(arg(3)); }
        else if (ieq(model, "ACRY07")) { m.kind = RateModelO::ACRY07; m.a = parse_double(arg(2)); m.b = parse_double(arg(3)); }
        else if (ieq(model, "ACABY02")) { m.kind = RateModelO::ACABY02; m.a = parse_double(arg(2)); }
        else if (ieq(model, "ACLBPL07")) {
            m.kind = RateModelO::ACLBPL07; m.a = parse_double(arg(2)); m.b = parse_double(arg(3)); m.c = parse_double(arg(4)); m.d = parse_double(arg(5));
            if (2 * m.c * m.b < m.d * m.d) throw std::invalid_argument("IllegalArgumentException: Invalid parameters for CIR process: 0 rate may be achieved.");
        } else if (ieq(model, "IIDRK11")) m.kind = RateModelO::IIDRK11;        // BranchRelaxerParameters.java:108-113, built below once the tree is read
        else throw std::invalid_argument("IllegalArgumentException: Invalid rate model identifier: " + model);
        std::unique_ptr<NTree> t = file_exists(arg(0)) ? nw_read_tree(read_file(arg(0))) : nw_read_tree(arg(0));
        RK11Model rk; std::unique_ptr<NTree> rk_host;
        if (m.kind == RateModelO::IIDRK11) {
            strict_newick_checks(*t);
            rk_host = file_exists(arg(2)) ? nw_read_tree(read_file(arg(2))) : nw_read_tree(arg(2));
            strict_newick_checks(*rk_host);
            auto gs = read_guest_host_map(arg(3));
            rk.init(*t, *rk_host, gs, parse_double(arg(4)));
        }
        auto bfs = t->vertices_bfs(); int n = (int)bfs.size();
        if (m.strict_tree()) strict_newick_checks(*t);
        for (NVertex* v : bfs) if (v->children.size() == 1) throw TopologyException("Cannot create RTree from collapsable NewickTree.");
        for (NVertex* v : bfs) if (v->is_leaf() && !v->name) throw std::runtime_error("NullPointerException: Missing leaf name in vertex " + std::to_string(v->number) + ".");
        std::vector<double> orig(n, std::nan("")); bool has = false;
        for (NVertex* v : bfs) if (v->bl) { orig[v->number] = *v->bl; has = true; }
        if (!has) throw NewickIOException("Cannot create time map; times missing in tree.");
        int root = t->root->number, ignore = -1;
        if (std::isnan(orig[root])) { orig[root] = 0.0; ignore = root; }
        double mn = parse_double(p.get("-min", "1e-64")), mx = parse_double(p.get("-max", "1e64"));
        int maxAttempts = parse_int(p.get("-a", "10000")), attempts = 0;
        std::vector<double> rates(n);
        std::vector<int> topo(n), parent(n, -1);
        for (int i = 0; i < n; ++i) { topo[i] = bfs[i]->number; if (bfs[i]->parent) parent[bfs[i]->number] = bfs[i]->parent->number; }
        for (;;) {
            if (attempts > maxAttempts) { o.err += "Failed to create valid rates within max allowed attempts.\n"; if (st) { st->status = 1; st->attempts = attempts; } return; }
            if (m.kind == RateModelO::SAMPLES) { /* the reference re-reads the file on every call; same values */ }
            if (m.kind == RateModelO::IIDRK11) rk.rates(prng, rates);
            else if (m.kind >= RateModelO::ACTK98) m.rates_autocorrelated(topo, parent, orig, prng, rates);
            else for (int x = 0; x < n; ++x) rates[x] = m.sample(prng);
            ++attempts;
            bool ok = true;
            for (int x = 0; x < n; ++x) { if (x == ignore) continue; if (rates[x] < mn || rates[x] > mx) { ok = false; break; } }
            if (ok) break;
        }
        std::vector<double> rel(n); for (int x = 0; x < n; ++x) rel[x] = orig[x] * rates[x];
        std::string tree; relax_to_string(*t, t->root, rel, orig, rates, p.has("-innms"), p.has("-x"), tree); tree.push_back(';');
        if (!p.has("-o")) o.out += tree + "\n";
        else {
            std::string of = trim_java(p.get("-o", ""));
            o.file(of, tree + "\n");
            std::string info = "# BRANCHLENGTHRELAXER\nArguments:\t" + arrays_to_string(args) + "\nAttempts:\t" + std::to_string(attempts) + "\n";
            info += "Original lengths:\t" + arr_to_string(orig) + "\nRates:\t" + arr_to_string(rates) + "\nRelaxed lengths:\t" + arr_to_string(rel) + "\n";
            info += "Model:\t" + m.name() + "\nModel parameters:\n";
            for (auto& kv : m.params()) info += "\t" + kv.first + "\t" + kv.second + "\n";
            o.file(of + ".info", info);
        }
        if (st) { st->attempts = attempts; st->n_unpruned = n; double tot = 0; for (double v : rel) tot += v; st->total_time = tot; }
    } catch (std::exception& e) {
        o.err += std::string(e.what()) + "\n\nUse option -h or --help to show usage.\n";
        if (st) st->status = 2;
    }
}

}  // namespace oracle
