// ORACLE — test infrastructure only (see oracle/README.md). PARITY UNPINNED.
//
// Host-tree model: RBTree arrays, vertex/arc times, epoch discretisation and transfer-recipient sampling. Restates
//   topology/RBTree.java:56-85,231-258                      (arrays from Newick numbering, getLeaves, naive LCA)
//   io/PrIMENewickTree.java:733-779, io/PrIMENewickTreeVerifier.java:237-270   (times from branch lengths)
//   topology/RBTreeEpochDiscretiser.java:95-106,146-229,264-344,620-632,662-683 (update, getDistance, getters, toString, sampleRoot)
//   topology/Epoch.java:63-82,232-329                       (times, findIndexOfArc, sampleArc, setDistance)
//   java.math.BigDecimal valueOf/add/multiply/compareTo and TreeMap.higherEntry as used there (JDK specification).
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>
#include "jmath.hpp"
#include "jrandom.hpp"
#include "newick.hpp"
#include "jpow.hpp"

namespace oracle {

// test hook: how often Epoch.sampleArc found no eligible recipient (returned -5) since the counter was last reset
inline long& no_recipient_hits() { static thread_local long n = 0; return n; }


struct TopologyException : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- non-negative arbitrary-precision decimal (the subset of BigDecimal the path uses) --------------------------------
struct Dec {
    std::vector<uint32_t> mag;   // base 1e9, little endian, no leading zero limbs (empty == 0)
    int scale = 0;               // value = mag * 10^-scale
    static Dec zero() { return Dec(); }
    void trim() { while (!mag.empty() && mag.back() == 0) mag.pop_back(); }
    void mul_small(uint32_t m) {
        uint64_t carry = 0;
        for (auto& limb : mag) { uint64_t v = (uint64_t)limb * m + carry; limb = (uint32_t)(v % 1000000000ull); carry = v / 1000000000ull; }
        while (carry) { mag.push_back((uint32_t)(carry % 1000000000ull)); carry /= 1000000000ull; }
    }
    void add_small(uint32_t a) {
        uint64_t carry = a;
        for (size_t i = 0; i < mag.size() && carry; ++i) { uint64_t v = (uint64_t)mag[i] + carry; mag[i] = (uint32_t)(v % 1000000000ull); carry = v / 1000000000ull; }
        if (carry) mag.push_back((uint32_t)carry);
    }
    // BigDecimal.valueOf(double) == new BigDecimal(Double.toString(d)); d finite, >= 0.
    static Dec from_double(double d) {
        std::string s = java_double_to_string(d);
        Dec r; int exp10 = 0; size_t epos = s.find('E');
        std::string m = s;
        if (epos != std::string::npos) { exp10 = std::atoi(s.c_str() + epos + 1); m = s.substr(0, epos); }
        int frac = 0; bool seen = false;
        for (char c : m) {
            if (c == '.') { seen = true; continue; }
            r.mul_small10(); r.add_small((uint32_t)(c - '0'));
            if (seen) ++frac;
        }
        r.scale = frac - exp10;
        r.trim();
        if (r.scale < 0) { r.pow10_up(-r.scale); r.scale = 0; }
        return r;
    }
    void mul_small10() { if (mag.empty()) return; mul_small(10); }
    void pow10_up(int k) { while (k >= 9) { mul_small_1e9(); k -= 9; } static const uint32_t p[9] = {1,10,100,1000,10000,100000,1000000,10000000,100000000}; if (k > 0 && !mag.empty()) mul_small(p[k]); }
    void mul_small_1e9() { if (!mag.empty()) mag.insert(mag.begin(), 0u); }
    static void align(Dec& a, Dec& b) {
        if (a.scale < b.scale) { a.pow10_up(b.scale - a.scale); a.scale = b.scale; }
        else if (b.scale < a.scale) { b.pow10_up(a.scale - b.scale); b.scale = a.scale; }
    }
    static Dec add(Dec a, Dec b) {
        align(a, b);
        Dec r; r.scale = a.scale;
        size_t n = std::max(a.mag.size(), b.mag.size());
        uint64_t carry = 0;
        for (size_t i = 0; i < n || carry; ++i) {
            uint64_t v = carry + (i < a.mag.size() ? a.mag[i] : 0) + (i < b.mag.size() ? b.mag[i] : 0);
            r.mag.push_back((uint32_t)(v % 1000000000ull)); carry = v / 1000000000ull;
        }
        r.trim(); return r;
    }
    static Dec mul(const Dec& a, const Dec& b) {
        Dec r; r.scale = a.scale + b.scale;
        if (a.mag.empty() || b.mag.empty()) return r;
        std::vector<uint64_t> acc(a.mag.size() + b.mag.size() + 1, 0);
        for (size_t i = 0; i < a.mag.size(); ++i) {
            uint64_t carry = 0;
            for (size_t j = 0; j < b.mag.size() || carry; ++j) {
                uint64_t cur = acc[i + j] + carry + (j < b.mag.size() ? (uint64_t)a.mag[i] * b.mag[j] : 0);
                acc[i + j] = cur % 1000000000ull; carry = cur / 1000000000ull;
            }
        }
        r.mag.resize(acc.size());
        for (size_t i = 0; i < acc.size(); ++i) r.mag[i] = (uint32_t)acc[i];
        r.trim(); return r;
    }
    static int cmp(Dec a, Dec b) {      // compareTo: numeric
        align(a, b);
        if (a.mag.size() != b.mag.size()) return a.mag.size() < b.mag.size() ? -1 : 1;
        for (size_t i = a.mag.size(); i-- > 0;) if (a.mag[i] != b.mag[i]) return a.mag[i] < b.mag[i] ? -1 : 1;
        return 0;
    }
};

// TreeMap<BigDecimal,Integer> restricted to the usage pattern "put(non-decreasing running total, value)".
struct CumMap {
    std::vector<std::pair<Dec, int>> e;
    void put(const Dec& k, int v) { if (!e.empty() && Dec::cmp(e.back().first, k) == 0) e.back().second = v; else e.push_back({k, v}); }
    bool empty() const { return e.empty(); }
    const Dec& last_key() const { return e.back().first; }
    int higher(const Dec& r) const { for (auto& kv : e) if (Dec::cmp(kv.first, r) > 0) return kv.second; throw std::runtime_error("NullPointerException: higherEntry == null"); }
};

struct EpochO {
    int no = 0;
    std::vector<int> arcs;
    double times[3] = {0, 0, 0};      // nmin = nmax = nroot = 1  =>  [lo, mid, up]   (Epoch.java:70-81)
    int transferedToArc = -1;
    double lower() const { return times[0]; }
    double upper() const { return times[2]; }
    int n_arcs() const { return (int)arcs.size(); }
    int find_index_of_arc(int a) const { int i = 0; while (arcs[i] != a) ++i; return i; }
};

class HostModel {
public:
    int k = 0, root = -1;
    std::vector<int> parent, left, right, host_tag;
    std::vector<std::optional<std::string>> names;     // leaf names only
    std::vector<double> vt, at;                        // TimesMap vertex / arc times
    std::vector<EpochO> epochs;
    std::vector<int> v2e;                              // vertexToEpoch
    std::optional<std::string> tree_name;              // RBTree name = host.getTreeName()

    // GuestTreeInHostTreeCreator ctor (apps/SaGePhy/GuestTreeInHostTreeCreator.java:73-86) + readers.
    HostModel(const NTree& t, std::optional<std::string> treeName, const std::optional<double>& stem) : tree_name(std::move(treeName)) {
        auto bfs = t.vertices_bfs();
        k = (int)bfs.size();
        parent.assign(k, -1); left.assign(k, -1); right.assign(k, -1); host_tag.assign(k, -10);
        names.assign(k, std::nullopt);
        for (NVertex* v : bfs) if (!(v->children.empty() || v->children.size() == 2))
            throw TopologyException("Cannot create RBTree from non-bifurcating NewickTree.");
        for (NVertex* v : bfs) {
            int i = v->number;
            if (!v->parent) { parent[i] = -1; root = i; } else parent[i] = v->parent->number;
            if (!v->is_leaf()) { left[i] = v->children[0]->number; right[i] = v->children[1]->number; }
            host_tag[i] = v->get_host_vertex();
            if (v->is_leaf()) {
                if (!v->name) throw std::runtime_error("NullPointerException: Missing leaf name in vertex " + std::to_string(i) + ".");
                names[i] = v->name;
            }
        }
        // times (BL selector; VT/AT tags can never be present because they need the "[&&PRIME " prefix)
        bool has = false; at.assign(k, std::nan(""));
        for (NVertex* v : bfs) if (v->bl) { at[v->number] = *v->bl; has = true; }
        if (!has) throw NewickIOException("Cannot create time map; times missing in tree.");
        vt.assign(k, std::nan(""));
        for (int i = k - 1; i > 0; --i) {
            NVertex* v = bfs[i]; int x = v->number;
            if (v->is_leaf()) vt[x] = 0.0;
            if (!std::isnan(vt[x])) { double tt = vt[x] + at[x]; int anc = v->parent->number; if (std::isnan(vt[anc])) vt[anc] = tt; }
        }
        if (k == 1) vt[0] = 0.0;
        if (stem) at[root] = *stem;
        if (at[root] <= 0.0) at[root] = 1.0e-64;
        update();
    }

    bool is_root(int x) const { return parent[x] == -1; }
    bool is_leaf(int x) const { return left[x] == -1; }

    void leaves_l2r(std::vector<int>& q, int x) const { if (is_leaf(x)) q.push_back(x); else { leaves_l2r(q, left[x]); leaves_l2r(q, right[x]); } }

    void update() {      // RBTreeEpochDiscretiser.update :146-221
        epochs.assign((k + 1) / 2, EpochO());
        v2e.assign(k, 0);
        std::vector<int> q; leaves_l2r(q, root);
        for (int x : q) v2e[x] = 0;
        int xLo = q.front(); double tLo = vt[xLo], tUp; int epochNo = 0;
        const double MIN_SPLICE_DELTA = 0.0001;
        while (q.size() > 1) {
            int xUpIdx = -1; tUp = 1.7976931348623157e308;
            for (int j = 0; j < (int)q.size(); ++j) { int xp = parent[q[j]]; if (vt[xp] < tUp) { xUpIdx = j; tUp = vt[xp]; } }
            if (tUp + MIN_SPLICE_DELTA < tLo) tUp = tLo + MIN_SPLICE_DELTA;
            set_epoch(epochNo, q, tLo, tUp);
            v2e[xLo] = epochNo;
            int par = parent[q[xUpIdx]];
            q.erase(q.begin() + xUpIdx); q.insert(q.begin() + xUpIdx, par); q.erase(q.begin() + xUpIdx + 1);
            xLo = par; tLo = tUp; ++epochNo;
        }
        tUp = vt[root] + at[root];
        if (tUp + MIN_SPLICE_DELTA < tLo) tUp = tLo + MIN_SPLICE_DELTA;
        set_epoch(epochNo, q, tLo, tUp);
        v2e[xLo] = epochNo;
    }
    void set_epoch(int no, const std::vector<int>& arcs, double lo, double up) {
        EpochO& e = epochs[no]; e.no = no; e.arcs = arcs;
        double step = (up - lo) / 1;
        e.times[0] = lo; e.times[1] = lo + step / 2.0 + 0 * step; e.times[2] = up;
    }

    double vertex_time(int x) const { return epochs[v2e[x]].lower(); }          // discretised
    double vertex_upper_time(int x) const { return epochs[v2e[x]].upper(); }
    double tip_to_leaf_time() const { return epochs.back().upper(); }
    int epoch_no_above(int x) const { return v2e[x]; }
    std::vector<int> leaves() const { std::vector<int> l; for (int i = 0; i < k; ++i) if (left[i] == -1) l.push_back(i); return l; }

    int lca(int x, int y) const {      // RBTree.getLCA :243-258
        std::vector<int> visited;
        for (;;) {
            if (x != -1) { if (std::find(visited.begin(), visited.end(), x) != visited.end()) return x; visited.push_back(x); x = parent[x]; }
            if (y != -1) { if (std::find(visited.begin(), visited.end(), y) != visited.end()) return y; visited.push_back(y); y = parent[y]; }
        }
    }
    double distance(int x, int y, double eventTime) const { return 2.0 * (vertex_time(lca(x, y)) - eventTime); }

    // Epoch.setDistance :301-329
    CumMap set_distance(const EpochO& ep, int excludeArc, double eventTime, const std::string& bias, double rate,
                        const std::vector<int>* emptyArcs, int hostArc, bool include) const {
        Dec total; CumMap m;
        for (int x : ep.arcs) {
            int arcHost = host_tag[x];
            if (x != excludeArc && (include ? (arcHost == hostArc) : (arcHost != hostArc))) {
                if (!emptyArcs || std::find(emptyArcs->begin(), emptyArcs->end(), x) == emptyArcs->end()) {
                    double v;
                    if (bias == "exponential") v = rate * j_pow_e(-rate * distance(excludeArc, x, eventTime));
                    else v = 1.0 / distance(excludeArc, x, eventTime);
                    if (std::isinf(v)) v = 1.7976931348623157e308;
                    total = Dec::add(total, Dec::from_double(v));
                    m.put(total, x);
                }
            }
        }
        return m;
    }
    // Math.pow(Math.E, y) through the fdlibm pow restatement (jpow.hpp)
    static double j_pow_e(double y) { return j_pow(2.718281828459045, y); }

    // Epoch.sampleArc :249-296. Returns -5 when there is no candidate.
    int sample_arc(EpochO& ep, JavaMT& prng, int excludeArc, int fromArcIdx, double eventTime, const std::string& bias,
                   double rate, const std::vector<int>* emptyArcs, int hostArc, bool include) const {
        int arc;
        if (ep.arcs.size() == 1 && excludeArc == ep.arcs[0])
            throw std::invalid_argument("IllegalArgumentException: Cannot exclude arc from sampling (it's the only one!): " + std::to_string(excludeArc));
        if (bias != "none") {
            CumMap dist = set_distance(ep, excludeArc, eventTime, bias, rate, emptyArcs, hostArc, include);
            if (dist.empty()) { ++no_recipient_hits(); return -5; }
            Dec total = dist.last_key();
            Dec rnd = Dec::mul(total, Dec::from_double(prng.nextDouble()));
            arc = dist.higher(rnd);
        } else {
            std::vector<int> same;
            for (int x : ep.arcs) {
                int arcHost = host_tag[x];
                if (x != excludeArc && (!emptyArcs || std::find(emptyArcs->begin(), emptyArcs->end(), x) == emptyArcs->end()) &&
                    (include ? (arcHost == hostArc) : (arcHost != hostArc))) same.push_back(x);
            }
            if (same.empty()) { ++no_recipient_hits(); return -5; }
            arc = same[prng.nextInt((int)same.size())];
            // the reference's re-draw loop never iterates: Arrays.asList(int[]).indexOf(Integer) == -1 != fromArc (Epoch.java:279-291)
            (void)fromArcIdx;
        }
        ep.transferedToArc = arc;
        return arc;
    }

    int sample_root(JavaMT& prng, double gene_birth) const {     // RBTreeEpochDiscretiser.sampleRoot :662-683
        Dec total; CumMap m;
        for (int i = 0; i < k; ++i) {
            double v = gene_birth * j_pow_e(-gene_birth * (tip_to_leaf_time() - vertex_time(i)));
            total = Dec::add(total, Dec::from_double(v));
            m.put(total, i);
        }
        Dec rnd = Dec::mul(m.last_key(), Dec::from_double(prng.nextDouble()));
        return m.higher(rnd);
    }

    // RBTreeEpochDiscretiser.toString :620-632 (host-tree header of the .guest2host files)
    void to_string_rec(int x, std::string& sb) const {
        if (!is_leaf(x)) { sb.push_back('('); to_string_rec(left[x], sb); sb.push_back(','); to_string_rec(right[x], sb); sb.push_back(')'); }
        if (names[x]) sb += *names[x];
        if (!std::isnan(at[x])) { sb.push_back(':'); sb += java_double_to_string(at[x]); }
        int below = is_root(x) ? v2e[x] : (v2e[parent[x]] - 1);
        sb += "[ID=" + std::to_string(x) + " NT=" + java_double_to_string(vt[x]) + " EPOCHS=(" + std::to_string(v2e[x]) + "..." + std::to_string(below) + ")]";
    }
    std::string to_string() const {
        std::string sb; to_string_rec(root, sb);
        sb += "[NAME=" + (tree_name ? *tree_name : std::string("null")) + " DISCTYPE=RBTreeEpochDiscretiser NMIN=1 NMAX=1 DELTAT=NaN NROOT=1];";
        return sb;
    }
};

}  // namespace oracle
