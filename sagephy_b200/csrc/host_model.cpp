#include "host_model.hpp"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "hostfmt.hpp"

namespace sgp {

namespace {

struct Cursor {
    const std::string& s; size_t p = 0;
    explicit Cursor(const std::string& str) : s(str) {}
    bool done() const { return p >= s.size(); }
    char look() const { return p < s.size() ? s[p] : '\0'; }
    char take() { if (p >= s.size()) throw SgpError("Error parsing Newick tree: unexpected end of input."); return s[p++]; }
};

// Whitespace outside [...] is dropped; reading stops after the second ';' (the reference's strip(str, 1) keeps going
// while count <= 1). '\r' is kept, as in the reference.
std::string squeeze(const std::string& in) {
    std::string out; out.reserve(in.size() + 1);
    int semis = 0; bool in_meta = false;
    for (size_t i = 0; i < in.size() && semis <= 1; ++i) {
        char c = in[i];
        if (c == '[') in_meta = true; else if (c == ']') in_meta = false; else if (c == ';') ++semis;
        if (in_meta || !(c == ' ' || c == '\t' || c == '\n')) out.push_back(c);
    }
    out.push_back(';');
    return out;
}

struct Builder {
    FlatTree t;
    int fresh() {
        int id = (int)t.parent.size();
        t.parent.push_back(-1); t.first_child.push_back(-1); t.next_sibling.push_back(-1); t.n_children.push_back(0);
        t.name.emplace_back(); t.has_name.push_back(0); t.bl.push_back(std::nan("")); t.has_bl.push_back(0);
        t.meta.emplace_back(); t.has_meta.push_back(0); t.host_tag.push_back(-10);
        return id;
    }
};

bool read_meta(Cursor& c, std::string& out) {
    if (c.look() != '[') return false;
    out.clear(); char ch;
    do { ch = c.take(); out.push_back(ch); } while (ch != ']');
    return true;
}

void read_label(Cursor& c, FlatTree& t, int v) {
    std::string nm; char ch = c.look();
    while (!c.done() && ch != ',' && ch != ';' && ch != ')' && ch != '[' && ch != ':') { nm.push_back(c.take()); ch = c.look(); }
    if (!nm.empty()) {
        if (nm[0] == '\'') {
            if (nm.back() != '\'') throw SgpError("Unmatched quote in leaf name: " + nm);
            nm = nm.size() >= 2 ? nm.substr(1, nm.size() - 2) : std::string();
        }
        t.name[v] = nm; t.has_name[v] = 1;
    }
    if (c.look() == ':') {
        c.take();
        if (c.done()) throw SgpError("Expected branch length when reading Newick tree.");
        std::string num; num.push_back(c.take()); ch = c.look();
        while (!c.done() && ch != ',' && ch != ';' && ch != ')' && ch != '[') { num.push_back(c.take()); ch = c.look(); }
        char* end = nullptr; double val = std::strtod(num.c_str(), &end);
        if (end == num.c_str() || *end) throw SgpError("NumberFormatException: For input string: \"" + num + "\"");
        t.bl[v] = val; t.has_bl[v] = 1;
    }
    std::string m;
    if (read_meta(c, m)) {
        t.meta[v] = m; t.has_meta[v] = 1;
        // first occurrence of HOST="?<digits>"?
        for (size_t pos = m.find("HOST="); pos != std::string::npos; pos = m.find("HOST=", pos + 5)) {
            size_t k = pos + 5; if (k < m.size() && m[k] == '"') ++k;
            size_t st = k; while (k < m.size() && m[k] >= '0' && m[k] <= '9') ++k;
            if (k > st) { t.host_tag[v] = std::atoi(m.substr(st, k - st).c_str()); break; }
        }
    } else {
        t.host_tag[v] = 0;
    }
}

// Returns the temporary id of the subtree root; 'order' receives ids in completion (= post-order) order.
int read_subtree(Cursor& c, Builder& b, std::vector<int>& order) {
    int v = b.fresh();
    if (c.look() == '(') {
        c.take();
        int last = -1, count = 0;
        for (;;) {
            int ch = read_subtree(c, b, order);
            b.t.parent[ch] = v;
            if (last < 0) b.t.first_child[v] = ch; else b.t.next_sibling[last] = ch;
            last = ch; ++count;
            if (count == 1) { if (c.done() || c.take() != ',') throw SgpError("Expected comma when reading Newick tree."); continue; }
            if (c.look() == ',') { c.take(); continue; }
            break;
        }
        b.t.n_children[v] = count;
        if (c.done() || c.take() != ')') throw SgpError("Expected right parenthesis when reading Newick tree.");
    }
    read_label(c, b.t, v);
    order.push_back(v);
    return v;
}

}  // namespace

FlatTree parse_newick(const std::string& text) {
    std::string sq = squeeze(text);
    Cursor c(sq);
    Builder b; std::vector<int> order;
    int tmp_root = read_subtree(c, b, order);
    std::string tm; bool has_tm = read_meta(c, tm);
    if (c.done() || c.take() != ';') throw SgpError("Expected semi-colon when reading Newick tree.");
    // renumber: temporary id -> post-order number
    FlatTree& s = b.t; int n = (int)order.size();
    std::vector<int> num(n);
    for (int i = 0; i < n; ++i) num[order[i]] = i;
    FlatTree t; t.n = n; t.root = num[tmp_root];
    t.parent.assign(n, -1); t.first_child.assign(n, -1); t.next_sibling.assign(n, -1); t.n_children.assign(n, 0);
    t.name.resize(n); t.has_name.assign(n, 0); t.bl.assign(n, std::nan("")); t.has_bl.assign(n, 0); t.meta.resize(n); t.has_meta.assign(n, 0);
    t.host_tag.assign(n, -10);
    for (int o = 0; o < n; ++o) {
        int v = num[o];
        t.parent[v] = s.parent[o] < 0 ? -1 : num[s.parent[o]];
        t.first_child[v] = s.first_child[o] < 0 ? -1 : num[s.first_child[o]];
        t.next_sibling[v] = s.next_sibling[o] < 0 ? -1 : num[s.next_sibling[o]];
        t.n_children[v] = s.n_children[o];
        t.name[v] = s.name[o]; t.has_name[v] = s.has_name[o]; t.bl[v] = s.bl[o]; t.has_bl[v] = s.has_bl[o];
        t.meta[v] = s.meta[o]; t.has_meta[v] = s.has_meta[o]; t.host_tag[v] = s.host_tag[o];
    }
    t.tree_meta = tm; t.has_tree_meta = has_tm;
    t.bfs.push_back(t.root);
    for (size_t i = 0; i < t.bfs.size(); ++i) for (int x = t.first_child[t.bfs[i]]; x >= 0; x = t.next_sibling[x]) t.bfs.push_back(x);
    return t;
}

void HostModel::build(const FlatTree& t, const std::string* name, const double* stem, bool need_lca) {
    nV = t.n; root = t.root;
    has_tree_name = name != nullptr; tree_name = name ? *name : std::string();
    parent = t.parent; left.assign(nV, -1); right.assign(nV, -1); host_tag = t.host_tag; leaf_name.assign(nV, std::string());
    for (int v = 0; v < nV; ++v) {
        if (t.n_children[v] != 0 && t.n_children[v] != 2) throw SgpError("Cannot create RBTree from non-bifurcating NewickTree.");
        if (t.n_children[v] == 2) { left[v] = t.first_child[v]; right[v] = t.next_sibling[left[v]]; }
        else {
            if (!t.has_name[v]) throw SgpError("Missing leaf name in vertex " + std::to_string(v) + ".");
            leaf_name[v] = t.name[v];
        }
    }
    leaves.clear(); for (int v = 0; v < nV; ++v) if (left[v] < 0) leaves.push_back(v);
    nLeaves = (int)leaves.size();
    // arc times = branch lengths; vertex times by the "first writer wins, scanning the BFS list backwards" rule
    bool any = false; at_raw.assign(nV, std::nan(""));
    for (int v = 0; v < nV; ++v) if (t.has_bl[v]) { at_raw[v] = t.bl[v]; any = true; }
    if (!any) throw SgpError("Cannot create time map; times missing in tree.");
    vt_raw.assign(nV, std::nan(""));
    for (int i = nV - 1; i > 0; --i) {
        int v = t.bfs[i];
        if (t.n_children[v] == 0) vt_raw[v] = 0.0;
        if (!std::isnan(vt_raw[v])) { int a = parent[v]; if (std::isnan(vt_raw[a])) vt_raw[a] = vt_raw[v] + at_raw[v]; }
    }
    if (nV == 1) vt_raw[0] = 0.0;
    if (stem) at_raw[root] = *stem;
    if (at_raw[root] <= 0.0) at_raw[root] = 1.0e-64;

    // epochs: contemporaries kept in planar order; merge the first pair whose parent is lowest
    nE = (nV + 1) / 2;
    ep_lo.assign(nE, 0); ep_mid.assign(nE, 0); ep_up.assign(nE, 0); ep_off.assign(nE + 1, 0); ep_arcs.clear(); v2e.assign(nV, 0);
    std::vector<int> row;
    {   // leaves left to right (iterative DFS)
        std::vector<int> st{root};
        while (!st.empty()) { int v = st.back(); st.pop_back(); if (left[v] < 0) row.push_back(v); else { st.push_back(right[v]); st.push_back(left[v]); } }
    }
    auto close_epoch = [&](int e, double lo, double up) {
        ep_lo[e] = lo; ep_up[e] = up;
        double step = (up - lo) / 1;
        ep_mid[e] = lo + step / 2.0 + 0 * step;
        ep_off[e] = (int)ep_arcs.size();
        ep_arcs.insert(ep_arcs.end(), row.begin(), row.end());
        max_arcs = std::max(max_arcs, (int)row.size());
    };
    const double kMinSplice = 0.0001;
    int low_vertex = row.front(); double t_lo = vt_raw[low_vertex], t_up; int e = 0;
    while (row.size() > 1) {
        int pick = -1; t_up = 1.7976931348623157e308;
        for (int j = 0; j < (int)row.size(); ++j) { double pt = vt_raw[parent[row[j]]]; if (pt < t_up) { pick = j; t_up = pt; } }
        if (pick < 0) throw SgpError("Host tree times are not usable (NaN vertex time).");
        if (t_up + kMinSplice < t_lo) t_up = t_lo + kMinSplice;
        close_epoch(e, t_lo, t_up);
        v2e[low_vertex] = e;
        int par = parent[row[pick]];
        if (pick + 1 >= (int)row.size()) throw SgpError("Host tree is not consistent with a planar epoch ordering.");
        row[pick] = par; row.erase(row.begin() + pick + 1);
        low_vertex = par; t_lo = t_up; ++e;
    }
    t_up = vt_raw[root] + at_raw[root];
    // A gene tree read without a stem length has a NaN top time (DomainTreeGen does not override it as GuestTreeGen does for its
    // host): with it no lineage ever reaches a vertex and every attempt grows until memory runs out. Refuse instead.
    if (std::isnan(t_up)) throw SgpError("Newick tree has no usable top time (stem length missing on the root)");
    if (t_up + kMinSplice < t_lo) t_up = t_lo + kMinSplice;
    close_epoch(e, t_lo, t_up);
    v2e[low_vertex] = e;
    ep_off[nE] = (int)ep_arcs.size();
    vtime.assign(nV, 0); for (int v = 0; v < nV; ++v) vtime[v] = ep_lo[v2e[v]];

    if (need_lca) {
        // time of the last common ancestor of every vertex pair, O(V^2): in pre-order a subtree is a contiguous range, and the
        // LCA of a vertex of the left subtree of v with a vertex of its right subtree is v (as is the LCA of v with any descendant)
        std::vector<int> ord; ord.reserve((size_t)nV); std::vector<int> tin(nV, 0), tout(nV, 0);
        {
            std::vector<int> st{root};
            while (!st.empty()) {
                const int v = st.back(); st.pop_back();
                tin[v] = (int)ord.size(); ord.push_back(v);
                if (left[v] >= 0) { st.push_back(right[v]); st.push_back(left[v]); }
            }
            for (int i = nV - 1; i >= 0; --i) { const int v = ord[(size_t)i]; tout[v] = left[v] >= 0 ? tout[right[v]] : tin[v] + 1; }
        }
        lca_time.assign((size_t)nV * nV, 0.0);
        for (int v = 0; v < nV; ++v) {
            const double tv = vtime[v];
            for (int i = tin[v]; i < tout[v]; ++i) { const int x = ord[(size_t)i]; lca_time[(size_t)v * nV + x] = lca_time[(size_t)x * nV + v] = tv; }
            if (left[v] < 0) continue;
            const int l = left[v], r = right[v];
            for (int i = tin[l]; i < tout[l]; ++i) {
                const int x = ord[(size_t)i]; double* row = &lca_time[(size_t)x * nV];
                for (int j = tin[r]; j < tout[r]; ++j) { const int y = ord[(size_t)j]; row[y] = tv; lca_time[(size_t)y * nV + x] = tv; }
            }
        }
    }

    // header = Newick of the host with [ID= NT= EPOCHS=] tags and the discretiser's tree tag
    header.clear();
    {
        struct Fr { int v; int stage; };
        std::vector<Fr> st{{root, 0}};
        while (!st.empty()) {
            Fr& f = st.back(); int v = f.v;
            if (left[v] >= 0 && f.stage == 0) { header.push_back('('); f.stage = 1; st.push_back({left[v], 0}); continue; }
            if (left[v] >= 0 && f.stage == 1) { header.push_back(','); f.stage = 2; st.push_back({right[v], 0}); continue; }
            if (left[v] >= 0) header.push_back(')');
            header += leaf_name[v];
            if (!std::isnan(at_raw[v])) { header.push_back(':'); header += jdouble(at_raw[v]); }
            int below = parent[v] < 0 ? v2e[v] : v2e[parent[v]] - 1;
            header += "[ID=" + std::to_string(v) + " NT=" + jdouble(vt_raw[v]) + " EPOCHS=(" + std::to_string(v2e[v]) + "..." + std::to_string(below) + ")]";
            st.pop_back();
        }
        header += "[NAME=" + (has_tree_name ? tree_name : std::string("null")) + " DISCTYPE=RBTreeEpochDiscretiser NMIN=1 NMAX=1 DELTAT=NaN NROOT=1];";
    }
}

}  // namespace sgp
