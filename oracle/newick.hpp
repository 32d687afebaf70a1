// ORACLE — test infrastructure only (see oracle/README.md). PARITY UNPINNED.
//
// Newick vertex / tree / reader / writer, restating:
//   io/NewickVertex.java:55-75,249-293,367-375,412-435   (fields, predicates, renumber, toString)
//   io/NewickTree.java:41-52,212-224,300-337,457-465      (ctor, BFS vertex list, names, branch lengths, toString)
//   io/NewickStringAlgorithms.java:27-48                  (strip)
//   io/NewickTreeReader.java:99-109,158-296               (readTree, readSubtree, readInfo, readName, readMeta, readBranchLength)
#pragma once
#include <cmath>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>
#include "jmath.hpp"

namespace oracle {

struct NewickIOException : std::runtime_error { using std::runtime_error::runtime_error; };

struct NVertex {
    int number = -1;
    std::optional<std::string> name;
    std::optional<double> bl;
    std::optional<std::string> meta;
    std::optional<int> host_vertex;       // NewickVertex.host_vertex (Integer, may be null)
    NVertex* parent = nullptr;
    std::vector<NVertex*> children;
    bool is_leaf() const { return children.empty(); }
    int get_host_vertex() const { return host_vertex ? *host_vertex : -10; }   // NewickVertex.java:102-108
};

// Owns its vertices.
struct NTree {
    std::vector<std::unique_ptr<NVertex>> pool;
    NVertex* root = nullptr;
    std::optional<std::string> meta;
    NVertex* make() { pool.emplace_back(new NVertex()); return pool.back().get(); }

    // NewickTree.getVerticesAsList: BFS, root first (io/NewickTree.java:212-224).
    std::vector<NVertex*> vertices_bfs() const {
        std::vector<NVertex*> q; q.push_back(root);
        for (size_t i = 0; i < q.size(); ++i) for (NVertex* c : q[i]->children) q.push_back(c);
        return q;
    }
    int size() const { return (int)pool.size(); }
};

inline int nv_renumber(NVertex* v, int num) {      // io/NewickVertex.java:367-375
    for (NVertex* c : v->children) num = nv_renumber(c, num);
    v->number = num;
    return num + 1;
}

inline void nv_to_string(const NVertex* v, std::string& sb) {   // io/NewickVertex.java:412-435
    if (!v->children.empty()) {
        sb.push_back('(');
        for (size_t i = 0; i < v->children.size(); ++i) { if (i) sb.push_back(','); nv_to_string(v->children[i], sb); }
        sb.push_back(')');
    }
    if (v->name) sb += *v->name;
    if (v->bl) { sb.push_back(':'); sb += java_double_to_string(*v->bl); }
    if (v->meta) sb += *v->meta;
}

inline std::string ntree_to_string(const NTree& t) {   // io/NewickTree.java:457-465
    std::string s; nv_to_string(t.root, s);
    if (t.meta) s += *t.meta;
    s.push_back(';');
    return s;
}

// ---- reader ----------------------------------------------------------------------------------------------------------
struct CharQ {
    std::string s; size_t p = 0;
    bool empty() const { return p >= s.size(); }
    char peek() const { return p < s.size() ? s[p] : '\0'; }
    char get() { if (p >= s.size()) throw NewickIOException("Unexpected end of Newick input."); return s[p++]; }
};

inline CharQ nw_strip(const std::string& str, int maxSemi) {     // NewickStringAlgorithms.strip
    CharQ q; int count = 0; bool inMeta = false;
    for (size_t i = 0; i < str.size() && count <= maxSemi; ++i) {
        char c = str[i];
        if (c == '[') inMeta = true; else if (c == ']') inMeta = false; else if (c == ';') ++count;
        if (inMeta || (c != ' ' && c != '\t' && c != '\n')) q.s.push_back(c);
    }
    return q;
}

inline std::optional<std::string> nw_read_meta(CharQ& q) {
    if (q.peek() == '[') {
        std::string m; char c;
        do { c = q.get(); m.push_back(c); } while (c != ']');
        return m;
    }
    return std::nullopt;
}

inline void nw_read_info(CharQ& q, NVertex* n) {     // NewickTreeReader.readInfo :211-226
    // name
    std::string name; char c = q.peek();
    while (!q.empty() && c != ',' && c != ';' && c != ')' && c != '[' && c != ':') { name.push_back(q.get()); c = q.peek(); }
    if (name.empty()) n->name = std::nullopt;
    else if (name[0] == '\'') {
        if (name.back() == '\'' ) n->name = name.substr(1, name.size() >= 2 ? name.size() - 2 : 0);
        else throw NewickIOException("Unmatched quote in leaf name: " + name);
    } else n->name = name;
    // branch length
    n->bl = std::nullopt;
    if (q.peek() == ':') {
        q.get();
        if (q.empty()) throw NewickIOException("Expected branch length when reading Newick tree.");
        std::string num; num.push_back(q.get()); c = q.peek();
        while (!q.empty() && c != ',' && c != ';' && c != ')' && c != '[') { num.push_back(q.get()); c = q.peek(); }
        char* end = nullptr;
        double v = std::strtod(num.c_str(), &end);
        if (end == num.c_str() || *end != '\0') throw NewickIOException("NumberFormatException: For input string: \"" + num + "\"");
        n->bl = v;
    }
    n->meta = nw_read_meta(q);
    n->number = -1;
    // HOST="?([0-9]+)"? — first match anywhere in the meta; no meta at all => 0.
    if (n->meta) {
        const std::string& m = *n->meta;
        size_t pos = 0; bool found = false;
        while ((pos = m.find("HOST=", pos)) != std::string::npos) {
            size_t k = pos + 5;
            if (k < m.size() && m[k] == '"') ++k;
            size_t st = k;
            while (k < m.size() && m[k] >= '0' && m[k] <= '9') ++k;
            if (k > st) { n->host_vertex = std::atoi(m.substr(st, k - st).c_str()); found = true; break; }
            pos += 5;
        }
        (void)found;
    } else {
        n->host_vertex = 0;
    }
}

inline NVertex* nw_read_subtree(CharQ& q, NTree& t) {    // NewickTreeReader.readSubtree :171-190
    NVertex* n = t.make();
    if (q.peek() == '(') {
        q.get();
        n->children.push_back(nw_read_subtree(q, t));
        if (q.empty() || q.get() != ',') throw NewickIOException("Expected comma when reading Newick tree.");
        for (;;) {
            n->children.push_back(nw_read_subtree(q, t));
            if (q.peek() == ',') { q.get(); continue; }
            break;
        }
        for (NVertex* c : n->children) c->parent = n;
        if (q.empty() || q.get() != ')') throw NewickIOException("Expected right parenthesis when reading Newick tree.");
        nw_read_info(q, n);
    } else {
        nw_read_info(q, n);
    }
    return n;
}

// NewickTreeReader.readTree(String, doSort=false): strip to 1 semicolon, append ';', parse, renumber post-order.
inline std::unique_ptr<NTree> nw_read_tree(const std::string& str) {
    auto t = std::make_unique<NTree>();
    CharQ q = nw_strip(str, 1);
    q.s.push_back(';');
    t->root = nw_read_subtree(q, *t);
    t->meta = nw_read_meta(q);
    if (q.empty() || q.get() != ';') throw NewickIOException("Expected semi-colon when reading Newick tree.");
    nv_renumber(t->root, 0);
    return t;
}

}  // namespace oracle
