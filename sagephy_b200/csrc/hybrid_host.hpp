// `-hybrid <post-hyb timespan> <post-hyb dup fact> <post-hyb loss fact>` of GuestTreeGen / DomainTreeGen.
//
// What the reference does with it (paths relative to src/main/java/uc/sgp/sagephy/): the host positional is read as a GML graph
// (apps/SaGePhy/GuestTreeGenParameters.java:182-197, DomainTreeGenParameters.java:205-220; io/GMLReader.java:62-231,
// io/GMLGraph.java:57-140, io/GMLNode.java:49-118, io/GMLEdge.java:56-139), turned into a hybridisation DAG and validated
// (topology/HybridGraph.java:73-160,215-292, topology/DiscretisedArc.java:42-61), the creator checks its rates
// (apps/SaGePhy/GuestTreeInHybridGraphCreator.java:57-82) - and then the first call of createGuestVertex (:272) builds
// `new GuestVertex(event, X, null, eventTime, branchTime, null)`, whose constructor dereferences that last argument
// (apps/SaGePhy/GuestVertex.java:89-100: `guestTree.getRBTree().getName()`): a NullPointerException before any vertex exists.
// GuestTreeMachina.sampleGuestTree (GuestTreeMachina.java:69-115) does not catch it, so `main` prints the stack trace and
// "Use option -h or --help to show usage." and exits 0 without writing a file (GuestTreeGen.java:104-107). In this fork the mode
// therefore cannot produce a tree for any input; being a drop-in means ending the same way: every check the reference makes
// before that point raises its message, and a graph that passes them all ends with the NullPointerException. No replicate
// is run (there is nothing to accelerate) and no file is written, whatever -n says.
//
// Single pass over the text: the GML is tokenised into flat node / edge tables (the reference builds a tree of key-value pairs
// first); only the keys the hybrid graph reads are kept.
#pragma once
#include <cmath>
#include <set>
#include <string>
#include <utility>
#include <vector>
#include "host_model.hpp"

namespace sgp {
namespace hybrid {

struct Node { bool has_id = false; int id = 0; bool has_name = false; std::string name; int time_kind = 0; double time = 0; };   // time_kind: 0 none, 1 number, 2 other
struct Edge { bool has_id = false; int id = 0; bool has_src = false, has_dst = false; int src = 0, dst = 0; };
struct Graph { int directed = -1; std::vector<Node> nodes; std::vector<Edge> edges; };

inline bool ieq(const std::string& a, const char* b) {
    size_t i = 0; for (; i < a.size() && b[i]; ++i) if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
    return i == a.size() && !b[i];
}
// Character.isWhitespace for the ASCII range (space, \t \n \v \f \r and the separators FS..US)
inline bool jspace(int c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31); }

struct Scanner {
    std::string q; size_t at = 0;
    std::vector<Graph> graphs;                               // every top-level "graph [...]" is built (GMLFactory.getGraphs), the first one is used
    std::string deferred;                                    // first error of the GMLGraph / GMLNode / GMLEdge constructors (they run after the whole text has been parsed)
    void defer(const std::string& why) { if (deferred.empty()) deferred = "uc.sgp.sagephy.io.GMLIOException: " + why; }
    int peek() const { return at < q.size() ? (unsigned char)q[at] : -1; }
    [[noreturn]] void fail(const std::string& why) {
        throw SgpError("uc.sgp.sagephy.io.GMLIOException: Error parsing GML file. First 100 remaining unparsed characters:\n" + q.substr(std::min(at, q.size()), 100) +
                       "\nCaused by: uc.sgp.sagephy.io.GMLIOException: " + why);
    }
    // scope: 0 top level, 1 a graph, 2 a node of it, 3 an edge of it, 4 anything else (skipped, but still checked for syntax)
    void list(int scope, Node* nd, Edge* ed) {
        for (;;) {
            while (jspace(peek())) ++at;
            if (peek() < 0 || peek() == ']') return;
            int c = peek();
            if (!((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'))) { ++at; fail("Error parsing GML file. Expected [A-Za-z]."); }
            std::string key; while (std::isalnum(peek()) && peek() < 128) key.push_back(q[at++]);
            while (jspace(peek())) ++at;
            if (peek() == '[') {
                ++at;
                int inner = 4; Node n; Edge e;
                if (scope == 0 && ieq(key, "graph")) { inner = 1; graphs.emplace_back(); }
                else if (scope == 1 && ieq(key, "node")) inner = 2;
                else if (scope == 1 && ieq(key, "edge")) inner = 3;
                list(inner, &n, &e);
                if (peek() != ']') fail("Error parsing GML file. Expected ].");
                ++at;
                if (inner == 2) graphs.back().nodes.push_back(n);
                if (inner == 3) {
                    if (!e.has_src || !e.has_dst) defer("Invalid GML edge with ID" + (e.has_id ? std::to_string(e.id) : std::string("null")) + ": source and target attributes are required.");
                    else graphs.back().edges.push_back(e);
                }
                if (scope >= 1 && scope <= 3 && inner == 4) typed(scope, key, 'L', "", nd, ed);
                continue;
            }
            if (peek() == '"') {
                ++at; std::string s; while (peek() >= 0 && peek() != '"') s.push_back(q[at++]);
                if (peek() != '"') fail("Error parsing GML file. Expected \" to end string.");
                ++at;
                typed(scope, key, 'S', s, nd, ed);
                continue;
            }
            std::string num;
            while ((peek() >= '0' && peek() <= '9') || peek() == '.' || peek() == 'E' || peek() == 'e' || peek() == '+' || peek() == '-') num.push_back(q[at++]);
            char kind = num.find('.') != std::string::npos ? 'R' : 'I';
            char* end = nullptr;
            if (kind == 'R') { std::strtod(num.c_str(), &end); if (num.empty() || *end) fail("Error parsing GML file. Invalid number format."); }
            else {
                bool ok = !num.empty(); size_t d = (num[0] == '-' || num[0] == '+') ? 1 : 0; if (d >= num.size()) ok = false;
                for (size_t k = d; k < num.size(); ++k) if (num[k] < '0' || num[k] > '9') ok = false;
                if (ok) { long long v = std::strtoll(num.c_str(), nullptr, 10); if (v > 2147483647ll || v < -2147483648ll) ok = false; }
                if (!ok) fail("Error parsing GML file. Invalid number format.");
            }
            typed(scope, key, kind, num, nd, ed);
        }
    }
    // the typed keys of GMLGraph / GMLNode / GMLEdge (first occurrence wins, a wrong type raises)
    void typed(int scope, const std::string& key, char kind, const std::string& v, Node* nd, Edge* ed) {
        auto want = [&](char k, const char* what, const char* type) { if (kind != k) defer(std::string("Invalid attribute for GML ") + what + ": expected " + type + "."); };
        if (scope == 0) {
            if (ieq(key, "graph")) want('L', "graph", "list [...]");
        } else if (scope == 1) {
            if (ieq(key, "node")) want('L', "node", "list [...]");
            else if (ieq(key, "edge")) want('L', "edge", "list [...]");
            else if (ieq(key, "id")) want('I', "ID", "integer");
            else if (ieq(key, "name")) want('S', "name", "string");
            else if (ieq(key, "label")) want('S', "label", "string");
            else if (ieq(key, "comment")) want('S', "comment", "string");
            else if (ieq(key, "creator")) want('S', "Creator", "string");
            else if (ieq(key, "version")) want('I', "version", "integer");
            else if (ieq(key, "directed")) { want('I', "directed", "integer"); Graph& g = graphs.back(); if (kind == 'I' && g.directed < 0) g.directed = std::atoi(v.c_str()) != 0; }
        } else if (scope == 2) {
            if (ieq(key, "id")) { want('I', "ID", "integer"); if (!nd->has_id) { nd->has_id = true; nd->id = std::atoi(v.c_str()); } }
            else if (ieq(key, "name")) { want('S', "name", "string"); if (!nd->has_name) { nd->has_name = true; nd->name = v; } }
            else if (ieq(key, "label")) want('S', "label", "string");
            else if (ieq(key, "comment")) want('S', "comment", "string");
            else if (ieq(key, "edgeAnchor")) want('S', "edge anchor", "string");
            else if (ieq(key, "graphics")) want('L', "graphics", "list [...]");
            else if (ieq(key, "LabelGraphics")) want('L', "LabelGraphics", "list [...]");
            else if (ieq(key, "time") && nd->time_kind == 0) { nd->time_kind = (kind == 'I' || kind == 'R') ? 1 : 2; if (nd->time_kind == 1) nd->time = std::strtod(v.c_str(), nullptr); }
        } else if (scope == 3) {
            if (ieq(key, "id")) { want('I', "ID", "integer"); if (!ed->has_id) { ed->has_id = true; ed->id = std::atoi(v.c_str()); } }
            else if (ieq(key, "name")) want('S', "name", "string");
            else if (ieq(key, "label")) want('S', "label", "string");
            else if (ieq(key, "comment")) want('S', "comment", "string");
            else if (ieq(key, "source")) { want('I', "source", "integer"); if (!ed->has_src) { ed->has_src = true; ed->src = std::atoi(v.c_str()); } }
            else if (ieq(key, "target")) { want('I', "target", "integer"); if (!ed->has_dst) { ed->has_dst = true; ed->dst = std::atoi(v.c_str()); } }
            else if (ieq(key, "graphics")) want('L', "graphics", "list [...]");
            else if (ieq(key, "LabelGraphics")) want('L', "LabelGraphics", "list [...]");
        }
    }
};

// GMLReader.readGML(String): trim, split at '\n', drop lines whose first non-blank character is '#', and queue what is left
// WITHOUT the line breaks (so a token at the end of a line runs into the first token of the next line unless blanks separate them).
inline std::string queue_text(const std::string& text) {
    size_t b = 0, e = text.size();
    while (b < e && (unsigned char)text[b] <= ' ') ++b;
    while (e > b && (unsigned char)text[e - 1] <= ' ') --e;
    std::string out;
    for (size_t i = b; i <= e;) {
        size_t j = text.find('\n', i); if (j == std::string::npos || j > e) j = e;
        bool drop = false;
        for (size_t k = i; k < j; ++k) { if (text[k] == '#') { drop = true; break; } if (!jspace((unsigned char)text[k])) break; }
        if (!drop) out.append(text, i, j - i);
        i = j + 1;
    }
    return out;
}

inline const double kEps = 1e-10;   // DiscretisedArc.EPS

// GMLFactory.getGraphs(GMLReader.readGML(...)).get(0): syntax first, then the typed keys, then "is there a graph at all".
inline Graph read_first_graph(const std::string& gml_text) {
    Scanner sc; sc.q = queue_text(gml_text);
    sc.list(0, nullptr, nullptr);
    if (sc.peek() >= 0) sc.fail("GML file has trailing characters after supposed end.");
    if (!sc.deferred.empty()) throw SgpError(sc.deferred);
    if (sc.graphs.empty()) throw SgpError("java.lang.IndexOutOfBoundsException: Index 0 out of bounds for length 0");
    return sc.graphs.front();
}

// new GuestTreeInHybridGraphCreator(...) and the first createGuestVertex; never returns.
[[noreturn]] inline void create_and_sample(Graph& g, double lambda, double mu, double lambda_fact, double mu_fact, double timespan, double rho) {
    auto iae = [](const std::string& m) { return SgpError("java.lang.IllegalArgumentException: " + m); };
    if (g.directed < 0) throw SgpError("java.lang.NullPointerException");                                                   // unboxing `Boolean directed == null`
    if (!g.directed) throw iae("Invalid hybrid graph: graph is not directed.");
    const int n = (int)g.nodes.size();
    std::vector<double> t(n, 0.0); std::vector<std::string> name(n); std::vector<char> named(n, 0);
    std::set<int> ids;
    for (const Node& v : g.nodes) {
        if (!v.has_id) throw iae("Invalid hybrid graph: Vertex lacks ID.");
        if (!ids.insert(v.id).second) throw iae("Invalid hybrid graph: Duplicate vertex with ID " + std::to_string(v.id) + ".");
        if (v.id < 0 || v.id >= n) throw SgpError("java.lang.ArrayIndexOutOfBoundsException: Index " + std::to_string(v.id) + " out of bounds for length " + std::to_string(n));
        name[v.id] = v.name; named[v.id] = v.has_name;
        if (v.time_kind == 0) throw iae("Invalid hybrid graph: No absolute time specified for vertex with ID " + std::to_string(v.id) + ".");
        if (v.time_kind == 2) throw iae("Invalid hybrid graph: Time specified for vertex with ID " + std::to_string(v.id) + " is not a number.");
        t[v.id] = v.time;
    }
    std::vector<std::vector<std::pair<int, double>>> in(n), out(n);   // (other end, arc time), in insertion order
    std::set<std::pair<int, int>> have;
    for (const Edge& a : g.edges) {
        const int x = a.src, y = a.dst;
        if (x < 0 || y < 0 || x >= n || y >= n) throw iae("Invalid hybrid graph: Invalid arc source or target ID (" + std::to_string(x) + "," + std::to_string(y) + ").");
        const double tx = t[x]; double ty = t[y];
        if (std::fabs(tx - ty) <= kEps) { ty = tx - kEps / 2.0; t[y] = ty; }
        const double at = tx - ty;
        if (at < 0) throw iae("Invalid arc time: Timespan is less than 0. Time of source must be greater than or equal to time of target.");
        if (x == y) throw iae("loops not allowed");                                                  // SimpleDirectedGraph.addEdge
        if (!have.insert({x, y}).second) continue;                                                   // a parallel arc is not added
        const double span = std::fabs(at) <= kEps ? 0.0 : at;                                        // times = {sourcet, sourcet} for a hybrid arc
        out[x].push_back({y, span}); in[y].push_back({x, span});
    }
    int source = -1; std::vector<int> sinks;
    for (int x = 0; x < n; ++x) { if (in[x].empty() && source < 0) source = x; if (out[x].empty()) sinks.push_back(x); }
    if (source < 0) throw SgpError("java.util.NoSuchElementException");                              // sources.iterator().next()
    {   // ConnectivityInspector: weakly connected
        std::vector<char> seen(n, 0); std::vector<int> st{0}; seen[0] = 1; int cnt = 0;
        while (!st.empty()) { int x = st.back(); st.pop_back(); ++cnt; for (auto& p : out[x]) if (!seen[p.first]) { seen[p.first] = 1; st.push_back(p.first); } for (auto& p : in[x]) if (!seen[p.first]) { seen[p.first] = 1; st.push_back(p.first); } }
        if (cnt != n) throw iae("Invalid hybrid graph: Graph is not connected.");
    }
    for (int x = 0; x < n; ++x) {                                                                     // HybridGraph.getVertexTyp
        const size_t ni = in[x].size(), no = out[x].size();
        const std::string xs = std::to_string(x);
        if (ni == 0) { if (no != 1) throw iae("Invalid stem tip vertex " + xs + ": must have exactly one child (or did you forget the incoming edge?)."); continue; }
        if (ni == 1) {
            const double tp = in[x][0].second;
            if (no == 0) { if (tp <= kEps) throw iae("Invalid arc time: Must be greater than 0 for arc ending in " + xs + "."); continue; }
            if (no == 1) {
                const double tc = out[x][0].second;
                if (std::fabs(tp) <= kEps) continue;                                                  // autopolyploidic hybrid
                if (std::fabs(tc) <= kEps) continue;                                                  // extinct hybrid donor
                throw iae("Invalid arc times for vertex " + xs + ". One of the surrounding arcs must have time span 0.");
            }
            {   // two (or more: the first two) children
                if (tp <= kEps) throw iae("Invalid arc time: Must be greater than 0 for arc ending in " + xs + ".");
                const double t1 = out[x][0].second, t2 = out[x][1].second;
                if (t1 > kEps && t2 > kEps) continue;                                                 // speciation
                if ((t1 > kEps && std::fabs(t2) <= kEps) || (t2 > kEps && std::fabs(t1) <= kEps)) continue;   // hybrid donor
                throw iae("Invalid vertex " + xs + ": Wrong degree or times of outgoing arcs.");
            }
        }
        if (ni == 2 && no == 1) {
            if (std::fabs(in[x][0].second) <= kEps && std::fabs(in[x][1].second) <= kEps) continue;   // allopolyploidic hybrid
            throw iae("Invalid outgoing arc times for vertex " + xs + ": Must be greater than 0.");
        }
        throw iae("Invalid vertex " + xs + ": Wrong degree of in- or outgoing arcs.");
    }
    for (int x : sinks) if (!named[x] || name[x].empty()) throw iae("Invalid hybrid graph: Missing name for leaf vertex " + std::to_string(x) + ".");
    // GuestTreeInHybridGraphCreator.java:74-81
    if (lambda < 0 || mu < 0 || lambda_fact < 0 || mu_fact < 0 || timespan < 0) throw iae("Cannot have rate or rate factor less than 0.");
    if (rho < 0 || rho > 1) throw iae("Cannot have leaf sampling probability outside [0,1].");
    // GuestVertex.java:100 <- GuestTreeInHybridGraphCreator.java:272 <- :95 <- GuestTreeMachina.java:92
    throw SgpError("java.lang.NullPointerException");
}

}  // namespace hybrid
}  // namespace sgp
