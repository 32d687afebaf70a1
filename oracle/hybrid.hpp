// ORACLE — test infrastructure only (see oracle/README.md). PARITY UNPINNED.
//
// `-hybrid` host graphs. Restates, in the reference's own order,
//   io/GMLReader.java:62-231 (text -> key/value tree; comment lines dropped, lines queued WITHOUT their line breaks),
//   misc/CharQueue.java (peek() on an empty queue is CharacterIterator.DONE),
//   io/GMLFactory.java:26-38 (every top-level "graph" list becomes a GMLGraph), io/GMLGraph.java:57-140,
//   io/GMLNode.java:49-118,244-275, io/GMLEdge.java:56-139 (typed keys, first occurrence wins, the rest are attributes),
//   topology/HybridGraph.java:73-160 (constructor), :215-292 (getVertexTyp), topology/DAG.java:254-275 (sources / sinks),
//   topology/DiscretisedArc.java:42-61 (arc time), jgrapht SimpleDirectedGraph.addEdge (no loops, no parallel arcs),
//   apps/SaGePhy/GuestTreeInHybridGraphCreator.java:57-82 (creator), :86-96 (first lineage), :258-272 (createGuestVertex),
//   apps/SaGePhy/GuestVertex.java:89-100 (constructor: `guestTree.getRBTree().getName()` on the null the creator passes).
// The last item means that the reference cannot build a single guest vertex in this mode: every run that gets past the
// argument and graph checks ends in a NullPointerException, which GuestTreeGen.main / DomainTreeGen.main catch (stack trace,
// "Use option -h or --help to show usage.", exit 0, no files). The restatement therefore stops there too; the exception
// class names are what `Throwable.toString()` prints as the first line of the stack trace (JDK >= 9 wording for the
// index exceptions).
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {
namespace gml {

struct KV;
using List = std::vector<KV>;
enum class Type { INTEGER, REAL, STRING, LIST };
struct KV { std::string key; Type type; long ival = 0; double dval = 0; std::string sval; std::shared_ptr<List> list; };

struct GMLIOException : std::runtime_error { using std::runtime_error::runtime_error; };

struct CharQueue {
    std::string s; size_t i = 0;
    static constexpr int DONE = 0xFFFF;
    bool isEmpty() const { return i >= s.size(); }
    int peek() const { return isEmpty() ? DONE : (unsigned char)s[i]; }
    int get() { if (isEmpty()) throw std::runtime_error("java.lang.NullPointerException"); return (unsigned char)s[i++]; }
    void put(const std::string& t) { s += t; }
};

inline bool isWhitespace(int c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31); }   // Character.isWhitespace, ASCII part

List readList(CharQueue& q);

inline std::string readKey(CharQueue& q) {
    std::string sb;
    int c = q.get();
    if (!((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'))) throw GMLIOException("Error parsing GML file. Expected [A-Za-z].");
    sb.push_back((char)c);
    while ((q.peek() >= 'A' && q.peek() <= 'Z') || (q.peek() >= 'a' && q.peek() <= 'z') || (q.peek() >= '0' && q.peek() <= '9')) sb.push_back((char)q.get());
    return sb;
}

inline std::string readString(CharQueue& q) {
    q.get();
    std::string sb;
    while (q.peek() != '"' && !q.isEmpty()) sb.push_back((char)q.get());
    if (q.peek() != '"') throw GMLIOException("Error parsing GML file. Expected \" to end string.");
    q.get();
    return sb;     // StringEscapeUtils.unescapeHtml4 is not applied here: names only matter for being non-empty
}

inline KV readKeyValue(CharQueue& q) {
    KV kv; kv.key = readKey(q);
    while (isWhitespace(q.peek())) q.get();
    if (q.peek() == '[') {
        q.get();
        kv.type = Type::LIST; kv.list = std::make_shared<List>(readList(q));
        if (q.peek() != ']') throw GMLIOException("Error parsing GML file. Expected ].");
        q.get();
        return kv;
    }
    if (q.peek() == '"') { kv.type = Type::STRING; kv.sval = readString(q); return kv; }
    std::string s;
    while ((q.peek() >= '0' && q.peek() <= '9') || q.peek() == '.' || q.peek() == 'E' || q.peek() == 'e' || q.peek() == '+' || q.peek() == '-') s.push_back((char)q.get());
    if (s.find('.') != std::string::npos) {
        char* end = nullptr; double d = std::strtod(s.c_str(), &end);
        if (*end) throw GMLIOException("Error parsing GML file. Invalid number format.");
        kv.type = Type::REAL; kv.dval = d; return kv;
    }
    // new Integer(s): optional sign, decimal digits, 32-bit range
    size_t d0 = (!s.empty() && (s[0] == '-' || s[0] == '+')) ? 1 : 0;
    bool ok = s.size() > d0;
    for (size_t k = d0; k < s.size(); ++k) ok = ok && s[k] >= '0' && s[k] <= '9';
    long long v = ok ? std::strtoll(s.c_str(), nullptr, 10) : 0;
    if (!ok || v > 2147483647ll || v < -2147483648ll) throw GMLIOException("Error parsing GML file. Invalid number format.");
    kv.type = Type::INTEGER; kv.ival = (long)v;
    return kv;
}

inline List readList(CharQueue& q) {
    List list;
    while (true) {
        while (isWhitespace(q.peek())) q.get();
        if (q.isEmpty() || q.peek() == ']') break;
        list.push_back(readKeyValue(q));
    }
    return list;
}

inline List readGML(const std::string& str0) {
    // String.trim(): code points <= ' ' at both ends
    size_t b = 0, e = str0.size();
    while (b < e && (unsigned char)str0[b] <= ' ') ++b;
    while (e > b && (unsigned char)str0[e - 1] <= ' ') --e;
    std::string str = str0.substr(b, e - b);
    CharQueue q;
    size_t i = 0;
    while (i <= str.size()) {
        size_t j = str.find('\n', i); if (j == std::string::npos) j = str.size();
        std::string ln = str.substr(i, j - i);
        bool ignore = false;
        for (char c : ln) { if (c == '#') { ignore = true; break; } if (!isWhitespace((unsigned char)c)) { ignore = false; break; } }
        if (!ignore) q.put(ln);
        i = j + 1;
    }
    try {
        List list = readList(q);
        if (!q.isEmpty()) throw GMLIOException("GML file has trailing characters after supposed end.");
        return list;
    } catch (std::exception& ex) {
        std::string rest; int n = 0; while (n < 100 && !q.isEmpty()) { rest.push_back((char)q.get()); ++n; }
        throw GMLIOException("uc.sgp.sagephy.io.GMLIOException: Error parsing GML file. First 100 remaining unparsed characters:\n" + rest +
                             "\nCaused by: uc.sgp.sagephy.io.GMLIOException: " + ex.what());
    }
}

inline bool eqic(const std::string& a, const char* b) {
    size_t n = std::char_traits<char>::length(b); if (a.size() != n) return false;
    for (size_t i = 0; i < n; ++i) if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
    return true;
}
inline void expect(const KV& kv, Type t, const char* what) {
    if (kv.type != t) throw GMLIOException(std::string("uc.sgp.sagephy.io.GMLIOException: Invalid attribute for GML ") + what + ": expected " +
                                           (t == Type::INTEGER ? "integer" : t == Type::STRING ? "string" : "list [...]") + ".");
}

struct GMLNode {
    bool hasId = false; int id = 0; bool hasName = false; std::string name; List attributes;
    bool f_label = false, f_comment = false, f_anchor = false, f_gfx = false, f_lgfx = false;
    explicit GMLNode(const List& values) {
        for (const KV& kv : values) {
            if (eqic(kv.key, "id")) { expect(kv, Type::INTEGER, "ID"); if (!hasId) { hasId = true; id = (int)kv.ival; } else attributes.push_back(kv); }
            else if (eqic(kv.key, "name")) { expect(kv, Type::STRING, "name"); if (!hasName) { hasName = true; name = kv.sval; } else attributes.push_back(kv); }
            else if (eqic(kv.key, "label")) { expect(kv, Type::STRING, "label"); if (!f_label) f_label = true; else attributes.push_back(kv); }
            else if (eqic(kv.key, "comment")) { expect(kv, Type::STRING, "comment"); if (!f_comment) f_comment = true; else attributes.push_back(kv); }
            else if (eqic(kv.key, "edgeAnchor")) { expect(kv, Type::STRING, "edge anchor"); if (!f_anchor) f_anchor = true; else attributes.push_back(kv); }
            else if (eqic(kv.key, "graphics")) { expect(kv, Type::LIST, "graphics"); if (!f_gfx) f_gfx = true; else attributes.push_back(kv); }
            else if (eqic(kv.key, "LabelGraphics")) { expect(kv, Type::LIST, "LabelGraphics"); if (!f_lgfx) f_lgfx = true; else attributes.push_back(kv); }
            else attributes.push_back(kv);
        }
    }
    const KV* attribute(const char* key) const { for (const KV& kv : attributes) if (eqic(kv.key, key)) return &kv; return nullptr; }
};

struct GMLEdge {
    bool hasId = false; int id = 0; bool hasSource = false, hasTarget = false; int source = 0, target = 0;
    explicit GMLEdge(const List& values) {
        bool f_name = false, f_label = false, f_comment = false, f_gfx = false, f_lgfx = false;
        for (const KV& kv : values) {
            if (eqic(kv.key, "id")) { expect(kv, Type::INTEGER, "ID"); if (!hasId) { hasId = true; id = (int)kv.ival; } }
            else if (eqic(kv.key, "name")) { expect(kv, Type::STRING, "name"); f_name = true; }
            else if (eqic(kv.key, "label")) { expect(kv, Type::STRING, "label"); f_label = true; }
            else if (eqic(kv.key, "comment")) { expect(kv, Type::STRING, "comment"); f_comment = true; }
            else if (eqic(kv.key, "source")) { expect(kv, Type::INTEGER, "source"); if (!hasSource) { hasSource = true; source = (int)kv.ival; } }
            else if (eqic(kv.key, "target")) { expect(kv, Type::INTEGER, "target"); if (!hasTarget) { hasTarget = true; target = (int)kv.ival; } }
            else if (eqic(kv.key, "graphics")) { expect(kv, Type::LIST, "graphics"); f_gfx = true; }
            else if (eqic(kv.key, "LabelGraphics")) { expect(kv, Type::LIST, "LabelGraphics"); f_lgfx = true; }
        }
        (void)f_name; (void)f_label; (void)f_comment; (void)f_gfx; (void)f_lgfx;
        if (!hasSource || !hasTarget)
            throw GMLIOException("uc.sgp.sagephy.io.GMLIOException: Invalid GML edge with ID" + (hasId ? std::to_string(id) : std::string("null")) + ": source and target attributes are required.");
    }
};

struct GMLGraph {
    std::vector<GMLNode> nodes; std::vector<GMLEdge> edges; int directed = -1;   // -1: null
    explicit GMLGraph(const List& values) {
        bool f_id = false, f_name = false, f_label = false, f_comment = false, f_creator = false, f_version = false;
        for (const KV& kv : values) {
            if (eqic(kv.key, "node")) { expect(kv, Type::LIST, "node"); nodes.emplace_back(*kv.list); }
            else if (eqic(kv.key, "edge")) { expect(kv, Type::LIST, "edge"); edges.emplace_back(*kv.list); }
            else if (eqic(kv.key, "id")) { expect(kv, Type::INTEGER, "ID"); f_id = true; }
            else if (eqic(kv.key, "name")) { expect(kv, Type::STRING, "name"); f_name = true; }
            else if (eqic(kv.key, "label")) { expect(kv, Type::STRING, "label"); f_label = true; }
            else if (eqic(kv.key, "comment")) { expect(kv, Type::STRING, "comment"); f_comment = true; }
            else if (eqic(kv.key, "Creator")) { expect(kv, Type::STRING, "Creator"); f_creator = true; }
            else if (eqic(kv.key, "version")) { expect(kv, Type::INTEGER, "version"); f_version = true; }
            else if (eqic(kv.key, "directed")) { expect(kv, Type::INTEGER, "directed"); if (directed < 0) directed = kv.ival == 0 ? 0 : 1; }
        }
        (void)f_id; (void)f_name; (void)f_label; (void)f_comment; (void)f_creator; (void)f_version;
    }
};

// GMLFactory.getGraphs(gml).get(0)
inline GMLGraph first_graph(const List& gml) {
    std::vector<GMLGraph> graphs;
    for (const KV& kv : gml) if (eqic(kv.key, "graph")) { expect(kv, Type::LIST, "graph"); graphs.emplace_back(*kv.list); }
    if (graphs.empty()) throw std::out_of_range("java.lang.IndexOutOfBoundsException: Index 0 out of bounds for length 0");
    return graphs[0];
}

}  // namespace gml

// topology/HybridGraph.java:73-160 + GuestTreeInHybridGraphCreator.java:57-82,86-96,258-272 + GuestVertex.java:100
[[noreturn]] inline void hybrid_creator_and_first_vertex(const gml::GMLGraph& rawG, double lambda, double mu, double lambdaChangeFact, double muChangeFact,
                                                         double postHybTimespan, double rho) {
    const double EPS = 1e-10;
    auto IAE = [](const std::string& m) { return std::invalid_argument("java.lang.IllegalArgumentException: " + m); };
    if (rawG.directed < 0) throw std::runtime_error("java.lang.NullPointerException");
    if (!rawG.directed) throw IAE("Invalid hybrid graph: graph is not directed.");
    const int n = (int)rawG.nodes.size();
    std::set<int> uniq; std::vector<std::string> vertexNames(n); std::vector<bool> hasName(n, false); std::vector<double> vertexTimes(n, 0.0);
    for (const gml::GMLNode& v : rawG.nodes) {
        if (!v.hasId) throw IAE("Invalid hybrid graph: Vertex lacks ID.");
        const int x = v.id;
        if (!uniq.insert(x).second) throw IAE("Invalid hybrid graph: Duplicate vertex with ID " + std::to_string(x) + ".");
        if (x < 0 || x >= n) throw std::out_of_range("java.lang.ArrayIndexOutOfBoundsException: Index " + std::to_string(x) + " out of bounds for length " + std::to_string(n));
        vertexNames[x] = v.name; hasName[x] = v.hasName;
        const gml::KV* t = v.attribute("time");
        if (!t) throw IAE("Invalid hybrid graph: No absolute time specified for vertex with ID " + std::to_string(x) + ".");
        if (t->type == gml::Type::INTEGER) vertexTimes[x] = (double)t->ival;
        else if (t->type == gml::Type::REAL) vertexTimes[x] = t->dval;
        else throw IAE("Invalid hybrid graph: Time specified for vertex with ID " + std::to_string(x) + " is not a number.");
    }
    struct Arc { int x, y; double arcTime; };
    std::vector<Arc> arcs; std::vector<std::vector<int>> incoming(n), outgoing(n);
    for (const gml::GMLEdge& a : rawG.edges) {
        const int x = a.source, y = a.target;
        if (x < 0 || y < 0 || x >= n || y >= n) throw IAE("Invalid hybrid graph: Invalid arc source or target ID (" + std::to_string(x) + "," + std::to_string(y) + ").");
        double tx = vertexTimes[x], ty = vertexTimes[y];
        if (std::fabs(tx - ty) <= EPS) { ty = tx - EPS / 2.0; vertexTimes[y] = ty; }
        const double at = tx - ty;                                      // DiscretisedArc constructor
        if (at < 0) throw IAE("Invalid arc time: Timespan is less than 0. Time of source must be greater than or equal to time of target.");
        const double arcTime = std::fabs(at) <= EPS ? 0.0 : at;         // times = {sourcet, sourcet}: getArcTime() = 0
        if (x == y) throw IAE("loops not allowed");
        bool parallel = false; for (int k : outgoing[x]) if (arcs[k].y == y) parallel = true;
        if (parallel) continue;
        arcs.push_back({x, y, arcTime}); outgoing[x].push_back((int)arcs.size() - 1); incoming[y].push_back((int)arcs.size() - 1);
    }
    // update(): sources / sinks; source = sources.iterator().next()
    std::vector<int> sources, sinks;
    for (int x = 0; x < n; ++x) { if (incoming[x].empty()) sources.push_back(x); if (outgoing[x].empty()) sinks.push_back(x); }
    if (sources.empty()) throw std::runtime_error("java.util.NoSuchElementException");
    // ConnectivityInspector.connectedSets().size() != 1
    {
        std::vector<int> comp(n, -1); int ncomp = 0;
        for (int s = 0; s < n; ++s) if (comp[s] < 0) {
            std::vector<int> st{s}; comp[s] = ncomp;
            while (!st.empty()) { int x = st.back(); st.pop_back(); for (int k : outgoing[x]) if (comp[arcs[k].y] < 0) { comp[arcs[k].y] = ncomp; st.push_back(arcs[k].y); } for (int k : incoming[x]) if (comp[arcs[k].x] < 0) { comp[arcs[k].x] = ncomp; st.push_back(arcs[k].x); } }
            ++ncomp;
        }
        if (ncomp != 1) throw IAE("Invalid hybrid graph: Graph is not connected.");
    }
    for (int x = 0; x < n; ++x) {   // getVertexTyp
        const int indegree = (int)incoming[x].size(), outdegree = (int)outgoing[x].size();
        const std::string xs = std::to_string(x);
        if (indegree == 0) { if (outdegree != 1) throw IAE("Invalid stem tip vertex " + xs + ": must have exactly one child (or did you forget the incoming edge?)."); continue; }
        if (indegree == 1) {
            const double par = arcs[incoming[x][0]].arcTime;
            if (outdegree == 0) { if (par <= EPS) throw IAE("Invalid arc time: Must be greater than 0 for arc ending in " + xs + "."); continue; }
            if (outdegree == 1) {
                const double tp = par, tc = arcs[outgoing[x][0]].arcTime;
                if (std::fabs(tp) <= EPS) continue;
                if (std::fabs(tp) > EPS && std::fabs(tc) <= EPS) continue;
                throw IAE("Invalid arc times for vertex " + xs + ". One of the surrounding arcs must have time span 0.");
            }
            if (par <= EPS) throw IAE("Invalid arc time: Must be greater than 0 for arc ending in " + xs + ".");
            const double t1 = arcs[outgoing[x][0]].arcTime, t2 = arcs[outgoing[x][1]].arcTime;
            if (t1 > EPS && t2 > EPS) continue;
            if ((t1 > EPS && std::fabs(t2) <= EPS) || (t2 > EPS && std::fabs(t1) <= EPS)) continue;
            throw IAE("Invalid vertex " + xs + ": Wrong degree or times of outgoing arcs.");
        }
        if (indegree == 2 && outdegree == 1) {
            const double t1 = arcs[incoming[x][0]].arcTime, t2 = arcs[incoming[x][1]].arcTime;
            if (std::fabs(t1) <= EPS && std::fabs(t2) <= EPS) continue;
            throw IAE("Invalid outgoing arc times for vertex " + xs + ": Must be greater than 0.");
        }
        throw IAE("Invalid vertex " + xs + ": Wrong degree of in- or outgoing arcs.");
    }
    for (int x : sinks) if (!hasName[x] || vertexNames[x].empty()) throw IAE("Invalid hybrid graph: Missing name for leaf vertex " + std::to_string(x) + ".");
    // creator (stem override and the 0-stem hack change arc times only)
    if (lambda < 0 || mu < 0 || lambdaChangeFact < 0 || muChangeFact < 0 || postHybTimespan < 0) throw IAE("Cannot have rate or rate factor less than 0.");
    if (rho < 0 || rho > 1) throw IAE("Cannot have leaf sampling probability outside [0,1].");
    // createUnprunedTree -> createGuestVertex -> new GuestVertex(event, X, null, eventTime, branchTime, null) -> guestTree.getRBTree()
    throw std::runtime_error("java.lang.NullPointerException");
}

}  // namespace oracle
