// Shared helpers of the app front-ends (JCommander-like option scan, Java-style number parsing).
#pragma once
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "engine.hpp"
#include "hostfmt.hpp"

namespace sgp {
namespace cli {

struct Opt { const char* shortn; const char* longn; int arity; };

struct Parsed {
    std::map<std::string, std::vector<std::string>> val;   // keyed by short name
    std::map<std::string, int> val_index;                  // index in args of the (last) value token
    std::vector<std::string> positional;
    bool on(const char* k) const { return val.count(k) != 0; }
    std::string str(const char* k, const char* dflt) const { auto it = val.find(k); return it == val.end() ? std::string(dflt) : it->second.back(); }
};

inline Parsed scan(const std::vector<std::string>& args, const Opt* table, int n_opts) {
    Parsed p;
    for (size_t i = 0; i < args.size(); ++i) {
        const std::string& tok = args[i];
        const Opt* o = nullptr;
        for (int k = 0; k < n_opts; ++k) if (tok == table[k].shortn || tok == table[k].longn) { o = &table[k]; break; }
        if (o) {
            if (o->arity == 0) { p.val[o->shortn].push_back("true"); continue; }
            if (i + o->arity >= args.size()) throw SgpError("Expected a value after parameter " + tok);
            for (int k = 0; k < o->arity; ++k) { p.val[o->shortn].push_back(args[++i]); p.val_index[o->shortn] = (int)i; }
            continue;
        }
        const bool looks_numeric = tok.size() > 1 && tok[0] == '-' && ((tok[1] >= '0' && tok[1] <= '9') || tok[1] == '.');
        if (tok.size() > 1 && tok[0] == '-' && !looks_numeric) throw SgpError("Unknown option: " + tok);
        p.positional.push_back(tok);
    }
    return p;
}

inline double to_double(const std::string& s) {
    size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
    if (b == std::string::npos) throw SgpError("NumberFormatException: empty String");
    std::string t = s.substr(b, e - b + 1);
    char* end = nullptr; double v = std::strtod(t.c_str(), &end);
    if (end == t.c_str() || *end) throw SgpError("NumberFormatException: For input string: \"" + s + "\"");
    return v;
}
inline int to_int(const std::string& s) {
    char* end = nullptr; long v = std::strtol(s.c_str(), &end, 10);
    if (end == s.c_str() || *end) throw SgpError("NumberFormatException: For input string: \"" + s + "\"");
    return (int)v;
}
inline long long to_i64(const std::string& s) {
    char* end = nullptr; long long v = std::strtoll(s.c_str(), &end, 10);
    if (end == s.c_str() || *end) throw SgpError("seed must be an integer in the signed 64-bit range: \"" + s + "\"");
    return v;
}
inline std::string java_trim(const std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e && (unsigned char)s[b] <= ' ') ++b;
    while (e > b && (unsigned char)s[e - 1] <= ' ') --e;
    return s.substr(b, e - b);
}
inline std::vector<int> load_sizes(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw SgpError("FileNotFoundException: " + path);
    std::vector<int> v; std::string line;
    while (std::getline(f, line)) v.push_back(to_int(line));
    if (v.empty()) throw SgpError("leaf sizes file is empty: " + path);
    return v;
}
inline long long random_seed() {
    unsigned long long v = 0;
    std::ifstream f("/dev/urandom", std::ios::binary);
    if (f) f.read((char*)&v, sizeof v);
    return (long long)(v >> 2);     // keep seed + replicate index inside int64
}

inline std::string slurp(const std::string& path) { std::ifstream f(path, std::ios::binary); std::stringstream ss; ss << f.rdbuf(); return ss.str(); }
inline bool exists(const std::string& path) { std::ifstream f(path); return (bool)f; }
inline std::string file_name(const std::string& path) { size_t k = path.find_last_of('/'); return k == std::string::npos ? path : path.substr(k + 1); }


}  // namespace cli
}  // namespace sgp
