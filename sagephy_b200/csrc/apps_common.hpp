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
            if (i + o->arity >= args.size()) throw SgpError("ParameterException: Expected a value after parameter " + tok);
            for (int k = 0; k < o->arity; ++k) { p.val[o->shortn].push_back(args[++i]); p.val_index[o->shortn] = (int)i; }
            continue;
        }
        const bool looks_numeric = tok.size() > 1 && tok[0] == '-' && ((tok[1] >= '0' && tok[1] <= '9') || tok[1] == '.');
        if (tok.size() > 1 && tok[0] == '-' && !looks_numeric) throw SgpError("ParameterException: Unknown option: " + tok);
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
// new BigInteger(String) (GuestTreeMachina.java:47, BranchRelaxerParameters.java:82): optional sign, decimal digits, any length.
inline SeedHost parse_seed(const std::string& s) {
    size_t i = 0; bool neg = false;
    if (i < s.size() && (s[i] == '-' || s[i] == '+')) { neg = s[i] == '-'; ++i; }
    if (i >= s.size()) throw SgpError("NumberFormatException: Zero length BigInteger");
    std::vector<uint32_t> mag;                      // little-endian base 2^32
    for (size_t k = i; k < s.size(); ++k) {
        if (s[k] < '0' || s[k] > '9') throw SgpError("NumberFormatException: For input string: \"" + s + "\"");
        unsigned long long carry = (unsigned)(s[k] - '0');
        for (auto& w : mag) { unsigned long long t = (unsigned long long)w * 10ull + carry; w = (uint32_t)t; carry = t >> 32; }
        if (carry) mag.push_back((uint32_t)carry);
    }
    while (!mag.empty() && mag.back() == 0) mag.pop_back();
    if (mag.empty()) neg = false;
    SeedHost h; h.negative = neg;
    auto word = [&](size_t k) { return k < mag.size() ? mag[k] : 0u; };
    unsigned __int128 low = 0; for (int k = 3; k >= 0; --k) low = (low << 32) | word((size_t)k);
    if (neg) low = ~low + 1;                         // two's complement mod 2^128 (higher magnitude words do not reach the low 128 bits)
    h.lo = (unsigned long long)low; h.hi = (unsigned long long)(low >> 64);
    // wide: |seed| >= 2^120 + 2^64
    bool wide = mag.size() > 4;
    if (!wide && mag.size() == 4) wide = mag[3] > 0x01000000u || (mag[3] == 0x01000000u && mag[2] >= 1u);
    h.wide = wide;
    std::string dec = s.substr(i); size_t nz = dec.find_first_not_of('0'); dec = nz == std::string::npos ? "0" : dec.substr(nz);
    if (wide) {
        const std::string low19 = dec.substr(dec.size() - 19); std::string hs = dec.substr(0, dec.size() - 19);
        h.dec_lo19 = std::strtoull(low19.c_str(), nullptr, 10);
        h.hs[0] = hs;
        std::string t = hs;                          // hs + 1 (non-negative seed: carry) or hs - 1 (negative seed: borrow)
        if (!neg) { int k = (int)t.size() - 1; while (k >= 0 && t[k] == '9') t[k--] = '0'; if (k >= 0) ++t[k]; else t.insert(t.begin(), '1'); }
        else { int k = (int)t.size() - 1; while (k >= 0 && t[k] == '0') t[k--] = '9'; --t[k]; size_t z = t.find_first_not_of('0'); t = z == std::string::npos ? "0" : t.substr(z); }
        h.hs[1] = t;
    }
    // Philox key: the seed itself when it fits 64 bits (as before), otherwise both halves folded
    const bool fits64 = !wide && ((h.hi == 0ull && !neg) || (h.hi == ~0ull && neg && (long long)h.lo < 0));
    h.philox_key = fits64 ? h.lo : (h.lo ^ (h.hi * 0x9E3779B97F4A7C15ull));
    h.text = (neg ? "-" : "") + dec;
    return h;
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
inline SeedHost random_seed() {
    unsigned long long v = 0;
    std::ifstream f("/dev/urandom", std::ios::binary);
    if (f) f.read((char*)&v, sizeof v);
    return parse_seed(std::to_string(v >> 2));
}

inline std::string slurp(const std::string& path) { std::ifstream f(path, std::ios::binary); std::stringstream ss; ss << f.rdbuf(); return ss.str(); }
inline bool exists(const std::string& path) { std::ifstream f(path); return (bool)f; }
inline std::string file_name(const std::string& path) { size_t k = path.find_last_of('/'); return k == std::string::npos ? path : path.substr(k + 1); }


}  // namespace cli
}  // namespace sgp
