#include "hostfmt.hpp"
#include "engine.hpp"
#include "rng.cuh"
#include "d2s_tables.inc"

namespace sgp {

const unsigned long long* d2s_g_table() { return SGP_D2S_G_INIT; }
int d2s_g_table_len() { return (int)(sizeof(SGP_D2S_G_INIT) / sizeof(SGP_D2S_G_INIT[0])); }
const double* p10_table() { return SGP_P10_INIT; }
int p10_table_len() { return (int)(sizeof(SGP_P10_INIT) / sizeof(SGP_P10_INIT[0])); }

const MathTables& host_math_tables() {
    static const MathTables t{SGP_D2S_G_INIT, SGP_P10_INIT};
    return t;
}

namespace { struct StrSink { std::string s; void put(char c) { s.push_back(c); } void skip(int) {} char* reserve(int) { return nullptr; } static constexpr bool kCounts = false; static constexpr bool kRandomAccess = false; }; }

std::string jdouble(double v) { StrSink k; put_double(k, host_math_tables(), v); return k.s; }
double host_round_sig(double v, int n) { return round_sig(host_math_tables(), v, n); }

void host_philox_probe(int variant, const uint32_t* ctr, const uint32_t* key, long long n, uint32_t* out) {
    for (long long i = 0; i < n; ++i) {
        const uint32_t* c = ctr + 4 * i; const uint32_t* k = key + 2 * i;
        const U4 r = variant == 0 ? philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1])
                                  : philox4x32_10_rk(c[0], c[1], c[2], c[3], philox_keys((uint64_t)k[0] | ((uint64_t)k[1] << 32)));
        out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
    }
}

void host_seed_probe(const SeedHost& h, unsigned long long i, uint32_t key[4], std::string& text) {
    std::string blob; SeedSpec s{}; s.lo = h.lo; s.hi = h.hi; s.wide = h.wide ? 1 : 0; s.negative = h.negative ? 1 : 0; s.dec_lo19 = h.dec_lo19;
    for (int k = 0; k < 2; ++k) { s.hs_off[k] = (int)blob.size(); s.hs_len[k] = (int)h.hs[k].size(); blob += h.hs[k]; }
    seed_words(s, i, key);
    StrSink k; put_seed(k, s, blob.c_str(), i); text = k.s;
}

}  // namespace sgp
