#include "hostfmt.hpp"
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

}  // namespace sgp
