// ORACLE — test infrastructure only. Never linked into, imported by or executed from the product path
// (sagephy_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg use it.
//
// CPU restatement of the JDK / SaGePhy numeric primitives the hot path depends on.
// PARITY UNPINNED: the reference has no tests or golden vectors and no JVM exists in this environment
// (SURVEY.md §0.2, §0.3, §8c), so the JDK semantics below are restated from the published algorithms:
//   * StrictMath.log / log10 / exp  = fdlibm 5.3 e_log.c / e_log10.c / e_exp.c (Sun Microsystems, freely
//     redistributable algorithm; HotSpot's Math.* intrinsics may differ from it by <= 1 ulp on some CPUs).
//   * Math.round(double)            = Java 8 "round half up" (floor(x + 1/2) evaluated exactly).
//   * Double.toString(double)       = JDK >= 19 specification (shortest decimal that rounds to the double, at least two
//                                     digits, closest to the true value); rendered with Java's layout rules.
//   * NumberManipulation.roundToSignificantFigures: reference src/main/java/uc/sgp/sagephy/math/NumberManipulation.java:16-27
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace oracle {

inline int32_t hi_word(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int32_t)(u >> 32); }
inline uint32_t lo_word(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (uint32_t)u; }
inline double with_hi(double x, int32_t hi) {
    uint64_t u; std::memcpy(&u, &x, 8);
    u = (u & 0xffffffffull) | ((uint64_t)(uint32_t)hi << 32);
    std::memcpy(&x, &u, 8); return x;
}

// fdlibm __ieee754_log. Call site in the reference: ExponentialDistribution.getQuantile,
// src/main/java/uc/sgp/sagephy/math/ExponentialDistribution.java:79-81 (Math.log(1.0 - p)).
inline double j_log(double x) {
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                        two54 = 1.80143985094819840000e+16,
                        Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                        Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                        Lg7 = 1.479819860511658591e-01;
    volatile double zero = 0.0;
    double hfsq, f, s, z, R, w, t1, t2, dk;
    int32_t k, hx, i, j;
    uint32_t lx;
    hx = hi_word(x); lx = lo_word(x);
    k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -two54 / zero;
        if (hx < 0) return (x - x) / zero;
        k -= 54; x *= two54; hx = hi_word(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    x = with_hi(x, hx | (i ^ 0x3ff00000));
    k += (i >> 20);
    f = x - 1.0;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) { if (k == 0) return 0.0; dk = (double)k; return dk * ln2_hi + dk * ln2_lo; }
        R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        dk = (double)k; return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    s = f / (2.0 + f);
    dk = (double)k;
    z = s * s;
    i = hx - 0x6147a;
    w = z * z;
    j = 0x6b851 - hx;
    t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    R = t2 + t1;
    if (i > 0) {
        hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

// fdlibm __ieee754_log10 (only its ceil() is consumed, NumberManipulation.java:21).
inline double j_log10(double x) {
    static const double two54 = 1.80143985094819840000e+16, ivln10 = 4.34294481903251816668e-01,
                        log10_2hi = 3.01029995663611771306e-01, log10_2lo = 3.69423907715893078616e-13;
    volatile double zero = 0.0;
    double y, z;
    int32_t i, k, hx;
    uint32_t lx;
    hx = hi_word(x); lx = lo_word(x);
    k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -two54 / zero;
        if (hx < 0) return (x - x) / zero;
        k -= 54; x *= two54; hx = hi_word(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    i = (int32_t)(((uint32_t)k & 0x80000000u) >> 31);
    hx = (hx & 0x000fffff) | ((0x3ff - i) << 20);
    y = (double)(k + i);
    x = with_hi(x, hx);
    z = y * log10_2lo + ivln10 * j_log(x);
    return z + y * log10_2hi;
}

// fdlibm __ieee754_exp. Call sites: LogNormalDistribution.sampleValue (math/LogNormalDistribution.java:196-205),
// Gamma / ChiSquared internals.
inline double j_exp(double x) {
    static const double one = 1.0, halF[2] = {0.5, -0.5}, huge = 1.0e+300, twom1000 = 9.33263618503218878990e-302,
                        o_threshold = 7.09782712893383973096e+02, u_threshold = -7.45133219101941108420e+02,
                        ln2HI[2] = {6.93147180369123816490e-01, -6.93147180369123816490e-01},
                        ln2LO[2] = {1.90821492927058770002e-10, -1.90821492927058770002e-10},
                        invln2 = 1.44269504088896338700e+00,
                        P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
                        P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                        P5 = 4.13813679705723846039e-08;
    double y, hi = 0, lo = 0, c, t;
    int32_t k = 0, xsb;
    uint32_t hx;
    hx = (uint32_t)hi_word(x);
    xsb = (int32_t)((hx >> 31) & 1);
    hx &= 0x7fffffff;
    if (hx >= 0x40862E42) {
        if (hx >= 0x7ff00000) {
            if (((hx & 0xfffff) | lo_word(x)) != 0) return x + x;
            return (xsb == 0) ? x : 0.0;
        }
        if (x > o_threshold) return huge * huge;
        if (x < u_threshold) return twom1000 * twom1000;
    }
    if (hx > 0x3fd62e42) {
        if (hx < 0x3FF0A2B2) { hi = x - ln2HI[xsb]; lo = ln2LO[xsb]; k = 1 - xsb - xsb; }
        else { k = (int32_t)(invln2 * x + halF[xsb]); t = k; hi = x - t * ln2HI[0]; lo = t * ln2LO[0]; }
        x = hi - lo;
    } else if (hx < 0x3e300000) {
        if (huge + x > one) return one + x;
    } else k = 0;
    t = x * x;
    c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return one - ((x * c) / (c - 2.0) - x);
    y = one - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) { y = with_hi(y, hi_word(y) + (k << 20)); return y; }
    y = with_hi(y, hi_word(y) + ((k + 1000) << 20)); return y * twom1000;
}

// Math.pow(10, p) for integer p: taken as the correctly rounded power of ten (exact for 0 <= p <= 22, the only
// range reachable for branch lengths >= 1e-14; fdlibm pow is < 1 ulp elsewhere — assumption recorded in DESIGN.md).
inline double j_pow10(int p) {
    struct Table { double v[701]; Table() { for (int i = 0; i <= 700; ++i) { char buf[32]; std::snprintf(buf, sizeof buf, "1e%d", i - 350); v[i] = std::strtod(buf, nullptr); } } };
    static const Table t;
    if (p < -350) return 0.0;
    if (p > 350) return HUGE_VAL;
    return t.v[p + 350];
}

// Java Math.round(double): (long) floor(x + 1/2) computed without double rounding of the sum.
inline int64_t j_round(double x) {
    if (std::isnan(x)) return 0;
    if (x >= 9.2233720368547758e18) return INT64_MAX;
    if (x <= -9.2233720368547758e18) return INT64_MIN;
    double f = std::floor(x);
    double d = x - f;                 // exact
    int64_t r = (int64_t)f;
    return d >= 0.5 ? r + 1 : r;
}

// NumberManipulation.roundToSignificantFigures (math/NumberManipulation.java:16-27).
inline double round_sig(double num, int n) {
    if (num == 0) return 0;
    const double d = std::ceil(j_log10(num < 0 ? -num : num));
    const int power = n - (int)d;
    const double magnitude = j_pow10(power);
    const int64_t shifted = j_round(num * magnitude);
    return (double)shifted / magnitude;
}

// ---- Double.toString -------------------------------------------------------------------------------------------------
// Shortest-digit search through correctly rounded printf/strtod round trips (glibc is correctly rounded), which is
// deliberately a different algorithm from the product's integer Schubfach/Ryu-style kernel code.
struct DecDigits { char d[20]; int n; int e10; };   // value = 0.d[0]d[1]... x 10^e10 ; n digits, no trailing zeros (n>=1)

inline bool dd_try(double x, int p, long long adj, DecDigits* out) {
    // p significant digits, correctly rounded, optionally nudged by adj units in the last place.
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.*e", p - 1, x);
    // buf = d.ddddde[+-]XX
    char digs[24] = {0}; int nd = 0;
    const char* s = buf;
    for (; *s && *s != 'e'; ++s) if (*s >= '0' && *s <= '9') digs[nd++] = *s;
    int ex = std::atoi(s + 1);
    if (adj != 0) {
        // big-number adjust on the digit string
        long long carry = adj;
        for (int i = nd - 1; i >= 0 && carry != 0; --i) {
            long long v = (digs[i] - '0') + carry;
            if (v >= 10) { digs[i] = (char)('0' + (v - 10)); carry = 1; }
            else if (v < 0) { digs[i] = (char)('0' + (v + 10)); carry = -1; }
            else { digs[i] = (char)('0' + v); carry = 0; }
        }
        if (carry > 0) { std::memmove(digs + 1, digs, nd - 1); digs[0] = '1'; ex += 1; }
        if (carry < 0) return false;
        if (digs[0] == '0') return false;
    }
    char back[64]; int pos = 0;
    back[pos++] = digs[0]; back[pos++] = '.';
    for (int i = 1; i < nd; ++i) back[pos++] = digs[i];
    if (nd == 1) back[pos++] = '0';
    pos += std::snprintf(back + pos, sizeof back - pos, "e%d", ex);
    double y = std::strtod(back, nullptr);
    if (y != x) return false;
    out->n = nd; std::memcpy(out->d, digs, nd); out->e10 = ex + 1;
    return true;
}

inline DecDigits shortest_digits(double x) {   // x finite, > 0
    DecDigits best{};
    for (int p = 2; p <= 17; ++p) {
        if (dd_try(x, p, 0, &best)) break;
        // asymmetric rounding interval (x a power of two): a farther neighbour may still round to x
        DecDigits a{}, b{};
        bool oka = dd_try(x, p, +1, &a), okb = dd_try(x, p, -1, &b);
        if (oka || okb) { best = oka ? a : b; break; }
    }
    while (best.n > 1 && best.d[best.n - 1] == '0') best.n--;
    return best;
}

inline std::string java_double_to_string(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "Infinity" : "-Infinity";
    std::string out;
    if (std::signbit(v)) { out.push_back('-'); v = -v; }
    if (v == 0.0) { out += "0.0"; return out; }
    DecDigits dd = shortest_digits(v);
    int e = dd.e10;                 // value = 0.DIGITS x 10^e
    if (v >= 1e-3 && v < 1e7) {
        if (e <= 0) {               // 0.000ddd
            out += "0.";
            for (int i = 0; i < -e; ++i) out.push_back('0');
            out.append(dd.d, dd.n);
        } else if (e >= dd.n) {     // ddd000.0
            out.append(dd.d, dd.n);
            for (int i = dd.n; i < e; ++i) out.push_back('0');
            out += ".0";
        } else {
            out.append(dd.d, e); out.push_back('.'); out.append(dd.d + e, dd.n - e);
        }
    } else {
        out.push_back(dd.d[0]); out.push_back('.');
        if (dd.n == 1) out.push_back('0'); else out.append(dd.d + 1, dd.n - 1);
        out.push_back('E');
        out += std::to_string(e - 1);
    }
    return out;
}

}  // namespace oracle
