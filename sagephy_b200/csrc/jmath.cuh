// Java-exact scalar math for host and device code.
//
// Every function here must produce bit-identical results on the CPU (host build of the same source) and on sm_100a:
// arithmetic goes through XADD/XMUL/... which map to the non-contractable __dadd_rn/__dmul_rn/... intrinsics on the
// device, so nvcc never fuses them into FMAs (Java doubles are unfused IEEE-754).
//
//   jlog / jlog10 / jexp : fdlibm 5.3 algorithms (= java.lang.StrictMath), see DESIGN.md "numerics".
//   jround               : Java 8 Math.round (round half up)
//   round_sig            : reference math/NumberManipulation.java:16-27 (roundToSignificantFigures)
//   d2s_*                : java.lang.Double.toString (JDK >= 19: Schubfach shortest decimal) incl. Java's layout rules
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define SGP_HD __host__ __device__ __forceinline__
#define SGP_HD_NOINL __host__ __device__ __noinline__
#else
#define SGP_HD inline
#define SGP_HD_NOINL inline
#endif

#if defined(__CUDA_ARCH__)
#define XADD(a, b) __dadd_rn((a), (b))
#define XSUB(a, b) __dsub_rn((a), (b))
#define XMUL(a, b) __dmul_rn((a), (b))
#define XDIV(a, b) __ddiv_rn((a), (b))
#else
#define XADD(a, b) ((a) + (b))
#define XSUB(a, b) ((a) - (b))
#define XMUL(a, b) ((a) * (b))
#define XDIV(a, b) ((a) / (b))
#endif

namespace sgp {

SGP_HD int32_t dbl_hi(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (int32_t)(u >> 32);
#endif
}
SGP_HD uint32_t dbl_lo(double x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (uint32_t)u;
#endif
}
SGP_HD double dbl_make(int32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, (int32_t)lo);
#else
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | lo; double x; memcpy(&x, &u, 8); return x;
#endif
}
SGP_HD uint64_t dbl_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
SGP_HD double dbl_inf() { return dbl_make(0x7ff00000, 0u); }
SGP_HD double dbl_nan() { return dbl_make(0x7ff80000, 0u); }

// ---- fdlibm log --------------------------------------------------------------------------------------------------------
SGP_HD double jlog(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10, two54 = 1.80143985094819840000e+16,
                 Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    int32_t hx = dbl_hi(x); uint32_t lx = dbl_lo(x);
    int32_t k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -dbl_inf();
        if (hx < 0) return dbl_nan();
        k -= 54; x = XMUL(x, two54); hx = dbl_hi(x); lx = dbl_lo(x);
    }
    if (hx >= 0x7ff00000) return XADD(x, x);
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;
    x = dbl_make(hx | (i ^ 0x3ff00000), lx);
    k += (i >> 20);
    double f = XSUB(x, 1.0);
    double dk = (double)k;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) { if (k == 0) return 0.0; return XADD(XMUL(dk, ln2_hi), XMUL(dk, ln2_lo)); }
        double R = XMUL(XMUL(f, f), XSUB(0.5, XMUL(0.33333333333333333, f)));
        if (k == 0) return XSUB(f, R);
        return XSUB(XMUL(dk, ln2_hi), XSUB(XSUB(R, XMUL(dk, ln2_lo)), f));
    }
    double s = XDIV(f, XADD(2.0, f));
    double z = XMUL(s, s);
    i = hx - 0x6147a;
    double w = XMUL(z, z);
    int32_t j = 0x6b851 - hx;
    double t1 = XMUL(w, XADD(Lg2, XMUL(w, XADD(Lg4, XMUL(w, Lg6)))));
    double t2 = XMUL(z, XADD(Lg1, XMUL(w, XADD(Lg3, XMUL(w, XADD(Lg5, XMUL(w, Lg7)))))));
    i |= j;
    double R = XADD(t2, t1);
    if (i > 0) {
        double hfsq = XMUL(XMUL(0.5, f), f);
        if (k == 0) return XSUB(f, XSUB(hfsq, XMUL(s, XADD(hfsq, R))));
        return XSUB(XMUL(dk, ln2_hi), XSUB(XSUB(hfsq, XADD(XMUL(s, XADD(hfsq, R)), XMUL(dk, ln2_lo))), f));
    }
    if (k == 0) return XSUB(f, XMUL(s, XSUB(f, R)));
    return XSUB(XMUL(dk, ln2_hi), XSUB(XSUB(XMUL(s, XSUB(f, R)), XMUL(dk, ln2_lo)), f));
}

SGP_HD double jlog10(double x) {
    const double two54 = 1.80143985094819840000e+16, ivln10 = 4.34294481903251816668e-01,
                 log10_2hi = 3.01029995663611771306e-01, log10_2lo = 3.69423907715893078616e-13;
    int32_t hx = dbl_hi(x); uint32_t lx = dbl_lo(x);
    int32_t k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -dbl_inf();
        if (hx < 0) return dbl_nan();
        k -= 54; x = XMUL(x, two54); hx = dbl_hi(x); lx = dbl_lo(x);
    }
    if (hx >= 0x7ff00000) return XADD(x, x);
    k += (hx >> 20) - 1023;
    int32_t i = (int32_t)(((uint32_t)k & 0x80000000u) >> 31);
    hx = (hx & 0x000fffff) | ((0x3ff - i) << 20);
    double y = (double)(k + i);
    x = dbl_make(hx, lx);
    double z = XADD(XMUL(y, log10_2lo), XMUL(ivln10, jlog(x)));
    return XADD(z, XMUL(y, log10_2hi));
}

SGP_HD double jexp(double x) {
    const double one = 1.0, huge = 1.0e+300, twom1000 = 9.33263618503218878990e-302,
                 o_threshold = 7.09782712893383973096e+02, u_threshold = -7.45133219101941108420e+02,
                 ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00,
                 P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    double hi = 0, lo = 0; int32_t k = 0;
    uint32_t hx = (uint32_t)dbl_hi(x);
    int32_t xsb = (int32_t)((hx >> 31) & 1);
    hx &= 0x7fffffff;
    if (hx >= 0x40862E42) {
        if (hx >= 0x7ff00000) {
            if (((hx & 0xfffff) | dbl_lo(x)) != 0) return XADD(x, x);
            return (xsb == 0) ? x : 0.0;
        }
        if (x > o_threshold) return XMUL(huge, huge);
        if (x < u_threshold) return XMUL(twom1000, twom1000);
    }
    if (hx > 0x3fd62e42) {
        if (hx < 0x3FF0A2B2) {
            hi = XSUB(x, xsb ? -ln2HI : ln2HI); lo = xsb ? -ln2LO : ln2LO; k = 1 - xsb - xsb;
        } else {
            k = (int32_t)XADD(XMUL(invln2, x), xsb ? -0.5 : 0.5);
            double t = (double)k;
            hi = XSUB(x, XMUL(t, ln2HI)); lo = XMUL(t, ln2LO);
        }
        x = XSUB(hi, lo);
    } else if (hx < 0x3e300000) {
        if (XADD(huge, x) > one) return XADD(one, x);
    } else k = 0;
    double t = XMUL(x, x);
    double c = XSUB(x, XMUL(t, XADD(P1, XMUL(t, XADD(P2, XMUL(t, XADD(P3, XMUL(t, XADD(P4, XMUL(t, P5))))))))));
    if (k == 0) return XSUB(one, XSUB(XDIV(XMUL(x, c), XSUB(c, 2.0)), x));
    double y = XSUB(one, XSUB(XSUB(lo, XDIV(XMUL(x, c), XSUB(2.0, c))), hi));
    if (k >= -1021) return dbl_make(dbl_hi(y) + (k << 20), dbl_lo(y));
    y = dbl_make(dbl_hi(y) + ((k + 1000) << 20), dbl_lo(y));
    return XMUL(y, twom1000);
}

// ---- tables (defined once per translation unit that needs them; see tables.cuh) ---------------------------------------
struct MathTables {
    const unsigned long long* g;   // Schubfach 126-bit powers of ten, (g1,g0) pairs for k = K_MIN..K_MAX
    const double* p10;             // 10^p for p = P10_MIN..P10_MAX (correctly rounded; 0 / +inf outside the double range)
};
constexpr int kD2sKMin = -324, kD2sKMax = 292, kP10Min = -340, kP10Max = 340;

SGP_HD double jpow10(const MathTables& T, int p) {
    if (p < kP10Min) return 0.0;
    if (p > kP10Max) return dbl_inf();
    return T.p10[p - kP10Min];
}

SGP_HD int64_t jround(double x) {
    if (x != x) return 0;
    if (x >= 9.2233720368547758e18) return INT64_MAX;
    if (x <= -9.2233720368547758e18) return INT64_MIN;
    double f = floor(x);
    double d = XSUB(x, f);
    int64_t r = (int64_t)f;
    return d >= 0.5 ? r + 1 : r;
}

SGP_HD double round_sig(const MathTables& T, double num, int n) {
    if (num == 0) return 0;
    const double d = ceil(jlog10(num < 0 ? -num : num));
    const int power = n - (int)d;
    const double magnitude = jpow10(T, power);
    const int64_t shifted = jround(XMUL(num, magnitude));
    return XDIV((double)shifted, magnitude);
}

// ---- Double.toString -----------------------------------------------------------------------------------------------------
SGP_HD uint64_t umulh64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
// multiplyHigh on non-negative longs == unsigned high multiply
SGP_HD uint64_t d2s_rop(uint64_t g1, uint64_t g0, uint64_t cp) {
    uint64_t x1 = umulh64(g0, cp);
    uint64_t y0 = g1 * cp;
    uint64_t y1 = umulh64(g1, cp);
    uint64_t z = (y0 >> 1) + x1;
    uint64_t w = y1 + (z >> 63);
    const uint64_t M63 = 0x7fffffffffffffffull;
    return w | (((z & M63) + M63) >> 63);
}
SGP_HD int d2s_flog10pow2(int e) { return (int)(((int64_t)e * 661971961083LL) >> 41); }
SGP_HD int d2s_flog10_34pow2(int e) { return (int)(((int64_t)e * 661971961083LL + (-274743187321LL)) >> 41); }
SGP_HD int d2s_flog2pow10(int e) { return (int)(((int64_t)e * 913124641741LL) >> 38); }

// x (!= 0) without its trailing decimal zeros; *k = how many there were (<= 9). Four fixed steps, no loop.
SGP_HD uint32_t strip_zeros_u32(uint32_t x, int* k_out) {
    int k = 0; uint32_t q;
    q = x / 100000000u; if (q * 100000000u == x) { x = q; k += 8; }
    q = x / 10000u;     if (q * 10000u == x)     { x = q; k += 4; }
    q = x / 100u;       if (q * 100u == x)       { x = q; k += 2; }
    q = x / 10u;        if (q * 10u == x)        { x = q; k += 1; }
    *k_out = k;
    return x;
}
// 10^k for 0 <= k <= 9 as a product of selected factors (no table, no loop)
SGP_HD uint32_t pow10_u32(int k) {
    uint32_t p = (k & 1) ? 10u : 1u;
    if (k & 2) p *= 100u;
    if (k & 4) p *= 10000u;
    if (k & 8) p *= 100000000u;
    return p;
}

// Shortest decimal: v (finite, > 0) == f * 10^e with f free of trailing zeros. Mirrors the published Schubfach
// procedure that Double.toString is specified by since JDK 19 (R. Giulietti, "The Schubfach way to render doubles").
SGP_HD void d2s_decimal(const MathTables& T, double v, uint64_t* f_out, int* e_out) {
    const int Q_MIN = -1074; const uint64_t C_MIN = 1ull << 52, C_TINY = 3;
    uint64_t bits = dbl_bits(v);
    uint64_t t = bits & ((1ull << 52) - 1);
    int bq = (int)((bits >> 52) & 0x7ff);
    uint64_t c; int q; int dk = 0;
    uint64_t f = 0; int e = 0; bool done = false;
    if (bq > 0) {
        int mq = -Q_MIN + 1 - bq;
        c = C_MIN | t;
        if (0 < mq && mq < 53) {
            uint64_t ff = c >> mq;
            if ((ff << mq) == c) { f = ff; e = 0; done = true; }
        }
        q = -mq;
    } else {
        if (t < C_TINY) { c = 10 * t; q = Q_MIN; dk = -1; } else { c = t; q = Q_MIN; }
    }
    if (!done) {
        int out = (int)(c & 1);
        uint64_t cb = c << 2, cbr = cb + 2, cbl; int k;
        if (c != C_MIN || q == Q_MIN) { cbl = cb - 2; k = d2s_flog10pow2(q); }
        else { cbl = cb - 1; k = d2s_flog10_34pow2(q); }
        int h = q + d2s_flog2pow10(-k) + 2;
        uint64_t g1 = T.g[(k - kD2sKMin) * 2], g0 = T.g[(k - kD2sKMin) * 2 + 1];
        uint64_t vb = d2s_rop(g1, g0, cb << h), vbl = d2s_rop(g1, g0, cbl << h), vbr = d2s_rop(g1, g0, cbr << h);
        uint64_t s = vb >> 2;
        bool decided = false;
        if (s >= 100) {
            uint64_t sp10 = 10 * (s / 10);
            uint64_t tp10 = sp10 + 10;
            bool upin = vbl + out <= (sp10 << 2);
            bool wpin = (tp10 << 2) + out <= vbr;
            if (upin != wpin) { f = upin ? sp10 : tp10; e = k; decided = true; }
        }
        if (!decided) {
            uint64_t tt = s + 1;
            bool uin = vbl + out <= (s << 2);
            bool win = (tt << 2) + out <= vbr;
            if (uin != win) { f = uin ? s : tt; }
            else {
                int64_t cmp = (int64_t)(vb - ((s + tt) << 1));
                f = (cmp < 0 || (cmp == 0 && (s & 1) == 0)) ? s : tt;
            }
            e = k + dk;
        }
    }
    // strip trailing zeros: 32-bit limbs (branch lengths rounded to 8 significant digits end in nine zeros, and a 64-bit division
    // per stripped digit made this the most expensive part of the conversion on the device) and a fixed sequence of four
    // divisibility tests (10^8, 10^4, 10^2, 10) instead of a loop per digit: lanes of a warp strip different numbers of zeros
    {
        const uint64_t hi = f / 1000000000ull; uint32_t lo = (uint32_t)(f - hi * 1000000000ull);
        if (hi == 0 || lo == 0) {
            uint64_t rest = 0;
            if (hi != 0) { rest = hi; e += 9; }          // f = hi x 10^9
            else rest = lo;
            if (rest <= 0xffffffffull) {
                int k; f = strip_zeros_u32((uint32_t)rest, &k); e += k;
            } else {
                f = rest;
                while (f >= 10) { const uint64_t qd = f / 10; if (qd * 10 != f) break; f = qd; ++e; }
            }
        } else {
            int k; lo = strip_zeros_u32(lo, &k);
            if (k) { f = hi * (uint64_t)pow10_u32(9 - k) + lo; e += k; }
        }
    }
    *f_out = f; *e_out = e;
}

// Branch-free sums of comparisons: lanes of a warp print numbers of different lengths, and a comparison chain that compiles to
// branches was 18 % of the instructions of the write kernels.
SGP_HD int dec_len_u32(uint32_t v) {
    int n = 1 + (v >= 10u) + (v >= 100u) + (v >= 1000u) + (v >= 10000u);
    if (v >= 100000u) n += (v >= 100000u) + (v >= 1000000u) + (v >= 10000000u) + (v >= 100000000u) + (v >= 1000000000u);
    return n;
}
// The same in a third of the code, for the size pass: bit length -> floor(log10) estimate (x 1233 / 4096) -> one comparison with a
// power of ten from constant memory. The sum of comparisons above compiles to ~50 instructions per call site, and with some forty
// call sites inlined into the counting instantiation of the formatting code it was 46 % of the finish kernel's instructions
// (static count) - the reason that kernel did not fit the instruction cache. The write kernels keep the sum: there the table load
// (lanes index it differently) cost more than the instructions it saved (+3 % on C3, +5 % on the relaxer's emitter).
#if defined(__CUDACC__)
static __constant__ uint32_t kPow10U32[10] = {1u, 10u, 100u, 1000u, 10000u, 100000u, 1000000u, 10000000u, 100000000u, 1000000000u};
#endif
SGP_HD int dec_len_u32_compact(uint32_t v) {
#if defined(__CUDA_ARCH__)
    const uint32_t w = v | 1u;                                   // same number of digits as v, and 1 for v == 0
    const int t = ((32 - __clz((int)w)) * 1233) >> 12;          // floor(log10(w)) or one more
    return t + (w >= kPow10U32[t] ? 1 : 0);
#else
    return dec_len_u32(v);
#endif
}
SGP_HD int dec_len_u64(uint64_t f) {
    if (f <= 0xffffffffull) return dec_len_u32((uint32_t)f);
    const uint64_t hi = f / 1000000000ull;          // < 2^35
    if (hi <= 0xffffffffull) return 9 + dec_len_u32((uint32_t)hi);
    return 19 + (f >= 10000000000000000000ull);
}
SGP_HD int dec_len_u64_compact(uint64_t f) {
    if (f <= 0xffffffffull) return dec_len_u32_compact((uint32_t)f);
    const uint64_t hi = f / 1000000000ull;          // < 2^35
    if (hi <= 0xffffffffull) return 9 + dec_len_u32_compact((uint32_t)hi);
    return 19 + (f >= 10000000000000000000ull);
}
SGP_HD int dec_len_i32(int v) { int n = 0; uint32_t u; if (v < 0) { n = 1; u = (uint32_t)(-(int64_t)v); } else u = (uint32_t)v; return n + dec_len_u32(u); }

// Byte sinks: the same formatting code runs in the "size" pass (CountSink) and the "write" pass (MemSink).
struct CountSink {
    long long n = 0;
    SGP_HD void put(char) { ++n; }
    SGP_HD void skip(int k) { n += k; }
    SGP_HD char* reserve(int) { return nullptr; }
    static constexpr bool kCounts = true;
    static constexpr bool kRandomAccess = false;
};
struct MemSink {
    char* p;
    SGP_HD void put(char c) { *p++ = c; }
    SGP_HD void skip(int) {}
    SGP_HD char* reserve(int n) { char* q = p; p += n; return q; }     // random-access window of n bytes
    static constexpr bool kCounts = false;
    static constexpr bool kRandomAccess = true;
};

#if defined(__CUDACC__)
// Sink on a shared-memory staging buffer: byte stores as st.shared (a plain char* that may also point to global memory compiles
// to generic stores, which is what the formatting code would otherwise use for every character).
struct SmemSink {
    uint32_t a;               // address in the shared window
    struct Ref { uint32_t a; __device__ __forceinline__ void operator=(char c) const { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"((int)c)); } };
    struct Win { uint32_t a; __device__ __forceinline__ Ref operator[](int i) const { return Ref{a + (uint32_t)i}; } };
    __device__ __forceinline__ void put(char c) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"((int)c)); ++a; }
    __device__ __forceinline__ void skip(int) {}
    __device__ __forceinline__ Win reserve(int n) { Win w{a}; a += (uint32_t)n; return w; }
    static constexpr bool kCounts = false;
    static constexpr bool kRandomAccess = true;
};
#endif

template <class Sink> SGP_HD void put_str(Sink& s, const char* z) { while (*z) s.put(*z++); }
// String literal: the length is a compile-time constant (the size pass adds it in one step) and the unrolled stores carry the
// characters as immediates — no per-character load, test and branch as in put_str.
template <class Sink, int N> SGP_HD void put_lit(Sink& s, const char (&z)[N]) {
    if (Sink::kCounts) { s.skip(N - 1); return; }
#pragma unroll
    for (int i = 0; i < N - 1; ++i) s.put(z[i]);
}
// n decimal digits of v (v < 10^n, n <= 10), most significant first, 32-bit arithmetic only
template <class Sink> SGP_HD void put_u32_digits(Sink& s, uint32_t v, int n) {
    if (Sink::kRandomAccess) {
        auto d = s.reserve(n);
        for (int i = n - 1; i >= 0; --i) { uint32_t q = v / 10u; d[i] = (char)('0' + (v - q * 10u)); v = q; }
        return;
    }
    char tmp[10];
    for (int i = n - 1; i >= 0; --i) { uint32_t q = v / 10u; tmp[i] = (char)('0' + (v - q * 10u)); v = q; }
    for (int i = 0; i < n; ++i) s.put(tmp[i]);
}
template <class Sink> SGP_HD void put_u32(Sink& s, uint32_t v) {
    if (Sink::kCounts) { s.skip(dec_len_u32_compact(v)); return; }
    const int n = dec_len_u32(v);
    put_u32_digits(s, v, n);
}
// values beyond 32 bits are rare (wide seeds): out of line, so that the common path stays small
template <class Sink> SGP_HD_NOINL void put_u64_wide(Sink& s, uint64_t v) {
    char tmp[20]; int n = 0;
    while (v) { const uint64_t q = v / 10; tmp[n++] = (char)('0' + (int)(v - q * 10)); v = q; }
    while (n) s.put(tmp[--n]);
}
template <class Sink> SGP_HD void put_u64(Sink& s, uint64_t v) {
    if (v <= 0xffffffffull) put_u32(s, (uint32_t)v); else put_u64_wide(s, v);
}
template <class Sink> SGP_HD void put_i32(Sink& s, int v) {
    if (v < 0) { s.put('-'); put_u32(s, (uint32_t)(-(int64_t)v)); } else put_u32(s, (uint32_t)v);
}
template <class Sink> SGP_HD void put_i64(Sink& s, long long v) {
    if (v >= 0 && v <= 0xffffffffll) { put_u32(s, (uint32_t)v); return; }
    if (v < 0) { s.put('-'); put_u64(s, (uint64_t)(-(v + 1)) + 1ull); } else put_u64(s, (uint64_t)v);
}

// Layout part of java.lang.Double.toString for a finite v > 0 whose shortest decimal is f x 10^e (f < 10^18): plain notation
// for 1e-3 <= v < 1e7, computerized scientific notation otherwise. One pass over the digits: digit i (0 = most significant)
// lands at window position lead + i + (i >= split), everything else ('0', '.', padding zeros, exponent) is placed around them.
// n_hint > 0: the number of digits of f, if the caller knows it (then the size pass does no 64-bit arithmetic at all).
template <class Sink> SGP_HD void put_double_digits(Sink& s, double v, uint64_t f, int e, int n_hint = 0) {
    const bool plain = v >= 1e-3 && v < 1e7;
    if (Sink::kCounts && n_hint > 0) {
        const int n = n_hint, pe = n + e;
        s.skip(plain ? (pe <= 0 ? 2 - pe + n : pe >= n ? pe + 2 : n + 1) : (n == 1 ? 3 : n + 1) + 1 + dec_len_i32(pe - 1));
        return;
    }
    const uint64_t hi64 = f / 1000000000ull;
    uint32_t l0 = (uint32_t)(f - hi64 * 1000000000ull);                  // low 9 digits
    uint32_t l1 = hi64 < 1000000000ull ? (uint32_t)hi64 : (uint32_t)(hi64 % 1000000000ull);
    const int n = n_hint > 0 ? n_hint : (hi64 ? 9 + dec_len_u32(l1) : dec_len_u32(l0));
    const int pe = n + e;          // v = 0.DIGITS x 10^pe
    int lead, split, total, exp10 = 0;
    if (plain) {
        if (pe <= 0) { lead = 2 - pe; split = 99; total = lead + n; }        // 0.000DIGITS
        else if (pe >= n) { lead = 0; split = 99; total = pe + 2; }          // DIGITS000.0
        else { lead = 0; split = pe; total = n + 1; }                        // DIG.ITS
    } else {
        exp10 = pe - 1;
        lead = 0; split = 1; total = (n == 1 ? 3 : n + 1) + 1 + dec_len_i32(exp10);      // D.IGITSE-x
    }
    if (Sink::kCounts) { s.skip(total); return; }
    auto put_all = [&](auto& w) {
        int pos = n - 1;
        for (int i = 0; i < 9 && pos >= 0; ++i, --pos) { const uint32_t q = l0 / 10u; w[lead + pos + (pos >= split ? 1 : 0)] = (char)('0' + (l0 - q * 10u)); l0 = q; }
        for (; pos >= 0; --pos) { const uint32_t q = l1 / 10u; w[lead + pos + (pos >= split ? 1 : 0)] = (char)('0' + (l1 - q * 10u)); l1 = q; }
        if (plain) {
            if (pe <= 0) { w[0] = '0'; w[1] = '.'; if (lead > 2) w[2] = '0'; if (lead > 3) w[3] = '0'; for (int z = 4; z < lead; ++z) w[z] = '0'; }     // (1e-3 <= v: lead <= 4)
            else if (pe >= n) { for (int z = n; z < pe; ++z) w[z] = '0'; w[pe] = '.'; w[pe + 1] = '0'; }
            else w[pe] = '.';
        } else {
            w[1] = '.';
            int p = n + 1;
            if (n == 1) { w[2] = '0'; p = 3; }
            w[p++] = 'E';
            uint32_t u = (uint32_t)exp10;
            if (exp10 < 0) { w[p++] = '-'; u = (uint32_t)(-exp10); }
            const int m = dec_len_u32(u);
            for (int i = m - 1; i >= 0; --i) { const uint32_t q = u / 10u; w[p + i] = (char)('0' + (u - q * 10u)); u = q; }
        }
    };
    if (Sink::kRandomAccess) { auto w = s.reserve(total); put_all(w); }
    else { char tmp[32]; char* w = tmp; put_all(w); for (int i = 0; i < total; ++i) s.put(tmp[i]); }
}

// java.lang.Double.toString
template <class Sink> SGP_HD void put_double(Sink& s, const MathTables& T, double v) {
    if (v != v) { put_str(s, "NaN"); return; }
    if (dbl_hi(v) < 0) { s.put('-'); v = -v; }
    if (v == dbl_inf()) { put_str(s, "Infinity"); return; }
    if (v == 0.0) { put_str(s, "0.0"); return; }
    uint64_t f; int e;
    d2s_decimal(T, v, &f, &e);
    put_double_digits(s, v, f, e);
}

}  // namespace sgp
