// ORACLE — test infrastructure only (see oracle/README.md). PARITY UNPINNED.
// fdlibm 5.3 __ieee754_pow (= java.lang.StrictMath.pow), restated from the published algorithm. Constants are given by
// their IEEE-754 bit patterns; tests check the restatement against libm's pow to within 1 ulp.
#pragma once
#include <cmath>
#include "jmath.hpp"

namespace oracle {

inline double dbl_from_words(uint32_t hi, uint32_t lo) { uint64_t u = ((uint64_t)hi << 32) | lo; double d; std::memcpy(&d, &u, 8); return d; }
inline double with_lo(double x, uint32_t lo) { uint64_t u; std::memcpy(&u, &x, 8); u = (u & 0xffffffff00000000ull) | lo; std::memcpy(&x, &u, 8); return x; }

inline double j_pow(double x, double y) {
    static const double bp[2] = {1.0, 1.5};
    static const double dp_h[2] = {0.0, dbl_from_words(0x3FE2B803, 0x40000000)};
    static const double dp_l[2] = {0.0, dbl_from_words(0x3E4CFDEB, 0x43CFD006)};
    static const double zero = 0.0, one = 1.0, two = 2.0, two53 = 9007199254740992.0, huge = 1.0e300, tiny = 1.0e-300,
        L1 = dbl_from_words(0x3FE33333, 0x33333303), L2 = dbl_from_words(0x3FDB6DB6, 0xDB6FABFF), L3 = dbl_from_words(0x3FD55555, 0x518F264D),
        L4 = dbl_from_words(0x3FD17460, 0xA91D4101), L5 = dbl_from_words(0x3FCD864A, 0x93C9DB65), L6 = dbl_from_words(0x3FCA7E28, 0x4A454EEF),
        P1 = dbl_from_words(0x3FC55555, 0x5555553E), P2 = dbl_from_words(0xBF66C16C, 0x16BEBD93), P3 = dbl_from_words(0x3F11566A, 0xAF25DE2C),
        P4 = dbl_from_words(0xBEBBBD41, 0xC5D26BF1), P5 = dbl_from_words(0x3E663769, 0x72BEA4D0),
        lg2 = dbl_from_words(0x3FE62E42, 0xFEFA39EF), lg2_h = dbl_from_words(0x3FE62E43, 0x00000000), lg2_l = dbl_from_words(0xBE205C61, 0x0CA86C39),
        ovt = 8.0085662595372944372e-0017,
        cp = dbl_from_words(0x3FEEC709, 0xDC3A03FD), cp_h = dbl_from_words(0x3FEEC709, 0xE0000000), cp_l = dbl_from_words(0xBE3E2FE0, 0x145B01F5),
        ivln2 = dbl_from_words(0x3FF71547, 0x652B82FE), ivln2_h = dbl_from_words(0x3FF71547, 0x60000000), ivln2_l = dbl_from_words(0x3E54AE0B, 0xF85DDF44);
    double z, ax, z_h, z_l, p_h, p_l, y1, t1, t2, r, s, t, u, v, w;
    int32_t i, j, k, yisint, n, hx, hy, ix, iy; uint32_t lx, ly;
    hx = hi_word(x); lx = lo_word(x); hy = hi_word(y); ly = lo_word(y);
    ix = hx & 0x7fffffff; iy = hy & 0x7fffffff;
    if ((iy | ly) == 0) return one;
    if (ix > 0x7ff00000 || ((ix == 0x7ff00000) && (lx != 0)) || iy > 0x7ff00000 || ((iy == 0x7ff00000) && (ly != 0))) return x + y;
    yisint = 0;
    if (hx < 0) {
        if (iy >= 0x43400000) yisint = 2;
        else if (iy >= 0x3ff00000) {
            k = (iy >> 20) - 0x3ff;
            if (k > 20) { j = (int32_t)(ly >> (52 - k)); if (((uint32_t)j << (52 - k)) == ly) yisint = 2 - (j & 1); }
            else if (ly == 0) { j = iy >> (20 - k); if ((j << (20 - k)) == iy) yisint = 2 - (j & 1); }
        }
    }
    if (ly == 0) {
        if (iy == 0x7ff00000) {
            if (((ix - 0x3ff00000) | lx) == 0) return y - y;
            else if (ix >= 0x3ff00000) return (hy >= 0) ? y : zero;
            else return (hy < 0) ? -y : zero;
        }
        if (iy == 0x3ff00000) { if (hy < 0) return one / x; else return x; }
        if (hy == 0x40000000) return x * x;
        if (hy == 0x3fe00000) { if (hx >= 0) return std::sqrt(x); }
    }
    ax = std::fabs(x);
    if (lx == 0) {
        if (ix == 0x7ff00000 || ix == 0 || ix == 0x3ff00000) {
            z = ax;
            if (hy < 0) z = one / z;
            if (hx < 0) {
                if (((ix - 0x3ff00000) | yisint) == 0) z = (z - z) / (z - z);
                else if (yisint == 1) z = -z;
            }
            return z;
        }
    }
    if ((((hx >> 31) + 1) | yisint) == 0) return (x - x) / (x - x);
    if (iy > 0x41e00000) {
        if (iy > 0x43f00000) {
            if (ix <= 0x3fefffff) return (hy < 0) ? huge * huge : tiny * tiny;
            if (ix >= 0x3ff00000) return (hy > 0) ? huge * huge : tiny * tiny;
        }
        if (ix < 0x3fefffff) return (hy < 0) ? huge * huge : tiny * tiny;
        if (ix > 0x3ff00000) return (hy > 0) ? huge * huge : tiny * tiny;
        t = ax - one;
        w = (t * t) * (0.5 - t * (0.3333333333333333333333 - t * 0.25));
        u = ivln2_h * t;
        v = t * ivln2_l - w * ivln2;
        t1 = u + v;
        t1 = with_lo(t1, 0);
        t2 = v - (t1 - u);
    } else {
        double ss, s2, s_h, s_l, t_h, t_l;
        n = 0;
        if (ix < 0x00100000) { ax *= two53; n -= 53; ix = hi_word(ax); }
        n += ((ix) >> 20) - 0x3ff;
        j = ix & 0x000fffff;
        ix = j | 0x3ff00000;
        if (j <= 0x3988E) k = 0;
        else if (j < 0xBB67A) k = 1;
        else { k = 0; n += 1; ix -= 0x00100000; }
        ax = with_hi(ax, ix);
        u = ax - bp[k];
        v = one / (ax + bp[k]);
        ss = u * v;
        s_h = ss;
        s_h = with_lo(s_h, 0);
        t_h = zero;
        t_h = with_hi(t_h, ((ix >> 1) | 0x20000000) + 0x00080000 + (k << 18));
        t_l = ax - (t_h - bp[k]);
        s_l = v * ((u - s_h * t_h) - s_h * t_l);
        s2 = ss * ss;
        r = s2 * s2 * (L1 + s2 * (L2 + s2 * (L3 + s2 * (L4 + s2 * (L5 + s2 * L6)))));
        r += s_l * (s_h + ss);
        s2 = s_h * s_h;
        t_h = 3.0 + s2 + r;
        t_h = with_lo(t_h, 0);
        t_l = r - ((t_h - 3.0) - s2);
        u = s_h * t_h;
        v = s_l * t_h + t_l * ss;
        p_h = u + v;
        p_h = with_lo(p_h, 0);
        p_l = v - (p_h - u);
        z_h = cp_h * p_h;
        z_l = cp_l * p_h + p_l * cp + dp_l[k];
        t = (double)n;
        t1 = (((z_h + z_l) + dp_h[k]) + t);
        t1 = with_lo(t1, 0);
        t2 = z_l - (((t1 - t) - dp_h[k]) - z_h);
    }
    s = one;
    if ((((hx >> 31) + 1) | (yisint - 1)) == 0) s = -one;
    y1 = y;
    y1 = with_lo(y1, 0);
    p_l = (y - y1) * t1 + y * t2;
    p_h = y1 * t1;
    z = p_l + p_h;
    j = hi_word(z); i = (int32_t)lo_word(z);
    if (j >= 0x40900000) {
        if (((j - 0x40900000) | i) != 0) return s * huge * huge;
        else { if (p_l + ovt > z - p_h) return s * huge * huge; }
    } else if ((j & 0x7fffffff) >= 0x4090cc00) {
        if (((j - (int32_t)0xc090cc00) | i) != 0) return s * tiny * tiny;
        else { if (p_l <= z - p_h) return s * tiny * tiny; }
    }
    i = j & 0x7fffffff;
    k = (i >> 20) - 0x3ff;
    n = 0;
    if (i > 0x3fe00000) {
        n = j + (0x00100000 >> (k + 1));
        k = ((n & 0x7fffffff) >> 20) - 0x3ff;
        t = zero;
        t = with_hi(t, (n & ~(0x000fffff >> k)));
        n = ((n & 0x000fffff) | 0x00100000) >> (20 - k);
        if (j < 0) n = -n;
        p_h -= t;
    }
    t = p_l + p_h;
    t = with_lo(t, 0);
    u = t * lg2_h;
    v = (p_l - (t - p_h)) * lg2 + t * lg2_l;
    z = u + v;
    w = v - (z - u);
    t = z * z;
    t1 = z - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    r = (z * t1) / (t1 - two) - (w + z * w);
    z = one - (r - z);
    j = hi_word(z);
    j += (n << 20);
    if ((j >> 20) <= 0) z = std::ldexp(z, n);
    else z = with_hi(z, hi_word(z) + (n << 20));
    return s * z;
}

}  // namespace oracle
