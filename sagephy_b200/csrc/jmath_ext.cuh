// More Java-exact math for host and device: fdlibm pow (= StrictMath.pow) and the reference's own quantile functions
// (math/Normal.java:17-86 Voutier inverse normal, math/Gamma.java:44-57,127-204 Lanczos lnGamma + AS 32,
// math/ChiSquared.java:23-84 AS 91). Same non-contractable arithmetic discipline as jmath.cuh.
#pragma once
#include "jmath.cuh"

namespace sgp {

SGP_HD double dbl_with_lo(double x, uint32_t lo) { return dbl_make(dbl_hi(x), lo); }
SGP_HD double dbl_with_hi(double x, int32_t hi) { return dbl_make(hi, dbl_lo(x)); }
SGP_HD double jsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(x);
#else
    return sqrt(x);
#endif
}
SGP_HD double jldexp(double x, int n) { return ldexp(x, n); }

SGP_HD double jpow(double x, double y) {
    const double zero = 0.0, one = 1.0, two = 2.0, two53 = 9007199254740992.0, huge = 1.0e300, tiny = 1.0e-300;
    const double L1 = dbl_make(0x3FE33333, 0x33333303), L2 = dbl_make(0x3FDB6DB6, 0xDB6FABFF), L3 = dbl_make(0x3FD55555, 0x518F264D),
                 L4 = dbl_make(0x3FD17460, 0xA91D4101), L5 = dbl_make(0x3FCD864A, 0x93C9DB65), L6 = dbl_make(0x3FCA7E28, 0x4A454EEF),
                 P1 = dbl_make(0x3FC55555, 0x5555553E), P2 = dbl_make((int32_t)0xBF66C16C, 0x16BEBD93), P3 = dbl_make(0x3F11566A, 0xAF25DE2C),
                 P4 = dbl_make((int32_t)0xBEBBBD41, 0xC5D26BF1), P5 = dbl_make(0x3E663769, 0x72BEA4D0),
                 lg2 = dbl_make(0x3FE62E42, 0xFEFA39EF), lg2_h = dbl_make(0x3FE62E43, 0x00000000), lg2_l = dbl_make((int32_t)0xBE205C61, 0x0CA86C39),
                 ovt = 8.0085662595372944372e-0017,
                 cp = dbl_make(0x3FEEC709, 0xDC3A03FD), cp_h = dbl_make(0x3FEEC709, 0xE0000000), cp_l = dbl_make((int32_t)0xBE3E2FE0, 0x145B01F5),
                 ivln2 = dbl_make(0x3FF71547, 0x652B82FE), ivln2_h = dbl_make(0x3FF71547, 0x60000000), ivln2_l = dbl_make(0x3E54AE0B, 0xF85DDF44);
    double z, ax, z_h, z_l, p_h, p_l, y1, t1, t2, r, s, t, u, v, w;
    int32_t i, j, k, yisint, n;
    int32_t hx = dbl_hi(x), hy = dbl_hi(y); uint32_t lx = dbl_lo(x), ly = dbl_lo(y);
    int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if ((iy | ly) == 0) return one;
    if (ix > 0x7ff00000 || ((ix == 0x7ff00000) && (lx != 0)) || iy > 0x7ff00000 || ((iy == 0x7ff00000) && (ly != 0))) return XADD(x, y);
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
            if (((ix - 0x3ff00000) | lx) == 0) return XSUB(y, y);
            else if (ix >= 0x3ff00000) return (hy >= 0) ? y : zero;
            else return (hy < 0) ? -y : zero;
        }
        if (iy == 0x3ff00000) { if (hy < 0) return XDIV(one, x); else return x; }
        if (hy == 0x40000000) return XMUL(x, x);
        if (hy == 0x3fe00000) { if (hx >= 0) return jsqrt(x); }
    }
    ax = fabs(x);
    if (lx == 0) {
        if (ix == 0x7ff00000 || ix == 0 || ix == 0x3ff00000) {
            z = ax;
            if (hy < 0) z = XDIV(one, z);
            if (hx < 0) {
                if (((ix - 0x3ff00000) | yisint) == 0) z = dbl_nan();
                else if (yisint == 1) z = -z;
            }
            return z;
        }
    }
    if ((((hx >> 31) + 1) | yisint) == 0) return dbl_nan();
    if (iy > 0x41e00000) {
        if (iy > 0x43f00000) {
            if (ix <= 0x3fefffff) return (hy < 0) ? XMUL(huge, huge) : XMUL(tiny, tiny);
            if (ix >= 0x3ff00000) return (hy > 0) ? XMUL(huge, huge) : XMUL(tiny, tiny);
        }
        if (ix < 0x3fefffff) return (hy < 0) ? XMUL(huge, huge) : XMUL(tiny, tiny);
        if (ix > 0x3ff00000) return (hy > 0) ? XMUL(huge, huge) : XMUL(tiny, tiny);
        t = XSUB(ax, one);
        w = XMUL(XMUL(t, t), XSUB(0.5, XMUL(t, XSUB(0.3333333333333333333333, XMUL(t, 0.25)))));
        u = XMUL(ivln2_h, t);
        v = XSUB(XMUL(t, ivln2_l), XMUL(w, ivln2));
        t1 = XADD(u, v);
        t1 = dbl_with_lo(t1, 0);
        t2 = XSUB(v, XSUB(t1, u));
    } else {
        double ss, s2, s_h, s_l, t_h, t_l;
        n = 0;
        if (ix < 0x00100000) { ax = XMUL(ax, two53); n -= 53; ix = dbl_hi(ax); }
        n += ((ix) >> 20) - 0x3ff;
        j = ix & 0x000fffff;
        ix = j | 0x3ff00000;
        if (j <= 0x3988E) k = 0;
        else if (j < 0xBB67A) k = 1;
        else { k = 0; n += 1; ix -= 0x00100000; }
        ax = dbl_with_hi(ax, ix);
        const double bpk = k ? 1.5 : 1.0;
        const double dp_hk = k ? dbl_make(0x3FE2B803, 0x40000000) : 0.0, dp_lk = k ? dbl_make(0x3E4CFDEB, 0x43CFD006) : 0.0;
        u = XSUB(ax, bpk);
        v = XDIV(one, XADD(ax, bpk));
        ss = XMUL(u, v);
        s_h = dbl_with_lo(ss, 0);
        t_h = dbl_make(((ix >> 1) | 0x20000000) + 0x00080000 + (k << 18), 0);
        t_l = XSUB(ax, XSUB(t_h, bpk));
        s_l = XMUL(v, XSUB(XSUB(u, XMUL(s_h, t_h)), XMUL(s_h, t_l)));
        s2 = XMUL(ss, ss);
        r = XMUL(XMUL(s2, s2), XADD(L1, XMUL(s2, XADD(L2, XMUL(s2, XADD(L3, XMUL(s2, XADD(L4, XMUL(s2, XADD(L5, XMUL(s2, L6)))))))))));
        r = XADD(r, XMUL(s_l, XADD(s_h, ss)));
        s2 = XMUL(s_h, s_h);
        t_h = XADD(XADD(3.0, s2), r);
        t_h = dbl_with_lo(t_h, 0);
        t_l = XSUB(r, XSUB(XSUB(t_h, 3.0), s2));
        u = XMUL(s_h, t_h);
        v = XADD(XMUL(s_l, t_h), XMUL(t_l, ss));
        p_h = XADD(u, v);
        p_h = dbl_with_lo(p_h, 0);
        p_l = XSUB(v, XSUB(p_h, u));
        z_h = XMUL(cp_h, p_h);
        z_l = XADD(XADD(XMUL(cp_l, p_h), XMUL(p_l, cp)), dp_lk);
        t = (double)n;
        t1 = XADD(XADD(XADD(z_h, z_l), dp_hk), t);
        t1 = dbl_with_lo(t1, 0);
        t2 = XSUB(z_l, XSUB(XSUB(XSUB(t1, t), dp_hk), z_h));
    }
    s = one;
    if ((((hx >> 31) + 1) | (yisint - 1)) == 0) s = -one;
    y1 = dbl_with_lo(y, 0);
    p_l = XADD(XMUL(XSUB(y, y1), t1), XMUL(y, t2));
    p_h = XMUL(y1, t1);
    z = XADD(p_l, p_h);
    j = dbl_hi(z); i = (int32_t)dbl_lo(z);
    if (j >= 0x40900000) {
        if (((j - 0x40900000) | i) != 0) return XMUL(XMUL(s, huge), huge);
        else { if (XADD(p_l, ovt) > XSUB(z, p_h)) return XMUL(XMUL(s, huge), huge); }
    } else if ((j & 0x7fffffff) >= 0x4090cc00) {
        if (((j - (int32_t)0xc090cc00) | i) != 0) return XMUL(XMUL(s, tiny), tiny);
        else { if (p_l <= XSUB(z, p_h)) return XMUL(XMUL(s, tiny), tiny); }
    }
    i = j & 0x7fffffff;
    k = (i >> 20) - 0x3ff;
    n = 0;
    if (i > 0x3fe00000) {
        n = j + (0x00100000 >> (k + 1));
        k = ((n & 0x7fffffff) >> 20) - 0x3ff;
        t = dbl_make((n & ~(0x000fffff >> k)), 0);
        n = ((n & 0x000fffff) | 0x00100000) >> (20 - k);
        if (j < 0) n = -n;
        p_h = XSUB(p_h, t);
    }
    t = XADD(p_l, p_h);
    t = dbl_with_lo(t, 0);
    u = XMUL(t, lg2_h);
    v = XADD(XMUL(XSUB(p_l, XSUB(t, p_h)), lg2), XMUL(t, lg2_l));
    z = XADD(u, v);
    w = XSUB(v, XSUB(z, u));
    t = XMUL(z, z);
    t1 = XSUB(z, XMUL(t, XADD(P1, XMUL(t, XADD(P2, XMUL(t, XADD(P3, XMUL(t, XADD(P4, XMUL(t, P5))))))))));
    r = XSUB(XDIV(XMUL(z, t1), XSUB(t1, two)), XADD(w, XMUL(z, w)));
    z = XSUB(one, XSUB(r, z));
    j = dbl_hi(z);
    j += (n << 20);
    if ((j >> 20) <= 0) z = jldexp(z, n);
    else z = dbl_with_hi(z, dbl_hi(z) + (n << 20));
    return XMUL(s, z);
}

SGP_HD double jnormal_cdf(double x) {
    if (x < -39) return 0.0;
    if (x > 9) return 1.0;
    const double b1 = 0.319381530, b2 = -0.356563782, b3 = 1.781477937, b4 = -1.821255978, b5 = 1.330274429, p = 0.2316419, c = 0.39894228;
    const double e = jexp(XDIV(XMUL(-x, x), 2.0));
    const double t = XDIV(1.0, x >= 0.0 ? XADD(1.0, XMUL(p, x)) : XSUB(1.0, XMUL(p, x)));
    const double poly = XADD(XMUL(t, XADD(XMUL(t, XADD(XMUL(t, XADD(XMUL(t, b5), b4)), b3)), b2)), b1);
    const double tail = XMUL(XMUL(XMUL(c, e), t), poly);
    return x >= 0.0 ? XSUB(1.0, tail) : tail;
}
SGP_HD double jnormal_quantile(double p) {
    if (0.025 <= p && p <= 0.975) {
        const double a0 = 0.151015505647689, a1 = -0.5303572634357367, a2 = 1.365020122861334, b0 = 0.132089632343748, b1 = -0.7607324991323768;
        const double q = XSUB(p, 0.5), r = XMUL(q, q);
        return XMUL(q, XADD(a2, XDIV(XADD(XMUL(a1, r), a0), XADD(XADD(XMUL(r, r), XMUL(b1, r)), b0))));
    }
    if (1e-50 < p && p < 1.0 - 1e-16) {
        const double c3 = -1.000182518730158122, cp0 = 16.682320830719986527, cp1 = 4.120411523939115059, cp2 = 0.029814187308200211,
                     d0 = 7.173787663925508066, d1 = 8.759693508958633869;
        const bool lowtail = p < 0.5;
        const double pp = lowtail ? p : XSUB(1.0, p);
        const double r = jsqrt(jlog(XDIV(1.0, XMUL(pp, pp))));
        const double val = XADD(XADD(XMUL(c3, r), cp2), XDIV(XADD(XMUL(cp1, r), cp0), XADD(XADD(XMUL(r, r), XMUL(d1, r)), d0)));
        return lowtail ? val : -val;
    }
    return p < 0.5 ? -dbl_inf() : dbl_inf();
}
SGP_HD double jln_gamma(double xx) {
    const double LANCZOS[14] = {57.1562356658629235, -59.5979603554754912, 14.1360979747417471, -0.491913816097520199, .339946499848118887e-4,
        .465236289270485756e-4, -.983744753048795646e-4, .158088703224912494e-3, -.210264441724104883e-3, .217439618115212643e-3,
        -.164318106536763890e-3, .844182239838527433e-4, -.261908384015814087e-4, .368991826595316234e-5};
    const double SQRT_2_PI = 2.5066282746310002;      // Math.sqrt(2 * Math.PI)
    double x = xx, y = xx;
    double tmp = XADD(x, 5.24218750000000000);
    tmp = XSUB(XMUL(XADD(x, 0.5), jlog(tmp)), tmp);
    double ser = 0.999999999999997092;
    for (int j = 0; j < 14; ++j) { y = XADD(y, 1.0); ser = XADD(ser, XDIV(LANCZOS[j], y)); }
    return XADD(tmp, jlog(XDIV(XMUL(SQRT_2_PI, ser), x)));
}
// g = jln_gamma(k): the reference recomputes it on every call (math/Gamma.java:135); callers that evaluate many quantiles of one
// distribution pass the value they computed once — same function, same bits.
SGP_HD double jinc_gamma_ratio(double x, double k, double g) {
    if (x == 0.0) return 0.0;
    const double accuracy = 1.0e-8, overflow = 1.0e30, xbig = 1.0E6, alimit = 1000.0;
    double g_in = 0.0;
    const double factor = jexp(XSUB(XSUB(XMUL(k, jlog(x)), x), g));
    if (k > alimit) {
        const double pn1 = XMUL(XMUL(3.0, jsqrt(k)), XSUB(XADD(jpow(XDIV(x, k), 1.0 / 3.0), XDIV(1.0, XMUL(9.0, k))), 1.0));
        return jnormal_cdf(pn1);
    }
    if (x > xbig) return 1.0;
    if (x > 1.0 && x >= k) {
        double a = XSUB(1.0, k), b = XADD(XADD(a, x), 1.0), term = 0.0;
        double pn0 = 1.0, pn1 = x, pn2 = XADD(x, 1.0), pn3 = XMUL(x, b), pn4, pn5;
        g_in = XDIV(pn2, pn3);
        double dif, rn;
        for (;;) {
            a = XADD(a, 1.0); b = XADD(b, 2.0); term = XADD(term, 1.0);
            const double an = XMUL(a, term);
            pn4 = XSUB(XMUL(b, pn2), XMUL(an, pn0)); pn5 = XSUB(XMUL(b, pn3), XMUL(an, pn1));
            if (pn5 != 0.0) {
                rn = XDIV(pn4, pn5); dif = fabs(XSUB(g_in, rn));
                if (dif <= accuracy) { if (dif <= XMUL(accuracy, rn)) return XSUB(1.0, XMUL(factor, g_in)); }
                g_in = rn;
            }
            pn0 = pn2; pn1 = pn3; pn2 = pn4; pn3 = pn5;
            if (fabs(pn4) >= overflow) { pn0 = XDIV(pn0, overflow); pn1 = XDIV(pn1, overflow); pn2 = XDIV(pn2, overflow); pn3 = XDIV(pn3, overflow); }
        }
    }
    g_in = 1.0; double term = 1.0, rn = k;
    do { rn = XADD(rn, 1.0); term = XMUL(term, XDIV(x, rn)); g_in = XADD(g_in, term); } while (term > accuracy);
    return XDIV(XMUL(g_in, factor), k);
}
// ChiSquared.quantile; *err = 1 when the reference would throw (p outside [2e-6, 0.999998])
// G = jln_gamma(0.5 * k) and AA = jlog(2.0) (= 0x3FE62E42FEFA39EF), both constants of the distribution
SGP_HD double jchi2_quantile_pre(double p, double k, double G, double AA, int* err) {
    if (p < 2e-6 || p > 0.999998 || k <= 0.0) { *err = 1; return dbl_nan(); }
    const double E = 0.5e-6;
    const double XX = XMUL(0.5, k), C = XSUB(XX, 1.0);
    double ch;
    if (k < XMUL(-1.24, jlog(p))) {
        ch = jpow(XMUL(XMUL(p, XX), jexp(XADD(G, XMUL(XX, AA)))), XDIV(1.0, XX));
        if (ch < E) return ch;
    } else if (k <= 0.32) {
        ch = 0.4; const double a = jlog(XSUB(1.0, p)); double Q;
        do {
            Q = ch;
            const double P1 = XADD(1.0, XMUL(ch, XADD(4.67, ch))), P2 = XMUL(ch, XADD(6.73, XMUL(ch, XADD(6.66, ch))));
            const double T = XSUB(XADD(-0.5, XDIV(XADD(4.67, XMUL(2.0, ch)), P1)), XDIV(XADD(6.73, XMUL(ch, XADD(13.32, XMUL(3.0, ch)))), P2));
            ch = XSUB(ch, XDIV(XSUB(1.0, XDIV(XMUL(jexp(XADD(XADD(XADD(a, G), XMUL(0.5, ch)), XMUL(C, AA))), P2), P1)), T));
        } while (fabs(XSUB(XDIV(Q, ch), 1.0)) > 0.01);
    } else {
        const double X = jnormal_quantile(p), P1 = XDIV(0.222222, k);
        ch = XMUL(k, jpow(XSUB(XADD(XMUL(X, jsqrt(P1)), 1.0), P1), 3.0));
        if (ch > XADD(XMUL(2.2, k), 6.0)) ch = XMUL(-2.0, XADD(XSUB(jlog(XSUB(1.0, p)), XMUL(C, jlog(XMUL(0.5, ch)))), G));
    }
    double Q;
    do {
        Q = ch;
        const double P1 = XMUL(0.5, ch), P2 = XSUB(p, jinc_gamma_ratio(P1, XX, G));
        const double T = XMUL(P2, jexp(XSUB(XADD(XADD(XMUL(XX, AA), G), P1), XMUL(C, jlog(ch)))));
        const double B = XDIV(T, ch), A = XSUB(XMUL(0.5, T), XMUL(B, C));
        const double S1 = XDIV(XADD(210.0, XMUL(A, XADD(140.0, XMUL(A, XADD(105.0, XMUL(A, XADD(84.0, XMUL(A, XADD(70.0, XMUL(60.0, A)))))))))), 420.0);
        const double S2 = XDIV(XADD(420.0, XMUL(A, XADD(735.0, XMUL(A, XADD(966.0, XMUL(A, XADD(1141.0, XMUL(1278.0, A)))))))), 2520.0);
        const double S3 = XDIV(XADD(210.0, XMUL(A, XADD(462.0, XMUL(A, XADD(707.0, XMUL(932.0, A)))))), 2520.0);
        const double S4 = XDIV(XADD(XADD(252.0, XMUL(A, XADD(672.0, XMUL(1182.0, A)))), XMUL(C, XADD(294.0, XMUL(A, XADD(889.0, XMUL(1740.0, A)))))), 5040.0);
        const double S5 = XDIV(XADD(XADD(84.0, XMUL(264.0, A)), XMUL(C, XADD(175.0, XMUL(606.0, A)))), 2520.0);
        const double S6 = XDIV(XADD(120.0, XMUL(C, XADD(346.0, XMUL(127.0, C)))), 5040.0);
        const double inner = XSUB(S1, XMUL(B, XSUB(S2, XMUL(B, XSUB(S3, XMUL(B, XSUB(S4, XMUL(B, XSUB(S5, XMUL(B, S6))))))))));
        ch = XADD(ch, XMUL(T, XSUB(XADD(1.0, XMUL(XMUL(0.5, T), S1)), XMUL(XMUL(B, C), inner))));
    } while (fabs(XSUB(XDIV(Q, ch), 1.0)) > E);
    return ch;
}
SGP_HD double jchi2_quantile(double p, double k, int* err) {
    if (p < 2e-6 || p > 0.999998 || k <= 0.0) { *err = 1; return dbl_nan(); }
    return jchi2_quantile_pre(p, k, jln_gamma(XMUL(0.5, k)), jlog(2.0), err);
}

}  // namespace sgp
