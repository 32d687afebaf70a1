// Random streams for the samplers.
//
//  * MtStream    — seed-replay mode. MT19937 exactly as the reference seeds and tempers it
//                  (math/PRNG.java:29-37, math/MersenneTwisterRNG2.java:100-202) with java.util.Random's
//                  nextDouble / nextInt(n) on top.  One replicate owns one 624-word state that lives in global memory,
//                  interleaved across the 32 lanes of a warp (word i of lane l at base[i*32 + l]) so that lanes walking
//                  their states in step touch one 128-byte line per access.
//  * PhiloxStream — throughput mode. Philox4x32-10 keyed by (seed), counter = (draw block, attempt, replicate):
//                  a pure function of its coordinates, so any attempt of any replicate can be regenerated anywhere
//                  (count pass -> build pass, any GPU of the shard).
#pragma once
#include <cstdint>
#include "jmath.cuh"

namespace sgp {

// ---- Philox4x32-10 -------------------------------------------------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };

SGP_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
#if defined(__CUDA_ARCH__)
        // one IMAD.WIDE per product (plain C++ 64-bit products make nvcc append a redundant high-word add)
        asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(M0));
        asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(M1));
#else
        { const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
          hi0 = (uint32_t)(p0 >> 32); lo0 = (uint32_t)p0; hi1 = (uint32_t)(p1 >> 32); lo1 = (uint32_t)p1; }
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return U4{c0, c1, c2, c3};
}

// Same function with the ten round keys precomputed (rk[2r], rk[2r+1] = key + r * Weyl): callers that keep `rk` in the kernel
// parameter block get the keys as constant-bank operands instead of recomputing the schedule for every block.
struct PhiloxKeys { uint32_t rk[20]; };
SGP_HD PhiloxKeys philox_keys(uint64_t seed) {
    PhiloxKeys K; uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { K.rk[2 * r] = k0; K.rk[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return K;
}
SGP_HD U4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& K) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
#if defined(__CUDA_ARCH__)
        asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(M0));
        asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(M1));
#else
        { const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
          hi0 = (uint32_t)(p0 >> 32); lo0 = (uint32_t)p0; hi1 = (uint32_t)(p1 >> 32); lo1 = (uint32_t)p1; }
#endif
        uint32_t n0 = hi1 ^ c1 ^ K.rk[2 * r], n1 = lo1, n2 = hi0 ^ c3 ^ K.rk[2 * r + 1], n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
    return U4{c0, c1, c2, c3};
}

// Two 32-bit words -> a double in [0,1) laid out like java.util.Random.nextDouble: (26 high bits, 27 high bits) * 2^-53.
SGP_HD double u01_from_words(uint32_t a, uint32_t b) {
    uint64_t m = ((uint64_t)(a >> 6) << 27) + (uint64_t)(b >> 5);
    return (double)(int64_t)m * 1.1102230246251565e-16;   // 2^-53 (exact scaling)
}

struct PhiloxStream {
    uint32_t k0, k1;          // key = 64-bit seed
    uint32_t rep_lo, rep_hi;  // replicate id
    uint32_t attempt;
    uint32_t block;           // next counter block
    U4 buf; int have;         // unread words in buf (consumed x,y,z,w)
    SGP_HD void init(uint64_t seed, uint64_t rep, uint32_t attempt_) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); rep_lo = (uint32_t)rep; rep_hi = (uint32_t)(rep >> 32);
        attempt = attempt_; block = 0; have = 0;
    }
    // out of line: the general engine draws at dozens of sites, and a Philox block inlined at each of them (~60 instructions) made
    // its kernels miss the instruction cache
    SGP_HD_NOINL void refill() { buf = philox4x32_10(block++, attempt, rep_lo, rep_hi, k0, k1); have = 4; }
    SGP_HD uint32_t next_word() {
        if (have == 0) refill();
        uint32_t r = have == 4 ? buf.x : have == 3 ? buf.y : have == 2 ? buf.z : buf.w;
        --have;
        return r;
    }
    SGP_HD int32_t next(int bits) { return (int32_t)(next_word() >> (32 - bits)); }
    SGP_HD double next_double() { uint32_t a = next_word(); uint32_t b = next_word(); return u01_from_words(a, b); }
    SGP_HD int32_t next_int(int32_t n) {
        if ((n & -n) == n) return (int32_t)(((int64_t)n * (int64_t)next(31)) >> 31);
        int32_t bits, val;
        do { bits = next(31); val = bits % n; } while ((int32_t)((uint32_t)bits - (uint32_t)val + (uint32_t)(n - 1)) < 0);
        return val;
    }
    SGP_HD void begin_attempt(uint32_t a) { attempt = a; block = 0; have = 0; }
    // drops the unread words of the current block: the next draw starts a fresh block. The general engine does this before every
    // vertex, so that all lanes of a warp refill at the same point of the step instead of wherever their earlier draws left them.
    SGP_HD void align() { have = 0; }
    SGP_HD uint64_t position() const { return 0; }
};

// BigInteger(seed).toByteArray() right-aligned in 16 bytes -> 4 big-endian words, for seeds in the int64 range.
SGP_HD void seed_words_i64(long long s, uint32_t key[4]) {
    unsigned long long u = (unsigned long long)s;
    if (s < 0) {
        int L = 1;      // minimal two's-complement byte length
        while (L < 8 && s < -(1LL << (8 * L - 1))) ++L;
        if (L < 8) u &= ((1ULL << (8 * L)) - 1ULL);
    }
    key[0] = 0; key[1] = 0; key[2] = (uint32_t)(u >> 32); key[3] = (uint32_t)u;
}

// The seed of a run as the reference sees it: `new BigInteger(String)` of ANY size, of which PRNG.get16bytes keeps the low 16
// bytes of the minimal two's-complement form, zero-padded on the left when shorter (math/PRNG.java:29-37,58-60;
// GuestTreeMachina.java:47). Replicate i runs with seed + i in big-integer arithmetic, so the device carries
//   * the low 128 bits of the two's-complement seed (enough for the MT key of every seed + i), and
//   * `wide` = 1 when |seed| >= 2^120 + 2^64: seed + i then needs >= 16 bytes for every 0 <= i < 2^63, the key is simply
//     (seed + i) mod 2^128 and the decimal text of seed + i is <leading digits, with or without a carry/borrow><19 digits>;
//     otherwise seed + i fits a signed __int128 with room to spare and everything is done in that type.
struct SeedSpec {
    unsigned long long lo, hi;        // low 128 bits of the two's-complement seed
    int32_t wide, negative;
    unsigned long long dec_lo19;      // wide: the 19 least significant decimal digits of |seed| ...
    int32_t hs_off[2], hs_len[2];     // ... and the digits above them: blob offset/length without ([0]) and with ([1]) a carry (borrow if negative)
};
SGP_HD void seed_words(const SeedSpec& s, unsigned long long i, uint32_t key[4]) {
    unsigned __int128 v = (((unsigned __int128)s.hi << 64) | (unsigned __int128)s.lo) + (unsigned __int128)i;
    if (!s.wide) {
        const __int128 sv = (__int128)v;
        int L = 1;      // minimal two's-complement byte length of seed + i
        while (L < 16 && (sv < -((__int128)1 << (8 * L - 1)) || sv >= ((__int128)1 << (8 * L - 1)))) ++L;
        if (L < 16) v &= (((unsigned __int128)1 << (8 * L)) - 1);
    }
    key[0] = (uint32_t)(v >> 96); key[1] = (uint32_t)(v >> 64); key[2] = (uint32_t)(v >> 32); key[3] = (uint32_t)v;
}
// decimal text of seed + i (what a separate reference run of replicate i would print in its "Arguments:" line)
template <class Sink> SGP_HD void put_seed(Sink& out, const SeedSpec& s, const char* blob, unsigned long long i) {
    const unsigned long long P19 = 10000000000000000000ull;
    if (!s.wide) {
        __int128 sv = (__int128)((((unsigned __int128)s.hi << 64) | (unsigned __int128)s.lo) + (unsigned __int128)i);
        if (sv >= -(__int128)0x7fffffffffffffffLL && sv <= (__int128)0x7fffffffffffffffLL) { put_i64(out, (long long)sv); return; }
        if (sv < 0) { out.put('-'); sv = -sv; }
        const unsigned __int128 m = (unsigned __int128)sv;
        if (m < (unsigned __int128)P19) { put_u64(out, (unsigned long long)m); return; }
        const unsigned __int128 top = m / P19; const unsigned long long low = (unsigned long long)(m - top * P19);
        put_u64(out, (unsigned long long)top);       // < 2^121 / 10^19 < 2^64
        out.put((char)('0' + (int)(low / 1000000000000000000ull))); const unsigned long long r = low % 1000000000000000000ull;
        put_u32_digits(out, (uint32_t)(r / 1000000000ull), 9); put_u32_digits(out, (uint32_t)(r % 1000000000ull), 9);
        return;
    }
    int c; unsigned long long low;
    if (!s.negative) { c = i >= P19 - s.dec_lo19 ? 1 : 0; low = c ? i - (P19 - s.dec_lo19) : s.dec_lo19 + i; }
    else { c = i > s.dec_lo19 ? 1 : 0; low = c ? s.dec_lo19 + (P19 - i) : s.dec_lo19 - i; out.put('-'); }
    for (int k = 0; k < s.hs_len[c]; ++k) out.put(blob[s.hs_off[c] + k]);
    out.put((char)('0' + (int)(low / 1000000000000000000ull))); const unsigned long long r = low % 1000000000000000000ull;
    put_u32_digits(out, (uint32_t)(r / 1000000000ull), 9); put_u32_digits(out, (uint32_t)(r % 1000000000ull), 9);
}
SGP_HD int seed_dec_len(const SeedSpec& s, const char* blob, unsigned long long i) { CountSink c; put_seed(c, s, blob, i); return (int)c.n; }

// ---- MT19937 (Java / uncommons flavour) ------------------------------------------------------------------------------------
// seed16: the 16 big-endian bytes PRNG.get16bytes produced, as 4 big-endian words.
struct MtStream {
    uint32_t* mt;      // strided state: element i at mt[i * stride]
    int stride;
    int idx;
    uint64_t drawn;    // words drawn so far (lets the build pass fast-forward to the accepted attempt)
    SGP_HD uint32_t& at(int i) { return mt[(size_t)i * stride]; }
    SGP_HD void align() {}      // (replay: the stream is the reference's, word for word)

    SGP_HD_NOINL void seed(const uint32_t key[4]) {
        at(0) = 19650218u;
        for (int i = 1; i < 624; ++i) { uint32_t p = at(i - 1); at(i) = 1812433253u * (p ^ (p >> 30)) + (uint32_t)i; }
        int i = 1, j = 0;
        for (int k = 624; k > 0; --k) {
            uint32_t p = at(i - 1);
            at(i) = (at(i) ^ ((p ^ (p >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            ++i; ++j;
            if (i >= 624) { at(0) = at(623); i = 1; }
            if (j >= 4) j = 0;
        }
        for (int k = 623; k > 0; --k) {
            uint32_t p = at(i - 1);
            at(i) = (at(i) ^ ((p ^ (p >> 30)) * 1566083941u)) - (uint32_t)i;
            ++i;
            if (i >= 624) { at(0) = at(623); i = 1; }
        }
        at(0) = 0x80000000u;
        idx = 624; drawn = 0;
    }
    SGP_HD_NOINL void twist() {        // out of line for the same reason as PhiloxStream::refill
        int kk; uint32_t y;
        for (kk = 0; kk < 624 - 397; ++kk) { y = (at(kk) & 0x80000000u) | (at(kk + 1) & 0x7fffffffu); at(kk) = at(kk + 397) ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfu : 0u); }
        for (; kk < 623; ++kk) { y = (at(kk) & 0x80000000u) | (at(kk + 1) & 0x7fffffffu); at(kk) = at(kk + (397 - 624)) ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfu : 0u); }
        y = (at(623) & 0x80000000u) | (at(0) & 0x7fffffffu); at(623) = at(396) ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfu : 0u);
        idx = 0;
    }
    SGP_HD uint32_t next_word() {
        if (idx >= 624) twist();
        uint32_t y = at(idx++);
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        ++drawn;
        return y;
    }
    // discard n words (used to fast-forward to the start of the accepted attempt)
    SGP_HD void skip(uint64_t n) {
        while (n > 0) {
            if (idx >= 624) twist();
            uint64_t take = (uint64_t)(624 - idx); if (take > n) take = n;
            idx += (int)take; drawn += take; n -= take;
        }
    }
    SGP_HD int32_t next(int bits) { return (int32_t)(next_word() >> (32 - bits)); }
    SGP_HD double next_double() { int64_t a = next(26), b = next(27); return (double)((a << 27) + b) * 1.1102230246251565e-16; }
    SGP_HD int32_t next_int(int32_t n) {
        if ((n & -n) == n) return (int32_t)(((int64_t)n * (int64_t)next(31)) >> 31);
        int32_t bits, val;
        do { bits = next(31); val = bits % n; } while ((int32_t)((uint32_t)bits - (uint32_t)val + (uint32_t)(n - 1)) < 0);
        return val;
    }
    SGP_HD void begin_attempt(uint32_t) {}     // one stream is shared by all attempts of a run (GuestTreeMachina.java:47,94)
    SGP_HD uint64_t position() const { return drawn; }
};

}  // namespace sgp
