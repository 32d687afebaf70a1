// ORACLE — test infrastructure only (see oracle/README.md). PARITY UNPINNED (no reference tests / no JVM here).
//
// Restates the reference's PRNG stack:
//   PRNG(BigInteger) -> get16bytes           src/main/java/uc/sgp/sagephy/math/PRNG.java:29-37,58-60
//   MersenneTwisterRNG2(byte[16]) seeding    src/main/java/uc/sgp/sagephy/math/MersenneTwisterRNG2.java:100-148
//   MersenneTwisterRNG2.next(bits)           src/main/java/uc/sgp/sagephy/math/MersenneTwisterRNG2.java:164-202
//   java.util.Random.nextDouble/nextInt(n)/nextGaussian/nextBoolean on top of next(bits)  (JDK specification)
//   uncommons-maths 1.2.2 BinaryUtils.convertBytesToInts = big-endian 4-byte groups (pom.xml:123-129)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <stdexcept>

namespace oracle {

// BigInteger(seedString).toByteArray() -> minimal two's-complement big-endian bytes, then PRNG.get16bytes.
// Seeds are supported over the signed 128-bit range (wider values would be truncated to the low 16 bytes by the
// reference as well, but need arbitrary precision parsing that nothing on the path uses).
inline void seed_to_16_bytes(const std::string& dec, uint8_t out[16]) {
    bool neg = false; size_t i = 0;
    if (i < dec.size() && (dec[i] == '-' || dec[i] == '+')) { neg = dec[i] == '-'; ++i; }
    if (i >= dec.size()) throw std::invalid_argument("NumberFormatException: Zero length BigInteger");
    unsigned __int128 mag = 0;
    for (; i < dec.size(); ++i) {
        if (dec[i] < '0' || dec[i] > '9') throw std::invalid_argument("NumberFormatException: For input string: \"" + dec + "\"");
        mag = mag * 10 + (unsigned)(dec[i] - '0');
    }
    unsigned __int128 tc = neg ? (~mag + 1) : mag;
    uint8_t full[17];
    full[0] = neg && mag != 0 ? 0xFF : 0x00;       // sign extension byte
    for (int b = 0; b < 16; ++b) full[1 + b] = (uint8_t)(tc >> (8 * (15 - b)));
    // minimal length: strip redundant leading sign bytes
    int start = 0;
    while (start < 16) {
        uint8_t cur = full[start], nxt = full[start + 1];
        if ((cur == 0x00 && (nxt & 0x80) == 0) || (cur == 0xFF && (nxt & 0x80) != 0)) ++start; else break;
    }
    int len = 17 - start;
    std::memset(out, 0, 16);
    int from = len > 16 ? len - 16 : 0;
    int to = len < 16 ? 16 - len : 0;
    int n = len < 16 ? len : 16;
    std::memcpy(out + to, full + start + from, n);
}

class JavaMT {
public:
    static constexpr int N = 624, M = 397;
    uint32_t mt[N];
    int idx;
    uint64_t words_drawn = 0;
    bool have_gauss = false; double next_gauss = 0;

    explicit JavaMT(const std::string& seed_decimal) { uint8_t b[16]; seed_to_16_bytes(seed_decimal, b); init(b); }
    explicit JavaMT(const uint8_t b[16]) { init(b); }

    void init(const uint8_t b[16]) {
        uint32_t key[4];
        for (int i = 0; i < 4; ++i)
            key[i] = ((uint32_t)b[4 * i] << 24) | ((uint32_t)b[4 * i + 1] << 16) | ((uint32_t)b[4 * i + 2] << 8) | b[4 * i + 3];
        mt[0] = 19650218u;
        for (idx = 1; idx < N; ++idx) mt[idx] = 1812433253u * (mt[idx - 1] ^ (mt[idx - 1] >> 30)) + (uint32_t)idx;
        int i = 1, j = 0;
        for (int k = N; k > 0; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            ++i; ++j;
            if (i >= N) { mt[0] = mt[N - 1]; i = 1; }
            if (j >= 4) j = 0;
        }
        for (int k = N - 1; k > 0; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            ++i;
            if (i >= N) { mt[0] = mt[N - 1]; i = 1; }
        }
        mt[0] = 0x80000000u;
        // idx == N here: first draw twists.
    }

    uint32_t next_word() {
        if (idx >= N) {
            static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
            int kk; uint32_t y;
            for (kk = 0; kk < N - M; ++kk) { y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + M] ^ (y >> 1) ^ mag01[y & 1]; }
            for (; kk < N - 1; ++kk) { y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ mag01[y & 1]; }
            y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu); mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ mag01[y & 1];
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        ++words_drawn;
        return y;
    }
    inline int32_t next(int bits) { return (int32_t)(next_word() >> (32 - bits)); }

    double nextDouble() {
        int64_t a = next(26), b = next(27);
        return (double)((a << 27) + b) * 0x1.0p-53;
    }
    int32_t nextInt(int32_t n) {
        if (n <= 0) throw std::invalid_argument("IllegalArgumentException: n must be positive");
        if ((n & -n) == n) return (int32_t)(((int64_t)n * (int64_t)next(31)) >> 31);
        int32_t bits, val;
        do {
            bits = next(31);
            val = bits % n;
        } while ((int32_t)((uint32_t)bits - (uint32_t)val + (uint32_t)(n - 1)) < 0);
        return val;
    }
    bool nextBoolean() { return next(1) != 0; }
    double nextGaussian();   // defined in relax.hpp (needs StrictMath.log/sqrt)
};

}  // namespace oracle
