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
#include <vector>
#include <algorithm>

namespace oracle {

// BigInteger(seedString).toByteArray() -> minimal two's-complement big-endian bytes, then PRNG.get16bytes
// (math/PRNG.java:29-37: keep the LAST 16 bytes, or right-align a shorter array in 16 zero bytes). Any number of digits.
inline std::vector<uint8_t> decimal_to_twos_complement(const std::string& dec) {
    bool neg = false; size_t i = 0;
    if (i < dec.size() && (dec[i] == '-' || dec[i] == '+')) { neg = dec[i] == '-'; ++i; }
    if (i >= dec.size()) throw std::invalid_argument("NumberFormatException: Zero length BigInteger");
    std::vector<uint8_t> mag;        // big-endian magnitude
    for (; i < dec.size(); ++i) {
        if (dec[i] < '0' || dec[i] > '9') throw std::invalid_argument("NumberFormatException: For input string: \"" + dec + "\"");
        int carry = dec[i] - '0';
        for (size_t k = mag.size(); k-- > 0;) { int t = mag[k] * 10 + carry; mag[k] = (uint8_t)(t & 0xff); carry = t >> 8; }
        while (carry) { mag.insert(mag.begin(), (uint8_t)(carry & 0xff)); carry >>= 8; }
    }
    while (!mag.empty() && mag.front() == 0) mag.erase(mag.begin());
    if (mag.empty()) return std::vector<uint8_t>{0};
    std::vector<uint8_t> b(mag.size() + 1, 0);            // one spare sign byte
    std::copy(mag.begin(), mag.end(), b.begin() + 1);
    if (neg) {                                            // negate: invert and add one
        for (auto& x : b) x = (uint8_t)~x;
        for (size_t k = b.size(); k-- > 0;) { if (++b[k] != 0) break; }
    }
    // minimal length: drop a leading byte that only repeats the sign of the next one
    while (b.size() > 1 && ((b[0] == 0x00 && (b[1] & 0x80) == 0) || (b[0] == 0xFF && (b[1] & 0x80) != 0))) b.erase(b.begin());
    return b;
}
inline void seed_to_16_bytes(const std::string& dec, uint8_t out[16]) {
    const std::vector<uint8_t> b = decimal_to_twos_complement(dec);
    std::memset(out, 0, 16);
    const int len = (int)b.size();
    const int from = len > 16 ? len - 16 : 0, to = len < 16 ? 16 - len : 0, n = len < 16 ? len : 16;
    std::memcpy(out + to, b.data() + from, (size_t)n);
}
// decimal string + small non-negative offset, in decimal-string arithmetic (replicate i of a batch == run with -s seed+i)
inline std::string decimal_add(const std::string& dec, unsigned long long add) {
    bool neg = false; size_t i = 0;
    if (i < dec.size() && (dec[i] == '-' || dec[i] == '+')) { neg = dec[i] == '-'; ++i; }
    std::string d = dec.substr(i); size_t nz = d.find_first_not_of('0'); d = nz == std::string::npos ? "0" : d.substr(nz);
    std::string a = std::to_string(add);
    auto cmp = [](const std::string& x, const std::string& y) { return x.size() != y.size() ? (x.size() < y.size() ? -1 : 1) : x.compare(y); };
    auto plus = [](const std::string& x, const std::string& y) {
        std::string r; int c = 0;
        for (int p = (int)x.size() - 1, q = (int)y.size() - 1; p >= 0 || q >= 0 || c; --p, --q) { int t = c + (p >= 0 ? x[p] - '0' : 0) + (q >= 0 ? y[q] - '0' : 0); r.insert(r.begin(), (char)('0' + t % 10)); c = t / 10; }
        return r;
    };
    auto minus = [](const std::string& x, const std::string& y) {      // x >= y
        std::string r; int bw = 0;
        for (int p = (int)x.size() - 1, q = (int)y.size() - 1; p >= 0; --p, --q) { int t = (x[p] - '0') - bw - (q >= 0 ? y[q] - '0' : 0); bw = t < 0; if (t < 0) t += 10; r.insert(r.begin(), (char)('0' + t)); }
        size_t z = r.find_first_not_of('0'); return z == std::string::npos ? std::string("0") : r.substr(z);
    };
    if (!neg || d == "0") return plus(d, a);
    if (cmp(d, a) > 0) return "-" + minus(d, a);
    return minus(a, d);
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
