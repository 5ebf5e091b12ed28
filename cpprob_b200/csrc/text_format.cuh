// cpprob-b200: exact `%.15e` formatting of doubles on the GPU — the device half of the posterior-file
// text stage.
//
// The reference prints every real value with `std::scientific`, `precision(15)`
// (/root/reference src/cpprob/state.cpp:262-267), i.e. 16 significant digits, correctly rounded
// (round-half-even on the exact binary value, as glibc's printf does).  format_e15 reproduces those bytes
// with integer arithmetic only:
//
//   v = m * 2^e2 (m < 2^53);  k = floor(log10 v);  D = round(v * 10^(15-k)) in [10^15, 10^16)
//   v * 10^q = m * P[q] * 2^(e2 + Pe[q]),  P[q] the 128-bit mantissa of 10^q truncated (pow10_table.inc)
//
// The 192-bit product is split at the binary point into the integer part I and a fraction whose top bit
// is the rounding bit.  For 0 <= q <= 55 the table entry is exact and so is the decision, ties included.
// Otherwise the true fraction lies in [F, F + 2^53 units of the last limb): "round up" is certain when the
// rounding bit is set, "round down" is certain unless every fraction bit from the rounding bit's neighbour
// down to bit 53 is one — a window of relative width 2^-74.  Such a value is reported as *ambiguous* and
// the caller lets the host re-format that one record with std::to_chars; for continuous data this
// happens with probability ~1e-22 per value, and never changes the length of the text.
//
// The function is __host__ __device__: tests/test_text_format.py checks it on the CPU against
// printf("%.15e") for 10^7 doubles (random bit patterns, integers, powers of ten, exact ties,
// subnormals), and tests/test_files_gpu.py checks the GPU-written files byte for byte.
#ifndef CPPROB_B200_TEXT_FORMAT_CUH
#define CPPROB_B200_TEXT_FORMAT_CUH

#include <cstdint>
#include <cstring>

#include "pow10_table.inc"

#if defined(__CUDACC__)
#define CPPROB_FMT_HD __host__ __device__ __forceinline__
#else
#define CPPROB_FMT_HD inline
#endif

namespace cpprob {
namespace text {

struct pow10_entry {
    std::uint64_t hi, lo;
    int exp;
};

#if defined(__CUDACC__)
static __device__ const pow10_entry d_pow10[] = {CPPROB_POW10_ROWS};
#endif
static const pow10_entry h_pow10[] = {CPPROB_POW10_ROWS};

CPPROB_FMT_HD pow10_entry pow10_of(int q)
{
#if defined(__CUDA_ARCH__)
    return d_pow10[q - CPPROB_POW10_QMIN];
#else
    return h_pow10[q - CPPROB_POW10_QMIN];
#endif
}

CPPROB_FMT_HD void mul64(std::uint64_t a, std::uint64_t b, std::uint64_t & hi, std::uint64_t & lo)
{
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    const unsigned __int128 p = static_cast<unsigned __int128>(a) * b;
    lo = static_cast<std::uint64_t>(p);
    hi = static_cast<std::uint64_t>(p >> 64);
#endif
}

CPPROB_FMT_HD int bit_length64(std::uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return 64 - __clzll(static_cast<long long>(x));
#else
    return x ? 64 - __builtin_clzll(x) : 0;
#endif
}

// 192-bit value r2:r1:r0 (r2 most significant)
struct u192 {
    std::uint64_t r0, r1, r2;
    CPPROB_FMT_HD bool bit(int i) const
    {
        const std::uint64_t w = i < 64 ? r0 : (i < 128 ? r1 : r2);
        return (w >> (i & 63)) & 1u;
    }
    // low 64 bits of (this >> sh), 0 <= sh < 192
    CPPROB_FMT_HD std::uint64_t shr(int sh) const
    {
        const int w = sh >> 6, b = sh & 63;
        const std::uint64_t a = w == 0 ? r0 : (w == 1 ? r1 : r2);
        const std::uint64_t c = w == 0 ? r1 : (w == 1 ? r2 : 0);
        return b == 0 ? a : ((a >> b) | (c << (64 - b)));
    }
    // any of the bits [0, n) set?
    CPPROB_FMT_HD bool any_below(int n) const
    {
        if (n <= 0) return false;
        if (n >= 128) return (r0 | r1) != 0 || (n > 128 && (r2 & ((n >= 192) ? ~0ull : ((1ull << (n - 128)) - 1))) != 0);
        if (n >= 64) return r0 != 0 || (n > 64 && (r1 & ((1ull << (n - 64)) - 1)) != 0);
        return (r0 & ((1ull << n) - 1)) != 0;
    }
    // all of the bits [lo, hi) set?  (empty range: true)
    CPPROB_FMT_HD bool all_ones(int lo, int hi) const
    {
        for (int i = lo; i < hi; ++i) {
            if (!bit(i)) return false;
        }
        return true;
    }
};

CPPROB_FMT_HD char * put_2digits(char * p, unsigned v)   // v < 100
{
    p[0] = static_cast<char>('0' + v / 10);
    p[1] = static_cast<char>('0' + v % 10);
    return p + 2;
}

CPPROB_FMT_HD char * put_8digits(char * p, unsigned v)   // v < 10^8, zero padded
{
    const unsigned a = v / 10000, b = v % 10000;
    p = put_2digits(p, a / 100);
    p = put_2digits(p, a % 100);
    p = put_2digits(p, b / 100);
    return put_2digits(p, b % 100);
}

CPPROB_FMT_HD char * put_int(char * p, long long v)
{
    char tmp[24];
    int n = 0;
    unsigned long long u = v < 0 ? 0ull - static_cast<unsigned long long>(v) : static_cast<unsigned long long>(v);
    do {
        tmp[n++] = static_cast<char>('0' + u % 10);
        u /= 10;
    } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}

// Writes printf("%.15e", v) at p (at most 24 bytes), returns the end.  *ambiguous is set (never cleared)
// when the last digit could not be decided (see the header comment).
CPPROB_FMT_HD char * format_e15(char * p, double v, bool * ambiguous)
{
    std::uint64_t bits;
#if defined(__CUDA_ARCH__)
    bits = static_cast<std::uint64_t>(__double_as_longlong(v));
#else
    std::memcpy(&bits, &v, sizeof bits);
#endif
    const bool neg = bits >> 63;
    const int ex = static_cast<int>((bits >> 52) & 0x7FF);
    const std::uint64_t frac = bits & ((1ull << 52) - 1);
    if (neg) *p++ = '-';
    if (ex == 0x7FF) {
        const char * s = frac ? "nan" : "inf";
        p[0] = s[0]; p[1] = s[1]; p[2] = s[2];
        return p + 3;
    }
    std::uint64_t D;
    int k;
    if (ex == 0 && frac == 0) {
        D = 0;
        k = 0;
    } else {
        const std::uint64_t m = ex ? (frac | (1ull << 52)) : frac;
        const int e2 = ex ? ex - 1075 : -1074;
        const int t = e2 + bit_length64(m) - 1;                                   // floor(log2 v)
        k = static_cast<int>((static_cast<long long>(t) * 1292913986LL) >> 32);   // floor(t log10 2): k or k - 1
        std::uint64_t I = 0;
        bool round_bit = false, sticky = false, exact = false;
        int sh = 0;
        u192 prod{0, 0, 0};
        // k from the binary exponent is the decimal exponent or one less, so the scaled value is >= 10^15 in
        // exact arithmetic; a second round with k + 1 is needed when it is >= 10^16
        for (int attempt = 0; attempt < 2; ++attempt) {
            const int q = 15 - k;
            const pow10_entry P = pow10_of(q);
            exact = q >= 0 && q <= 55;
            std::uint64_t l1, l0, h1, h0;
            mul64(m, P.lo, l1, l0);
            mul64(m, P.hi, h1, h0);
            prod.r0 = l0;
            prod.r1 = l1 + h0;
            prod.r2 = h1 + (prod.r1 < l1 ? 1u : 0u);
            sh = -(e2 + P.exp);                                                   // value = prod * 2^-sh, 64 < sh < 192
            I = prod.shr(sh);
            if (I < 10000000000000000ull) break;
            ++k;
        }
        bool up;
        if (I == 999999999999999ull) {
            // only the truncation of the table entry can push the integer part below 10^15: v is 10^k itself
            I = 1000000000000000ull;
            up = false;
        } else {
            round_bit = prod.bit(sh - 1);
            sticky = prod.any_below(sh - 1);
            if (exact) {
                up = round_bit && (sticky || (I & 1u));                           // ties to even
            } else {
                up = round_bit;                                                   // true fraction > computed >= 1/2
                if (!round_bit && prod.all_ones(53, sh - 1)) *ambiguous = true;   // within 2^53 ulps below 1/2
            }
        }
        D = I + (up ? 1u : 0u);
        if (D == 10000000000000000ull) { D = 1000000000000000ull; ++k; }
    }
    const unsigned hi8 = static_cast<unsigned>(D / 100000000ull), lo8 = static_cast<unsigned>(D % 100000000ull);
    char digits[16];
    put_8digits(digits, hi8);
    put_8digits(digits + 8, lo8);
    *p++ = digits[0];
    *p++ = '.';
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int i = 1; i < 16; ++i) *p++ = digits[i];
    *p++ = 'e';
    *p++ = k < 0 ? '-' : '+';
    const unsigned ak = static_cast<unsigned>(k < 0 ? -k : k);
    if (ak >= 100) {
        *p++ = static_cast<char>('0' + ak / 100);
        return put_2digits(p, ak % 100);
    }
    return put_2digits(p, ak);
}

}  // namespace text
}  // namespace cpprob
#endif  // CPPROB_B200_TEXT_FORMAT_CUH
