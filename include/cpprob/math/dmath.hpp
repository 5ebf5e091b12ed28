// cpprob-b200: fp64 elementary functions used by the samplers, log-pdfs and the weight reduction.
//
// On the device the hot ones (log of a unit-interval uniform, sqrt, sin/cos of 2*pi*u, exp of a
// shifted log-weight) are branch-free DFMA chains: the SIS kernels are bound by the FP64 pipe
// (64 lanes/SM, one warp instruction per 2 cycles per SM sub-partition), so every instruction that
// is not an FP64-pipe instruction competes with it only for issue slots, and every FP64 instruction
// that can be avoided is time saved.  Polynomial coefficients are near-minimax (tools/gen_poly.py)
// and live in __constant__ memory so that they reach the DFMAs as uniform-register operands instead
// of two immediate moves each.  Everything has a host twin (libm) used only by the structure probe /
// dry run.
//
// Accuracy (checked on the GPU in tests/test_dmath_gpu.py against numpy/mpmath):
//   log_unit, log     <= 2 ulp
//   sqrt_pos          <= 1 ulp (Newton from the hardware seed + exact residual step)
//   sincos_2pi, cos_2pi  abs error <= 2^-52
//   exp_weight        <= 2 ulp on (-708, 709); exactly 0 at or below -708.4 (and for -inf)
#ifndef CPPROB_MATH_DMATH_HPP
#define CPPROB_MATH_DMATH_HPP

#include <cmath>
#include <cstdint>
#include <limits>

#include "cpprob/hd.hpp"

namespace cpprob {
namespace dm {

constexpr double pi      = 3.141592653589793238462643383279502884;
constexpr double two_pi  = 6.283185307179586476925286766559005768;

CPPROB_HD double neg_inf() { return -std::numeric_limits<double>::infinity(); }

#if defined(__CUDACC__)
namespace tbl {
// atanh(s)/s = 1 + z*P(z), z = s^2 <= 0.02944; degree 6, error 2^-57.6 (gen_poly.py "LOG_P")
static __constant__ double log_p[7] = {
    0x1.5555555555558p-2, 0x1.99999999952e2p-3, 0x1.2492492df148dp-3, 0x1.c71c62e5800a1p-4,
    0x1.7462b4ab2ef6bp-4, 0x1.39fe606542ddep-4, 0x1.2b584aae78a57p-4};
// exp(r) = 1 + r + r^2 Q(r), |r| <= ln2/2; degree 9, error 2^-56 (gen_poly.py "EXP_Q")
static __constant__ double exp_q[10] = {
    0x1.0000000000001p-1, 0x1.5555555555556p-3, 0x1.5555555553d63p-5, 0x1.11111111109b3p-7,
    0x1.6c16c1788bd90p-10, 0x1.a01a01a7c41d5p-13, 0x1.a019b90d2ae7ap-16, 0x1.71de0dae63bb3p-19,
    0x1.289185613a3d6p-22, 0x1.af38a9b0ec855p-26};
// sinpi(r) = r*pi + r*z*S(z), cospi(r) = 1 + z*C(z), z = r^2, |r| <= 1/4 (gen_poly.py)
// laid out as pairs {S_i, C_i}, S padded with a zero top coefficient, so one Horner chain serves both
static __constant__ double sincos_sc[7][2] = {
    {-0x1.4abbce625be52p+2, -0x1.3bd3cc9be45dep+2},
    {0x1.466bc6775a476p+1, 0x1.03c1f081b5ac0p+2},
    {-0x1.32d2cce500387p-1, -0x1.55d3c7e3cb241p+0},
    {0x1.50783208843ebp-4, 0x1.e1f5068688d5bp-3},
    {-0x1.e3027dea82bd7p-8, -0x1.a6d1eef479be1p-6},
    {0x1.e4a9d9166f052p-12, 0x1.f9ce245cada0bp-10},
    {0.0, -0x1.b2f3eb054afcdp-14}};
// scalar constants, kept next to the tables so that they too arrive as uniform-register operands
static __constant__ double k_ln2_hi = 6.93147180369123816490e-01;   // 32 trailing zero bits: e*ln2_hi exact
static __constant__ double k_ln2_lo = 1.90821492927058770002e-10;
static __constant__ double k_log2e = 1.4426950408889634074;
static __constant__ double k_round_magic = 6755399441055744.0;      // 2^52 + 2^51: x + magic rounds x to an integer
static __constant__ double k_pi = 3.141592653589793238462643383279502884;
// exp_weight_tab: 2^(j/4096) (tools/gen_exp2_table.py), copied to shared memory by exp2_table_load()
#include "cpprob/math/exp2_table.inc"
constexpr int kExpTabBits = 12;
constexpr int kExpTabSize = 1 << kExpTabBits;
static __device__ const double exp2_tab[kExpTabSize] = {CPPROB_EXP2_TABLE_ROWS};
static __constant__ double k_log2e_tab = 1.4426950408889634074 * kExpTabSize;
static __constant__ double k_ln2_hi_tab = 6.93147180369123816490e-01 / kExpTabSize;   // exact scalings of k_ln2_hi / k_ln2_lo
static __constant__ double k_ln2_lo_tab = 1.90821492927058770002e-10 / kExpTabSize;
static __constant__ double k_exp_c1 = 1.0 / 6.0;
}  // namespace tbl
#endif

#if CPPROB_ON_DEVICE
// ---------------------------------------------------------------------------------------------
// Device implementations.
// ---------------------------------------------------------------------------------------------
namespace detail {

// log(m) + e*ln2 given s = (m-1)/(m+1)
__device__ __forceinline__ double log_tail(double s, int e)
{
    const double z = s * s;
    double p = tbl::log_p[6];
    p = fma(p, z, tbl::log_p[5]);
    p = fma(p, z, tbl::log_p[4]);
    p = fma(p, z, tbl::log_p[3]);
    p = fma(p, z, tbl::log_p[2]);
    p = fma(p, z, tbl::log_p[1]);
    p = fma(p, z, tbl::log_p[0]);
    const double s2 = s + s;
    const double t = s2 * z;
    const double de = static_cast<double>(e);
    const double inner = fma(t, p, fma(de, tbl::k_ln2_lo, s2));
    return fma(de, tbl::k_ln2_hi, inner);
}
}  // namespace detail

// Natural log of a normal, finite, positive argument (no zero / denormal / inf / nan handling):
// the sampler's log of a unit-interval uniform.  m in [sqrt(1/2), sqrt(2)), log x = e ln2 + 2 atanh(s).
__device__ __forceinline__ double log_unit(double x)
{
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    const int e = (hi - 0x3FE6A09F) >> 20;          // arithmetic shift = floor
    hi -= e << 20;
    const double m = __hiloint2double(hi, lo);
    const double num = m - 1.0;
    const double den = m + 1.0;
    // 1/den, den in [1.7, 2.42): hardware seed on the high word (~2^-20) + 2 Newton steps
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
    double t = fma(-den, y, 1.0);
    y = fma(y, t, y);
    t = fma(-den, y, 1.0);
    y = fma(y, t, y);
    double s = num * y;
    s = fma(fma(-den, s, num), y, s);               // residual step: s = num/den to < 1 ulp
    return detail::log_tail(s, e);
}

// General natural log, IEEE special cases included, written without opaque intrinsics so that the
// compiler folds it for constant arguments and hoists it for loop-invariant ones (the log-pdf
// normalisers: log(2 pi sigma^2), log(b-a), log(lambda) ...).
__device__ __forceinline__ double log(double x)
{
    const bool tiny = x < 2.2250738585072014e-308;                 // denormal, zero or negative
    const double xs = tiny ? x * 18014398509481984.0 : x;          // * 2^54
    long long bits = __double_as_longlong(xs);
    int hi = static_cast<int>(bits >> 32);
    const int e = (hi - 0x3FE6A09F) >> 20;
    hi -= e << 20;
    bits = (static_cast<long long>(hi) << 32) | (bits & 0xffffffffLL);
    const double m = __longlong_as_double(bits);
    const double s = (m - 1.0) / (m + 1.0);
    double r = detail::log_tail(s, tiny ? e - 54 : e);
    r = (x == 0.0) ? neg_inf() : r;
    r = (x < 0.0) ? __longlong_as_double(0x7ff8000000000000LL) : r;
    r = (x == std::numeric_limits<double>::infinity()) ? x : r;
    r = (x != x) ? x : r;
    return r;
}

// sqrt of a positive normal number (no zero / denormal / inf / negative handling).
__device__ __forceinline__ double sqrt_pos(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));        // ~2^-20 on the high word
    const double e = fma(x, -(y * y), 1.0);
    const double p = fma(e, 0.375, 0.5);
    y = fma(p, y * e, y);                                           // third-order step: ~2^-58
    const double g = x * y;
    const double h = 0.5 * y;
    return fma(fma(g, -g, x), h, g);                                // exact-residual correction
}

__device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }

namespace detail {
// quarter-turn reduction of the angle 2*pi*u, u in [0,1]: k = rint(4u) in {0..4}, r in [-1/4, 1/4]
// with 2*pi*u = (pi/2) k + pi r.  Both steps are exact.
__device__ __forceinline__ void reduce_2pi(double u, int & k, double & r)
{
    const double magic = tbl::k_round_magic;
    const double km = fma(u, 4.0, magic);
    k = __double2loint(km);
    const double kf = km - magic;
    r = fma(kf, -0.5, u + u);
}
}  // namespace detail

// sin(2*pi*u) and cos(2*pi*u), u in [0,1].
__device__ __forceinline__ void sincos_2pi(double u, double & s, double & c)
{
    int k;
    double r;
    detail::reduce_2pi(u, k, r);
    const double z = r * r;
    double ps = tbl::sincos_sc[5][0], pc = tbl::sincos_sc[6][1];
    pc = fma(pc, z, tbl::sincos_sc[5][1]);
#pragma unroll
    for (int i = 4; i >= 0; --i) {
        ps = fma(ps, z, tbl::sincos_sc[i][0]);
        pc = fma(pc, z, tbl::sincos_sc[i][1]);
    }
    const double sv = fma(r * z, ps, r * tbl::k_pi);
    const double cv = fma(z, pc, 1.0);
    // k: 0 -> (s, c) = (sv, cv); 1 -> (cv, -sv); 2 -> (-sv, -cv); 3 -> (-cv, sv); 4 == 0
    double so = (k & 1) ? cv : sv;
    double co = (k & 1) ? sv : cv;
    const int s_neg = (k >> 1) & 1;
    const int c_neg = ((k + 1) >> 1) & 1;
    so = __hiloint2double(__double2hiint(so) ^ (s_neg << 31), __double2loint(so));
    co = __hiloint2double(__double2hiint(co) ^ (c_neg << 31), __double2loint(co));
    s = so;
    c = co;
}

// One of sin / cos of 2*pi*u through a single Horner chain with per-lane selected coefficients:
// the selects go to the ALU pipe, the FP64 pipe sees one polynomial instead of two.
template<bool WantCos>
__device__ __forceinline__ double sin_or_cos_2pi(double u)
{
    int k;
    double r;
    detail::reduce_2pi(u, k, r);
    const double z = r * r;
    const bool use_cos_poly = WantCos ? !(k & 1) : (k & 1);
    const int neg = WantCos ? (((k + 1) >> 1) & 1) : ((k >> 1) & 1);
    double p = use_cos_poly ? tbl::sincos_sc[6][1] : tbl::sincos_sc[6][0];
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        const double ci = use_cos_poly ? tbl::sincos_sc[i][1] : tbl::sincos_sc[i][0];
        p = fma(p, z, ci);
    }
    const double head = use_cos_poly ? 1.0 : r * tbl::k_pi;
    const double tail = use_cos_poly ? z : r * z;
    const double v = fma(tail, p, head);
    return __hiloint2double(__double2hiint(v) ^ (neg << 31), __double2loint(v));
}
__device__ __forceinline__ double cos_2pi(double u) { return sin_or_cos_2pi<true>(u); }
__device__ __forceinline__ double sin_2pi(double u) { return sin_or_cos_2pi<false>(u); }

__device__ __forceinline__ double exp(double x) { return ::exp(x); }

// exp(log_w - m_ref) for the weight reduction.  The argument is first clamped from below at -745 on its
// bit pattern (ALU pipe, no branch), so -inf and every x <= -708 give exactly 0 through the ordinary
// path (2^k below the normal range is flushed).  x >= 709.8, +inf and NaN are NOT handled here: the
// kernels count NaN log-weights, and a maximum far above m_ref triggers the engine's re-base pass, so
// such values never reach a reported estimate.
__device__ __forceinline__ double exp_weight(double x)
{
    // negative doubles order like their unsigned high words: hi(x) > hi(-745.0) (unsigned) <=> x < -745
    const unsigned xh = static_cast<unsigned>(__double2hiint(x));
    const bool low = xh > 0xC0874800u;                               // x < -745 (or -inf, or a negative NaN)
    const double xc = __hiloint2double(low ? static_cast<int>(0xC0874800u) : static_cast<int>(xh), low ? 0 : __double2loint(x));
    const double magic = tbl::k_round_magic;
    const double t = fma(xc, tbl::k_log2e, magic);
    const int k = __double2loint(t);
    const double kf = t - magic;
    double r = fma(kf, -tbl::k_ln2_hi, xc);
    r = fma(kf, -tbl::k_ln2_lo, r);
    double q = tbl::exp_q[9];
#pragma unroll
    for (int i = 8; i >= 0; --i) q = fma(q, r, tbl::exp_q[i]);
    const double e = fma(r * r, q, r) + 1.0;                        // in [0.70, 1.42]
    // times 2^k, built from the clamped biased exponent: k <= -1023 gives a factor of exactly 0, so
    // weights below the normal range are flushed without a select the compiler could turn into a branch
    const int biased = max(k + 1023, 0);
    return e * __hiloint2double(biased << 20, 0);
}

// Fast-path variant for the fused kernel: no clamp, 2^k applied by adding k to the exponent field.
// Valid only when x is finite and the result is a normal number (k >= -1021); the caller tracks the
// smallest k and the exponent field of x per chunk and recomputes the chunk with exp_weight otherwise.
__device__ __forceinline__ double exp_weight_unchecked(double x, int & k_out)
{
    const double magic = tbl::k_round_magic;
    const double t = fma(x, tbl::k_log2e, magic);
    const int k = __double2loint(t);
    const double kf = t - magic;
    double r = fma(kf, -tbl::k_ln2_hi, x);
    r = fma(kf, -tbl::k_ln2_lo, r);
    double q = tbl::exp_q[9];
#pragma unroll
    for (int i = 8; i >= 0; --i) q = fma(q, r, tbl::exp_q[i]);
    const double e = fma(r * r, q, r) + 1.0;
    k_out = k;
    return __hiloint2double(__double2hiint(e) + (k << 20), __double2loint(e));
}

__device__ __forceinline__ double lgamma(double x) { return ::lgamma(x); }
__device__ __forceinline__ double pow(double a, double b) { return ::pow(a, b); }
__device__ __forceinline__ double floor(double x) { return ::floor(x); }
__device__ __forceinline__ double fabs(double x) { return ::fabs(x); }

#else
// ---------------------------------------------------------------------------------------------
// Host twins (structure probe / dry run only).
// ---------------------------------------------------------------------------------------------
inline double log_unit(double u) { return std::log(u); }
inline double log(double x) { return std::log(x); }
inline double sqrt_pos(double x) { return std::sqrt(x); }
inline double sqrt(double x) { return std::sqrt(x); }
inline void sincos_2pi(double u, double & s, double & c)
{
    s = std::sin(two_pi * u);
    c = std::cos(two_pi * u);
}
inline double cos_2pi(double u) { return std::cos(two_pi * u); }
inline double sin_2pi(double u) { return std::sin(two_pi * u); }
inline double exp(double x) { return std::exp(x); }
inline double exp_weight(double x) { return x != x ? 0.0 : std::exp(x); }
inline double exp_weight_unchecked(double x, int & k_out) { k_out = 0; return std::exp(x); }
inline double lgamma(double x) { return std::lgamma(x); }
inline double pow(double a, double b) { return std::pow(a, b); }
inline double floor(double x) { return std::floor(x); }
inline double fabs(double x) { return std::fabs(x); }
#endif

#if defined(__CUDACC__)
// Table-assisted variant of exp_weight_unchecked for the fused kernel: exp(x) = 2^k 2^(j/4096) e^r with
// n = rint(4096 x / ln 2) = 4096 k + j and |r| <= ln2/8192, so e^r - 1 = r + r^2 (1/2 + r/6) to 2e-18:
// 8 FP64-pipe instructions instead of 16, plus one shared-memory load from a 32 KB table (the FP64 pipe's two
// issue cycles per instruction make that a good trade).  `tab` is what exp2_table_load()
// returned.  Contract: x finite and the result a normal number, i.e. x >= -707 (and below the overflow range, which
// the engine's re-base pass takes care of).  The caller checks that on the high word of x itself
// (exp_arg_too_low) — not on n, which is only the low 32 bits of rint(4096 x / ln 2) and wraps for |x| > 3.6e5 —
// and recomputes with exp_weight otherwise.  Non-finite x gives NaN.
// The shared-memory copy stores each entry with its high word pre-decremented by j << 8, so that adding n << 8
// (n = 4096 k + j) lands on hi(T[j]) + (k << 20): the 2^k scaling costs one integer multiply-add on the loaded word.
constexpr unsigned kExpTabBytes = tbl::kExpTabSize * sizeof(double);

__device__ __forceinline__ unsigned exp2_table_load()
{
    __shared__ double t[tbl::kExpTabSize];
    for (int i = threadIdx.x; i < tbl::kExpTabSize; i += blockDim.x) {
        const double v = tbl::exp2_tab[i];
        t[i] = __hiloint2double(__double2hiint(v) - (i << (20 - tbl::kExpTabBits)), __double2loint(v));
    }
    __syncthreads();
    unsigned base;
    asm volatile("{ .reg .u64 p; cvta.to.shared.u64 p, %1; cvt.u32.u64 %0, p; }" : "=r"(base) : "l"(t));
    return base;
}
__device__ __forceinline__ double exp_weight_tab(double x, unsigned tab)
{
    const double magic = tbl::k_round_magic;
    const double t = fma(x, tbl::k_log2e_tab, magic);
    const int n = __double2loint(t);
    const double kf = t - magic;
    double r = fma(kf, -tbl::k_ln2_hi_tab, x);
    r = fma(kf, -tbl::k_ln2_lo_tab, r);
    int t_lo, t_hi;
    asm("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(t_lo), "=r"(t_hi)
        : "r"(tab + ((static_cast<unsigned>(n) << 3) & ((tbl::kExpTabSize - 1u) << 3))));
    const double tj = __hiloint2double(t_hi + (n << (20 - tbl::kExpTabBits)), t_lo);          // 2^k 2^(j/4096)
    const double q = fma(r, tbl::k_exp_c1, 0.5);
    const double p = fma(r * r, q, r);
    return fma(tj, p, tj);
}
// Negative doubles order like their unsigned high words: the running maximum of exp_arg_key(x) over a set of
// arguments exceeds kExpArgKeyLimit iff one of them is below -707 (or is -inf / a negative NaN), whatever its
// magnitude.  Non-negative arguments have keys below 0x80000000 and never matter.
__device__ __forceinline__ unsigned exp_arg_key(double x) { return static_cast<unsigned>(__double2hiint(x)); }
constexpr unsigned kExpArgKeyLimit = 0xC0861800u;         // high word of -707.0
#endif  // __CUDACC__

}  // namespace dm
}  // namespace cpprob
#endif  // CPPROB_MATH_DMATH_HPP
