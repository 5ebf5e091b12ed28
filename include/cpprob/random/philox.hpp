// cpprob-b200: counter-based random streams for the SIS particle path.
//
// The reference draws every prior sample from ONE process-global std::mt19937 seeded from
// std::random_device (/root/reference: include/cpprob/utils.hpp:34-42, src/cpprob/utils.cpp:16-20;
// call site cpprob.hpp:72-74 `distr(get_rng())`).  A serial generator cannot feed 10^5 resident
// threads, and a random_device seed is not reproducible, so this engine replaces it with
// Philox4x32-10 (Salmon et al., SC'11):
//
//     key     = (seed_lo, seed_hi)
//     counter = (stream_lo, stream_hi, block_index, stream_tag)
//
// Particle -> stream map.  Particles are grouped in tiles of 512 consecutive global indices; the
// two particles p and p + 256 of a tile share one stream and consume it one after the other
// (p first):
//
//     stream(p) = (p >> 9) * 256 + (p & 255)        turn(p) = (p >> 8) & 1
//
// so that the half block left over by the first particle (the README model draws a single normal
// = two words) is used by the second one instead of being thrown away, while consecutive lanes
// still own consecutive particles (coalesced trace rows).  The draws of particle p are a pure function of
// (seed, p): the multiset of samples of a run is identical for any GPU count / grid size / chunk
// schedule.
//
// Word consumption rules (identical on host and device, they are part of the stream definition):
//   * next_u32()      takes one 32-bit word from the current block (4 per block);
//   * next_u32x4()    drops what is left of the current block and takes the whole next one;
//   * next_std_normal_x4() drops what is left of the current block and runs the four pair trials of the next two;
//   * next_uniform()  takes an aligned pair of words (52 random mantissa bits, result in (0,1));
//   * next_std_normal() takes an aligned pair of words (a, b) and runs one ziggurat trial on them
//     (8192 layers, tools/gen_ziggurat.py): layer = bits 3..15 of a, sign = bit 0 of a, u = (b : a >> 12)
//     * 2^-52.  99.94 % of the draws end there.  The others (wedges, the tail, rejected trials) take
//     all further randomness from a SIDE stream — same stream id, counter word 3 = tag | 0x80000000,
//     counter word 2 = (block index << 2 | word position) of the pair, key rotated by the trial
//     number — so the main stream's position never depends on how a draw went.
//     (-DCPPROB_NORMAL_BOX_MULLER selects the previous Box-Muller sampler: a whole block per pair
//     of variates, the sine branch cached; kept for A/B measurements only.)
#ifndef CPPROB_RANDOM_PHILOX_HPP
#define CPPROB_RANDOM_PHILOX_HPP

#include <cstdint>
#include <cstring>

#include "cpprob/hd.hpp"
#include "cpprob/math/dmath.hpp"

namespace cpprob {

// The ten round keys, expanded once on the host: inside the kernels they are kernel-parameter
// constants (operands straight from the constant bank) instead of 20 integer adds per block.
struct philox_keys {
    static constexpr std::uint32_t W0 = 0x9E3779B9u;
    static constexpr std::uint32_t W1 = 0xBB67AE85u;
    std::uint32_t k[20];

    philox_keys() = default;
    CPPROB_HD explicit philox_keys(std::uint64_t seed)
    {
        std::uint32_t a = static_cast<std::uint32_t>(seed), b = static_cast<std::uint32_t>(seed >> 32);
        for (int r = 0; r < 10; ++r) {
            k[2 * r] = a;
            k[2 * r + 1] = b;
            a += W0;
            b += W1;
        }
    }
    CPPROB_HD philox_keys(std::uint32_t k0, std::uint32_t k1)
    {
        for (int r = 0; r < 10; ++r) {
            k[2 * r] = k0;
            k[2 * r + 1] = k1;
            k0 += W0;
            k1 += W1;
        }
    }
};

struct philox4x32 {
    static constexpr std::uint32_t M0 = 0xD2511F53u;
    static constexpr std::uint32_t M1 = 0xCD9E8D57u;

    // One Philox4x32-10 block.  Verified against the Random123 known-answer vectors
    // (tests/test_philox.py on the host twin, tests/test_device_layer_gpu.py on the GPU).
    static CPPROB_HD void block(std::uint32_t c0, std::uint32_t c1, std::uint32_t c2, std::uint32_t c3,
                                const philox_keys & keys,
                                std::uint32_t & o0, std::uint32_t & o1, std::uint32_t & o2, std::uint32_t & o3)
    {
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int r = 0; r < 10; ++r) {
            const std::uint64_t p0 = static_cast<std::uint64_t>(M0) * c0;
            const std::uint64_t p1 = static_cast<std::uint64_t>(M1) * c2;
            const std::uint32_t n0 = static_cast<std::uint32_t>(p1 >> 32) ^ c1 ^ keys.k[2 * r];
            const std::uint32_t n1 = static_cast<std::uint32_t>(p1);
            const std::uint32_t n2 = static_cast<std::uint32_t>(p0 >> 32) ^ c3 ^ keys.k[2 * r + 1];
            const std::uint32_t n3 = static_cast<std::uint32_t>(p0);
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        }
        o0 = c0; o1 = c1; o2 = c2; o3 = c3;
    }
};

// particle -> (stream, turn), see the header comment
constexpr unsigned kPairStride = 256;
CPPROB_HD std::uint64_t stream_of_particle(std::uint64_t p) { return ((p >> 9) << 8) | (p & 255u); }
CPPROB_HD unsigned turn_of_particle(std::uint64_t p) { return static_cast<unsigned>(p >> 8) & 1u; }

namespace detail {
// 52 random mantissa bits -> (k + 0.5) * 2^-52, strictly inside (0,1), exactly representable.
CPPROB_HD double u52_to_open01(std::uint32_t hi_word, std::uint32_t lo_word)
{
    const std::uint32_t hi = (hi_word >> 12) | 0x3FF00000u;           // [1,2)
#if CPPROB_ON_DEVICE
    const double d = __hiloint2double(static_cast<int>(hi), static_cast<int>(lo_word));
#else
    const std::uint64_t bits = (static_cast<std::uint64_t>(hi) << 32) | lo_word;
    double d;
    std::memcpy(&d, &bits, sizeof d);
#endif
    return d - 0.99999999999999988897769753748434595763683319091796875;   // 1 - 2^-53 (exact)
}
// two 32-bit words stored in the bytes of one double slot of a model table (lo word first)
CPPROB_HD void pack_u32_pair(double * slot, std::uint32_t lo, std::uint32_t hi)
{
    std::uint32_t * w = reinterpret_cast<std::uint32_t *>(slot);
    w[0] = lo;
    w[1] = hi;
}
}  // namespace detail


// ------------------------------------------------------------------------------------------------
// Ziggurat (tools/gen_ziggurat.py).  X[i]: right edges of the layers (X[0] = V/f(r), X[1] = r, X[N] = 0),
// F[i] = exp(-X[i]^2/2).  On the device the fast path reads X[i] and X[i+1] with a per-lane index from a
// shared-memory copy (zig::load_shared() at kernel start, returns the table's shared-space address, which
// the streams keep in a register); X and F in global memory are only touched by the slow path.
//
// One trial on the word pair (a, b):   layer i = bits 3..15 of a (so that `a & 0xfff8` IS the byte offset of
// X[i]), sign = bit 0 of a, u = 0.m with the 52-bit mantissa m = b : (a >> 12) — 48 independent bits; the last
// four are also the top layer bits, i.e. a per-layer offset below 2^-48 that saves an instruction per draw —
// x = u X[i] rounded once, as fma(1 + u, X[i], -X[i]).  Fast accept: the high word of x is below the high word
// of X[i+1].  8192 layers make the fast path settle 99.94 % of the draws: what is left is warp-divergent work
// (a full warp waits for its one slow lane), so the table is sized to make it rare rather than cheap.
// ------------------------------------------------------------------------------------------------
namespace zig {
#include "cpprob/random/ziggurat_table.inc"
constexpr int N = CPPROB_ZIG_N;
constexpr double R = CPPROB_ZIG_R;
static_assert(N == 8192, "the bit layout of a trial (layer = bits 3..15 of the first word) is written for 8192 layers");
constexpr unsigned kSharedBytes = (N + 1) * sizeof(double);   // dynamic shared memory a kernel that draws normals is launched with

#if defined(__CUDACC__)
static __device__ const double d_x[N + 1] = {CPPROB_ZIG_X_ROWS};
static __device__ const double d_f[N + 1] = {CPPROB_ZIG_F_ROWS};
#endif
inline const double * h_x() { static const double t[N + 1] = {CPPROB_ZIG_X_ROWS}; return t; }
inline const double * h_f() { static const double t[N + 1] = {CPPROB_ZIG_F_ROWS}; return t; }

#if defined(__CUDACC__)
// Every kernel that may draw a normal calls this once (all threads of the CTA) and hands the result to its
// streams.  The address is produced by a volatile asm so that it stays in one register for the whole kernel
// instead of being re-derived (S2UR/ULEA...) at every table access.
__device__ __forceinline__ unsigned load_shared()
{
    extern __shared__ double cpprob_zig_shared[];      // zig::kSharedBytes, asked for at launch (64 KB: beyond the static limit)
    double * const t = cpprob_zig_shared;
    for (int i = threadIdx.x; i <= N; i += blockDim.x) t[i] = d_x[i];
    __syncthreads();
    unsigned base;
    asm volatile("{ .reg .u64 p; cvta.to.shared.u64 p, %1; cvt.u32.u64 %0, p; }" : "=r"(base) : "l"(t));
    return base;
}
#endif

CPPROB_HD double x_of(unsigned i)
{
#if CPPROB_ON_DEVICE
    return d_x[i];
#else
    return h_x()[i];
#endif
}
CPPROB_HD double f_of(unsigned i)
{
#if CPPROB_ON_DEVICE
    return d_f[i];
#else
    return h_f()[i];
#endif
}

CPPROB_HD unsigned layer_of(std::uint32_t a) { return (a >> 3) & (N - 1); }

// |x| of one trial: u X[i], u = 0.m, m = b : (a >> 12); rounded once
CPPROB_HD double trial_abs(std::uint32_t a, std::uint32_t b, double xi)
{
    const std::uint32_t hi = 0x3FF00000u | (b >> 12);
    const std::uint32_t lo = (b << 20) | (a >> 12);
#if CPPROB_ON_DEVICE
    const double d = __hiloint2double(static_cast<int>(hi), static_cast<int>(lo));
    return fma(d, xi, -xi);
#else
    const std::uint64_t bits = (static_cast<std::uint64_t>(hi) << 32) | lo;
    double d;
    std::memcpy(&d, &bits, sizeof d);
    return std::fma(d, xi, -xi);
#endif
}

CPPROB_HD std::uint32_t high_word(double x)
{
#if CPPROB_ON_DEVICE
    return static_cast<std::uint32_t>(__double2hiint(x));
#else
    std::uint64_t bits;
    std::memcpy(&bits, &x, sizeof bits);
    return static_cast<std::uint32_t>(bits >> 32);
#endif
}

CPPROB_HD double with_sign(double x, std::uint32_t a)
{
#if CPPROB_ON_DEVICE
    return __hiloint2double(__double2hiint(x) ^ static_cast<int>(a << 31), __double2loint(x));
#else
    return (a & 1u) ? -x : x;
#endif
}

// exp(-x^2/2) for the wedge test, 0 <= x < 4.6: the engine's own branch-free exp on the device (2 ulp)
CPPROB_HD double gauss_kernel(double x)
{
#if CPPROB_ON_DEVICE
    int k;
    return dm::exp_weight_unchecked(-0.5 * x * x, k);
#else
    return std::exp(-0.5 * x * x);
#endif
}

// The 0.06 % of draws the fast test does not settle.  Pure function of its arguments (the main stream is
// not advanced); not inlined, so the particle loops carry only a call.  Everything comes in by value — the key
// as its two seed words, from which the round keys are re-derived with integer adds — so that the call touches
// no memory but the two table reads of a wedge.
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
inline double slow_path(std::uint32_t k0, std::uint32_t k1, std::uint32_t s_lo, std::uint32_t s_hi, std::uint32_t where,
                        std::uint32_t tag, std::uint32_t a, std::uint32_t b, double x, double x_next)
{
    const philox_keys keys(k0, k1);
    std::uint32_t side = 0;                                   // blocks taken from the side stream so far
    std::uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0;
    for (;;) {
        const unsigned i = layer_of(a);
        if (x < x_next) return with_sign(x, a);               // inside the next layer's rectangle after all
        philox4x32::block(s_lo, s_hi, where, (tag | 0x80000000u) + (side++ << 8), keys, q0, q1, q2, q3);
        if (i == 0) {
            // beyond r in the base strip: the tail, by Marsaglia's exponential rejection
            for (;;) {
                const double u1 = detail::u52_to_open01(q0, q1), u2 = detail::u52_to_open01(q2, q3);
                const double xx = -dm::log(u1) / R;
                const double yy = -dm::log(u2);
                if (yy + yy > xx * xx) return with_sign(R + xx, a);
                philox4x32::block(s_lo, s_hi, where, (tag | 0x80000000u) + (side++ << 8), keys, q0, q1, q2, q3);
            }
        }
        // wedge of layer i: y uniform between f(X[i]) and f(X[i+1])
        const double f0 = f_of(i), f1 = f_of(i + 1);
        const double y = fma(detail::u52_to_open01(q0, q1), f1 - f0, f0);
        if (y < gauss_kernel(x)) return with_sign(x, a);
        // rejected: a fresh trial from the rest of the side block
        a = q2;
        b = q3;
        const unsigned j = layer_of(a);
        x = trial_abs(a, b, x_of(j));
        x_next = x_of(j + 1);
    }
}
}  // namespace zig

// One random stream.  All members live in registers on the device.
class philox_stream {
public:
    // zig_base: on the device, what zig::load_shared() returned in this kernel (unused on the host)
    CPPROB_HD philox_stream(const philox_keys & keys, std::uint64_t stream, unsigned zig_base = 0, std::uint32_t stream_tag = 0)
        : keys_(keys), s_lo_(static_cast<std::uint32_t>(stream)), s_hi_(static_cast<std::uint32_t>(stream >> 32)),
          tag_(stream_tag), blk_(0), pos_(4), w0_(0), w1_(0), w2_(0), w3_(0), zig_base_(zig_base),
          spare_(0.0), has_spare_(false)
    {
#if CPPROB_ON_DEVICE
        // keep the stream id in two registers: without this the compiler re-derives it from the kernel parameters
        // (a dozen integer instructions) in front of every refill, i.e. every one or two draws of a long trace
        asm volatile("" : "+r"(s_lo_), "+r"(s_hi_));
#endif
    }

    CPPROB_HD std::uint32_t next_u32()
    {
        if (pos_ >= 4) refill();
        // the pool is a shift register: the next unread word is always w0_ (three moves instead of a select chain)
        const std::uint32_t r = w0_;
        w0_ = w1_; w1_ = w2_; w2_ = w3_;
        ++pos_;
        return r;
    }

    // A whole block at once: the unread words of the current block are dropped and the four words of the next block
    // are handed out together (the caller uses them in order).  For loops that draw one word per trip: taking them four
    // at a time keeps the pool bookkeeping (position counter, refill test, shifts) out of the trip.
    CPPROB_HD void next_u32x4(std::uint32_t (&w)[4])
    {
        philox4x32::block(s_lo_, s_hi_, blk_, tag_, keys_, w[0], w[1], w[2], w[3]);
        ++blk_;
        pos_ = 4;
    }

    // Uniform double in the open interval (0,1) with 52 random bits.
    CPPROB_HD double next_uniform()
    {
        if (pos_ >= 3) refill();          // need an aligned pair: block words (0,1) or (2,3)
        if (pos_ == 1) { w0_ = w1_; w1_ = w2_; w2_ = w3_; pos_ = 2; }   // drop the odd word
        const double u = detail::u52_to_open01(w0_, w1_);
        w0_ = w2_; w1_ = w3_;
        pos_ += 2;
        return u;
    }

#if defined(CPPROB_NORMAL_BOX_MULLER)
    // Standard normal by Box-Muller on a whole block; the sine branch is cached.
    CPPROB_HD double next_std_normal()
    {
        if (has_spare_) { has_spare_ = false; return spare_; }
        refill();
        const double u1 = detail::u52_to_open01(w0_, w1_);
        const double u2 = detail::u52_to_open01(w2_, w3_);
        pos_ = 4;
        const double r = dm::sqrt_pos(-2.0 * dm::log_unit(u1));
        double s, c;
        dm::sincos_2pi(u2, s, c);
        spare_ = r * s;
        has_spare_ = true;
        return r * c;
    }
#else
    // Standard normal by the ziggurat: one aligned pair of words, two shared-memory loads, one DFMA.
    CPPROB_HD double next_std_normal()
    {
        if (pos_ >= 3) refill();
        if (pos_ == 1) { w0_ = w1_; w1_ = w2_; w2_ = w3_; pos_ = 2; }   // drop the odd word
        const std::uint32_t a = w0_, b = w1_;
        const std::uint32_t where = ((blk_ - 1u) << 2) | pos_;
        w0_ = w2_; w1_ = w3_;
        pos_ += 2;
#if CPPROB_ON_DEVICE
        double xi, xn;
        const unsigned addr = zig_base_ + (a & 0xfff8u);
        asm("ld.shared.f64 %0, [%1];" : "=d"(xi) : "r"(addr));
        asm("ld.shared.f64 %0, [%1+8];" : "=d"(xn) : "r"(addr));
#else
        const double xi = zig::x_of(zig::layer_of(a)), xn = zig::x_of(zig::layer_of(a) + 1);
#endif
        const double x = zig::trial_abs(a, b, xi);
        if (CPPROB_UNLIKELY(zig::high_word(x) >= zig::high_word(xn))) return zig::slow_path(keys_.k[0], keys_.k[1], s_lo_, s_hi_, where, tag_, a, b, x, xn);
        return zig::with_sign(x, a);
    }
#endif

#if !defined(CPPROB_NORMAL_BOX_MULLER)
    // Four standard normals at once: the unread words of the current block are dropped, the next TWO blocks are
    // generated side by side (two independent multiply chains for the scheduler to interleave) and each of their four
    // aligned word pairs runs one ziggurat trial — the same trial next_std_normal() runs on that pair.  For loops that
    // draw one normal per trip: the pool bookkeeping leaves the trip and four draws overlap.
    CPPROB_HD void next_std_normal_x4(double (&z)[4])
    {
        std::uint32_t w[8];
        philox4x32::block(s_lo_, s_hi_, blk_, tag_, keys_, w[0], w[1], w[2], w[3]);
        philox4x32::block(s_lo_, s_hi_, blk_ + 1u, tag_, keys_, w[4], w[5], w[6], w[7]);
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int j = 0; j < 4; ++j) {
            const std::uint32_t a = w[2 * j], b = w[2 * j + 1];
            const std::uint32_t where = ((blk_ + static_cast<std::uint32_t>(j >> 1)) << 2) | static_cast<std::uint32_t>((j & 1) << 1);
#if CPPROB_ON_DEVICE
            double xi, xn;
            const unsigned addr = zig_base_ + (a & 0xfff8u);
            asm("ld.shared.f64 %0, [%1];" : "=d"(xi) : "r"(addr));
            asm("ld.shared.f64 %0, [%1+8];" : "=d"(xn) : "r"(addr));
#else
            const double xi = zig::x_of(zig::layer_of(a)), xn = zig::x_of(zig::layer_of(a) + 1);
#endif
            const double x = zig::trial_abs(a, b, xi);
            if (CPPROB_UNLIKELY(zig::high_word(x) >= zig::high_word(xn))) z[j] = zig::slow_path(keys_.k[0], keys_.k[1], s_lo_, s_hi_, where, tag_, a, b, x, xn);
            else z[j] = zig::with_sign(x, a);
        }
        blk_ += 2u;
        pos_ = 4;
    }
#endif

    CPPROB_HD std::uint32_t blocks_used() const { return blk_; }

private:
    CPPROB_HD void refill()
    {
        philox4x32::block(s_lo_, s_hi_, blk_, tag_, keys_, w0_, w1_, w2_, w3_);
        ++blk_;
        pos_ = 0;
    }

    const philox_keys & keys_;
    std::uint32_t s_lo_, s_hi_, tag_, blk_, pos_;
    std::uint32_t w0_, w1_, w2_, w3_;
    unsigned zig_base_;
    double spare_;
    bool has_spare_;
};

}  // namespace cpprob
#endif  // CPPROB_RANDOM_PHILOX_HPP
