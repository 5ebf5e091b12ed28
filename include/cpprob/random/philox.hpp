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
// so that a Box-Muller pair left over by the first particle (the README model draws a single
// normal) is used by the second one instead of being thrown away, while consecutive lanes still
// own consecutive particles (coalesced trace rows).  The draws of particle p are a pure function of
// (seed, p): the multiset of samples of a run is identical for any GPU count / grid size / chunk
// schedule.
//
// Word consumption rules (identical on host and device, they are part of the stream definition):
//   * next_u32()      takes one 32-bit word from the current block (4 per block);
//   * next_uniform()  takes an aligned pair of words (52 random mantissa bits, result in (0,1));
//   * next_std_normal() returns the cached second Box-Muller variate if there is one; otherwise it
//     discards what is left of the current block, takes a whole fresh block, runs one Box-Muller
//     transform and caches the sine branch.
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
}  // namespace detail

// One random stream.  All members live in registers on the device.
class philox_stream {
public:
    CPPROB_HD philox_stream(const philox_keys & keys, std::uint64_t stream, std::uint32_t stream_tag = 0)
        : keys_(keys), s_lo_(static_cast<std::uint32_t>(stream)), s_hi_(static_cast<std::uint32_t>(stream >> 32)),
          tag_(stream_tag), blk_(0), pos_(4), w0_(0), w1_(0), w2_(0), w3_(0),
          spare_(0.0), has_spare_(false) {}

    CPPROB_HD std::uint32_t next_u32()
    {
        if (pos_ >= 4) refill();
        // the pool is a shift register: the next unread word is always w0_ (three moves instead of a select chain)
        const std::uint32_t r = w0_;
        w0_ = w1_; w1_ = w2_; w2_ = w3_;
        ++pos_;
        return r;
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
    double spare_;
    bool has_spare_;
};

}  // namespace cpprob
#endif  // CPPROB_RANDOM_PHILOX_HPP
