// cpprob-b200: posterior records -> text on the GPU.
//
// One thread formats one record `([(id v) (id v) ...] logw)\n` (the line StateInfer::dump_predicts writes,
// /root/reference src/cpprob/state.cpp:262-267 with the grammar of include/cpprob/serialization.hpp:41-46,
// 71-98) straight from the SoA trace rows.  Lines have variable length (signs, exponent digits, ints), so the
// text of a batch is produced in three steps, all on the device:
//   k_text_lengths : every record's length + per-CTA sums
//   k_text_scan    : exclusive scan of the CTA sums (one CTA)
//   k_text_write   : every record formatted again at its final offset in one contiguous byte buffer
// Formatting twice is cheaper than staging: a %.15e costs a few hundred integer instructions, the batch is
// bound by the PCIe copy of the text that follows.  Records whose last digit could not be decided
// (text_format.cuh) are listed for the host to re-format.
#ifndef CPPROB_B200_TEXT_KERNELS_CUH
#define CPPROB_B200_TEXT_KERNELS_CUH

#include "sis_kernels.cuh"
#include "text_format.cuh"

namespace cpprob {
namespace engine {

constexpr int kTextBlock = 256;

struct text_slot {      // one predict statement of the kind being written
    int id, row, width;
};

struct text_flag {      // an ambiguous record: where its line sits in the text buffer
    unsigned long long record, offset;
    unsigned length, pad;
};

struct text_args {
    const text_slot * slots;     // device, [n_slots]
    int n_slots;
    int is_int;
    const double * real_rows;
    const int * int_rows;
    const double * logw;
    unsigned long long stride, n;
    unsigned long long first_particle;   // global index of record 0 (only used by the test hook below)
    unsigned long long force_every;      // test hook: records with global index % force_every == 0 are reported as
                                         // ambiguous and their first byte is damaged, so the host fix-up is exercised
};

// sink that only counts
struct count_sink {
    unsigned long long n = 0;
    __device__ __forceinline__ void put(char) { ++n; }
    __device__ __forceinline__ void put(const char * s, int len) { (void)s; n += static_cast<unsigned>(len); }
};
// sink that writes bytes to global memory
struct write_sink {
    char * p;
    __device__ __forceinline__ void put(char c) { *p++ = c; }
    __device__ __forceinline__ void put(const char * s, int len)
    {
        for (int i = 0; i < len; ++i) p[i] = s[i];
        p += len;
    }
};

template<class Sink>
__device__ __forceinline__ void put_double(Sink & out, double v, bool * ambiguous)
{
    char buf[32];
    const char * e = text::format_e15(buf, v, ambiguous);
    out.put(buf, static_cast<int>(e - buf));
}
template<class Sink>
__device__ __forceinline__ void put_integer(Sink & out, long long v)
{
    char buf[24];
    const char * e = text::put_int(buf, v);
    out.put(buf, static_cast<int>(e - buf));
}

template<class Sink>
__device__ __forceinline__ void format_record(Sink & out, const text_args & a, unsigned long long i, bool * ambiguous)
{
    out.put('(');
    out.put('[');
    for (int s = 0; s < a.n_slots; ++s) {
        const text_slot sl = a.slots[s];
        if (s) out.put(' ');
        out.put('(');
        put_integer(out, sl.id);
        out.put(' ');
        if (a.is_int) {
            put_integer(out, a.int_rows[static_cast<unsigned long long>(sl.row) * a.stride + i]);
        } else if (sl.width == 1) {
            put_double(out, a.real_rows[static_cast<unsigned long long>(sl.row) * a.stride + i], ambiguous);
        } else {                                  // NDArray vector value: "[v0 v1 ...]"
            out.put('[');
            for (int c = 0; c < sl.width; ++c) {
                if (c) out.put(' ');
                put_double(out, a.real_rows[static_cast<unsigned long long>(sl.row + c) * a.stride + i], ambiguous);
            }
            out.put(']');
        }
        out.put(')');
    }
    out.put(']');
    out.put(' ');
    put_double(out, a.logw[i], ambiguous);
    out.put(')');
    out.put('\n');
}

// inclusive scan of one value per thread over the CTA; returns the exclusive prefix of this thread and the total
__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long * smem /*[kTextBlock/32]*/,
                                                                   unsigned long long & total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) smem[warp] = x;
    __syncthreads();
    unsigned long long warp_base = 0, t = 0;
#pragma unroll
    for (int w = 0; w < kTextBlock / 32; ++w) {
        if (w < warp) warp_base += smem[w];
        t += smem[w];
    }
    total = t;
    __syncthreads();
    return warp_base + x - v;
}

__global__ void __launch_bounds__(kTextBlock) k_text_lengths(const text_args a, unsigned * __restrict__ lengths,
                                                            unsigned long long * __restrict__ block_sums)
{
    __shared__ unsigned long long smem[kTextBlock / 32];
    const unsigned long long i = blockIdx.x * static_cast<unsigned long long>(kTextBlock) + threadIdx.x;
    unsigned long long len = 0;
    if (i < a.n) {
        count_sink sink;
        bool amb = false;
        format_record(sink, a, i, &amb);
        len = sink.n;
        lengths[i] = static_cast<unsigned>(len);
    }
    unsigned long long total;
    block_exclusive_scan(len, smem, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// exclusive scan of n_blocks sums in place (one CTA); out_total[0] = grand total
__global__ void __launch_bounds__(1024) k_text_scan(unsigned long long * __restrict__ block_sums, unsigned n_blocks,
                                                    unsigned long long * __restrict__ out_total)
{
    __shared__ unsigned long long part[1024];
    const unsigned per = (n_blocks + 1023u) / 1024u;
    const unsigned lo = threadIdx.x * per, hi = min(lo + per, n_blocks);
    unsigned long long s = 0;
    for (unsigned i = lo; i < hi; ++i) s += block_sums[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int t = 0; t < 1024; ++t) {
            const unsigned long long v = part[t];
            part[t] = acc;
            acc += v;
        }
        out_total[0] = acc;
    }
    __syncthreads();
    unsigned long long acc = part[threadIdx.x];
    for (unsigned i = lo; i < hi; ++i) {
        const unsigned long long v = block_sums[i];
        block_sums[i] = acc;
        acc += v;
    }
}

__global__ void __launch_bounds__(kTextBlock) k_text_write(const text_args a, const unsigned * __restrict__ lengths,
                                                          const unsigned long long * __restrict__ block_offsets,
                                                          char * __restrict__ text, text_flag * __restrict__ flags,
                                                          unsigned long long * __restrict__ n_flags, unsigned max_flags)
{
    __shared__ unsigned long long smem[kTextBlock / 32];
    const unsigned long long i = blockIdx.x * static_cast<unsigned long long>(kTextBlock) + threadIdx.x;
    const unsigned long long len = i < a.n ? lengths[i] : 0;
    unsigned long long total;
    const unsigned long long off = block_offsets[blockIdx.x] + block_exclusive_scan(len, smem, total);
    if (i < a.n) {
        write_sink sink{text + off};
        bool amb = false;
        format_record(sink, a, i, &amb);
        if (a.force_every && (a.first_particle + i) % a.force_every == 0) {
            amb = true;
            text[off] = '#';
        }
        if (amb) {
            const unsigned long long slot = atomicAdd(n_flags, 1ull);
            if (slot < max_flags) flags[slot] = text_flag{i, off, static_cast<unsigned>(len), 0u};
        }
    }
}

}  // namespace engine
}  // namespace cpprob
#endif  // CPPROB_B200_TEXT_KERNELS_CUH
