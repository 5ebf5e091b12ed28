// cpprob-b200: host/device portability macros.
//
// Every header under include/cpprob that is used by the per-particle model path is written once and
// compiled twice: by nvcc as __device__ code inside the sm_100a kernels (the product path) and by a
// plain C++14 host compiler for the one-shot *structure probe* (address discovery) and for dry runs.
// There is no host execution of the inference loop itself.
#ifndef CPPROB_HD_HPP
#define CPPROB_HD_HPP

#if defined(__CUDACC__)
#  define CPPROB_HD __host__ __device__ __forceinline__
#  define CPPROB_D  __device__ __forceinline__
#else
#  define CPPROB_HD inline
#  define CPPROB_D  inline
#endif

#if defined(__CUDA_ARCH__)
#  define CPPROB_ON_DEVICE 1
#else
#  define CPPROB_ON_DEVICE 0
#endif

#if defined(__GNUC__) || defined(__CUDACC__)
#  define CPPROB_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#  define CPPROB_UNLIKELY(x) (x)
#endif

#endif  // CPPROB_HD_HPP
