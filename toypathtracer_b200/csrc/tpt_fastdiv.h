// Division of 32-bit unsigned numbers (n < 2^31) by a launch-time constant with one multiply-high and a shift
// (Granlund & Montgomery): used where a kernel maps a flat path index to (pixel, sample) per lane.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TPT_FD_HD __host__ __device__ __forceinline__
#else
#define TPT_FD_HD inline
#endif

namespace tpt {

struct FastDiv { uint32_t mul, shift, d; };

TPT_FD_HD uint32_t fdiv(uint32_t n, const FastDiv f)
{
    if (f.d == 1) return n;
#if defined(__CUDA_ARCH__)
    return (uint32_t)__umulhi(n, f.mul) >> f.shift;
#else
    return (uint32_t)(((uint64_t)n * f.mul) >> 32) >> f.shift;
#endif
}

inline FastDiv make_fastdiv(uint32_t d)
{
    FastDiv f; f.d = d; f.mul = 0; f.shift = 0;
    if (d <= 1) return f;
    uint32_t l = 0; while ((1u << l) < d) ++l;                    // ceil(log2 d)
    const uint64_t m = ((1ull << (32 + l)) + d - 1) / d;            // may need 33 bits
    if (m >> 32) { const uint64_t m2 = ((1ull << (31 + l)) + d - 1) / d; f.mul = (uint32_t)m2; f.shift = l - 1; }  // exact for n < 2^31
    else { f.mul = (uint32_t)m; f.shift = l; }
    return f;
}

} // namespace tpt
