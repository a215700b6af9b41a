// Test-only host build of toypathtracer_b200/csrc/tpt_fastdiv.h
#include "../../toypathtracer_b200/csrc/tpt_fastdiv.h"
extern "C" long long check_fastdiv(uint32_t d, uint32_t seed, long long n)
{
    const tpt::FastDiv f = tpt::make_fastdiv(d);
    long long bad = 0;
    uint32_t s = seed | 1u;
    const uint32_t edges[] = {0u, 1u, d - 1, d, d + 1, 2 * d - 1, 2 * d, 0x7fffffffu, 0x7ffffffeu, 0x40000000u};
    for (uint32_t e : edges) if (e <= 0x7fffffffu && tpt::fdiv(e, f) != e / d) ++bad;
    for (long long i = 0; i < n; ++i)
    {
        s ^= s << 13; s ^= s >> 17; s ^= s << 15;
        const uint32_t v = s & 0x7fffffffu;
        if (tpt::fdiv(v, f) != v / d) ++bad;
    }
    return bad;
}
