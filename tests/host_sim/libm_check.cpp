// Test-only: host build of the product's glibc-faithful libm restatement (toypathtracer_b200/csrc/tpt_libm.cuh)
// checked exhaustively against the platform libm the reference links.
#include "../../toypathtracer_b200/csrc/tpt_libm.cuh"
#include <thread>
#include <vector>
#include <atomic>
#include <stdio.h>

static const float kPI = 3.1415926f;

extern "C" {
// all 2^24 values of RandomFloat01()*2.0f*kPI (Maths.cpp:42) == 2*kPI*eps2 (Test.cpp:115)
long long check_sincos_domain(long long* first_bad)
{
    long long bad = 0;
    for (uint32_t k = 0; k < (1u << 24); ++k)
    {
        float rf = k / 16777216.0f;
        float a = rf * 2.0f * kPI;
        float s, c;
        bool ok1 = tptlibm::sinf_glibc(a, &s), ok2 = tptlibm::cosf_glibc(a, &c);
        float rs = sinf(a), rc = cosf(a);
        if (!ok1 || !ok2 || tptlibm::f2u(s) != tptlibm::f2u(rs) || tptlibm::f2u(c) != tptlibm::f2u(rc))
        {
            if (!bad && first_bad) *first_bad = k;
            ++bad;
        }
    }
    return bad;
}
// every float in (-120,120) by bit pattern, stride `stride`
long long check_sincos_range(uint32_t stride)
{
    long long bad = 0;
    for (uint64_t u = 0; u < (1ull << 32); u += stride)
    {
        float a = tptlibm::u2f((uint32_t)u);
        float s, c;
        if (!(fabsf(a) < 120.0f)) continue;
        bool ok1 = tptlibm::sinf_glibc(a, &s), ok2 = tptlibm::cosf_glibc(a, &c);
        float rs = sinf(a), rc = cosf(a);
        if (!ok1 || !ok2 || tptlibm::f2u(s) != tptlibm::f2u(rs) || tptlibm::f2u(c) != tptlibm::f2u(rc)) ++bad;
    }
    return bad;
}
// powf(x, y) for all 2^32 x (NaN results compared as "both NaN")
long long check_powf_all_x(float y, int nthreads, unsigned long long* first_bad)
{
    std::atomic<long long> bad(0);
    std::atomic<unsigned long long> fb(~0ull);
    auto work = [&](int t) {
        long long mybad = 0;
        for (uint64_t u = t; u < (1ull << 32); u += nthreads)
        {
            float x = tptlibm::u2f((uint32_t)u);
            float a = tptlibm::powf_glibc(x, y), b = powf(x, y);
            bool same = (tptlibm::f2u(a) == tptlibm::f2u(b)) || (a != a && b != b);
            if (!same) { ++mybad; unsigned long long cur = fb.load(); while (u < cur && !fb.compare_exchange_weak(cur, u)) {} }
        }
        bad += mybad;
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    if (first_bad) *first_bad = fb.load();
    return bad.load();
}
// random (x,y) pairs from a xorshift stream
long long check_powf_random(long long n, uint32_t seed)
{
    long long bad = 0;
    uint32_t s = seed | 1;
    for (long long i = 0; i < n; ++i)
    {
        s ^= s << 13; s ^= s >> 17; s ^= s << 15; uint32_t a = s;
        s ^= s << 13; s ^= s >> 17; s ^= s << 15; uint32_t b = s;
        float x = tptlibm::u2f(a), y = tptlibm::u2f(b);
        float p = tptlibm::powf_glibc(x, y), q = powf(x, y);
        bool same = (tptlibm::f2u(p) == tptlibm::f2u(q)) || (p != p && q != q);
        if (!same) { if (bad < 5) printf("powf mismatch x=%a y=%a got %a want %a\n", x, y, p, q); ++bad; }
    }
    return bad;
}

// powf(x, y) over the 2^24 values RandomFloat01() can return (the REFGPU mode's pow(r, 1.0/3.0), ComputeShader.hlsl:33)
long long check_powf_rand01_domain(float y)
{
    long long bad = 0;
    for (uint32_t k = 0; k < (1u << 24); ++k)
    {
        float x = k / 16777216.0f;
        float a = tptlibm::powf_glibc(x, y), b = powf(x, y);
        if (tptlibm::f2u(a) != tptlibm::f2u(b)) ++bad;
    }
    return bad;
}
}
