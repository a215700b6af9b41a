// TEST INFRASTRUCTURE — not product code. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library built from this file.
//
// CPU restatement of the reference hot path with a RUNTIME scene (the reference's scene is a
// compile-time static array, Test.cpp:13-64, so it cannot run the 4096-sphere stress config).
// Plain scalar C++; strict IEEE (build with -ffp-contract=off, no fast-math); libm = the platform's
// glibc sinf/cosf/powf exactly as the reference binary uses them.
// PARITY PIN: tests/test_oracle.py requires this file to be BITWISE equal (all pixels, ray counts) to
// oracle/_ref/libtoyref.so (the unmodified reference) on the reference's own 46-sphere scene, and to the
// golden ray counts of SURVEY.md §9.2 (1280x720 frame0 = 16 809 105 ...).
//
// Each function cites the reference file:line it follows (paths relative to /root/reference/Cpp/Source).
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <thread>
#include <vector>
#include <atomic>
#include <chrono>

namespace {

// Maths.h:9
const float kPI = 3.1415926f;
// Test.cpp:71-73
const float kMinT = 0.001f;
const float kMaxT = 1.0e7f;
const int kMaxDepth = 10;

// Maths.h:250-285 (scalar float3; SURVEY §9.2: bitwise identical to the SSE float3 of Maths.h:23-115)
struct f3 { float x, y, z; };
inline f3 mk(float x, float y, float z) { f3 r = {x, y, z}; return r; }
inline f3 operator+(f3 a, f3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline f3 operator-(f3 a, f3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline f3 operator*(f3 a, f3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
inline f3 operator*(f3 a, float b) { return mk(a.x * b, a.y * b, a.z * b); }
inline f3 operator*(float a, f3 b) { return mk(a * b.x, a * b.y, a * b.z); }
// Maths.h:85 — SSE negation is (0 - a), which keeps +0 for a == +0 (unlike -a).
inline f3 neg(f3 a) { return mk(0.0f - a.x, 0.0f - a.y, 0.0f - a.z); }
// Maths.h:114-115: sum order (x + y) + z
inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// Maths.h:98-105 (SSE form): (a.zxy*b - a*b.zxy).zxy
inline f3 cross(f3 a, f3 b)
{
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// Maths.h:299-301
inline float length(f3 v) { return sqrtf(dot(v, v)); }
inline float sqLength(f3 v) { return dot(v, v); }
inline f3 normalize(f3 v) { return v * (1.0f / length(v)); }
// Maths.h:310-313
inline f3 reflect(f3 v, f3 n) { return v - (2 * dot(v, n)) * n; }
// Maths.h:315-326
inline bool refract(f3 v, f3 n, float nint, f3& outRefracted)
{
    float dt = dot(v, n);
    float discr = 1.0f - nint * nint * (1 - dt * dt);
    if (discr > 0)
    {
        outRefracted = nint * (v - n * dt) - n * sqrtf(discr);
        return true;
    }
    return false;
}
// Maths.h:327-332
inline float schlick(float cosine, float ri)
{
    float r0 = (1 - ri) / (1 + ri);
    r0 = r0 * r0;
    return r0 + (1 - r0) * powf(1 - cosine, 5);
}

// Maths.cpp:5-13
inline uint32_t XorShift32(uint32_t& state)
{
    uint32_t x = state;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 15;
    state = x;
    return x;
}
// Maths.cpp:15-18
inline float RandomFloat01(uint32_t& state) { return (XorShift32(state) & 0xFFFFFF) / 16777216.0f; }
// Maths.cpp:20-28. g++ evaluates the constructor arguments right to left:
// y takes the 1st draw, x the 2nd.
inline f3 RandomInUnitDisk(uint32_t& state)
{
    f3 p;
    do
    {
        float y = RandomFloat01(state);
        float x = RandomFloat01(state);
        p = 2.0f * mk(x, y, 0) - mk(1, 1, 0);
    } while (dot(p, p) >= 1.0);
    return p;
}
// Maths.cpp:30-37; right-to-left: z = 1st draw, y = 2nd, x = 3rd
inline f3 RandomInUnitSphere(uint32_t& state)
{
    f3 p;
    do
    {
        float z = RandomFloat01(state);
        float y = RandomFloat01(state);
        float x = RandomFloat01(state);
        p = 2.0f * mk(x, y, z) - mk(1, 1, 1);
    } while (sqLength(p) >= 1.0);
    return p;
}
// Maths.cpp:39-47
inline f3 RandomUnitVector(uint32_t& state)
{
    float z = RandomFloat01(state) * 2.0f - 1.0f;
    float a = RandomFloat01(state) * 2.0f * kPI;
    float r = sqrtf(1.0f - z * z);
    float x = r * cosf(a);
    float y = r * sinf(a);
    return mk(x, y, z);
}

// Maths.h:334-351
struct Ray { f3 orig, dir; };
struct Hit { f3 pos, normal; float t; };

// Maths.h:354-364 (20 B) and Test.cpp:36-44 (36 B): raw layouts as exported by GetSceneDesc.
struct SphereRaw { float cx, cy, cz, radius, invRadius; };
struct MaterialRaw { int type; float albedo[3]; float emissive[3]; float roughness; float ri; };
// Maths.h:444-449 (88 B)
struct CameraRaw { float origin[3], llc[3], horizontal[3], vertical[3], uu[3], vv[3], ww[3]; float lensRadius; };

struct Scene
{
    int count, simdCount;
    std::vector<float> cx, cy, cz, sqR, invR;   // Maths.h:368-404 SpheresSoA (padded to 4 with "impossible" spheres)
    std::vector<SphereRaw> spheres;
    std::vector<MaterialRaw> mats;
    std::vector<int> emissive;                  // Test.cpp:321-338
    CameraRaw cam;
    int simdTie;                                // 1: SSE HitSpheres semantics (reference build), 0: scalar
    int mitsuba;                                // DO_MITSUBA_COMPARE (Config.h:25) as a runtime switch
};

// Maths.cpp:50-203. Scalar loop (Maths.cpp:165-202) over simdCount spheres *including* the padded
// "impossible" ones, with the SSE build's tie rule (Maths.cpp:113-117,126-152): strict '<' inside a SIMD
// lane (= i % 4), then among lanes holding the global minimum the LOWEST LANE wins.
int HitSpheres(const Ray& r, const Scene& sc, float tMin, float tMax, Hit& outHit)
{
    float hitT = tMax;
    int id = -1;
    int n = sc.simdTie ? sc.simdCount : sc.count;
    for (int i = 0; i < n; ++i)
    {
        float coX = sc.cx[i] - r.orig.x;
        float coY = sc.cy[i] - r.orig.y;
        float coZ = sc.cz[i] - r.orig.z;
        float nb = coX * r.dir.x + coY * r.dir.y + coZ * r.dir.z;
        float c = coX * coX + coY * coY + coZ * coZ - sc.sqR[i];
        float discr = nb * nb - c;
        if (discr > 0)
        {
            float discrSq = sqrtf(discr);
            float t = nb - discrSq;
            if (t <= tMin)
                t = nb + discrSq;
            if (t > tMin)
            {
                bool take = t < hitT;
                if (sc.simdTie && !take && id != -1 && t == hitT && (i & 3) < (id & 3))
                    take = true;
                if (take) { id = i; hitT = t; }
            }
        }
    }
    if (id != -1)
    {
        outHit.pos = r.orig + r.dir * hitT;
        outHit.normal = (outHit.pos - mk(sc.cx[id], sc.cy[id], sc.cz[id])) * sc.invR[id];
        outHit.t = hitT;
    }
    return id;
}

inline f3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }

// Test.cpp:83-193
bool Scatter(const Scene& sc, int matId, const Ray& r_in, const Hit& rec, f3& attenuation, Ray& scattered,
             f3& outLightE, long long& rayCount, uint32_t& state, long long& padHits)
{
    const MaterialRaw& mat = sc.mats[matId];
    outLightE = mk(0, 0, 0);
    if (mat.type == 0) // Lambert, Test.cpp:86-136
    {
        f3 target = rec.pos + rec.normal + RandomUnitVector(state);
        scattered.orig = rec.pos;
        scattered.dir = normalize(target - rec.pos);
        f3 matAlbedo = ld3(mat.albedo);
        attenuation = matAlbedo;
        for (size_t j = 0; j < sc.emissive.size(); ++j)
        {
            int i = sc.emissive[j];
            if (matId == i) continue; // Test.cpp:100 (&mat == &smat)
            const MaterialRaw& smat = sc.mats[i];
            const SphereRaw& s = sc.spheres[i];
            f3 scn = mk(s.cx, s.cy, s.cz);
            f3 sw = normalize(scn - rec.pos);
            f3 su = normalize(cross(fabsf(sw.x) > 0.01f ? mk(0, 1, 0) : mk(1, 0, 0), sw));
            f3 sv = cross(sw, su);
            float cosAMax = sqrtf(1.0f - s.radius * s.radius / sqLength(rec.pos - scn));
            float eps1 = RandomFloat01(state), eps2 = RandomFloat01(state);
            float cosA = 1.0f - eps1 + eps1 * cosAMax;
            float sinA = sqrtf(1.0f - cosA * cosA);
            float phi = 2 * kPI * eps2;
            f3 l = su * (cosf(phi) * sinA) + sv * (sinf(phi) * sinA) + sw * cosA;
            Hit lightHit;
            ++rayCount;
            Ray sr; sr.orig = rec.pos; sr.dir = l;
            int hitID = HitSpheres(sr, sc, kMinT, kMaxT, lightHit);
            if (hitID == i)
            {
                float omega = 2 * kPI * (1 - cosAMax);
                f3 rdir = r_in.dir;
                f3 nl = dot(rec.normal, rdir) < 0 ? rec.normal : neg(rec.normal);
                f3 smatEmissive = ld3(smat.emissive);
                float d = dot(l, nl);
                float m = (0.0f < d) ? d : 0.0f; // std::max(0.0f, d)
                outLightE = outLightE + (matAlbedo * smatEmissive) * (m * omega / kPI);
            }
        }
        return true;
    }
    else if (mat.type == 1) // Metal, Test.cpp:137-150
    {
        f3 refl = reflect(r_in.dir, rec.normal);
        float roughness = mat.roughness;
        if (sc.mitsuba) roughness = 0;          // Test.cpp:143-145
        scattered.orig = rec.pos;
        scattered.dir = normalize(refl + roughness * RandomInUnitSphere(state));
        attenuation = ld3(mat.albedo);
        return dot(scattered.dir, rec.normal) > 0;
    }
    else if (mat.type == 2) // Dielectric, Test.cpp:151-186
    {
        f3 outwardN;
        f3 rdir = r_in.dir;
        f3 refl = reflect(rdir, rec.normal);
        float nint;
        attenuation = mk(1, 1, 1);
        f3 refr = mk(0, 0, 0);
        float reflProb;
        float cosine;
        if (dot(rdir, rec.normal) > 0)
        {
            outwardN = neg(rec.normal);
            nint = mat.ri;
            cosine = mat.ri * dot(rdir, rec.normal);
        }
        else
        {
            outwardN = rec.normal;
            nint = 1.0f / mat.ri;
            cosine = -dot(rdir, rec.normal);
        }
        if (refract(rdir, outwardN, nint, refr))
            reflProb = schlick(cosine, mat.ri);
        else
            reflProb = 1;
        scattered.orig = rec.pos;
        if (RandomFloat01(state) < reflProb)
            scattered.dir = normalize(refl);
        else
            scattered.dir = normalize(refr);
    }
    else
    {
        attenuation = mk(1, 0, 1);
        return false;
    }
    return true;
}

// Test.cpp:195-234. The recursion is unrolled into an explicit stack so the colour is still folded
// back-to-front exactly as `matE + lightE + attenuation * Trace(...)` does (Test.cpp:216).
f3 Trace(const Scene& sc, Ray r, long long& rayCount, uint32_t& state, long long& padHits)
{
    f3 e[kMaxDepth + 1], a[kMaxDepth + 1];
    int n = 0;
    bool doMaterialE = true;
    f3 result;
    for (int depth = 0;; ++depth)
    {
        Hit rec;
        ++rayCount;
        int id = HitSpheres(r, sc, kMinT, kMaxT, rec);
        if (id >= sc.count) ++padHits;
        if (id >= 0)
        {
            Ray scattered;
            f3 attenuation, lightE;
            // id >= count: the ray "hit" a padded impossible sphere (Maths.h:381-387; possible through
            // rounding when the ray points almost exactly at (10000,10000,10000)). The reference then reads
            // s_SphereMats[count] out of bounds (Test.cpp:205) — undefined behaviour. In the reference build
            // made by oracle/Makefile those bytes are .bss alignment padding (type = 0 = Lambert,
            // albedo.x = 0) followed by heap-pointer bits of s_SpheresSoA (ASLR-dependent albedo.yz /
            // emissive). We restate the CONTROL FLOW the reference binary takes (a Lambert scatter, so the RNG
            // stream and ray counts stay identical) with albedo = emissive = 0; the colour of such a pixel is
            // not reproducible by the reference itself. mats[count] holds that all-zero material.
            const MaterialRaw& mat = sc.mats[id >= sc.count ? sc.count : id];
            if (id >= sc.count) id = sc.count;
            f3 matE = ld3(mat.emissive);
            if (depth < kMaxDepth && Scatter(sc, id, r, rec, attenuation, scattered, lightE, rayCount, state, padHits))
            {
                if (!doMaterialE) matE = mk(0, 0, 0);
                doMaterialE = (mat.type != 0);
                e[n] = matE + lightE;
                a[n] = attenuation;
                ++n;
                r = scattered;
                continue;
            }
            result = matE; // Test.cpp:218-221
            break;
        }
        else
        {
            // sky, Test.cpp:224-232
            if (sc.mitsuba) { result = mk(0.15f, 0.21f, 0.3f); break; }   // Test.cpp:226-227
            float t = 0.5f * (r.dir.y + 1.0f);
            result = ((1.0f - t) * mk(1.0f, 1.0f, 1.0f) + t * mk(0.5f, 0.7f, 1.0f)) * 0.3f;
            break;
        }
    }
    for (int k = n - 1; k >= 0; --k)
        result = e[k] + a[k] * result;
    return result;
}

// Maths.h:437-442
Ray GetRay(const CameraRaw& c, float s, float t, uint32_t& state)
{
    f3 rd = c.lensRadius * RandomInUnitDisk(state);
    f3 offset = ld3(c.uu) * rd.x + ld3(c.vv) * rd.y;
    Ray r;
    r.orig = ld3(c.origin) + offset;
    r.dir = normalize(ld3(c.llc) + s * ld3(c.horizontal) + t * ld3(c.vertical) - ld3(c.origin) - offset);
    return r;
}

// Test.cpp:266-300, one row
struct PadLog { std::atomic<int> n; int cap; int* xyf; };

void TraceRow(const Scene& sc, int y, int frameCount, int w, int h, unsigned flags, int spp, float* backbuffer,
              long long& rayCountOut, long long& padHits, PadLog* padLog)
{
    float invWidth = 1.0f / w;
    float invHeight = 1.0f / h;
    float lerpFac = float(frameCount) / float(frameCount + 1);
    if (flags & 1) lerpFac *= 0.9f;      // kFlagAnimate, DO_ANIMATE_SMOOTHING (Config.h:23)
    if (!(flags & 2)) lerpFac = 0;       // !kFlagProgressive
    long long rayCount = 0;
    uint32_t state = ((uint32_t)y * 9781u + (uint32_t)frameCount * 6271u) | 1u; // Test.cpp:280
    float* bb = backbuffer + (size_t)y * w * 4;
    for (int x = 0; x < w; ++x)
    {
        f3 col = mk(0, 0, 0);
        long long padBefore = padHits;
        for (int s = 0; s < spp; s++)
        {
            float u = float(x + RandomFloat01(state)) * invWidth;
            float v = float((uint32_t)y + RandomFloat01(state)) * invHeight;
            Ray r = GetRay(sc.cam, u, v, state);
            col = col + Trace(sc, r, rayCount, state, padHits);
        }
        if (padHits != padBefore && padLog)
        {
            int k = padLog->n.fetch_add(1);
            if (k < padLog->cap) { padLog->xyf[3 * k] = x; padLog->xyf[3 * k + 1] = y; padLog->xyf[3 * k + 2] = frameCount; }
        }
        col = col * (1.0f / float(spp));
        f3 prev = mk(bb[0], bb[1], bb[2]);
        col = prev * lerpFac + col * (1 - lerpFac);
        bb[0] = col.x; bb[1] = col.y; bb[2] = col.z;
        bb += 4;
    }
    rayCountOut += rayCount;
}

} // namespace

static int g_mitsuba = 0;

extern "C" {

// DO_MITSUBA_COMPARE for the following orc_render* calls (constant sky, zero Metal roughness; the zero aperture is camera
// data supplied by the caller).
void orc_set_mitsuba(int on) { g_mitsuba = on; }

// spheres: count x {cx,cy,cz,radius,invRadius} (invRadius recomputed like UpdateTest, Test.cpp:325);
// mats: count x 36 B; cam: 88 B. Renders frames [frame0, frame0+nframes) like the reference shells do
// (UpdateTest + DrawTest per frame) into buf (w*h*4, caller-owned, read as `prev`).
// rays[i] = rays of frame i; pad_hits (optional) = number of path rays that hit a padded sphere (reference
// UB, see Trace()); pad_xyf (optional, pad_cap triples) = (x, y, frame) of the pixels those rays belong to.
// orc_render_rows: the same over the rows y_i = row0 + i*row_step, i in [0, num_rows) only (TraceRowJob(start,end)
// generalised the way tpt_draw's row shards are); the other rows of buf are not touched and not counted.
int orc_render_rows(const float* spheres, const void* mats, int count, const void* cam,
               int w, int h, int frame0, int nframes, unsigned flags, int spp, int simd_tie,
               float* buf, long long* rays, long long* pad_hits, double* seconds, int nthreads,
               int* pad_xyf, int pad_cap, int row0, int num_rows, int row_step)
{
    PadLog padLog; padLog.n = 0; padLog.cap = pad_xyf ? pad_cap : 0; padLog.xyf = pad_xyf;
    Scene sc;
    sc.count = count;
    sc.simdCount = (count + 3) / 4 * 4;
    sc.simdTie = simd_tie;
    sc.mitsuba = g_mitsuba;
    sc.cx.assign(sc.simdCount, 10000.0f); sc.cy = sc.cx; sc.cz = sc.cx;
    sc.sqR.assign(sc.simdCount, 0.0f); sc.invR.assign(sc.simdCount, 0.0f);
    sc.spheres.resize(count); sc.mats.resize(count + 1);
    memset(&sc.mats[count], 0, sizeof(MaterialRaw)); // the out-of-bounds "pad" material, see Trace()
    memcpy(sc.spheres.data(), spheres, (size_t)count * sizeof(SphereRaw));
    memcpy(sc.mats.data(), mats, (size_t)count * sizeof(MaterialRaw));
    memcpy(&sc.cam, cam, sizeof(CameraRaw));
    for (int i = 0; i < count; ++i) // Test.cpp:321-339
    {
        SphereRaw& s = sc.spheres[i];
        s.invRadius = 1.0f / s.radius;
        sc.cx[i] = s.cx; sc.cy[i] = s.cy; sc.cz[i] = s.cz;
        sc.sqR[i] = s.radius * s.radius;
        sc.invR[i] = s.invRadius;
        const MaterialRaw& m = sc.mats[i];
        if (m.emissive[0] > 0 || m.emissive[1] > 0 || m.emissive[2] > 0) sc.emissive.push_back(i);
    }
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads <= 0) nthreads = 1;
    long long padTotal = 0;
    for (int f = 0; f < nframes; ++f)
    {
        auto t0 = std::chrono::steady_clock::now();
        std::atomic<int> nextRow(0);
        std::atomic<long long> rc(0), ph(0);
        auto work = [&]() {
            long long myRays = 0, myPad = 0;
            for (;;)
            {
                int i0 = nextRow.fetch_add(4); // Test.cpp:359 min range 4 rows
                if (i0 >= num_rows) break;
                for (int i = i0; i < i0 + 4 && i < num_rows; ++i)
                    TraceRow(sc, row0 + i * row_step, frame0 + f, w, h, flags, spp, buf, myRays, myPad, &padLog);
            }
            rc += myRays; ph += myPad;
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        auto t1 = std::chrono::steady_clock::now();
        if (rays) rays[f] = rc.load();
        if (seconds) seconds[f] = std::chrono::duration<double>(t1 - t0).count();
        padTotal += ph.load();
    }
    if (pad_hits) *pad_hits = padTotal;
    return 0;
}

int orc_render(const float* spheres, const void* mats, int count, const void* cam,
               int w, int h, int frame0, int nframes, unsigned flags, int spp, int simd_tie,
               float* buf, long long* rays, long long* pad_hits, double* seconds, int nthreads,
               int* pad_xyf, int pad_cap)
{
    return orc_render_rows(spheres, mats, count, cam, w, h, frame0, nframes, flags, spp, simd_tie, buf, rays, pad_hits,
                           seconds, nthreads, pad_xyf, pad_cap, 0, h, 1);
}

// libm probes so tests can pin the product's device-side glibc restatement against the very libm
// this oracle (and oracle/_ref) links: out[i] = f(in[i]).
void orc_sinf(const float* in, float* out, long long n) { for (long long i = 0; i < n; ++i) out[i] = sinf(in[i]); }
void orc_cosf(const float* in, float* out, long long n) { for (long long i = 0; i < n; ++i) out[i] = cosf(in[i]); }
void orc_powf(const float* x, const float* y, float* out, long long n) { for (long long i = 0; i < n; ++i) out[i] = powf(x[i], y[i]); }

}
