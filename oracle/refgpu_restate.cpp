// TEST INFRASTRUCTURE — not product code. Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load the
// library built from this file.
//
// CPU restatement of the reference's GPU compute shader (Cpp/Windows/ComputeShader.hlsl; Cpp/Apple/Shaders.metal is the
// same algorithm) with a runtime scene: the oracle of TPT_MODE_REFGPU. One XorShift32 stream per pixel, analytic
// samplers, <= 10 path segments, front-to-back colour accumulation, lerp() blend, alpha = 1.
//
// PARITY UNPINNED: the reference has no golden vectors for its GPU path and no HLSL/Metal toolchain exists in the build
// image, so this file cannot be checked against the shader itself. It states the shader's source in scalar IEEE float
// arithmetic (every operation rounded separately in source order: -ffp-contract=off; dot(a,b) = (ax*bx + ay*by) + az*bz;
// normalize(v) = v * (1/sqrt(dot(v,v))); reflect(v,n) = v - (2*dot(v,n))*n; lerp(a,b,s) = a + s*(b-a); cos/sin/pow =
// the platform libm's cosf/sinf/powf). A real GPU evaluates these intrinsics with its own approximations, so agreement
// with a D3D11/Metal run can only ever be statistical; what this oracle pins bit for bit is the product's strict
// variant (tpt_refgpu.cuh, EXACT = true). tests/test_oracle.py cross-checks it statistically against the CPU path
// (same estimator up to the depth limit).
//
// Each function cites the shader line it follows (paths relative to /root/reference/Cpp/Windows/ComputeShader.hlsl).
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>
#include <atomic>

namespace {

struct f3 { float x, y, z; };
inline f3 mk(float x, float y, float z) { f3 r = {x, y, z}; return r; }
inline f3 operator+(f3 a, f3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline f3 operator-(f3 a, f3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline f3 operator*(f3 a, f3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
inline f3 operator*(f3 a, float b) { return mk(a.x * b, a.y * b, a.z * b); }
inline f3 operator*(float a, f3 b) { return mk(a * b.x, a * b.y, a * b.z); }
inline f3 neg(f3 a) { return mk(0.0f - a.x, 0.0f - a.y, 0.0f - a.z); }
inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline f3 cross(f3 a, f3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline f3 normalize(f3 v) { return v * (1.0f / sqrtf(dot(v, v))); }
inline f3 reflect(f3 v, f3 n) { return v - (2.0f * dot(v, n)) * n; }
inline f3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }

// :3-16
inline uint32_t RNG(uint32_t& state)
{
    uint32_t x = state;
    x ^= x << 13; x ^= x >> 17; x ^= x << 15;
    state = x;
    return x;
}
inline float RandomFloat01(uint32_t& state) { return (RNG(state) & 0xFFFFFF) / 16777216.0f; }
// :18-24
inline f3 RandomInUnitDisk(uint32_t& state)
{
    float a = RandomFloat01(state) * 2.0f * 3.1415926f;
    float cx = cosf(a), sy = sinf(a);
    float m = sqrtf(RandomFloat01(state));
    return mk(cx * m, sy * m, 0);
}
// :25-35
inline f3 RandomInUnitSphere(uint32_t& state)
{
    float z = RandomFloat01(state) * 2.0f - 1.0f;
    float t = RandomFloat01(state) * 2.0f * 3.1415926f;
    float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
    float x = r * cosf(t);
    float y = r * sinf(t);
    float m = powf(RandomFloat01(state), 1.0f / 3.0f);
    return mk(x * m, y * m, z * m);
}
// :36-44
inline f3 RandomUnitVector(uint32_t& state)
{
    float z = RandomFloat01(state) * 2.0f - 1.0f;
    float a = RandomFloat01(state) * 2.0f * 3.1415926f;
    float r = sqrtf(1.0f - z * z);
    float x = r * cosf(a);
    float y = r * sinf(a);
    return mk(x, y, z);
}
// :57-67
inline bool refract(f3 v, f3 n, float nint, f3& out)
{
    float dt = dot(v, n);
    float discr = 1.0f - nint * nint * (1 - dt * dt);
    if (discr > 0) { out = nint * (v - n * dt) - n * sqrtf(discr); return true; }
    return false;
}
// :68-74
inline float schlick(float cosine, float ri)
{
    float r0 = (1 - ri) / (1 + ri);
    r0 = r0 * r0;
    float x = 1 - cosine;
    x = fminf(fmaxf(x, 0.0f), 1.0f);       // saturate
    return r0 + (1 - r0) * powf(x, 5);
}

struct SphereRaw { float cx, cy, cz, radius, invRadius; };
struct MaterialRaw { int type; float albedo[3]; float emissive[3]; float roughness; float ri; };
struct CameraRaw { float origin[3], llc[3], horizontal[3], vertical[3], u[3], v[3], w[3]; float lensRadius; };
struct Ray { f3 orig, dir; };
struct Hit { f3 pos, normal; float t; };

struct Scene
{
    std::vector<SphereRaw> spheres;
    std::vector<MaterialRaw> mats;
    std::vector<int> emissives;
    CameraRaw cam;
    bool mitsuba;
};

// :121-126
Ray CameraGetRay(const CameraRaw& cam, float s, float t, uint32_t& state)
{
    f3 rd = cam.lensRadius * RandomInUnitDisk(state);
    f3 offset = ld3(cam.u) * rd.x + ld3(cam.v) * rd.y;
    Ray r;
    r.orig = ld3(cam.origin) + offset;
    r.dir = normalize(ld3(cam.llc) + s * ld3(cam.horizontal) + t * ld3(cam.vertical) - ld3(cam.origin) - offset);
    return r;
}

// :129-166
int HitSpheres(const Scene& sc, const Ray& r, float tMin, float tMax, Hit& outHit)
{
    float hitT = tMax;
    int id = -1;
    for (int i = 0; i < (int)sc.spheres.size(); ++i)
    {
        const SphereRaw& s = sc.spheres[i];
        f3 co = mk(s.cx, s.cy, s.cz) - r.orig;
        float nb = dot(co, r.dir);
        float c = dot(co, co) - s.radius * s.radius;
        float discr = nb * nb - c;
        if (discr > 0)
        {
            float discrSq = sqrtf(discr);
            float t = nb - discrSq;
            if (t <= tMin) t = nb + discrSq;
            if (t > tMin && t < hitT) { id = i; hitT = t; }
        }
    }
    if (id != -1)
    {
        const SphereRaw& s = sc.spheres[id];
        outHit.pos = r.orig + r.dir * hitT;
        outHit.normal = (outHit.pos - mk(s.cx, s.cy, s.cz)) * s.invRadius;
        outHit.t = hitT;
    }
    return id;
}

const float kMinT = 0.001f, kMaxT = 1.0e7f;
const int kMaxDepth = 10;

// :181-291
bool Scatter(const Scene& sc, int matID, const Ray& r_in, const Hit& rec, f3& attenuation, Ray& scattered, f3& outLightE,
             long long& rayCount, uint32_t& state)
{
    outLightE = mk(0, 0, 0);
    const MaterialRaw& mat = sc.mats[matID];
    if (mat.type == 0)
    {
        f3 target = rec.pos + rec.normal + RandomUnitVector(state);
        scattered.orig = rec.pos;
        scattered.dir = normalize(target - rec.pos);
        attenuation = ld3(mat.albedo);
        for (size_t j = 0; j < sc.emissives.size(); ++j)
        {
            int i = sc.emissives[j];
            if (matID == i) continue;
            const MaterialRaw& smat = sc.mats[i];
            const SphereRaw& s = sc.spheres[i];
            f3 scn = mk(s.cx, s.cy, s.cz);
            f3 sw = normalize(scn - rec.pos);
            f3 su = normalize(cross(fabsf(sw.x) > 0.01f ? mk(0, 1, 0) : mk(1, 0, 0), sw));
            f3 sv = cross(sw, su);
            float cosAMax = sqrtf(1.0f - s.radius * s.radius / dot(rec.pos - scn, rec.pos - scn));
            float eps1 = RandomFloat01(state), eps2 = RandomFloat01(state);
            float cosA = 1.0f - eps1 + eps1 * cosAMax;
            float sinA = sqrtf(1.0f - cosA * cosA);
            float phi = 2 * 3.1415926f * eps2;
            f3 l = su * cosf(phi) * sinA + sv * sinf(phi) * sinA + sw * cosA;     // :213, left to right
            Hit lightHit;
            ++rayCount;
            Ray sr; sr.orig = rec.pos; sr.dir = l;
            int hitID = HitSpheres(sc, sr, kMinT, kMaxT, lightHit);
            if (hitID == i)
            {
                float omega = 2 * 3.1415926f * (1 - cosAMax);
                f3 nl = dot(rec.normal, r_in.dir) < 0 ? rec.normal : neg(rec.normal);
                outLightE = outLightE + (ld3(mat.albedo) * ld3(smat.emissive)) * (fmaxf(0.0f, dot(l, nl)) * omega / 3.1415926f);
            }
        }
        return true;
    }
    else if (mat.type == 1)
    {
        f3 refl = reflect(r_in.dir, rec.normal);
        float roughness = sc.mitsuba ? 0.0f : mat.roughness;                       // :238-240
        scattered.orig = rec.pos;
        scattered.dir = normalize(refl + roughness * RandomInUnitSphere(state));
        attenuation = ld3(mat.albedo);
        return dot(scattered.dir, rec.normal) > 0;
    }
    else if (mat.type == 2)
    {
        f3 outwardN, rdir = r_in.dir, refl = reflect(rdir, rec.normal), refr = mk(0, 0, 0);
        float nint, reflProb, cosine;
        attenuation = mk(1, 1, 1);
        if (dot(rdir, rec.normal) > 0) { outwardN = neg(rec.normal); nint = mat.ri; cosine = mat.ri * dot(rdir, rec.normal); }
        else { outwardN = rec.normal; nint = 1.0f / mat.ri; cosine = -dot(rdir, rec.normal); }
        if (refract(rdir, outwardN, nint, refr)) reflProb = schlick(cosine, mat.ri);
        else reflProb = 1;
        scattered.orig = rec.pos;
        if (RandomFloat01(state) < reflProb) scattered.dir = normalize(refl);
        else scattered.dir = normalize(refr);
        return true;
    }
    attenuation = mk(1, 0, 1);
    scattered.orig = mk(0, 0, 0); scattered.dir = mk(0, 0, 1);
    return false;
}

// :293-346
f3 Trace(const Scene& sc, Ray r, long long& rayCount, uint32_t& state)
{
    f3 col = mk(0, 0, 0), curAtten = mk(1, 1, 1);
    bool doMaterialE = true;
    for (int depth = 0; depth < kMaxDepth; ++depth)
    {
        Hit rec;
        ++rayCount;
        int id = HitSpheres(sc, r, kMinT, kMaxT, rec);
        if (id >= 0)
        {
            Ray scattered;
            f3 attenuation, lightE;
            const MaterialRaw& mat = sc.mats[id];
            f3 matE = ld3(mat.emissive);
            if (Scatter(sc, id, r, rec, attenuation, scattered, lightE, rayCount, state))
            {
                if (!doMaterialE) matE = mk(0, 0, 0);
                doMaterialE = (mat.type != 0);
                col = col + curAtten * (matE + lightE);
                curAtten = curAtten * attenuation;
                r = scattered;
            }
            else { col = col + curAtten * matE; break; }
        }
        else
        {
            f3 skyCol;
            if (sc.mitsuba) skyCol = mk(0.15f, 0.21f, 0.3f);
            else
            {
                float t = 0.5f * (r.dir.y + 1.0f);
                skyCol = ((1.0f - t) * mk(1.0f, 1.0f, 1.0f) + t * mk(0.5f, 0.7f, 1.0f)) * 0.3f;
            }
            col = col + curAtten * skyCol;
            break;
        }
    }
    return col;
}

} // namespace

extern "C" {

// Frames [frame0, frame0 + nframes) of main() (:353-395) over the whole image, one dispatch per frame like
// Cpp/Windows/TestWin.cpp:253-301 (lerpFac computed by the host there, :271-276). buf: w*h*4 floats, row 0 = gid.y 0,
// read as srcImage and written as dstImage (alpha = 1). flags: kFlagAnimate = 1, kFlagProgressive = 2.
int rgo_render(const float* spheres, const void* mats, int count, const void* cam, int w, int h, int frame0, int nframes,
               unsigned flags, int spp, int mitsuba, float* buf, long long* rays, int nthreads)
{
    Scene sc;
    sc.spheres.resize(count); sc.mats.resize(count);
    memcpy(sc.spheres.data(), spheres, (size_t)count * sizeof(SphereRaw));
    memcpy(sc.mats.data(), mats, (size_t)count * sizeof(MaterialRaw));
    memcpy(&sc.cam, cam, sizeof(CameraRaw));
    sc.mitsuba = mitsuba != 0;
    for (int i = 0; i < count; ++i)
    {
        sc.spheres[i].invRadius = 1.0f / sc.spheres[i].radius;      // UpdateDerivedData, Maths.h:359
        const MaterialRaw& m = sc.mats[i];
        if (m.emissive[0] > 0 || m.emissive[1] > 0 || m.emissive[2] > 0) sc.emissives.push_back(i);
    }
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads <= 0) nthreads = 1;
    const float invWidth = 1.0f / w, invHeight = 1.0f / h;
    for (int f = 0; f < nframes; ++f)
    {
        const int frames = frame0 + f;
        float lerpFac = float(frames) / float(frames + 1);
        if (flags & 1) lerpFac *= 0.9f;
        if (!(flags & 2)) lerpFac = 0;
        std::atomic<int> nextRow(0);
        std::atomic<long long> total(0);
        auto work = [&]() {
            long long mine = 0;
            for (;;)
            {
                int y = nextRow.fetch_add(1);
                if (y >= h) break;
                for (int x = 0; x < w; ++x)
                {
                    long long rayCount = 0;
                    f3 col = mk(0, 0, 0);
                    uint32_t rngState = ((uint32_t)x * 1973u + (uint32_t)y * 9277u + (uint32_t)frames * 26699u) | 1u;   // :380
                    for (int s = 0; s < spp; s++)
                    {
                        float u = float((uint32_t)x + RandomFloat01(rngState)) * invWidth;
                        float v = float((uint32_t)y + RandomFloat01(rngState)) * invHeight;
                        Ray r = CameraGetRay(sc.cam, u, v, rngState);
                        col = col + Trace(sc, r, rayCount, rngState);
                    }
                    col = col * (1.0f / float(spp));
                    float* px = buf + ((size_t)y * w + x) * 4;
                    if (lerpFac != 0.0f)       // lerp(col, prev, 0) == col for finite prev; the product does not read prev then either
                    {
                        f3 prev = mk(px[0], px[1], px[2]);
                        col = col + lerpFac * (prev - col);
                    }
                    px[0] = col.x; px[1] = col.y; px[2] = col.z; px[3] = 1.0f;
                    mine += rayCount;
                }
            }
            total += mine;
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        if (rays) rays[f] = total.load();
    }
    return 0;
}

}
